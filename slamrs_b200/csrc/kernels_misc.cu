// Read-outs (k_export*, slam.rs:83-88), slot initialisation, the simulator's lidar on the device
// (k_sim_scan, simulator/src/sim.rs:134-159), per-device setup and the kernel-level test hooks.
#include "kernels_common.cuh"

namespace slamrs {

// =============================================================================== k_export
// estimated_likelihood (slam.rs:83-88 -> Map::likelihood, map.rs:50-52): hit counters of the
// estimate's grid -> probabilities. Formats: f64 (what GridMapMessage carries, node.rs:68-72),
// f32 (what the visualizer converts to, visualize.rs:247) and u8 (round(255 p)); `win` restricts
// the export to a window of the grid (e.g. the informed extent) to cut the D2H copy.
template <typename T>
__device__ __forceinline__ T export_value(double p);
template <> __device__ __forceinline__ double export_value<double>(double p) { return p; }
template <> __device__ __forceinline__ float export_value<float>(double p) { return (float)p; }
template <> __device__ __forceinline__ uint8_t export_value<uint8_t>(double p) {
    return (uint8_t)__double2int_rn(__dmul_rn(p, 255.0));
}

template <typename T>
__global__ void __launch_bounds__(256)
k_export(const uint32_t* __restrict__ cells, const SlotMeta* __restrict__ meta, size_t cells_per_grid,
         const StepCounters* __restrict__ counters, MapGeom geom, int4 win /* x0, y0, x1, y1 */, T* __restrict__ out) {
    const long long slot = counters->est_slot;
    if (slot < 0) return;  // another GPU owns the estimate
    const uint32_t* grid = cells + (size_t)slot * cells_per_grid;
    const SlotMeta sm = meta[slot];
    const uint32_t ww = (uint32_t)(win.z - win.x), wh = (uint32_t)(win.w - win.y);
    const uint32_t n = ww * wh;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t ry = i / ww, rx = i - ry * ww;
        const int x = win.x + (int)rx, y = win.y + (int)ry;
        // outside the informed extent every cell is the prior (and a windowed slot holds nothing else)
        const bool inside = x >= sm.x0 && x < sm.x1 && y >= sm.y0 && y < sm.y1;
        const uint32_t cell = inside ? grid[phys_index(geom, (uint32_t)x, (uint32_t)y)] : 0u;
        // a never-informed cell is exactly the prior: log-odds 0 -> 1 - 1/(1 + exp(0)) = 0.5
        out[i] = export_value<T>(cell == 0u ? 0.5 : log_odds_probability(cell_log_odds(cell)));  // Map::likelihood
    }
}

void launch_export(cudaStream_t stream, const uint32_t* cells, const SlotMeta* meta, size_t cells_per_grid,
                   const StepCounters* counters, MapGeom geom, int x0, int y0, int x1, int y1, int format, void* out) {
    const uint32_t n = (uint32_t)(x1 - x0) * (uint32_t)(y1 - y0);
    const int blocks = (int)max(1u, min((n + 255u) / 256u, 148u * 8u));
    const int4 win = make_int4(x0, y0, x1, y1);
    if (format == 1) k_export<float><<<blocks, 256, 0, stream>>>(cells, meta, cells_per_grid, counters, geom, win, (float*)out);
    else if (format == 2) k_export<uint8_t><<<blocks, 256, 0, stream>>>(cells, meta, cells_per_grid, counters, geom, win, (uint8_t*)out);
    else k_export<double><<<blocks, 256, 0, stream>>>(cells, meta, cells_per_grid, counters, geom, win, (double*)out);
}

// informed extent of the estimate's grid (empty -> 0,0,0,0)
__global__ void k_estimate_extent(const SlotMeta* __restrict__ meta, const StepCounters* __restrict__ counters, int* out4) {
    const long long slot = counters->est_slot;
    if (slot < 0) { out4[0] = out4[1] = out4[2] = out4[3] = -1; return; }
    const SlotMeta m = meta[slot];
    if (m.x1 <= m.x0 || m.y1 <= m.y0) { out4[0] = out4[1] = out4[2] = out4[3] = 0; return; }
    out4[0] = m.x0; out4[1] = m.y0; out4[2] = m.x1; out4[3] = m.y1;
}
void launch_estimate_extent(cudaStream_t stream, const SlotMeta* meta, const StepCounters* counters, int* out4) {
    k_estimate_extent<<<1, 1, 0, stream>>>(meta, counters, out4);
}

// The step counters reach the host through a kernel that stores them into the page-locked mirror (device-accessible
// under UVA) instead of a device-to-host memcpy: a memcpy would queue on the copy engine behind a pipelined 8 MB map
// read-out, and the host's wait for "step t+1 is done" would then include the copy of map t.
__global__ void k_publish_counters(const StepCounters* __restrict__ counters, StepCounters* host_mirror) {
    static_assert(sizeof(StepCounters) % sizeof(uint32_t) == 0, "copied word by word");
    const uint32_t* src = reinterpret_cast<const uint32_t*>(counters);
    uint32_t* dst = reinterpret_cast<uint32_t*>(host_mirror);
    for (uint32_t i = threadIdx.x; i < sizeof(StepCounters) / sizeof(uint32_t); i += blockDim.x) dst[i] = src[i];
    __threadfence_system();
}
void launch_publish_counters(cudaStream_t stream, const StepCounters* counters, StepCounters* host_mirror) {
    k_publish_counters<<<1, 128, 0, stream>>>(counters, host_mirror);
}

// one slot's grid in logical order
__global__ void __launch_bounds__(256)
k_export_slot(const uint32_t* __restrict__ grid, const SlotMeta* __restrict__ slot_meta, MapGeom geom, bool as_log_odds,
              void* __restrict__ out) {
    const SlotMeta sm = *slot_meta;
    const uint32_t n = geom.gw * geom.gh;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t y = i / geom.gw, x = i - y * geom.gw;
        const bool inside = (int)x >= sm.x0 && (int)x < sm.x1 && (int)y >= sm.y0 && (int)y < sm.y1;
        const uint32_t cell = inside ? grid[phys_index(geom, x, y)] : 0u;
        if (as_log_odds) reinterpret_cast<double*>(out)[i] = cell_log_odds(cell);
        else reinterpret_cast<uint32_t*>(out)[i] = cell;
    }
}
void launch_export_slot(cudaStream_t stream, const uint32_t* grid, const SlotMeta* slot_meta, MapGeom geom,
                        bool as_log_odds, void* out) {
    const uint32_t n = geom.gw * geom.gh;
    const int blocks = (int)min((n + 255u) / 256u, 148u * 8u);
    k_export_slot<<<blocks, 256, 0, stream>>>(grid, slot_meta, geom, as_log_odds, out);
}

// dense logical image -> the cells of its extent inside a (zeroed) slot
__global__ void __launch_bounds__(256)
k_import_slot(const uint32_t* __restrict__ image, uint32_t* __restrict__ grid, SlotMeta m, MapGeom geom) {
    const uint32_t bw = (uint32_t)(m.x1 - m.x0), bh = (uint32_t)(m.y1 - m.y0);
    const uint32_t n = bw * bh;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t y = (uint32_t)m.y0 + i / bw, x = (uint32_t)m.x0 + i % bw;
        grid[phys_index(geom, x, y)] = image[(size_t)y * geom.gw + x];
    }
}
void launch_import_slot(cudaStream_t stream, const uint32_t* image, uint32_t* grid, SlotMeta m, MapGeom geom) {
    const uint32_t n = (uint32_t)max(0, m.x1 - m.x0) * (uint32_t)max(0, m.y1 - m.y0);
    if (n == 0) return;
    k_import_slot<<<(int)min((n + 255u) / 256u, 148u * 8u), 256, 0, stream>>>(image, grid, m, geom);
}

// =============================================================================== init

__global__ void k_init_slots(int32_t* slot_of, uint32_t n_local, int32_t* spare_list, uint32_t n_spare,
                             StepCounters* counters, uint32_t rank, SlotMeta* meta, int32_t* alias_of) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_local + n_spare) {
        meta[i] = SlotMeta{0, 0, 0, 0, 0, 0, 0, 0};   // empty: every cell is prior
        alias_of[i] = (int32_t)i;                     // every grid is private
    }
    if (i < n_local) slot_of[i] = (int32_t)i;
    if (i < n_spare) spare_list[i] = (int32_t)(n_local + i);
    if (i == 0) {
        StepCounters c;
        memset(&c, 0, sizeof(c));
        c.est_slot = rank == 0 ? 0 : -1;  // before the first update: particle 0 (max_particle = 0, particle.rs:26)
        c.n_spare = n_spare;
        *counters = c;
    }
}
void launch_init_slots(cudaStream_t stream, int32_t* slot_of, uint32_t n_local, int32_t* spare_list, uint32_t n_spare,
                       StepCounters* counters, uint32_t rank, SlotMeta* meta, int32_t* alias_of) {
    const uint32_t n = n_local + n_spare;
    k_init_slots<<<(n + 255) / 256, 256, 0, stream>>>(slot_of, n_local, spare_list, n_spare, counters, rank, meta, alias_of);
}

// =============================================================================== k_init_uniform
__global__ void __launch_bounds__(256)
k_init_uniform(uint64_t seed, uint32_t first_particle, uint32_t n_local, float x0, float y0, float x1, float y1,
               float* __restrict__ pose) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_local) return;
    slamrs_stream::uniform_pose(seed, first_particle + i, (double)x0, (double)y0, (double)x1, (double)y1, &pose[3 * i],
                                &pose[3 * i + 1], &pose[3 * i + 2]);
}
void launch_init_uniform(cudaStream_t stream, uint64_t seed, uint32_t first_particle, uint32_t n_local, float x0, float y0,
                         float x1, float y1, float* pose) {
    k_init_uniform<<<(n_local + 255) / 256, 256, 0, stream>>>(seed, first_particle, n_local, x0, y0, x1, y1, pose);
}

// =============================================================================== k_sim_scan
// The simulator's lidar on the device (slamrs/simulator/src/sim.rs:134-159 against the line
// segments of scene/ray.rs:55-83): one thread per beam, nearest hit over all segments, beams whose
// ray hits nothing are dropped (sim.rs:138) by an ordered compaction, so the observation lands in
// the handle's device scan buffers in exactly the order the reference would publish it. f32
// throughout, no contraction, glibc-exact sin/cos: bit-identical to the CPU restatement.
__global__ void __launch_bounds__(1024)
k_sim_scan(const float* __restrict__ segments, uint32_t n_seg, float px, float py, float ptheta, uint32_t n_beams,
           float scanner_range, float* __restrict__ angle, float* __restrict__ dist, uint8_t* __restrict__ valid,
           uint32_t* __restrict__ out_count_maxbits /* [0] = measurements, [1] = bits of the largest distance */) {
    __shared__ uint32_t s_warp[33];
    __shared__ uint32_t s_base;
    if (threadIdx.x == 0) s_base = 0u;
    float maxd = 0.0f;
    __syncthreads();
    for (uint32_t b0 = 0; b0 < n_beams; b0 += blockDim.x) {
        const uint32_t b = b0 + threadIdx.x;
        bool have = false;
        float best = 0.0f, a = 0.0f;
        if (b < n_beams) {
            // (angle as f32).to_radians() = value * (PI_f32 / 180), generalised to 360/n_beams degree steps
            const float deg = __fmul_rn((float)b, __fdiv_rn(360.0f, (float)n_beams));
            a = __fmul_rn(deg, __fdiv_rn(3.14159265358979323846264338327950288f, 180.0f));
            float dx, dy;
            slamrs_libm::sincosf_exact(__fadd_rn(a, ptheta), &dy, &dx);
            const float x3 = px, y3 = py, x4 = __fadd_rn(px, dx), y4 = __fadd_rn(py, dy);
            for (uint32_t k = 0; k < n_seg; ++k) {
                const float x1 = segments[4 * k], y1 = segments[4 * k + 1], x2 = segments[4 * k + 2], y2 = segments[4 * k + 3];
                const float denom = __fsub_rn(__fmul_rn(__fsub_rn(x1, x2), __fsub_rn(y3, y4)),
                                              __fmul_rn(__fsub_rn(y1, y2), __fsub_rn(x3, x4)));
                if (denom == 0.0f) continue;   // parallel
                const float t = __fdiv_rn(__fsub_rn(__fmul_rn(__fsub_rn(x1, x3), __fsub_rn(y3, y4)),
                                                    __fmul_rn(__fsub_rn(y1, y3), __fsub_rn(x3, x4))), denom);
                const float u = __fdiv_rn(-__fsub_rn(__fmul_rn(__fsub_rn(x1, x2), __fsub_rn(y1, y3)),
                                                     __fmul_rn(__fsub_rn(y1, y2), __fsub_rn(x1, x3))), denom);
                if (t >= 0.0f && t <= 1.0f && u > 0.0f) {
                    if (!have || u < best) { best = u; have = true; }   // min_by keeps the earlier element on ties
                }
            }
        }
        uint32_t total;
        const uint32_t pos = s_base + block_excl_scan_u32(have ? 1u : 0u, s_warp, &total);
        if (have) {
            const bool hit = best < scanner_range;
            const float d = hit ? best : scanner_range;
            angle[pos] = a; dist[pos] = d; valid[pos] = hit ? 1 : 0;
            maxd = fmaxf(maxd, fabsf(d));
        }
        __syncthreads();
        if (threadIdx.x == 0) s_base += total;
        __syncthreads();
    }
    atomicMax(&out_count_maxbits[1], __float_as_uint(maxd));   // non-negative floats order like their bits
    if (threadIdx.x == 0) out_count_maxbits[0] = s_base;
}

void launch_sim_scan(cudaStream_t stream, const float* segments, uint32_t n_seg, float px, float py, float ptheta,
                     uint32_t n_beams, float scanner_range, float* angle, float* dist, uint8_t* valid,
                     uint32_t* out_count_maxbits) {
    k_sim_scan<<<1, 1024, 0, stream>>>(segments, n_seg, px, py, ptheta, n_beams, scanner_range, angle, dist, valid,
                                       out_count_maxbits);
}
// =============================================================================== per-device setup

cudaError_t configure_ray_kernels();
cudaError_t configure_resample_kernels();
cudaError_t configure_kernels() {
    cudaError_t e = configure_ray_kernels();
    if (e != cudaSuccess) return e;
    return configure_resample_kernels();
}

// =============================================================================== test hooks

__global__ void k_debug_raycast(const float* x0, const float* y0, const float* x1, const float* y1, uint32_t n_rays,
                                uint32_t gw, uint32_t gh, uint32_t extra, int32_t* out_xy, uint32_t cap,
                                uint32_t* out_count) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rays) return;
    uint32_t count = 0;
    int32_t* o = out_xy + (size_t)r * cap * 2;
    ray_walk(x0[r], y0[r], x1[r], y1[r], gw, gh, extra, [&](int x, int y) {
        if (count < cap) { o[2 * count] = x; o[2 * count + 1] = y; }
        count++;
    });
    out_count[r] = count;
}
void launch_debug_raycast(cudaStream_t stream, const float* x0, const float* y0, const float* x1, const float* y1,
                          uint32_t n_rays, uint32_t gw, uint32_t gh, uint32_t extra, int32_t* out_xy, uint32_t cap,
                          uint32_t* out_count) {
    k_debug_raycast<<<(n_rays + 127) / 128, 128, 0, stream>>>(x0, y0, x1, y1, n_rays, gw, gh, extra, out_xy, cap,
                                                             out_count);
}

__global__ void k_debug_sincos(const float* x, uint32_t n, float* s, float* c) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    // the fused form is what the kernels use; the separate forms must agree with it bit for bit
    float fs, fc;
    slamrs_libm::sincosf_exact(x[i], &fs, &fc);
    const float ss = slamrs_libm::sinf_exact(x[i]), cc = slamrs_libm::cosf_exact(x[i]);
    const bool same = (__float_as_uint(fs) == __float_as_uint(ss) || (fs != fs && ss != ss)) &&
                      (__float_as_uint(fc) == __float_as_uint(cc) || (fc != fc && cc != cc));
    s[i] = same ? fs : __int_as_float(0x7fc00001);
    c[i] = same ? fc : __int_as_float(0x7fc00001);
}
void launch_debug_sincos(cudaStream_t stream, const float* x, uint32_t n, float* s, float* c) {
    k_debug_sincos<<<(n + 255) / 256, 256, 0, stream>>>(x, n, s, c);
}

__global__ void k_debug_stream(uint64_t seed, uint64_t step, uint64_t first, uint64_t count, double* z, double* u) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) slamrs_stream::motion_normals(seed, step, (uint32_t)(first + i), &z[2 * i], &z[2 * i + 1]);
    if (i == 0) *u = slamrs_stream::resample_uniform(seed, step);
}
void launch_debug_stream(cudaStream_t stream, uint64_t seed, uint64_t step, uint64_t first, uint64_t count, double* z,
                         double* u) {
    const uint64_t n = count ? count : 1;
    k_debug_stream<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(seed, step, first, count, z, u);
}

}  // namespace slamrs
