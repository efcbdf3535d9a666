// Hand-written sm_100a kernels of the grid particle-filter SLAM step.
//
//   k_motion / k_likelihood  robot.rs:170-183 (sample), map.rs:113-145 (beam-endpoint likelihood),
//                        robot.rs:152-167 (motion pdf)  -> one thread / one warp per particle
//   k_ray_update         map.rs:71-106 + ray.rs:21-110 + map.rs:148-172 (integrate)
//                                                                  -> one CTA per particle,
//                        hit counters accumulated in a shared-memory window, 128-bit write-back
//   k_weights            particle.rs:49-56 (normalise) + :40-46 (argmax) + the running sum of :85-91
//   k_resample_indices   particle.rs:78-101 (systematic resampling indices)
//   k_plan               turns the index vector into "keep in place / copy / pull" work lists
//   k_copy               particle.rs:97-100 `value.clone()` -> streaming 128-bit grid copies
//   k_export             slam.rs:83-88 / map.rs:50-52 (counters -> probability grid)
//
// No tensor cores anywhere: nothing here is a dense contraction. The step is HBM-bound
// (grid copies) with a latency-bound cell walk in front of it.
#include "kernels.cuh"
#include "shared_stream.cuh"
#include <cooperative_groups.h>

namespace slamrs {

namespace cg = cooperative_groups;

// =============================================================================== helpers

__device__ __forceinline__ uint4 ld_stream_v4(const uint4* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream_v4(uint4* p, const uint4& v) {
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z),
                 "r"(v.w)
                 : "memory");
}

// exclusive prefix sum of one uint32 per thread over a 1024-thread CTA (warp shuffles + one
// shared array of 32 warp totals). Returns the exclusive prefix; *total gets the CTA sum.
__device__ __forceinline__ uint32_t block_excl_scan_u32(uint32_t v, uint32_t* warp_tot /*[33]*/, uint32_t* total) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    __syncthreads();  // protect warp_tot reuse across calls
    if (lane == 31) warp_tot[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        uint32_t w = lane < nw ? warp_tot[lane] : 0u;
        uint32_t winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += t;
        }
        warp_tot[lane] = winc - w;  // exclusive warp offsets
        if (lane == 31) warp_tot[32] = winc;
    }
    __syncthreads();
    *total = warp_tot[32];
    return warp_tot[wid] + inc - v;
}

// same for doubles (fixed combination order => deterministic)
__device__ __forceinline__ double block_excl_scan_f64(double v, double* warp_tot /*[33]*/, double* total) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    double inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc = __dadd_rn(t, inc);
    }
    __syncthreads();
    if (lane == 31) warp_tot[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        const double w = lane < nw ? warp_tot[lane] : 0.0;
        double winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const double t = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc = __dadd_rn(t, winc);
        }
        const double excl = __shfl_up_sync(0xffffffffu, winc, 1);
        warp_tot[lane] = lane == 0 ? 0.0 : excl;
        if (lane == 31) warp_tot[32] = winc;
    }
    __syncthreads();
    *total = warp_tot[32];
    // exclusive prefix of this thread = warp offset + (inclusive - own) within the warp
    const double within = __shfl_up_sync(0xffffffffu, inc, 1);
    return lane == 0 ? warp_tot[wid] : __dadd_rn(warp_tot[wid], within);
}

// =============================================================================== k_motion + k_likelihood
// k_motion: one THREAD per particle. Odometry::sample (robot.rs:170-183) and the motion log-density
// Odometry::probabiliy_of (robot.rs:152-167) need a few hundred scalar f64 operations per particle
// and nothing else; giving them a warp or a CTA would multiply the issued instructions by 32.
// The log-density is parked in ParticleResult::weight until k_likelihood folds it in.
__global__ void __launch_bounds__(128)
k_motion(OdomModel od, const float* __restrict__ pose_cur, const int32_t* __restrict__ slot_of,
         ParticleResult* __restrict__ results, uint32_t first_particle, uint32_t n_local,
         const double* __restrict__ z_draws, uint64_t seed, uint64_t step) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_local) return;
    const uint32_t gp = first_particle + p;      // global logical index
    const float ox = pose_cur[3 * p], oy = pose_cur[3 * p + 1], otheta = pose_cur[3 * p + 2];
    // Odometry::sample. statrs: sample = mean + std_dev * z.
    double z1, z2;
    if (z_draws) {
        z1 = z_draws[2 * (size_t)gp];
        z2 = z_draws[2 * (size_t)gp + 1];
    } else {
        slamrs_stream::motion_normals(seed, step, gp, &z1, &z2);
    }
    const float center_distance = (float)__dadd_rn(od.mean_c, __dmul_rn(od.std_c, z1));
    const float ntheta = __fadd_rn(otheta, (float)__dadd_rn(od.mean_t, __dmul_rn(od.std_t, z2)));
    float sn, cs;
    slamrs_libm::sincosf_exact(ntheta, &sn, &cs);
    const float nx = __fadd_rn(ox, __fmul_rn(cs, center_distance));
    const float ny = __fadd_rn(oy, __fmul_rn(sn, center_distance));
    // Odometry::probabiliy_of(old, new)
    const float dx = __fsub_rn(ox, nx), dy = __fsub_rn(oy, ny);
    const float moved = __fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
    const double ad = angle_diff((double)otheta, (double)ntheta);
    ParticleResult r;
    r.weight = __dadd_rn(log(normal_pdf((double)moved, od.mean_c, od.std_c)), log(normal_pdf(ad, od.mean_t, od.std_t)));
    r.x = nx; r.y = ny; r.theta = ntheta;
    r.slot = slot_of[p];                    // physical slot, read by other ranks' planners
    results[gp] = r;
}

// k_likelihood: one WARP per particle, lanes over beams. Map::probability_of (map.rs:113-145): one
// gather per valid beam from the PRE-update grid. LK_UNROLL gathers are in flight per lane before
// the first exp/log. A never-informed cell (counters 0 -> log-odds 0 -> p = 0.5) contributes
// log(1/1) = 0 and skips the transcendental work. Each lane adds its terms in beam order, the 32
// lane sums are combined by a fixed butterfly: deterministic, order-independent of scheduling.
constexpr int LK_WARPS = 4;
constexpr int LK_UNROLL = 4;

__global__ void __launch_bounds__(LK_WARPS * 32, 2048 / (LK_WARPS * 32))   // every particle of an 8,192-shard resident at once
k_likelihood(MapGeom geom, ScanDevice scan, const uint32_t* __restrict__ cells, const SlotMeta* __restrict__ meta,
             size_t cells_per_grid, ParticleResult* __restrict__ results, uint32_t first_particle, uint32_t n_local,
             const double* __restrict__ term_table,
             ParticleResult* const* __restrict__ peer_results, uint32_t peer_offset, uint32_t rank, uint32_t world) {
    const uint32_t p = blockIdx.x * LK_WARPS + (threadIdx.x >> 5);
    if (p >= n_local) return;
    const int lane = threadIdx.x & 31;
    const ParticleResult r = results[first_particle + p];
    const float nx = r.x, ny = r.y, ntheta = r.theta;
    const uint32_t* grid = cells + (size_t)r.slot * cells_per_grid;
    const int shift = meta[r.slot].ox;   // row rotation of this particle's slot

    double lp = log(1.0);
    for (uint32_t base = 0; base < scan.n_beams; base += 32u * LK_UNROLL) {
        uint32_t cell[LK_UNROLL];
#pragma unroll
        for (int u = 0; u < LK_UNROLL; ++u) {
            const uint32_t b = base + (uint32_t)u * 32u + (uint32_t)lane;
            cell[u] = 0u;
            if (b < scan.n_beams && scan.valid[b]) {
                float ex, ey;
                beam_endpoint(nx, ny, ntheta, scan.angle[b], scan.dist[b], &ex, &ey);
                const float gx = world_to_grid(ex, geom.pos_x, geom.res);
                const float gy = world_to_grid(ey, geom.pos_y, geom.res);
                if (grid_is_valid(gx, gy, geom.gw, geom.gh)) {
                    const size_t column = (size_t)f32_as_usize(gx), row = (size_t)f32_as_usize(gy);
                    // index(): map.rs:201-204, then the slot's row rotation
                    cell[u] = __ldg(&grid[row * geom.gh + phys_col(geom, (uint32_t)column, shift)]);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < LK_UNROLL; ++u) {
            if (cell[u] != 0u) {
                const uint32_t nf = cell[u] & 0xffffu, no = cell[u] >> 16;
                const double term = (nf < LK_TABLE_NF && no < LK_TABLE_NO) ? __ldg(&term_table[nf * LK_TABLE_NO + no])
                                                                           : beam_log_term(cell[u]);
                lp = __dadd_rn(lp, term);
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) lp = __dadd_rn(lp, __shfl_xor_sync(0xffffffffu, lp, o));
    // weight.prob().value(), slam.rs:71: exp(log p(z|x,m) + log p(x'|x,u))
    ParticleResult out = r;
    out.weight = exp(__dadd_rn(lp, r.weight));
    if (lane == 0) results[first_particle + p] = out;
    // The exchange step, fused: lane q stores the finished record straight into GPU q's copy of
    // the population array over NVLink (24 bytes per particle and peer), so that after one
    // peer barrier every GPU holds every particle's weight, pose and slot.
    if (peer_results != nullptr && (uint32_t)lane < world && (uint32_t)lane != rank)
        peer_results[lane][peer_offset + first_particle + p] = out;
}

// =============================================================================== k_peer_barrier
// Stream-ordered barrier across the GPUs of one box through peer-mapped flags: lane q publishes
// this rank's epoch into GPU q's flag array (release, system scope: every write this GPU issued
// before, including the peer stores of earlier kernels in the stream, is visible first) and then
// waits until GPU q's epoch has arrived here (acquire). Epochs only grow, so flags are never reset.
// A bounded wait (timeout_ns, one minute by default: ranks are driven by independent host threads
// that may lag) turns a lost peer into an error instead of a hung GPU.
__global__ void __launch_bounds__(64)
k_peer_barrier(unsigned long long* const* __restrict__ peer_flags, unsigned long long* my_flags, uint32_t rank,
               uint32_t world, unsigned long long epoch, unsigned long long timeout_ns, StepCounters* counters) {
    const uint32_t q = threadIdx.x;
    if (q >= world) return;
    __threadfence_system();
    unsigned long long* theirs = peer_flags[q] + rank;
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(theirs), "l"(epoch) : "memory");
    // once a barrier has given up the handle is poisoned (the error is reported at the next sync):
    // later barriers publish their epoch, so that healthy peers keep going, but do not wait again
    if (*reinterpret_cast<volatile unsigned long long*>(&counters->barrier_timeout) != 0ull) return;
    const unsigned long long* mine = my_flags + q;
    unsigned long long t0, now, seen = 0ull;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(mine) : "memory");
        if (seen >= epoch) break;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        if (now - t0 > timeout_ns) { counters->barrier_timeout = 1ull; break; }
        __nanosleep(200);
    }
    __threadfence_system();
}

void launch_peer_barrier(cudaStream_t stream, unsigned long long* const* peer_flags, unsigned long long* my_flags,
                         uint32_t rank, uint32_t world, unsigned long long epoch, unsigned long long timeout_ns,
                         StepCounters* counters) {
    k_peer_barrier<<<1, 64, 0, stream>>>(peer_flags, my_flags, rank, world, epoch, timeout_ns, counters);
}

void launch_motion_likelihood(cudaStream_t stream, MapGeom geom, OdomModel od, ScanDevice scan,
                              const float* pose_cur, const int32_t* slot_of, const uint32_t* cells,
                              const SlotMeta* meta, size_t cells_per_grid, ParticleResult* results, uint32_t first_particle,
                              uint32_t n_local, const double* z_draws, uint64_t seed, uint64_t step,
                              const double* term_table,
                              ParticleResult* const* peer_results, uint32_t peer_offset, uint32_t rank, uint32_t world) {
    k_motion<<<(n_local + 127u) / 128u, 128, 0, stream>>>(od, pose_cur, slot_of, results, first_particle, n_local,
                                                         z_draws, seed, step);
    k_likelihood<<<(n_local + LK_WARPS - 1) / LK_WARPS, LK_WARPS * 32, 0, stream>>>(geom, scan, cells, meta, cells_per_grid,
                                                                                   results, first_particle, n_local,
                                                                                   term_table, peer_results, peer_offset,
                                                                                   rank, world);
}

__global__ void k_fill_term_table(double* __restrict__ table) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= LK_TABLE_NF * LK_TABLE_NO) return;
    const uint32_t nf = i / LK_TABLE_NO, no = i % LK_TABLE_NO;
    table[i] = (nf | no) ? beam_log_term(nf | (no << 16)) : 0.0;   // entry (0,0) is never read (prior cells are skipped)
}
void launch_fill_term_table(cudaStream_t stream, double* table) {
    k_fill_term_table<<<(LK_TABLE_NF * LK_TABLE_NO + 255) / 256, 256, 0, stream>>>(table);
}

// =============================================================================== k_ray_update

constexpr int RAY_MAX_SMEM = 220 * 1024;  // window budget; the HW limit is 227 KB per CTA

// saturating packed add straight to global memory, for the (rare) cells outside the window
__device__ __forceinline__ void global_cell_add(uint32_t* addr, uint32_t inc, bool* saturated) {
    uint32_t old = *addr;
    for (;;) {
        const uint32_t nv = cell_sat_add(old, inc, saturated);
        if (nv == old) return;
        const uint32_t seen = atomicCAS(addr, old, nv);
        if (seen == old) return;
        old = seen;
    }
}

constexpr int RAY_MAX_THREADS = 512;

// touched-extent bookkeeping of one CTA: s_ext = {xmin, ymin, xmax, ymax} (inclusive cells)
__device__ __forceinline__ void ext_init(int* s_ext) {
    if (threadIdx.x == 0) { s_ext[0] = 0x7fffffff; s_ext[1] = 0x7fffffff; s_ext[2] = -1; s_ext[3] = -1; }
}
__device__ __forceinline__ void ext_add(int* s_ext, int xmin, int ymin, int xmax, int ymax) {
    if (xmax < xmin) return;
    atomicMin(&s_ext[0], xmin); atomicMin(&s_ext[1], ymin);
    atomicMax(&s_ext[2], xmax); atomicMax(&s_ext[3], ymax);
}
// union the CTA's touched extent into the slot's box (x aligned to 8 cells); one thread, after a barrier
constexpr int BOX_ALIGN = 8;   // x alignment of extents in cells: one 256-bit access
// Row rotation of the slot a ray kernel writes: the slot's own once it holds a grid; for an empty slot
// (first scan of a lineage) the one that puts the leftmost cell any ray can reach on a page
// boundary. s_shift[0] receives it; all threads of the CTA call this, with a barrier inside.
__device__ __forceinline__ int ray_slot_shift(const MapGeom& geom, const ScanDevice& scan, const SlotMeta* meta, float px,
                                              float py, float ptheta, int cx0, int* s_shift) {
    const SlotMeta m = *meta;
    const bool empty = m.x1 <= m.x0 || m.y1 <= m.y0;
    if (threadIdx.x == 0) s_shift[0] = empty ? 0x7fffffff : m.ox;
    __syncthreads();
    if (empty && geom.page_cells) {   // uniform over the CTA
        int xmin = cx0;
        for (uint32_t b = threadIdx.x; b < scan.n_beams; b += blockDim.x) {
            float ex, ey;
            beam_endpoint(px, py, ptheta, scan.angle[b], scan.dist[b], &ex, &ey);
            const float gx = floorf(world_to_grid(ex, geom.pos_x, geom.res));
            // a ray reaches at most two cells beyond its endpoint cell (map.rs:97); NaN / far-out -> 0
            const int reach = (gx >= 2.0f && gx < 1.0e6f) ? (int)gx - 2 : 0;
            xmin = min(xmin, reach);
        }
        atomicMin(&s_shift[0], max(0, xmin));
        __syncthreads();
        if (threadIdx.x == 0) s_shift[0] = align_shift(geom, s_shift[0] & ~7);
        __syncthreads();
    } else if (empty) {
        if (threadIdx.x == 0) s_shift[0] = 0;
        __syncthreads();
    }
    return s_shift[0];
}

__device__ __forceinline__ void ext_commit(const int* s_ext, SlotMeta* meta, int gw, int shift) {
    if (s_ext[2] < s_ext[0]) return;
    SlotMeta b = *meta;
    b.ox = shift;
    const int am = BOX_ALIGN - 1;
    const int x0 = s_ext[0] & ~am, x1 = min(gw, (s_ext[2] + 1 + am) & ~am), y0 = s_ext[1], y1 = s_ext[3] + 1;
    if (b.x1 <= b.x0) { b.x0 = x0; b.y0 = y0; b.x1 = x1; b.y1 = y1; }
    else { b.x0 = min(b.x0, x0); b.y0 = min(b.y0, y0); b.x1 = max(b.x1, x1); b.y1 = max(b.y1, y1); }
    *meta = b;
}
constexpr int RAY_MAX_RADIUS = 150;                    // rows of the window: 2 * radius + 1
constexpr int RAY_MAX_ROWS = 2 * RAY_MAX_RADIUS + 1;
constexpr int RAY_WB_BATCH = 6;                        // write-back: global loads in flight per thread

// exact integer square root of a small non-negative integer (same code on host and device so
// that the host's shared-memory sizing and the kernel's row table agree)
__host__ __device__ inline int isqrt_small(int v) {
    int r = (int)sqrtf((float)v);
    while (r * r > v) r--;
    while ((r + 1) * (r + 1) <= v) r++;
    return r;
}

// The window is a DISC of cells around the start cell (rays cannot leave it), stored row by row:
// row dy holds x in [cx - hw, cx + hw], hw = floor(sqrt(R^2 - dy^2)), widened to multiples of 4
// cells for 128-bit write-back. A disc needs pi/4 of the bounding square, which is what lets a
// 6 m range at 5 cm cells (radius 124) fit in one CTA's shared memory.
__host__ __device__ inline int ray_window_cells_upper_bound(int radius, bool vec) {
    int total = 0;
    for (int dy = -radius; dy <= radius; ++dy) total += 2 * isqrt_small(radius * radius - dy * dy) + 1 + (vec ? 6 : 0);
    return total;
}

template <bool kVector>
__global__ void __launch_bounds__(RAY_MAX_THREADS)
k_ray_update(MapGeom geom, ScanDevice scan, const ParticleResult* __restrict__ results, uint32_t first_particle,
             const uint32_t* __restrict__ alive_list,
             const int32_t* __restrict__ slot_of, uint32_t* __restrict__ cells, SlotMeta* __restrict__ meta,
             size_t cells_per_grid, int radius, StepCounters* counters) {
    extern __shared__ __align__(16) uint32_t s_win[];
    __shared__ int s_row_off[RAY_MAX_ROWS + 1];   // first window cell of each row (+ total at [wh])
    __shared__ int s_row_x[RAY_MAX_ROWS];         // x0 | (width << 16)
    __shared__ int s_ext[4];
    __shared__ int s_shift[1];
    if ((unsigned long long)blockIdx.x >= counters->n_alive) return;
    const uint32_t p = alive_list[blockIdx.x];
    ext_init(s_ext);
    const ParticleResult r = results[first_particle + p];
    const float px = r.x, py = r.y, ptheta = r.theta;
    uint32_t* grid = cells + (size_t)slot_of[p] * cells_per_grid;

    // Map::integrate, map.rs:71-73: ray start in grid coordinates
    const float sx = world_to_grid(px, geom.pos_x, geom.res);
    const float sy = world_to_grid(py, geom.pos_y, geom.res);
    const long long lcx = f32_as_isize(floorf(sx)), lcy = f32_as_isize(floorf(sy));
    // every ray starts in the same cell; outside the grid nothing is emitted (ray.rs:88-92)
    if (lcx < 0 || lcx >= (long long)geom.gw || lcy < 0 || lcy >= (long long)geom.gh) return;
    const int cx = (int)lcx, cy = (int)lcy;
    const int shift = ray_slot_shift(geom, scan, &meta[slot_of[p]], px, py, ptheta, cx, s_shift);

    // ---- row table of the disc window, clipped to the grid
    const int wy0 = max(0, cy - radius), wy1 = min((int)geom.gh, cy + radius + 1);
    const int wh = wy1 - wy0;
    for (int ly = threadIdx.x; ly < wh; ly += blockDim.x) {
        const int dy = wy0 + ly - cy;
        const int hw = isqrt_small(radius * radius - dy * dy);
        int x0 = max(0, cx - hw), x1 = min((int)geom.gw, cx + hw + 1);
        if (kVector) {
            x0 &= ~3;
            x1 = min((int)geom.gw, (x1 + 3) & ~3);
        }
        s_row_x[ly] = x0 | ((x1 - x0) << 16);
    }
    __syncthreads();
    if (threadIdx.x < 32) {  // exclusive prefix sum of the row widths, 32 rows per round
        int carry = 0;
        for (int base = 0; base < wh; base += 32) {
            const int ly = base + (int)threadIdx.x;
            const int w = ly < wh ? (s_row_x[ly] >> 16) : 0;
            int inc = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, inc, o);
                if ((int)threadIdx.x >= o) inc += t;
            }
            if (ly < wh) s_row_off[ly] = carry + inc - w;
            carry += __shfl_sync(0xffffffffu, inc, 31);
        }
        if (threadIdx.x == 0) s_row_off[wh] = carry;
    }
    __syncthreads();
    const int wcells = s_row_off[wh];

    if (kVector) {
        uint4* w4 = reinterpret_cast<uint4*>(s_win);
        for (int i = threadIdx.x; i < (wcells >> 2); i += blockDim.x) w4[i] = make_uint4(0u, 0u, 0u, 0u);
    } else {
        for (int i = threadIdx.x; i < wcells; i += blockDim.x) s_win[i] = 0u;
    }
    __syncthreads();

    bool saturated = false;
    uint32_t spilled = 0;
    for (uint32_t b = threadIdx.x; b < scan.n_beams; b += blockDim.x) {
        const float dist = scan.dist[b];
        float ex, ey;
        beam_endpoint(px, py, ptheta, scan.angle[b], dist, &ex, &ey);
        const float gx = world_to_grid(ex, geom.pos_x, geom.res);
        const float gy = world_to_grid(ey, geom.pos_y, geom.res);
        // measured distance in cells (map.rs:84) and the per-ray form of inverse_sensor_model
        const RayClassifier cls = make_ray_classifier(__fdiv_rn(dist, geom.res), scan.valid[b] != 0);
        // apply_measurement, map.rs:88-106 (additional_steps = 2)
        ray_walk_acc(sx, sy, gx, gy, geom.gw, geom.gh, 2u, [&](int x, int y, float acc) {
            const uint32_t inc = classify_cell(cls, acc);
            if (inc != 0u) {
                const int ly = y - wy0;
                bool in_window = false;
                if ((unsigned)ly < (unsigned)wh) {
                    const int rx = s_row_x[ly];
                    const int lx = x - (rx & 0xffff);
                    if ((unsigned)lx < (unsigned)(rx >> 16)) {
                        atomicAdd(&s_win[s_row_off[ly] + lx], inc);
                        in_window = true;
                    }
                }
                if (!in_window) {  // beyond the window (range larger than shared memory allows)
                    global_cell_add(&grid[(size_t)y * geom.gh + phys_col(geom, (uint32_t)x, shift)], inc, &saturated);
                    ext_add(s_ext, x, y, x, y);
                    spilled++;
                }
            }
        });
    }
    __syncthreads();

    // ---- write-back: grid += window, saturating per 16-bit counter, untouched groups skipped.
    // RAY_WB_BATCH independent global loads are issued per thread before the first dependent store.
    int exmin = 0x7fffffff, eymin = 0x7fffffff, exmax = -1, eymax = -1;   // this thread's touched extent
    if (kVector) {
        const int total4 = wcells >> 2;
        const uint4* win4 = reinterpret_cast<const uint4*>(s_win);
        for (int base = threadIdx.x; base < total4; base += blockDim.x * RAY_WB_BATCH) {
            uint4 d[RAY_WB_BATCH], v[RAY_WB_BATCH];
            uint4* gp[RAY_WB_BATCH];
            bool nz[RAY_WB_BATCH];
#pragma unroll
            for (int j = 0; j < RAY_WB_BATCH; ++j) {
                const int i = base + j * (int)blockDim.x;
                nz[j] = false;
                if (i < total4) {
                    d[j] = win4[i];
                    nz[j] = (d[j].x | d[j].y | d[j].z | d[j].w) != 0u;
                    if (nz[j]) {
                        int lo = 0, hi = wh;   // row containing window cell 4*i
                        while (hi - lo > 1) {
                            const int mid = (lo + hi) >> 1;
                            if (s_row_off[mid] <= 4 * i) lo = mid; else hi = mid;
                        }
                        const int lx = 4 * i - s_row_off[lo];
                        const int gx0 = (s_row_x[lo] & 0xffff) + lx;
                        exmin = min(exmin, gx0); exmax = max(exmax, gx0 + 3);
                        eymin = min(eymin, wy0 + lo); eymax = max(eymax, wy0 + lo);
                        gp[j] = reinterpret_cast<uint4*>(grid + (size_t)(wy0 + lo) * geom.gh + phys_col(geom, (uint32_t)gx0, shift));
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < RAY_WB_BATCH; ++j)
                if (nz[j]) v[j] = *gp[j];
#pragma unroll
            for (int j = 0; j < RAY_WB_BATCH; ++j) {
                if (nz[j]) {
                    v[j].x = cell_sat_add(v[j].x, d[j].x, &saturated);
                    v[j].y = cell_sat_add(v[j].y, d[j].y, &saturated);
                    v[j].z = cell_sat_add(v[j].z, d[j].z, &saturated);
                    v[j].w = cell_sat_add(v[j].w, d[j].w, &saturated);
                    *gp[j] = v[j];
                }
            }
        }
    } else {
        for (int ly = threadIdx.x >> 5; ly < wh; ly += blockDim.x >> 5) {
            const int rx = s_row_x[ly], x0 = rx & 0xffff, w = rx >> 16, off = s_row_off[ly];
            for (int c = threadIdx.x & 31; c < w; c += 32) {
                const uint32_t d = s_win[off + c];
                if (d != 0u) {
                    uint32_t* g = grid + (size_t)(wy0 + ly) * geom.gh + phys_col(geom, (uint32_t)(x0 + c), shift);
                    *g = cell_sat_add(*g, d, &saturated);
                    exmin = min(exmin, x0 + c); exmax = max(exmax, x0 + c);
                    eymin = min(eymin, wy0 + ly); eymax = max(eymax, wy0 + ly);
                }
            }
        }
    }
    ext_add(s_ext, exmin, eymin, exmax, eymax);
    __syncthreads();
    if (threadIdx.x == 0) ext_commit(s_ext, &meta[slot_of[p]], (int)geom.gw, shift);
    if (saturated) atomicAdd(&counters->saturated, 1ull);
    if (spilled) atomicAdd(&counters->spilled, (unsigned long long)spilled);
}

// ------------------------------------------------------------------------------- packed variant
// Same algorithm with a 2-byte window cell: bits 0..10 = free hits (a cell can be crossed by at
// most n_beams <= 2047 rays per scan), bits 11..15 = occupied hits (<= 31; a 32nd hit in one scan
// takes the exact global path). Half the shared memory per particle => two CTAs per SM at a 6 m /
// 5 cm window, which is what hides the latency of the serial cell walk. Free hits (the vast
// majority) are single fire-and-forget shared-memory adds; the walk itself is written with the
// row lookup hoisted to y-steps.
constexpr uint32_t PK_FREE_BITS = 11;
constexpr uint32_t PK_FREE_MASK = (1u << PK_FREE_BITS) - 1u;
constexpr uint32_t PK_OCC_MAX = 31;
constexpr uint32_t RAY_PACKED_MAX_BEAMS = PK_FREE_MASK;

__host__ __device__ inline int ray_window_cells_upper_bound_packed(int radius) {
    int total = 0;
    for (int dy = -radius; dy <= radius; ++dy) total += 2 * isqrt_small(radius * radius - dy * dy) + 1 + 14;
    return total;
}

// Facts the packed kernel's fast walk relies on:
//  * |x0 - centre_x| and |y0 - centre_y| never decrease along a walk (x only moves by x_inc, y
//    only by y_inc, away from the start cell), IEEE rounding is monotone, so
//    acc = fl(fl(dx^2) + fl(dy^2)) is non-decreasing along the ray. The inverse sensor model is
//    therefore a free run (acc < free_below), then an occupied run (acc <= prior_above, hits
//    only), then prior cells that add nothing: two tight loops and an early exit replace the
//    per-cell three-way classification.
//  * a ray whose endpoint cell is (ax, ay) cells away from the start cell visits only cells within
//    (ax + 2, ay + 2) of it; if that corner is inside the disc window and the disc is inside the
//    grid, no per-cell window or grid test is needed.
__global__ void __maxnreg__(80)   // 2 CTAs of 384 threads per SM; blocks have at most RAY_MAX_THREADS threads
k_ray_update_packed(MapGeom geom, ScanDevice scan, const ParticleResult* __restrict__ results, uint32_t first_particle,
             const uint32_t* __restrict__ alive_list,
                    const int32_t* __restrict__ slot_of, uint32_t* __restrict__ cells, SlotMeta* __restrict__ meta,
                    size_t cells_per_grid, int radius, StepCounters* counters) {
    extern __shared__ __align__(16) uint32_t s_win[];   // two 16-bit cells per word
    __shared__ int2 s_row[RAY_MAX_ROWS + 1];            // .x = first window cell of the row, .y = x0 | width << 16
    __shared__ uint32_t s_rowb[RAY_MAX_ROWS];           // shared byte address of column x = 0 of the row
    __shared__ int s_ext[4];
    __shared__ int s_shift[1];
    if ((unsigned long long)blockIdx.x >= counters->n_alive) return;
    const uint32_t p = alive_list[blockIdx.x];
    ext_init(s_ext);
    const ParticleResult r = results[first_particle + p];
    const float px = r.x, py = r.y, ptheta = r.theta;
    uint32_t* grid = cells + (size_t)slot_of[p] * cells_per_grid;

    const float sx = world_to_grid(px, geom.pos_x, geom.res);
    const float sy = world_to_grid(py, geom.pos_y, geom.res);
    const long long lcx = f32_as_isize(floorf(sx)), lcy = f32_as_isize(floorf(sy));
    if (lcx < 0 || lcx >= (long long)geom.gw || lcy < 0 || lcy >= (long long)geom.gh) return;
    const int cx0 = (int)lcx, cy0 = (int)lcy;
    const int gw = (int)geom.gw, gh = (int)geom.gh;
    const uint32_t win_base = (uint32_t)__cvta_generic_to_shared(s_win);
    const int slot_shift = ray_slot_shift(geom, scan, &meta[slot_of[p]], px, py, ptheta, cx0, s_shift);

    // ---- row table of the disc window (x ranges aligned to 8 cells = one 128-bit group)
    const int wy0 = max(0, cy0 - radius), wy1 = min(gh, cy0 + radius + 1);
    const int wh = wy1 - wy0;
    for (int ly = threadIdx.x; ly < wh; ly += blockDim.x) {
        const int dy = wy0 + ly - cy0;
        const int hw = isqrt_small(radius * radius - dy * dy);
        const int x0 = max(0, cx0 - hw) & ~7;
        const int x1 = min(gw, (min(gw, cx0 + hw + 1) + 7) & ~7);
        s_row[ly].y = x0 | ((x1 - x0) << 16);
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        int carry = 0;
        for (int base = 0; base < wh; base += 32) {
            const int ly = base + (int)threadIdx.x;
            const int w = ly < wh ? (s_row[ly].y >> 16) : 0;
            int inc = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, inc, o);
                if ((int)threadIdx.x >= o) inc += t;
            }
            if (ly < wh) {
                const int first = carry + inc - w;
                s_row[ly].x = first;
                s_rowb[ly] = win_base + 2u * (uint32_t)(first - (s_row[ly].y & 0xffff));
            }
            carry += __shfl_sync(0xffffffffu, inc, 31);
        }
        if (threadIdx.x == 0) s_row[wh] = make_int2(carry, 0);
    }
    __syncthreads();
    const int wcells = s_row[wh].x;
    {
        uint4* w4 = reinterpret_cast<uint4*>(s_win);
        for (int i = threadIdx.x; i < (wcells >> 3); i += blockDim.x) w4[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    __syncthreads();

    const bool disc_in_grid = cx0 - radius >= 0 && cx0 + radius < gw && cy0 - radius >= 0 && cy0 + radius < gh;
    bool saturated = false;
    uint32_t spilled = 0;
    for (uint32_t b = threadIdx.x; b < scan.n_beams; b += blockDim.x) {
        const float dist = scan.dist[b];
        float ex, ey;
        beam_endpoint(px, py, ptheta, scan.angle[b], dist, &ex, &ey);
        const float x1 = world_to_grid(ex, geom.pos_x, geom.res);
        const float y1 = world_to_grid(ey, geom.pos_y, geom.res);
        const RayClassifier cls = make_ray_classifier(__fdiv_rn(dist, geom.res), scan.valid[b] != 0);

        // GridRayIterator::new (ray.rs:21-77) -- identical arithmetic to ray_walk_acc
        const float delta_x = fabsf(__fsub_rn(x1, sx)), delta_y = fabsf(__fsub_rn(y1, sy));
        const float fx0 = floorf(sx), fy0 = floorf(sy);
        unsigned long long n = 1ull + 2ull;   // additional_steps = 2 (map.rs:97)
        unsigned long long ax = 0ull, ay = 0ull;   // |endpoint cell - start cell| per axis
        int x_inc, y_inc;
        float error;
        if (delta_x == 0.0f) {
            x_inc = 0;
            error = __int_as_float(0x7f800000);
        } else if (x1 > sx) {
            x_inc = 1;
            ax = (unsigned long long)f32_as_isize(__fsub_rn(floorf(x1), (float)cx0));
            error = __fmul_rn(__fsub_rn(__fadd_rn(fx0, 1.0f), sx), delta_y);
        } else {
            x_inc = -1;
            ax = (unsigned long long)(long long)cx0 - (unsigned long long)f32_as_isize(floorf(x1));
            error = __fmul_rn(__fsub_rn(sx, fx0), delta_y);
        }
        if (delta_y == 0.0f) {
            y_inc = 0;
            error = __fsub_rn(error, __int_as_float(0x7f800000));
        } else if (y1 > sy) {
            y_inc = 1;
            ay = (unsigned long long)f32_as_isize(floorf(y1)) - (unsigned long long)(long long)cy0;
            error = __fsub_rn(error, __fmul_rn(__fsub_rn(__fadd_rn(fy0, 1.0f), sy), delta_x));
        } else {
            y_inc = -1;
            ay = (unsigned long long)(long long)cy0 - (unsigned long long)f32_as_isize(floorf(y1));
            error = __fsub_rn(error, __fmul_rn(__fsub_rn(sy, fy0), delta_x));
        }
        n += ax + ay;   // wrapping isize arithmetic, then `as usize`
        const unsigned long long cap = (unsigned long long)geom.gw + geom.gh + 8ull;
        int remaining = (int)(n < cap ? n : cap);

        const float x_step = (float)x_inc, y_step = (float)y_inc;
        float cxf = __fadd_rn((float)cx0, 0.5f), cyf = __fadd_rn((float)cy0, 0.5f);
        float dxs = __fsub_rn(sx, cxf), dys = __fsub_rn(sy, cyf);
        float dx2 = __fmul_rn(dxs, dxs), dy2 = __fmul_rn(dys, dys);

        // every cell of this ray inside the window and the grid?
        bool fast = false;
        if (disc_in_grid && ax < 4096ull && ay < 4096ull) {
            const int cxa = (int)ax + 2, cya = (int)ay + 2;
            fast = cxa * cxa + cya * cya <= radius * radius;
        }
        if (fast) {
            uint32_t x2 = 2u * (uint32_t)cx0;         // twice the current column
            const uint32_t x_inc2 = (uint32_t)(2 * x_inc);
            int ly = cy0 - wy0;
            uint32_t rb = s_rowb[ly];
            // free run
            while (remaining > 0) {
                const float acc = __fadd_rn(dx2, dy2);
                if (!(acc < cls.free_below)) break;
                const uint32_t c2 = rb + x2;           // shared byte address of the 16-bit window cell
                asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(c2 & ~3u), "r"((c2 & 2u) ? 0x10000u : 1u) : "memory");
                if (error > 0.0f) {
                    error = __fsub_rn(error, delta_x);
                    cyf = __fadd_rn(cyf, y_step);
                    dys = __fsub_rn(sy, cyf);
                    dy2 = __fmul_rn(dys, dys);
                    ly += y_inc;
                    rb = s_rowb[ly];
                } else {
                    error = __fadd_rn(error, delta_y);
                    cxf = __fadd_rn(cxf, x_step);
                    dxs = __fsub_rn(sx, cxf);
                    dx2 = __fmul_rn(dxs, dxs);
                    x2 += x_inc2;
                }
                remaining -= 1;
            }
            // occupied run (hits only): bounded 5-bit field per scan, exact global path beyond it
            if (cls.mid_inc != 0u) {
                while (remaining > 0) {
                    const float acc = __fadd_rn(dx2, dy2);
                    if (acc > cls.prior_above) break;
                    const uint32_t c2 = rb + x2;
                    const uint32_t shift = ((c2 & 2u) << 3) + PK_FREE_BITS;
                    uint32_t old;
                    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(old) : "r"(c2 & ~3u) : "memory");
                    bool done = false;
                    for (;;) {
                        if (((old >> shift) & PK_OCC_MAX) == PK_OCC_MAX) break;
                        uint32_t seen;
                        asm volatile("atom.shared.cas.b32 %0, [%1], %2, %3;"
                                     : "=r"(seen) : "r"(c2 & ~3u), "r"(old), "r"(old + (1u << shift)) : "memory");
                        if (seen == old) { done = true; break; }
                        old = seen;
                    }
                    if (!done) {
                        const int x = (int)(x2 >> 1), y = wy0 + ly;
                        global_cell_add(&grid[(size_t)y * geom.gh + phys_col(geom, (uint32_t)x, slot_shift)], CELL_OCC_INC, &saturated);
                        ext_add(s_ext, x, y, x, y);
                        spilled++;
                    }
                    if (error > 0.0f) {
                        error = __fsub_rn(error, delta_x);
                        cyf = __fadd_rn(cyf, y_step);
                        dys = __fsub_rn(sy, cyf);
                        dy2 = __fmul_rn(dys, dys);
                        ly += y_inc;
                        rb = s_rowb[ly];
                    } else {
                        error = __fadd_rn(error, delta_y);
                        cxf = __fadd_rn(cxf, x_step);
                        dxs = __fsub_rn(sx, cxf);
                        dx2 = __fmul_rn(dxs, dxs);
                        x2 += x_inc2;
                    }
                    remaining -= 1;
                }
            }
            continue;
        }

        // general walk: per-cell window and grid tests (rays that may leave the window or the grid)
        int x = cx0, y = cy0;
        int ly = y - wy0;
        int2 row = s_row[ly];                      // the start cell is always inside the window
        int lx = x - (row.y & 0xffff);
        int row_w = row.y >> 16;
        bool inside = true;
        while (remaining > 0 && inside) {
            const float acc = __fadd_rn(dx2, dy2);
            const bool is_free = acc < cls.free_below;
            const bool is_mid = !is_free && !(acc > cls.prior_above) && (cls.mid_inc != 0u);
            if (is_free | is_mid) {
                const bool in_win = (unsigned)lx < (unsigned)row_w;
                const int cell = row.x + lx;
                const uint32_t addr = win_base + ((uint32_t)(cell >> 1) << 2);
                const uint32_t shift = (cell & 1) << 4;
                if (is_free & in_win) {
                    asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(addr), "r"(1u << shift) : "memory");
                } else {
                    bool done = false;
                    if (in_win) {   // occupied hit: bounded 5-bit field, exact path when it would overflow
                        uint32_t* wp = &s_win[cell >> 1];
                        uint32_t old = *wp;
                        for (;;) {
                            if (((old >> (shift + PK_FREE_BITS)) & PK_OCC_MAX) == PK_OCC_MAX) break;
                            const uint32_t seen = atomicCAS(wp, old, old + (1u << (shift + PK_FREE_BITS)));
                            if (seen == old) { done = true; break; }
                            old = seen;
                        }
                    }
                    if (!done) {
                        global_cell_add(&grid[(size_t)y * geom.gh + phys_col(geom, (uint32_t)x, slot_shift)],
                                        is_free ? CELL_FREE_INC : CELL_OCC_INC, &saturated);
                        ext_add(s_ext, x, y, x, y);
                        spilled++;
                    }
                }
            }
            // GridRayIterator::next (ray.rs:96-104)
            if (error > 0.0f) {
                y += y_inc;
                error = __fsub_rn(error, delta_x);
                cyf = __fadd_rn(cyf, y_step);
                dys = __fsub_rn(sy, cyf);
                dy2 = __fmul_rn(dys, dys);
                inside = (unsigned)y < (unsigned)gh;
                ly += y_inc;
                if ((unsigned)ly < (unsigned)wh) {
                    row = s_row[ly];
                    row_w = row.y >> 16;
                    lx = x - (row.y & 0xffff);
                } else {
                    row_w = 0;
                }
            } else {
                x += x_inc;
                error = __fadd_rn(error, delta_y);
                cxf = __fadd_rn(cxf, x_step);
                dxs = __fsub_rn(sx, cxf);
                dx2 = __fmul_rn(dxs, dxs);
                inside = (unsigned)x < (unsigned)gw;
                lx += x_inc;
            }
            remaining -= 1;
        }
    }
    __syncthreads();

    // ---- write-back, row by row: one warp per window row, one lane per 8-cell group (a 128-bit
    // shared load -> two 128-bit global RMWs); RAY_WB_ROWS rows are in flight per warp.
    constexpr int RAY_WB_ROWS = 4;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
    const uint4* win4 = reinterpret_cast<const uint4*>(s_win);
    int exmin = 0x7fffffff, eymin = 0x7fffffff, exmax = -1, eymax = -1;   // this thread's touched extent
    for (int ly0 = warp; ly0 < wh; ly0 += n_warps * RAY_WB_ROWS) {
        for (int g0 = 0; g0 < (RAY_MAX_RADIUS * 2 + 16) / 8; g0 += 32) {
            uint4 d[RAY_WB_ROWS], va[RAY_WB_ROWS], vb[RAY_WB_ROWS];
            uint4* gp[RAY_WB_ROWS];
            bool nz[RAY_WB_ROWS];
            bool any_row = false;
#pragma unroll
            for (int j = 0; j < RAY_WB_ROWS; ++j) {
                const int ly = ly0 + j * n_warps;
                nz[j] = false;
                if (ly < wh) {
                    const int2 row = s_row[ly];
                    const int groups = row.y >> 19;          // width / 8
                    const int g = g0 + lane;
                    any_row |= g0 < groups;
                    if (g < groups) {
                        d[j] = win4[(row.x >> 3) + g];
                        nz[j] = (d[j].x | d[j].y | d[j].z | d[j].w) != 0u;
                        if (nz[j]) {
                            const int gx0 = (row.y & 0xffff) + 8 * g;
                            exmin = min(exmin, gx0); exmax = max(exmax, gx0 + 7);
                            eymin = min(eymin, wy0 + ly); eymax = max(eymax, wy0 + ly);
                            gp[j] = reinterpret_cast<uint4*>(grid + (size_t)(wy0 + ly) * geom.gh +
                                                             phys_col(geom, (uint32_t)gx0, slot_shift));
                        }
                    }
                }
            }
            if (!any_row) break;   // warp-uniform: no row of this batch reaches group g0
#pragma unroll
            for (int j = 0; j < RAY_WB_ROWS; ++j)
                if (nz[j]) { va[j] = gp[j][0]; vb[j] = gp[j][1]; }
#pragma unroll
            for (int j = 0; j < RAY_WB_ROWS; ++j) {
                if (nz[j]) {
                    // group-level fast path: no occupied hit among the 8 window cells (a packed value is
                    // then the free count itself) and no counter of the 8 grid cells at or above 2^15
                    // (it cannot saturate by one scan's increment): eight plain adds
                    const uint32_t occ_any = (d[j].x | d[j].y | d[j].z | d[j].w) & 0xF800F800u;
                    const uint32_t high_any = (va[j].x | va[j].y | va[j].z | va[j].w | vb[j].x | vb[j].y | vb[j].z | vb[j].w) & 0x80008000u;
                    if ((occ_any | high_any) == 0u) {
                        va[j].x += d[j].x & 0xffffu; va[j].y += d[j].x >> 16;
                        va[j].z += d[j].y & 0xffffu; va[j].w += d[j].y >> 16;
                        vb[j].x += d[j].z & 0xffffu; vb[j].y += d[j].z >> 16;
                        vb[j].z += d[j].w & 0xffffu; vb[j].w += d[j].w >> 16;
                    } else {
                        auto apply = [&](uint32_t g, uint32_t packed16) {
                            const uint32_t delta = (packed16 & PK_FREE_MASK) | ((packed16 >> PK_FREE_BITS) << 16);
                            return cell_sat_add(g, delta, &saturated);
                        };
                        va[j].x = apply(va[j].x, d[j].x & 0xffffu); va[j].y = apply(va[j].y, d[j].x >> 16);
                        va[j].z = apply(va[j].z, d[j].y & 0xffffu); va[j].w = apply(va[j].w, d[j].y >> 16);
                        vb[j].x = apply(vb[j].x, d[j].z & 0xffffu); vb[j].y = apply(vb[j].y, d[j].z >> 16);
                        vb[j].z = apply(vb[j].z, d[j].w & 0xffffu); vb[j].w = apply(vb[j].w, d[j].w >> 16);
                    }
                    gp[j][0] = va[j];
                    gp[j][1] = vb[j];
                }
            }
        }
    }
    ext_add(s_ext, exmin, eymin, exmax, eymax);
    __syncthreads();
    if (threadIdx.x == 0) ext_commit(s_ext, &meta[slot_of[p]], (int)geom.gw, slot_shift);
    if (saturated) atomicAdd(&counters->saturated, 1ull);
    if (spilled) atomicAdd(&counters->spilled, (unsigned long long)spilled);
}

cudaError_t launch_ray_update(cudaStream_t stream, MapGeom geom, ScanDevice scan, const ParticleResult* results,
                              uint32_t first_particle, uint32_t n_local, const uint32_t* alive_list,
                              const int32_t* slot_of, uint32_t* cells, SlotMeta* meta,
                              size_t cells_per_grid, int radius_cells, StepCounters* counters,
                              uint64_t* window_cells, bool force_generic) {
    int threads = (int)((scan.n_beams + 31u) / 32u * 32u);
    threads = threads < 128 ? 128 : (threads > RAY_MAX_THREADS ? RAY_MAX_THREADS : threads);
    // preferred: the packed 16-bit window (two CTAs per SM at long range)
    if (!force_generic && geom.gw % 8u == 0u && cells_per_grid % 8u == 0u && scan.n_beams <= RAY_PACKED_MAX_BEAMS) {
        int radius = radius_cells < 1 ? 1 : (radius_cells > RAY_MAX_RADIUS ? RAY_MAX_RADIUS : radius_cells);
        while (radius > 1 && (size_t)ray_window_cells_upper_bound_packed(radius) * 2 > (size_t)RAY_MAX_SMEM) radius--;
        const size_t wmax = (size_t)ray_window_cells_upper_bound_packed(radius);
        *window_cells = wmax;
        k_ray_update_packed<<<n_local, threads, wmax * 2, stream>>>(geom, scan, results, first_particle, alive_list, slot_of, cells, meta,
                                                                  cells_per_grid, radius, counters);
        return cudaSuccess;
    }
    const bool vec = (geom.gw % 4u == 0u) && (cells_per_grid % 4u == 0u);
    // largest disc radius whose row-aligned window fits the shared-memory budget
    int radius = radius_cells < 1 ? 1 : (radius_cells > RAY_MAX_RADIUS ? RAY_MAX_RADIUS : radius_cells);
    while (radius > 1 && (size_t)ray_window_cells_upper_bound(radius, vec) * 4 > (size_t)RAY_MAX_SMEM) radius--;
    const size_t wmax = (size_t)ray_window_cells_upper_bound(radius, vec);
    const size_t smem = wmax * 4;
    *window_cells = wmax;
    if (vec)
        k_ray_update<true><<<n_local, threads, smem, stream>>>(geom, scan, results, first_particle, alive_list, slot_of, cells, meta,
                                                               cells_per_grid, radius, counters);
    else
        k_ray_update<false><<<n_local, threads, smem, stream>>>(geom, scan, results, first_particle, alive_list, slot_of, cells, meta,
                                                                cells_per_grid, radius, counters);
    return cudaSuccess;
}

// =============================================================================== k_weights

constexpr int W_THREADS = 1024;
constexpr int W_CLUSTER = 8;   // CTAs of the (portable-size) thread-block cluster that shares the reduction

// normalize_weights (particle.rs:49-56), the argmax of particle.rs:40-46 and the running sum of
// particle.rs:85-91 over the WHOLE population, on one thread-block cluster: 8 CTAs x 1024
// threads, each thread folds a contiguous chunk left to right, chunk sums are combined by a fixed
// shuffle tree inside the CTA and the 8 CTA totals are exchanged through distributed shared
// memory. The combination order depends only on N: bit-identical on every GPU and every run.
__global__ void __cluster_dims__(W_CLUSTER, 1, 1) __launch_bounds__(W_THREADS)
k_weights(const ParticleResult* __restrict__ results, uint32_t n, double* __restrict__ w_norm,
          double* __restrict__ cum, StepCounters* counters) {
    cg::cluster_group cluster = cg::this_cluster();
    const uint32_t crank = cluster.block_rank();
    __shared__ double s_warp[33];
    __shared__ double s_tot[3][W_CLUSTER];          // CTA totals (raw, normalised, squared), filled by the peers
    __shared__ long long s_key[32];
    __shared__ uint32_t s_arg[32];
    __shared__ long long s_ckey[W_CLUSTER];         // per-CTA argmax candidates (read by CTA 0)
    __shared__ uint32_t s_carg[W_CLUSTER];
    cluster.sync();   // every CTA of the cluster is running before its shared memory is written remotely
    const uint32_t gt = crank * W_THREADS + threadIdx.x;
    const uint32_t chunk = (n + W_CLUSTER * W_THREADS - 1) / (W_CLUSTER * W_THREADS);
    const uint32_t lo = min(n, gt * chunk), hi = min(n, lo + chunk);

    if (gt == 0) {   // per-step counters start from zero
        counters->clamped = 0ull; counters->saturated = 0ull; counters->spilled = 0ull;
        counters->n_alive = 0ull; counters->copy_bytes = 0ull; counters->copy_max_rows = 0ull;
    }

    // pass 1: sum of the raw weights
    double part = 0.0;
    for (uint32_t i = lo; i < hi; ++i) part = __dadd_rn(part, results[i].weight);
    double cta_sum;
    block_excl_scan_f64(part, s_warp, &cta_sum);
    if (threadIdx.x < W_CLUSTER) cluster.map_shared_rank(&s_tot[0][0], threadIdx.x)[crank] = cta_sum;
    cluster.sync();
    double sum = 0.0;
#pragma unroll
    for (int r = 0; r < W_CLUSTER; ++r) sum = __dadd_rn(sum, s_tot[0][r]);

    // pass 2: normalise, argmax candidate, chunk sums of the normalised weights
    double npart = 0.0, sqpart = 0.0;
    long long best_key = (long long)0x8000000000000000ull;
    uint32_t best_i = 0;
    bool have = false;
    for (uint32_t i = lo; i < hi; ++i) {
        const double w = __ddiv_rn(results[i].weight, sum);
        w_norm[i] = w;
        npart = __dadd_rn(npart, w);
        sqpart = __dadd_rn(sqpart, __dmul_rn(w, w));
        const long long k = total_order_key(w);
        if (!have || k >= best_key) { best_key = k; best_i = i; have = true; }  // last max wins
    }
    double cta_n, cta_sq;
    block_excl_scan_f64(sqpart, s_warp, &cta_sq);
    const double offset = block_excl_scan_f64(npart, s_warp, &cta_n);
    if (threadIdx.x < W_CLUSTER) cluster.map_shared_rank(&s_tot[1][0], threadIdx.x)[crank] = cta_n;
    if (threadIdx.x == 0) cluster.map_shared_rank(&s_tot[2][0], 0)[crank] = cta_sq;

    // argmax by f64::total_cmp, ties -> highest index (Iterator::max_by returns the last maximum)
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (!have) { best_key = (long long)0x8000000000000000ull; best_i = 0; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const long long ok = __shfl_down_sync(0xffffffffu, best_key, o);
        const uint32_t oi = __shfl_down_sync(0xffffffffu, best_i, o);
        const bool ohave = __shfl_down_sync(0xffffffffu, (int)have, o) != 0;
        if (ohave && (!have || ok > best_key || (ok == best_key && oi > best_i))) { best_key = ok; best_i = oi; have = true; }
    }
    if (lane == 0) { s_key[wid] = have ? best_key : (long long)0x8000000000000000ull; s_arg[wid] = have ? best_i : 0xffffffffu; }
    __syncthreads();
    if (threadIdx.x == 0) {
        long long bk = 0; uint32_t bi = 0xffffffffu; bool h = false;
        for (int w = 0; w < W_THREADS / 32; ++w) {
            if (s_arg[w] == 0xffffffffu) continue;
            if (!h || s_key[w] > bk || (s_key[w] == bk && s_arg[w] > bi)) { bk = s_key[w]; bi = s_arg[w]; h = true; }
        }
        cluster.map_shared_rank(&s_ckey[0], 0)[crank] = bk;
        cluster.map_shared_rank(&s_carg[0], 0)[crank] = bi;
    }
    cluster.sync();

    // running sum of the normalised weights (the `c += weight[i]` of particle.rs:85-91)
    double cta_off = 0.0;
    for (uint32_t r = 0; r < crank; ++r) cta_off = __dadd_rn(cta_off, s_tot[1][r]);
    double c = __dadd_rn(cta_off, offset);
    for (uint32_t i = lo; i < hi; ++i) {
        c = __dadd_rn(c, w_norm[i]);
        cum[i] = c;
    }
    if (gt == 0) {
        long long bk = 0; uint32_t bi = 0; bool h = false;
        for (int r = 0; r < W_CLUSTER; ++r) {
            if (s_carg[r] == 0xffffffffu) continue;
            if (!h || s_ckey[r] > bk || (s_ckey[r] == bk && s_carg[r] > bi)) { bk = s_ckey[r]; bi = s_carg[r]; h = true; }
        }
        counters->max_particle = bi;
        counters->sum = sum;
        // number_of_effective_particles (particle.rs:59-65) of the normalised weights, before resampling
        double sq = 0.0;
        for (int r = 0; r < W_CLUSTER; ++r) sq = __dadd_rn(sq, s_tot[2][r]);
        counters->n_eff = __ddiv_rn(1.0, sq);
    }
}

void launch_weights(cudaStream_t stream, const ParticleResult* results, uint32_t n_total, double* w_norm,
                    double* cum, StepCounters* counters) {
    k_weights<<<W_CLUSTER, W_THREADS, 0, stream>>>(results, n_total, w_norm, cum, counters);
}

// =============================================================================== k_resample_indices

// first i with !(u_m > cum[i]) for the zero-based new-particle index m0 (particle.rs:84-94)
__device__ __forceinline__ uint32_t resample_source(const double* __restrict__ cum, uint32_t n, double U, uint32_t m0,
                                                    bool* ran_off) {
    const double num = (double)n;
    const double r = __ddiv_rn(__dmul_rn(U, 1.0), num);   // particle.rs:84: r = rand::random::<f64>() * 1.0 / N
    // particle.rs:89: u = r + (m as f64 - 1.0) * 1.0 / N with m = m0 + 1
    const double u = __dadd_rn(r, __ddiv_rn(__dmul_rn(__dsub_rn((double)(m0 + 1u), 1.0), 1.0), num));
    // particle.rs:91-94: advance i while u > c. c is non-decreasing (weights >= 0), so the loop
    // stops at the first i with !(u > cum[i]); found here by bisection.
    uint32_t lo = 0, hi = n;  // answer in [lo, hi]; hi == n means "ran off the end"
    while (lo < hi) {
        const uint32_t mid = lo + ((hi - lo) >> 1);
        if (u > cum[mid]) lo = mid + 1; else hi = mid;
    }
    *ran_off = lo >= n;   // the reference would index out of bounds and panic; clamp and flag
    return lo >= n ? n - 1 : lo;
}

// Also builds the list of local particles that survive (some index selects them): a particle that
// no entry of the index vector selects is dropped by the resampler (particle.rs:88-104 builds the
// new generation only from old[i]); integrating the scan into its grid would be unobservable work,
// so the ray kernel runs on the survivors only. The thread of the FIRST new particle that selects a
// local source appends it (the index vector is non-decreasing, so "first" = differs from the
// predecessor's source, which comes from the neighbouring lane).
__global__ void __launch_bounds__(256)
k_resample_indices(const ParticleResult* __restrict__ results, const double* __restrict__ cum, uint32_t n,
                   const double* __restrict__ u01_caller, uint64_t seed, uint64_t step, uint32_t* __restrict__ idx,
                   float* __restrict__ pose_next, uint32_t first_particle, uint32_t n_local, bool build_alive,
                   uint32_t* __restrict__ alive_list, StepCounters* counters) {
    const uint32_t m0 = blockIdx.x * blockDim.x + threadIdx.x;  // zero-based new-particle index
    const double U = u01_caller ? *u01_caller : slamrs_stream::resample_uniform(seed, step);
    const int lane = threadIdx.x & 31;
    bool ran_off = false;
    uint32_t src_idx = 0xffffffffu;
    if (m0 < n) {
        src_idx = resample_source(cum, n, U, m0, &ran_off);
        if (ran_off) atomicAdd(&counters->clamped, 1ull);
        idx[m0] = src_idx;
        const ParticleResult src = results[src_idx];
        if (m0 >= first_particle && m0 < first_particle + n_local) {
            float* q = pose_next + 3 * (size_t)(m0 - first_particle);
            q[0] = src.x; q[1] = src.y; q[2] = src.theta;
        }
        // estimated_pose(), slam.rs:77-81: new generation indexed by the pre-resample argmax
        if ((unsigned long long)m0 == counters->max_particle) {
            counters->est_pose[0] = src.x; counters->est_pose[1] = src.y; counters->est_pose[2] = src.theta;
        }
    }
    if (!build_alive) return;   // uniform over the grid
    uint32_t prev = __shfl_up_sync(0xffffffffu, src_idx, 1);
    if (lane == 0 && m0 > 0 && m0 < n) {
        bool dummy;
        prev = resample_source(cum, n, U, m0 - 1u, &dummy);
    }
    const bool alive = m0 < n && (m0 == 0 || prev != src_idx) && src_idx >= first_particle &&
                       src_idx < first_particle + n_local;
    // warp-aggregated append (order is irrelevant: particles are independent)
    const unsigned mask = __ballot_sync(0xffffffffu, alive);
    if (mask) {
        const int leader = __ffs(mask) - 1;
        unsigned long long base = 0;
        if (lane == leader) base = atomicAdd(&counters->n_alive, (unsigned long long)__popc(mask));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (alive) alive_list[base + __popc(mask & ((1u << lane) - 1u))] = src_idx - first_particle;
    }
}

void launch_resample_indices(cudaStream_t stream, const ParticleResult* results, const double* cum,
                             uint32_t n_total, const double* u01_caller, uint64_t seed, uint64_t step,
                             uint32_t* idx, float* pose_next, uint32_t first_particle, uint32_t n_local,
                             bool build_alive, uint32_t* alive_list, StepCounters* counters) {
    k_resample_indices<<<(n_total + 255) / 256, 256, 0, stream>>>(results, cum, n_total, u01_caller, seed, step, idx,
                                                                 pose_next, first_particle, n_local, build_alive,
                                                                 alive_list, counters);
}

// =============================================================================== k_mark_alive
// The survivor list is normally built by k_resample_indices. This kernel builds it on its own:
// with all_particles the list is the identity (the reference's order of work: every particle's
// grid receives the scan).
__global__ void __launch_bounds__(256)
k_mark_alive(const uint32_t* __restrict__ idx, uint32_t n_total, uint32_t first_particle, uint32_t n_local,
             bool all_particles, uint32_t* __restrict__ alive_list, StepCounters* counters) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    bool alive = false;
    if (j < n_local) {
        if (all_particles) {
            alive = true;
        } else {
            const uint32_t v = first_particle + j;
            uint32_t lo = 0, hi = n_total;
            while (lo < hi) {
                const uint32_t mid = lo + ((hi - lo) >> 1);
                if (idx[mid] < v) lo = mid + 1; else hi = mid;
            }
            alive = lo < n_total && idx[lo] == v;
        }
    }
    // warp-aggregated append (order is irrelevant: particles are independent)
    const unsigned m = __ballot_sync(0xffffffffu, alive);
    if (m) {
        const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
        unsigned long long base = 0;
        if (lane == leader) base = atomicAdd(&counters->n_alive, (unsigned long long)__popc(m));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (alive) alive_list[base + __popc(m & ((1u << lane) - 1u))] = j;
    }
}

void launch_mark_alive(cudaStream_t stream, const uint32_t* idx, uint32_t n_total, uint32_t first_particle,
                       uint32_t n_local, bool all_particles, uint32_t* alive_list, StepCounters* counters) {
    k_mark_alive<<<(n_local + 255) / 256, 256, 0, stream>>>(idx, n_total, first_particle, n_local, all_particles,
                                                           alive_list, counters);
}

// =============================================================================== k_plan
// Turns the (non-decreasing) index vector into work for this rank's output range [lo, lo+S):
//   0  source is local and this is its first use here   -> the grid stays where it is
//   1  source is local, further use                     -> copy from the kept grid into a free slot
//   2  source lives on another GPU, first use here      -> copy over NVLink into a free slot
//   3  source lives on another GPU, further use         -> likewise (every 16th use re-reads the source)
// All copies form ONE list in output order; copies of one source are adjacent, and every
// COPY_FAN-th of them is a "leader": the copy kernel reads the source once per leader and stores it
// to the whole sub-run. Free slots = slots of local particles nobody here keeps + the persistent
// spare slots. A dropped slot whose grid another GPU copies from in this step ("unsafe") is not
// handed out now -- it joins the spare list of the next step -- so no rank ever writes a grid
// that a peer may still be reading, and one cross-GPU barrier per resampling is enough. That needs
// n_unsafe <= n_spare; otherwise the step reports SLAMRS_E_STAGING (raise spare_slots).

__device__ __forceinline__ uint32_t lower_bound_u32(const uint32_t* a, uint32_t lo, uint32_t hi, uint32_t v) {
    while (lo < hi) {
        const uint32_t mid = lo + ((hi - lo) >> 1);
        if (a[mid] < v) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// With n_local <= PLAN_STAGED_MAX_S the index range, the slot tables, the class bytes and the free
// list live in shared memory (22 bytes per particle): the planner is a chain of short sequential
// passes whose cost is load latency, and shared memory cuts that by an order of magnitude.
constexpr uint32_t PLAN_STAGED_MAX_S = 8192;
__host__ __device__ inline size_t plan_staged_bytes(uint32_t S) { return (size_t)S * 22u + 64u; }

__global__ void __launch_bounds__(1024) k_plan(PlanArgs a) {
    extern __shared__ __align__(16) unsigned char s_dyn[];
    __shared__ uint32_t s_warp[33];
    const uint32_t S = a.n_local, lo = a.rank * a.n_local, hi = lo + S;
    const uint32_t T = blockDim.x, t = threadIdx.x;
    const uint32_t chunk = (S + T - 1) / T;
    const uint32_t c0 = min(S, t * chunk), c1 = min(S, c0 + chunk);
    const uint32_t E = (uint32_t)a.counters->n_spare;

    // working arrays: shared memory when staged, the global scratch otherwise
    uint32_t* idx_l;       // idx[lo .. hi)
    int32_t* slot_old;     // read-only copy
    int32_t* slot_new;
    int32_t* free_list;    // [safe | spare | unsafe]
    uint8_t* keep;
    uint8_t* need;
    if (a.staged) {
        idx_l = reinterpret_cast<uint32_t*>(s_dyn);
        slot_old = reinterpret_cast<int32_t*>(idx_l + S);
        slot_new = slot_old + S;
        free_list = slot_new + S;          // 2 S + 1 entries (E <= S when staged)
        keep = reinterpret_cast<uint8_t*>(free_list + 2 * (size_t)S + 4);
        need = keep + S;
        for (uint32_t j = t; j < S; j += T) { idx_l[j] = a.idx[lo + j]; slot_old[j] = a.slot_old[j]; keep[j] = 0; }
    } else {
        idx_l = const_cast<uint32_t*>(a.idx) + lo;
        slot_old = const_cast<int32_t*>(a.slot_old);
        slot_new = a.slot_new;
        free_list = a.free_list;
        keep = reinterpret_cast<uint8_t*>(a.keep);
        need = reinterpret_cast<uint8_t*>(a.need);
        for (uint32_t j = t; j < S; j += T) keep[j] = 0;
    }
    __syncthreads();

    // ---- classify new particles
    uint32_t nA = 0, nR = 0;
    for (uint32_t m = t; m < S; m += T) {
        const uint32_t src = idx_l[m];
        const bool first = (m == 0) || (idx_l[m - 1] != src);
        const bool local = (src >= lo && src < hi);
        int cls;
        if (local && first) {
            cls = 0;
            keep[src - lo] = 1;
            slot_new[m] = slot_old[src - lo];
        } else if (local) cls = 1;
        else if (first) cls = 2;
        else cls = 3;
        need[m] = (uint8_t)cls;
        nA += (first ? 1u : 0u);
        nR += (cls == 2 ? 1u : 0u);
    }
    __syncthreads();

    // ---- classify old slots: 0 = kept, 1 = free & safe, 2 = free but read by another GPU this step.
    // A slot that is not kept has no local consumer; idx is non-decreasing, so its consumers (if
    // any) are all before this rank's range (v < idx[lo]) or all after it (v > idx[hi-1]).
    const uint32_t idx_first = idx_l[0], idx_last = idx_l[S - 1];
    for (uint32_t j = t; j < S; j += T) {
        int f = 0;
        if (!keep[j]) {
            f = 1;
            if (a.world > 1) {
                const uint32_t v = lo + j;
                if (v < idx_first && lo > 0) {
                    const uint32_t p = lower_bound_u32(a.idx, 0, lo, v);
                    if (p < lo && a.idx[p] == v) f = 2;
                } else if (v > idx_last && hi < a.n_total) {
                    const uint32_t p = lower_bound_u32(a.idx, hi, a.n_total, v);
                    if (p < a.n_total && a.idx[p] == v) f = 2;
                }
            }
        }
        keep[j] = (uint8_t)f;
    }
    __syncthreads();

    // ---- ordered compaction of the free slots: [safe | spare | unsafe]
    uint32_t n_safe_c = 0, n_unsafe_c = 0;
    for (uint32_t j = c0; j < c1; ++j) { n_safe_c += (keep[j] == 1); n_unsafe_c += (keep[j] == 2); }
    uint32_t n_safe, n_unsafe;
    uint32_t ps = block_excl_scan_u32(n_safe_c, s_warp, &n_safe);
    uint32_t pu = block_excl_scan_u32(n_unsafe_c, s_warp, &n_unsafe);
    for (uint32_t j = c0; j < c1; ++j) {
        if (keep[j] == 1) free_list[ps++] = slot_old[j];
        else if (keep[j] == 2) free_list[n_safe + E + pu++] = slot_old[j];
    }
    for (uint32_t e = t; e < E; e += T) free_list[n_safe + e] = a.spare_list[e];
    const uint32_t usable = n_safe + E;   // slots that may be written in this step

    // ---- ordered ranks of the consumers (every new particle that does not keep a grid in place)
    uint32_t n_cons_c = 0;
    for (uint32_t m = c0; m < c1; ++m) n_cons_c += (need[m] != 0);
    uint32_t n_cons;
    uint32_t pos = block_excl_scan_u32(n_cons_c, s_warp, &n_cons);
    __syncthreads();  // free_list complete

    // ---- the copy list. run_first = first position of the current source's run in this range.
    const uint32_t pos_start = pos;
    uint32_t n_lead_c = 0;
    const unsigned long long est_m = a.counters->max_particle - lo;   // >= S when another rank owns the estimate
    if (t == 0 && est_m >= S) a.counters->est_meta_ptr = 0ull;
    uint32_t run_first = c0 < c1 ? lower_bound_u32(idx_l, 0, S, idx_l[c0]) : 0u;
    for (uint32_t m = c0; m < c1; ++m) {
        const int cls = need[m];
        const uint32_t src = idx_l[m];
        if (m > c0 && idx_l[m - 1] != src) run_first = m;
        if (cls == 0) {
            // the published map (slam.rs:83-88) is this particle's grid: it stays in place
            if (m == est_m) a.counters->est_meta_ptr = (unsigned long long)(uintptr_t)(a.meta + slot_new[m]);
            continue;
        }
        if (pos < usable) {
            const int32_t dslot = free_list[pos];
            slot_new[m] = dslot;
            CopyItem it;
            if (cls == 1) {
                const int32_t sslot = slot_old[src - lo];
                it.src = a.cells + (size_t)sslot * a.cells_per_grid;
                it.src_meta = a.meta + sslot;
            } else {
                const uint32_t owner = src / S;
                const int32_t sslot = a.results[src].slot;
                it.src = a.peer_cells[owner] + (size_t)sslot * a.cells_per_grid;
                it.src_meta = a.peer_meta[owner] + sslot;
            }
            it.dst = a.cells + (size_t)dslot * a.cells_per_grid;
            it.dst_meta = a.meta + dslot;
            a.copies[pos] = it;
            if (m == est_m) a.counters->est_meta_ptr = (unsigned long long)(uintptr_t)it.src_meta;   // extent it will have
            // a local run keeps its first use in place, so its copies start one position later
            const uint32_t k = (cls == 1) ? (m - run_first - 1u) : (m - run_first);
            const bool lead = (k % COPY_FAN) == 0u;
            need[m] = (uint8_t)(lead ? 9 : 8);
            n_lead_c += lead;
        } else {
            // no writable slot left (SLAMRS_E_STAGING): nothing is copied, the filter state is invalid
            slot_new[m] = (cls == 1) ? slot_old[src - lo] : slot_old[0];
            need[m] = 10;
        }
        pos++;
    }
    // ordered list of leader positions within copies[]
    uint32_t n_lead;
    uint32_t pl = block_excl_scan_u32(n_lead_c, s_warp, &n_lead);
    pos = pos_start;
    for (uint32_t m = c0; m < c1; ++m) {
        const int cls = need[m];
        if (cls == 9) a.leaders[pl++] = pos;
        if (cls >= 8) pos++;
    }
    __syncthreads();
    // ---- next step's spare list: the usable slots nobody took, then this step's unsafe slots
    const uint32_t used = min(n_cons, usable);
    for (uint32_t e = t; e < E; e += T) {
        const uint32_t left = usable - used;   // = E - n_unsafe when nothing is short
        a.spare_list[e] = e < left ? free_list[used + e] : free_list[usable + (e - left)];
    }
    if (a.staged)
        for (uint32_t m = t; m < S; m += T) a.slot_new[m] = slot_new[m];

    uint32_t distinct, n_remote;
    block_excl_scan_u32(nA, s_warp, &distinct);
    block_excl_scan_u32(nR, s_warp, &n_remote);
    if (t == 0) {
        a.counters->n_copies = used;
        a.counters->n_leaders = n_lead;
        a.counters->n_pulls = n_remote;
        a.counters->distinct = distinct;
        a.counters->staging_short = (n_cons > usable) ? (unsigned long long)(n_cons - usable) : 0ull;
        const unsigned long long mp = a.counters->max_particle;
        a.counters->est_owner = mp / S;
        a.counters->est_slot = (mp >= lo && mp < hi) ? (long long)slot_new[mp - lo] : -1ll;
        if (a.history) {
            StepRecord r;
            r.step = a.step; r.n_copies = used; r.n_pulls = n_remote; r.distinct = distinct; r.n_leaders = n_lead;
            r.n_alive = 0; r.copy_bytes = 0; r.pad = 0;   // filled in by the step's last kernel (k_commit_boxes)
            a.history[a.step % STEP_HISTORY] = r;
        }
    }
}

bool plan_can_stage(uint32_t n_local, uint32_t n_spare_cap) {
    return n_local <= PLAN_STAGED_MAX_S && n_spare_cap <= n_local;
}

void launch_plan(cudaStream_t stream, const PlanArgs& a) {
    k_plan<<<1, 1024, a.staged ? plan_staged_bytes(a.n_local) : 0, stream>>>(a);
}

// =============================================================================== k_copy
// Grid copies (the `value.clone()` of particle.rs:97-100). Pure streaming: 128-bit loads that
// bypass L1, four in flight per thread, then 128-bit stores. Copies of the same source are
// adjacent in the list, so one work item = 16 KiB of a source grid fanned out to up to
// COPY_FAN destinations: the source is read once per sub-run instead of once per copy, which
// makes the kernel write-bound (D grids written, D / COPY_FAN + distinct sources read).
// CTAs stride over (leader, chunk) items; list lengths are read from device memory so that no
// host round trip sits between planning and copying.

constexpr int COPY_THREADS = 256;
constexpr int COPY_UNROLL = 2;
constexpr uint32_t COPY_ITEM_V8 = COPY_THREADS * COPY_UNROLL;  // 32-byte units per work item (16 KiB)
constexpr int COPY_CTAS_PER_SM = 32;  // measured on B200: 6.37 TB/s moved at 32/SM vs 5.72 TB/s at 8/SM (tools/bw_probe.cu)

struct alignas(32) V8 {
    uint4 a, b;
};
// 256-bit global accesses (sm_100: ld/st.global.v8.b32). Streaming: no L1 allocation.
__device__ __forceinline__ V8 ld_stream_v8(const V8* p) {
    V8 r;
    asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r.a.x), "=r"(r.a.y), "=r"(r.a.z), "=r"(r.a.w), "=r"(r.b.x), "=r"(r.b.y), "=r"(r.b.z), "=r"(r.b.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream_v8(V8* p, const V8& v) {
    asm volatile("st.global.L1::no_allocate.v8.u32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(v.a.x), "r"(v.a.y),
                 "r"(v.a.z), "r"(v.a.w), "r"(v.b.x), "r"(v.b.y), "r"(v.b.z), "r"(v.b.w)
                 : "memory");
}

__global__ void __launch_bounds__(COPY_THREADS)
k_copy(const CopyItem* __restrict__ items, const uint32_t* __restrict__ leaders,
       const unsigned long long* __restrict__ n_items, const unsigned long long* __restrict__ n_leaders,
       uint32_t v8_per_grid) {
    const unsigned long long n = *n_items;
    const unsigned long long nl = leaders ? *n_leaders : n;
    const uint32_t chunks = (v8_per_grid + COPY_ITEM_V8 - 1) / COPY_ITEM_V8;
    const unsigned long long total = nl * chunks;
    for (unsigned long long w = blockIdx.x; w < total; w += gridDim.x) {
        const unsigned long long q = w / chunks;
        const uint32_t c = (uint32_t)(w - q * chunks);
        const unsigned long long k = leaders ? leaders[q] : q;
        const CopyItem it = items[k];
        uint32_t fan = 1;
        if (leaders) {
            while (fan < COPY_FAN && k + fan < n && items[k + fan].src == it.src) fan++;
        }
        const V8* src = reinterpret_cast<const V8*>(it.src);
        const uint32_t base = c * COPY_ITEM_V8 + threadIdx.x;
        V8 v[COPY_UNROLL];
#pragma unroll
        for (int u = 0; u < COPY_UNROLL; ++u) {
            const uint32_t i = base + u * COPY_THREADS;
            if (i < v8_per_grid) v[u] = ld_stream_v8(src + i);
        }
        for (uint32_t f = 0; f < fan; ++f) {
            V8* dst = reinterpret_cast<V8*>(items[k + f].dst);
#pragma unroll
            for (int u = 0; u < COPY_UNROLL; ++u) {
                const uint32_t i = base + u * COPY_THREADS;
                if (i < v8_per_grid) st_stream_v8(dst + i, v[u]);
            }
        }
    }
}

void launch_copy(cudaStream_t stream, const CopyItem* items, const uint32_t* leaders,
                 const unsigned long long* n_items, const unsigned long long* n_leaders, size_t cells_per_grid,
                 int num_sms) {
    const uint32_t v8 = (uint32_t)(cells_per_grid / 8);  // cells_per_grid is a multiple of 32 cells
    k_copy<<<num_sms * COPY_CTAS_PER_SM, COPY_THREADS, 0, stream>>>(items, leaders, n_items, n_leaders, v8);
}

// =============================================================================== k_copy_boxed
// Extent-limited grid copy. A grid is zero outside its extent (SlotMeta), so cloning it means:
// copy the source's extent and clear whatever else the destination slot's previous tenant had
// informed. k_copy_prepare turns every fan-out sub-run into one CopyJob (source, source extent,
// destinations, U = union of the source extent and the destinations' old extents); k_copy_boxed
// then works on (job, band of rows of U) items, the number of bands per job chosen on the device
// so that every CTA gets several items. Inside a band the (row, 32-byte unit) pairs of U are
// linearised over the CTA's threads: each thread issues COPY_UNROLL independent 256-bit loads
// (zero outside the source extent) and stores each value to every destination of the sub-run.
// Bytes that really moved are counted on the device and are what the roofline in bench.py uses.

// All x quantities of a job are in 32-byte units on the ring of one physical grid row
// (ring size = row units when rows rotate, unbounded otherwise): an "arc" is (start, length).
struct alignas(16) CopyJob {
    const uint32_t* src;
    uint32_t fan;
    uint32_t rot;             // destination unit = (source unit + rot) & umask
    uint32_t n_start, n_len;  // arc of every destination that receives the source's extent
    int sy0, sy1;             // ... and its rows
    uint32_t u_start, u_len;  // arc written in every destination (new extent + old extents to clear)
    int uy0, uy1;             // ... and its rows
    uint32_t* dst[COPY_FAN];
};
static_assert(sizeof(CopyJob) % 16 == 0, "CopyJob is fetched as 16-byte pieces");
constexpr int COPY_JOB_V4 = (int)(sizeof(CopyJob) / 16);

__device__ __forceinline__ bool meta_empty(const SlotMeta& m) { return m.x1 <= m.x0 || m.y1 <= m.y0; }

// smallest arc (of those starting at either operand's start) that covers arcs a and b on the ring
__device__ __forceinline__ void arc_cover(uint32_t& a_start, uint32_t& a_len, uint32_t b_start, uint32_t b_len,
                                          uint32_t umask, uint32_t ring) {
    if (b_len == 0u) return;
    if (a_len == 0u) { a_start = b_start; a_len = b_len; return; }
    // 64-bit: with unrotated rows the "ring" is the whole 32-bit range and the sums may exceed it
    const unsigned long long l1 = max((unsigned long long)a_len, (unsigned long long)((b_start - a_start) & umask) + b_len);
    const unsigned long long l2 = max((unsigned long long)b_len, (unsigned long long)((a_start - b_start) & umask) + a_len);
    if (l2 < l1) { a_start = b_start; a_len = (uint32_t)min(l2, (unsigned long long)ring); }
    else { a_len = (uint32_t)min(l1, (unsigned long long)ring); }
    if (a_len >= ring) { a_start = 0u; a_len = ring; }
}

// one warp per job
__global__ void __launch_bounds__(256)
k_copy_prepare(const CopyItem* __restrict__ items, const uint32_t* __restrict__ leaders,
               const unsigned long long* __restrict__ n_items, const unsigned long long* __restrict__ n_leaders,
               CopyJob* __restrict__ jobs, MapGeom geom, StepCounters* counters) {
    const unsigned long long n = *n_items;
    const unsigned long long nl = leaders ? *n_leaders : n;
    const unsigned long long q = (unsigned long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (q >= nl) return;
    const int lane = threadIdx.x & 31;
    const uint32_t umask = geom.xmask == 0xffffffffu ? 0xffffffffu : (geom.xmask >> 3);
    const uint32_t ring = geom.xmask == 0xffffffffu ? 0xffffffffu : (geom.gw >> 3);
    const unsigned long long k = leaders ? leaders[q] : q;
    const bool have = lane < (int)COPY_FAN && k + lane < n && (leaders != nullptr || lane == 0);
    CopyItem it{};
    if (have) it = items[k + lane];
    const unsigned long long src0 = __shfl_sync(0xffffffffu, (unsigned long long)(uintptr_t)it.src, 0);
    const unsigned same = __ballot_sync(0xffffffffu, have && (unsigned long long)(uintptr_t)it.src == src0);
    const uint32_t fan = (uint32_t)(__ffs(~same) - 1);   // leading run of items that share the source
    SlotMeta sm{0, 0, 0, 0, 0, 0, 0, 0}, dm{0, 0, 0, 0, 0, 0, 0, 0};
    if (lane == 0) sm = *it.src_meta;
    if (lane < (int)fan) dm = *it.dst_meta;
    // the destination's old extent as a physical arc
    uint32_t o_start = 0u, o_len = 0u;
    int oy0 = 0x7fffffff, oy1 = -1;
    if (lane < (int)fan && !meta_empty(dm)) {
        o_start = (phys_col(geom, (uint32_t)dm.x0, dm.ox) >> 3) & umask;
        o_len = (uint32_t)(dm.x1 - dm.x0) >> 3;
        oy0 = dm.y0; oy1 = dm.y1;
    }
    // lane 0 folds the arcs (at most 17) and writes the job header
    uint32_t u_start = 0u, u_len = 0u, n_start = 0u, n_len = 0u, rot = 0u;
    int uy0 = 0x7fffffff, uy1 = -1, sy0 = 0, sy1 = 0;
    if (lane == 0 && !meta_empty(sm)) {
        const uint32_t s_start = (phys_col(geom, (uint32_t)sm.x0, sm.ox) >> 3) & umask;
        n_len = (uint32_t)(sm.x1 - sm.x0) >> 3;
        n_start = (phys_col(geom, (uint32_t)sm.x0, align_shift(geom, sm.x0)) >> 3) & umask;   // page-aligned
        rot = (n_start - s_start) & umask;
        sy0 = sm.y0; sy1 = sm.y1;
        u_start = n_start; u_len = n_len; uy0 = sy0; uy1 = sy1;
    }
    for (uint32_t f = 0; f < fan; ++f) {
        const uint32_t bs = __shfl_sync(0xffffffffu, o_start, (int)f), bl = __shfl_sync(0xffffffffu, o_len, (int)f);
        const int by0 = __shfl_sync(0xffffffffu, oy0, (int)f), by1 = __shfl_sync(0xffffffffu, oy1, (int)f);
        if (lane == 0) {
            arc_cover(u_start, u_len, bs, bl, umask, ring);
            if (bl) { uy0 = min(uy0, by0); uy1 = max(uy1, by1); }
        }
    }
    CopyJob* job = jobs + q;
    if (lane < (int)COPY_FAN) job->dst[lane] = lane < (int)fan ? it.dst : nullptr;
    if (lane == 0) {
        if (u_len == 0u || uy1 <= uy0) { u_start = u_len = 0u; uy0 = uy1 = 0; }
        job->src = it.src; job->fan = fan; job->rot = rot;
        job->n_start = n_start; job->n_len = n_len; job->sy0 = sy0; job->sy1 = sy1;
        job->u_start = u_start; job->u_len = u_len; job->uy0 = uy0; job->uy1 = uy1;
        if (uy1 > uy0) atomicMax(&counters->copy_max_rows, (unsigned long long)(uy1 - uy0));
    }
}

// One work item = (job, band of rows of U); the (row, 32-byte unit) pairs of the band are linearised
// over the CTA's threads, UNROLL independent 256-bit loads per thread, then every value is stored
// to each destination of the sub-run. The job of the next item is fetched into registers while the
// current item is copied. CTAs are single warps (BOX_THREADS): a band of ~7 rows x 32 units is a
// few hundred elements, and with more warps per CTA the two barriers per item dominate.
constexpr int BOX_THREADS = 32;
constexpr int BOX_CTAS_PER_SM = 256;
template <int UNROLL, int MINB, int THREADS>
__global__ void __launch_bounds__(THREADS, MINB)
k_copy_boxed(const CopyJob* __restrict__ jobs, const unsigned long long* __restrict__ n_jobs,
             uint32_t row_units /* 32-byte units per physical grid row */, uint32_t umask, uint32_t items_per_cta,
             StepCounters* counters) {
    __shared__ CopyJob s_job;
    __shared__ unsigned long long s_moved;
    if (threadIdx.x == 0) s_moved = 0ull;
    const unsigned long long nl = *n_jobs;
    if (nl == 0) return;
    // bands per job: about items_per_cta work items per CTA in total, at most one band per row
    const uint32_t max_rows = (uint32_t)counters->copy_max_rows;
    const unsigned long long want = ((unsigned long long)gridDim.x * items_per_cta + nl - 1ull) / nl;
    const uint32_t bands = (uint32_t)(want < 1ull ? 1ull : (want > max_rows ? (max_rows ? max_rows : 1u) : want));
    const uint32_t total = (uint32_t)min(nl * bands, 0xffffffffull);
    uint32_t moved = 0;   // 32-byte units read + written by this thread
    uint4 next_job = make_uint4(0u, 0u, 0u, 0u);
    if (threadIdx.x < COPY_JOB_V4 && blockIdx.x < total)
        next_job = reinterpret_cast<const uint4*>(jobs + blockIdx.x / bands)[threadIdx.x];
    for (uint32_t w = blockIdx.x; w < total; w += gridDim.x) {
        const uint32_t q = w / bands;
        const uint32_t band = w - q * bands;
        __syncthreads();   // the previous item's job is no longer read
        if (threadIdx.x < COPY_JOB_V4) {
            reinterpret_cast<uint4*>(&s_job)[threadIdx.x] = next_job;
            const unsigned long long wn = (unsigned long long)w + gridDim.x;
            if (wn < total) next_job = reinterpret_cast<const uint4*>(jobs + (uint32_t)wn / bands)[threadIdx.x];
        }
        __syncthreads();
        const int uy0 = s_job.uy0, uy1 = s_job.uy1;
        const int rows = uy1 - uy0;
        if (rows <= 0) continue;
        const int rows_per_band = (rows + (int)bands - 1) / (int)bands;
        const int r0 = uy0 + (int)band * rows_per_band, r1 = min(uy1, r0 + rows_per_band);
        if (r0 >= r1) continue;
        const uint32_t u_start = s_job.u_start, uw = s_job.u_len;
        const uint32_t n_start = s_job.n_start, n_len = s_job.n_len, rot = s_job.rot;
        const int sy0 = s_job.sy0, sy1 = s_job.sy1;
        const uint32_t fan = s_job.fan;
        const uint32_t count = (uint32_t)(r1 - r0) * uw;
        const V8* src = reinterpret_cast<const V8*>(s_job.src);
        for (uint32_t base = threadIdx.x; base < count; base += THREADS * UNROLL) {
            V8 v[UNROLL];
            uint32_t off[UNROLL];   // unit offset inside a destination grid (< 2^28)
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const uint32_t i = base + u * THREADS;
                off[u] = 0xffffffffu;
                v[u].a = make_uint4(0u, 0u, 0u, 0u); v[u].b = v[u].a;
                if (i < count) {
                    const uint32_t rr = i / uw;
                    const int ey = r0 + (int)rr;
                    const uint32_t du = (u_start + (i - rr * uw)) & umask;     // destination unit on the ring
                    off[u] = (uint32_t)ey * row_units + du;
                    if (((du - n_start) & umask) < n_len && ey >= sy0 && ey < sy1) {
                        v[u] = ld_stream_v8(src + ((uint32_t)ey * row_units + ((du - rot) & umask)));
                        moved++;
                    }
                }
            }
            for (uint32_t f = 0; f < fan; ++f) {
                V8* dst = reinterpret_cast<V8*>(s_job.dst[f]);
#pragma unroll
                for (int u = 0; u < UNROLL; ++u)
                    if (off[u] != 0xffffffffu) { st_stream_v8(dst + off[u], v[u]); moved++; }
            }
        }
    }
    // bytes actually moved, for the roofline: warp -> CTA -> one global atomic per CTA
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) moved += __shfl_down_sync(0xffffffffu, moved, o);
    if ((threadIdx.x & 31) == 0 && moved) atomicAdd(&s_moved, (unsigned long long)moved);
    __syncthreads();
    if (threadIdx.x == 0 && s_moved) atomicAdd(&counters->copy_bytes, s_moved * 32ull);
}

void launch_copy_boxed(cudaStream_t stream, const CopyItem* items, const uint32_t* leaders,
                       const unsigned long long* n_items, const unsigned long long* n_leaders, uint32_t max_items,
                       void* jobs, MapGeom geom, StepCounters* counters, int num_sms) {
    const uint32_t blocks = (max_items + 7u) / 8u;
    k_copy_prepare<<<blocks ? blocks : 1, 256, 0, stream>>>(items, leaders, n_items, n_leaders, (CopyJob*)jobs, geom, counters);
    // measured on B200 (gpurun_out/tune_copy4.log): 4 loads in flight per thread, 3 CTAs per SM
    // resident, grid oversubscribed 32x per SM for balance, ~6 items per CTA
    const uint32_t umask = geom.xmask == 0xffffffffu ? 0xffffffffu : (geom.xmask >> 3);
    // measured on B200 (profiles/r1_copy_tuning.md): one-warp CTAs (the per-item barriers cost more than
    // anything else in larger CTAs), 4 loads in flight per thread, grid oversubscribed for balance,
    // about 6 items per CTA
    k_copy_boxed<4, 24, BOX_THREADS><<<num_sms * BOX_CTAS_PER_SM, BOX_THREADS, 0, stream>>>(
        (const CopyJob*)jobs, leaders ? n_leaders : n_items, geom.gw / 8u, umask, 6u, counters);
}
size_t copy_job_bytes() { return sizeof(CopyJob); }

__global__ void k_commit_boxes(const CopyItem* __restrict__ items, const unsigned long long* __restrict__ n_items,
                               MapGeom geom, bool realign, StepCounters* counters, StepRecord* record) {
    const unsigned long long n = *n_items;
    for (unsigned long long k = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; k < n;
         k += (unsigned long long)gridDim.x * blockDim.x) {
        SlotMeta m = *items[k].src_meta;   // sources are never destinations of the same launch
        if (realign && m.x1 > m.x0) m.ox = align_shift(geom, m.x0);   // the rotation k_copy_prepare chose
        *items[k].dst_meta = m;
    }
    if (record && blockIdx.x == 0 && threadIdx.x == 0) {
        record->copy_bytes = counters->copy_bytes;
        record->n_alive = counters->n_alive;
        // informed extent of the published map, for the windowed read-out (sources are not written here)
        const SlotMeta* em = reinterpret_cast<const SlotMeta*>((uintptr_t)counters->est_meta_ptr);
        if (em == nullptr) { counters->est_box[0] = counters->est_box[1] = counters->est_box[2] = counters->est_box[3] = -1; }
        else {
            const SlotMeta m = *em;
            const bool empty = m.x1 <= m.x0 || m.y1 <= m.y0;
            counters->est_box[0] = empty ? 0 : m.x0; counters->est_box[1] = empty ? 0 : m.y0;
            counters->est_box[2] = empty ? 0 : m.x1; counters->est_box[3] = empty ? 0 : m.y1;
        }
    }
}
void launch_commit_boxes(cudaStream_t stream, const CopyItem* items, const unsigned long long* n_items, uint32_t max_items,
                         MapGeom geom, bool realign, StepCounters* counters, StepRecord* record) {
    const uint32_t blocks = (max_items + 255u) / 256u;
    k_commit_boxes<<<blocks ? blocks : 1, 256, 0, stream>>>(items, n_items, geom, realign, counters, record);
}

__global__ void k_account_full_copy(const unsigned long long* n_items, const unsigned long long* n_leaders,
                                    unsigned long long bytes_per_grid, StepCounters* counters) {
    counters->copy_bytes += bytes_per_grid * (*n_items + (n_leaders ? *n_leaders : *n_items));
}
void launch_account_full_copy(cudaStream_t stream, const unsigned long long* n_items, const unsigned long long* n_leaders,
                              size_t bytes_per_grid, StepCounters* counters) {
    k_account_full_copy<<<1, 1, 0, stream>>>(n_items, n_leaders, (unsigned long long)bytes_per_grid, counters);
}

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
    const int shift = meta[slot].ox;
    const uint32_t ww = (uint32_t)(win.z - win.x), wh = (uint32_t)(win.w - win.y);
    const uint32_t n = ww * wh;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t ry = i / ww, rx = i - ry * ww;
        const uint32_t cell = grid[(size_t)(win.y + ry) * geom.gw + phys_col(geom, (uint32_t)win.x + rx, shift)];
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

// one slot's grid in logical order (the row rotation undone)
__global__ void __launch_bounds__(256)
k_export_slot(const uint32_t* __restrict__ grid, const SlotMeta* __restrict__ slot_meta, MapGeom geom, bool as_log_odds,
              void* __restrict__ out) {
    const int shift = slot_meta->ox;
    const uint32_t n = geom.gw * geom.gh;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t y = i / geom.gw, x = i - y * geom.gw;
        const uint32_t cell = grid[(size_t)y * geom.gw + phys_col(geom, x, shift)];
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

// =============================================================================== init

__global__ void k_init_slots(int32_t* slot_of, uint32_t n_local, int32_t* spare_list, uint32_t n_spare,
                             StepCounters* counters, uint32_t rank, SlotMeta* meta) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_local + n_spare) meta[i] = SlotMeta{0, 0, 0, 0, 0, 0, 0, 0};   // empty: every cell is prior
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
                       StepCounters* counters, uint32_t rank, SlotMeta* meta) {
    const uint32_t n = n_local + n_spare;
    k_init_slots<<<(n + 255) / 256, 256, 0, stream>>>(slot_of, n_local, spare_list, n_spare, counters, rank, meta);
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

cudaError_t configure_kernels() {
    cudaError_t e = cudaFuncSetAttribute(k_ray_update<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, RAY_MAX_SMEM);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_ray_update_packed, cudaFuncAttributeMaxDynamicSharedMemorySize, RAY_MAX_SMEM);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_plan, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan_staged_bytes(PLAN_STAGED_MAX_S));
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(k_ray_update<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, RAY_MAX_SMEM);
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
