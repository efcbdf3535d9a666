// The grid copies of resampling (`value.clone()`, particle.rs:97-100): k_copy (whole grids),
// k_copy_prepare / k_copy_boxed / k_commit_boxes (informed extents, rotated rows).
#include "kernels_common.cuh"

namespace slamrs {

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

// A job works in PHYSICAL slot coordinates. x: 32-byte units on the ring of one slot row (ring size =
// row units when rows rotate, unbounded otherwise); y: rows on the ring of the slot's rows (ring size =
// slot height for windowed slots, unbounded otherwise). An "arc" is (start, length).
struct alignas(16) CopyJob {
    const uint32_t* src;
    uint32_t fan;
    uint32_t rot;               // destination unit = (source unit + rot) & umask
    uint32_t n_start, n_len;    // x arc of every destination that receives the source's extent
    uint32_t ny_start, ny_len;  // ... and its row arc (same rows in source and destination)
    uint32_t u_start, u_len;    // x arc written in every destination (new extent + old extents to clear)
    uint32_t uy_start, uy_len;  // ... and its row arc
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
    const uint32_t ring = geom.xmask == 0xffffffffu ? 0xffffffffu : (geom.pw >> 3);
    const uint32_t ymask = geom.ymask, yring = geom.ymask == 0xffffffffu ? 0xffffffffu : geom.ph;
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
    uint32_t o_start = 0u, o_len = 0u, oy_start = 0u, oy_len = 0u;
    if (lane < (int)fan && !meta_empty(dm)) {
        o_start = (phys_col(geom, (uint32_t)dm.x0, dm.ox) >> 3) & umask;
        o_len = (uint32_t)(dm.x1 - dm.x0) >> 3;
        oy_start = (uint32_t)dm.y0 & ymask;
        oy_len = (uint32_t)(dm.y1 - dm.y0);
    }
    // lane 0 folds the arcs (at most 17) and writes the job header
    uint32_t u_start = 0u, u_len = 0u, n_start = 0u, n_len = 0u, rot = 0u;
    uint32_t uy_start = 0u, uy_len = 0u, ny_start = 0u, ny_len = 0u;
    if (lane == 0 && !meta_empty(sm)) {
        const uint32_t s_start = (phys_col(geom, (uint32_t)sm.x0, sm.ox) >> 3) & umask;
        n_len = (uint32_t)(sm.x1 - sm.x0) >> 3;
        n_start = (phys_col(geom, (uint32_t)sm.x0, align_shift(geom, sm.x0)) >> 3) & umask;   // page-aligned
        rot = (n_start - s_start) & umask;
        ny_start = (uint32_t)sm.y0 & ymask; ny_len = (uint32_t)(sm.y1 - sm.y0);
        u_start = n_start; u_len = n_len; uy_start = ny_start; uy_len = ny_len;
    }
    for (uint32_t f = 0; f < fan; ++f) {
        const uint32_t bs = __shfl_sync(0xffffffffu, o_start, (int)f), bl = __shfl_sync(0xffffffffu, o_len, (int)f);
        const uint32_t bys = __shfl_sync(0xffffffffu, oy_start, (int)f), byl = __shfl_sync(0xffffffffu, oy_len, (int)f);
        if (lane == 0 && bl && byl) {
            arc_cover(u_start, u_len, bs, bl, umask, ring);
            arc_cover(uy_start, uy_len, bys, byl, ymask, yring);
        }
    }
    CopyJob* job = jobs + q;
    if (lane < (int)COPY_FAN) job->dst[lane] = lane < (int)fan ? it.dst : nullptr;
    if (lane == 0) {
        if (u_len == 0u || uy_len == 0u) { u_start = u_len = 0u; uy_start = uy_len = 0u; }
        job->src = it.src; job->fan = fan; job->rot = rot;
        job->n_start = n_start; job->n_len = n_len; job->ny_start = ny_start; job->ny_len = ny_len;
        job->u_start = u_start; job->u_len = u_len; job->uy_start = uy_start; job->uy_len = uy_len;
        if (uy_len) atomicMax(&counters->copy_max_rows, (unsigned long long)uy_len);
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
             uint32_t row_units /* 32-byte units per physical slot row */, uint32_t umask, uint32_t ymask,
             uint32_t items_per_cta, StepCounters* counters) {
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
        const uint32_t rows = s_job.uy_len;
        if (rows == 0u) continue;
        const uint32_t rows_per_band = (rows + bands - 1u) / bands;
        const uint32_t r0 = band * rows_per_band, r1 = min(rows, r0 + rows_per_band);   // offsets into the row arc
        if (r0 >= r1) continue;
        const uint32_t u_start = s_job.u_start, uw = s_job.u_len, uy_start = s_job.uy_start;
        const uint32_t n_start = s_job.n_start, n_len = s_job.n_len, rot = s_job.rot;
        const uint32_t ny_start = s_job.ny_start, ny_len = s_job.ny_len;
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
                    const uint32_t py = (uy_start + r0 + rr) & ymask;          // physical row (same in source and destination)
                    const uint32_t du = (u_start + (i - rr * uw)) & umask;     // destination unit on the ring
                    off[u] = py * row_units + du;
                    if (((du - n_start) & umask) < n_len && ((py - ny_start) & ymask) < ny_len) {
                        v[u] = ld_stream_v8(src + (py * row_units + ((du - rot) & umask)));
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
        (const CopyJob*)jobs, leaders ? n_leaders : n_items, geom.pw / 8u, umask, geom.ymask, 6u, counters);
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

}  // namespace slamrs
