// The resampler up to the copy lists: k_weights (particle.rs:40-56, 59-65, 85-91), k_resample_indices
// (particle.rs:78-101) with the survivor list, k_mark_alive, k_plan (slot tables and copy items).
#include "kernels_common.cuh"

namespace slamrs {

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
        counters->n_mat = 0ull; counters->n_mat_leaders = 0ull; counters->ray_cell_steps = 0ull;
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
    // Every consumer takes a free slot. Eager: every consumer is also a copy. Deferred: only the first
    // local use of a remote source is copied (over NVLink, after the barrier); every other consumer
    // becomes an alias of its source's slot (PlanArgs::alias_of) and moves no bytes now.
    uint32_t n_cons_c = 0, n_item_c = 0;
    for (uint32_t m = c0; m < c1; ++m) { n_cons_c += (need[m] != 0); n_item_c += (need[m] == 2); }
    uint32_t n_cons, n_items = 0;
    uint32_t pos = block_excl_scan_u32(n_cons_c, s_warp, &n_cons);
    uint32_t ipos = a.defer ? block_excl_scan_u32(n_item_c, s_warp, &n_items) : pos;   // position in copies[]
    __syncthreads();  // free_list complete

    // ---- the copy list. run_first = first position of the current source's run in this range.
    const uint32_t ipos_start = ipos;
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
        if (pos < usable && a.defer && cls != 2) {
            slot_new[m] = free_list[pos];
            need[m] = (uint8_t)(cls == 1 ? 11 : 12);   // alias of a local source / of this rank's copy of a remote one
            if (m == est_m && cls == 1) a.counters->est_meta_ptr = (unsigned long long)(uintptr_t)(a.meta + slot_old[src - lo]);
        } else if (pos < usable) {
            const int32_t dslot = free_list[pos];
            slot_new[m] = dslot;
            CopyItem it;
            if (cls == 1) {
                const int32_t sslot = slot_old[src - lo];
                it.src = a.cells + (size_t)sslot * a.cells_per_grid;
                it.src_meta = a.meta + sslot;
                it.src_bands = a.bands + (size_t)sslot * a.n_bands;
            } else {
                const uint32_t owner = src / S;
                const int32_t sslot = a.results[src].slot;
                it.src = a.peer_cells[owner] + (size_t)sslot * a.cells_per_grid;
                it.src_meta = a.peer_meta[owner] + sslot;
                it.src_bands = a.peer_bands[owner] + (size_t)sslot * a.n_bands;
            }
            it.dst = a.cells + (size_t)dslot * a.cells_per_grid;
            it.dst_meta = a.meta + dslot;
            it.dst_bands = a.bands + (size_t)dslot * a.n_bands;
            a.copies[ipos] = it;
            if (m == est_m) a.counters->est_meta_ptr = (unsigned long long)(uintptr_t)it.src_meta;   // extent it will have
            // a local run keeps its first use in place, so its copies start one position later
            const uint32_t k = (cls == 1) ? (m - run_first - 1u) : (m - run_first);
            const bool lead = a.defer || (k % COPY_FAN) == 0u;
            need[m] = (uint8_t)(lead ? 9 : 8);
            n_lead_c += lead;
            if (a.defer) ipos++;
        } else {
            // no writable slot left (SLAMRS_E_STAGING): nothing is copied, the filter state is invalid
            slot_new[m] = (cls == 1) ? slot_old[src - lo] : slot_old[0];
            need[m] = 10;
        }
        pos++;
        if (!a.defer) ipos = pos;
    }
    // ordered list of leader positions within copies[]
    uint32_t n_lead;
    uint32_t pl = block_excl_scan_u32(n_lead_c, s_warp, &n_lead);
    ipos = ipos_start;
    for (uint32_t m = c0; m < c1; ++m) {
        const int cls = need[m];
        if (cls == 9) a.leaders[pl++] = ipos;
        if (cls == 8 || cls == 9 || (!a.defer && cls == 10)) ipos++;
    }
    __syncthreads();   // slot_new complete
    // ---- deferred copies: the new particle's slot stands for its source's slot until it is written
    long long est_root = -1ll;
    if (a.defer) {
        for (uint32_t m = c0; m < c1; ++m) {
            const int cls = need[m];
            int32_t root = slot_new[m];
            if (cls == 11) root = slot_old[idx_l[m] - lo];
            else if (cls == 12) root = slot_new[lower_bound_u32(idx_l, 0, S, idx_l[m])];   // this rank's first use: the copy
            if (cls == 8 || cls == 9 || cls == 11 || cls == 12) a.alias_of[slot_new[m]] = root;
            if (m == est_m) {
                est_root = root;
                if (cls == 12) a.counters->est_meta_ptr = (unsigned long long)(uintptr_t)(a.meta + root);
            }
        }
        if (est_root >= 0) a.counters->est_slot = est_root;   // the published map is read from the cells it shares
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
        // staging short: positions in copies[] have holes (consumers without a slot), so the list is not
        // handed to the copy kernels at all -- the step fails with SLAMRS_E_STAGING and copies nothing
        const bool short_of_slots = n_cons > usable;
        const uint32_t n_copied = a.defer ? n_items : used;
        a.counters->n_copies = short_of_slots ? 0u : n_copied;
        a.counters->n_leaders = short_of_slots ? 0u : n_lead;
        a.counters->n_pulls = n_remote;
        a.counters->distinct = distinct;
        a.counters->staging_short = (n_cons > usable) ? (unsigned long long)(n_cons - usable) : 0ull;
        const unsigned long long mp = a.counters->max_particle;
        a.counters->est_owner = mp / S;
        if (!a.defer || !(mp >= lo && mp < hi))   // deferred: set above by the thread that owns the estimate
            a.counters->est_slot = (mp >= lo && mp < hi) ? (long long)slot_new[mp - lo] : -1ll;
        if (a.history) {
            StepRecord r;
            r.step = a.step; r.n_copies = n_copied; r.n_pulls = n_remote; r.distinct = distinct; r.n_leaders = n_lead;
            r.n_alive = 0; r.copy_bytes = 0; r.ray_cell_steps = 0;   // filled in by the step's last kernel (k_commit_boxes)
            a.history[a.step % STEP_HISTORY] = r;
        }
    }
}

// =============================================================================== k_materialize_list
// exclusive running maximum of one uint32 per thread over a 1024-thread CTA (0 = nothing before)
__device__ __forceinline__ uint32_t block_excl_scan_max_u32(uint32_t v, uint32_t* warp_tot /*[33]*/) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc = max(inc, t);
    }
    __syncthreads();  // protect warp_tot reuse across calls
    if (lane == 31) warp_tot[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        const uint32_t w = lane < nw ? warp_tot[lane] : 0u;
        uint32_t winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc = max(winc, t);
        }
        const uint32_t excl = __shfl_up_sync(0xffffffffu, winc, 1);
        warp_tot[lane] = lane == 0 ? 0u : excl;   // maximum over the earlier warps
    }
    __syncthreads();
    const uint32_t within = __shfl_up_sync(0xffffffffu, inc, 1);
    return max(warp_tot[wid], lane == 0 ? 0u : within);
}

// A clone that resampling created is an alias of its source's slot until somebody writes it
// (PlanArgs::alias_of). The particles whose grids this step writes -- the local particles the index
// vector selects, or all of them -- get their own cells first: one CopyItem (root slot -> own slot)
// per shared grid, in particle order. Clones of one source are neighbours in that order, so the
// fan-out leaders are found as in k_plan: the head of each run of equal roots and every COPY_FAN-th
// item after it. One CTA: the work is three short ordered passes over n_local entries.
__global__ void __launch_bounds__(1024)
k_materialize_list(const uint32_t* __restrict__ idx, uint32_t n_total, uint32_t first_particle, uint32_t S,
                   const int32_t* __restrict__ slot_of, int32_t* alias_of, uint32_t* cells, size_t cells_per_grid,
                   SlotMeta* meta, uint32_t* bands, uint32_t n_bands, CopyItem* __restrict__ items,
                   uint32_t* __restrict__ leaders, uint32_t* roots, StepCounters* counters) {
    __shared__ uint32_t s_warp[33];
    const uint32_t T = blockDim.x, t = threadIdx.x;
    const uint32_t chunk = (S + T - 1) / T;
    const uint32_t c0 = min(S, t * chunk), c1 = min(S, c0 + chunk);
    auto shared_and_written = [&](uint32_t j, int32_t* slot, int32_t* root) {
        if (idx != nullptr) {   // is local particle j selected by the index vector?
            const uint32_t v = first_particle + j;
            const uint32_t q = lower_bound_u32(idx, 0, n_total, v);
            if (q >= n_total || idx[q] != v) return false;
        }
        *slot = slot_of[j];
        *root = alias_of[*slot];
        return *root != *slot;
    };
    uint32_t cnt = 0;
    int32_t slot, root;
    for (uint32_t j = c0; j < c1; ++j) cnt += shared_and_written(j, &slot, &root) ? 1u : 0u;
    uint32_t total;
    uint32_t pos = block_excl_scan_u32(cnt, s_warp, &total);
    for (uint32_t j = c0; j < c1; ++j) {
        if (!shared_and_written(j, &slot, &root)) continue;
        CopyItem it;
        it.src = cells + (size_t)root * cells_per_grid;
        it.src_meta = meta + root;
        it.src_bands = bands + (size_t)root * n_bands;
        it.dst = cells + (size_t)slot * cells_per_grid;
        it.dst_meta = meta + slot;
        it.dst_bands = bands + (size_t)slot * n_bands;
        items[pos] = it;
        roots[pos] = (uint32_t)root;
        alias_of[slot] = slot;   // private from here on (the copy is issued right after this kernel)
        pos++;
    }
    __syncthreads();
    // leaders: run heads and every COPY_FAN-th item of a run
    const uint32_t ichunk = (total + T - 1) / T;
    const uint32_t i0 = min(total, t * ichunk), i1 = min(total, i0 + ichunk);
    uint32_t last_head = 0u;   // position + 1 of the last run head in this thread's items
    for (uint32_t i = i0; i < i1; ++i)
        if (i == 0u || roots[i] != roots[i - 1u]) last_head = i + 1u;
    const uint32_t before = block_excl_scan_max_u32(last_head, s_warp);
    uint32_t run_start = before ? before - 1u : 0u, n_lead_c = 0u;
    for (uint32_t i = i0; i < i1; ++i) {
        if (i == 0u || roots[i] != roots[i - 1u]) run_start = i;
        n_lead_c += ((i - run_start) % COPY_FAN) == 0u;
    }
    uint32_t n_lead;
    uint32_t pl = block_excl_scan_u32(n_lead_c, s_warp, &n_lead);
    run_start = before ? before - 1u : 0u;
    for (uint32_t i = i0; i < i1; ++i) {
        if (i == 0u || roots[i] != roots[i - 1u]) run_start = i;
        if (((i - run_start) % COPY_FAN) == 0u) leaders[pl++] = i;
    }
    if (t == 0) { counters->n_mat = total; counters->n_mat_leaders = n_lead; }
}

// The common case (survivors only): the survivor list is a few hundred entries long and already on the
// device, so the shared grids among them are listed by one thread per survivor, in no particular order
// and without fan-out grouping (every item reads its own source; the list is short).
__global__ void __launch_bounds__(256)
k_materialize_alive(const uint32_t* __restrict__ alive_list, const int32_t* __restrict__ slot_of, int32_t* alias_of,
                    uint32_t* cells, size_t cells_per_grid, SlotMeta* meta, uint32_t* bands, uint32_t n_bands,
                    CopyItem* __restrict__ items, StepCounters* counters) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    bool take = false;
    int32_t slot = 0, root = 0;
    if ((unsigned long long)i < counters->n_alive) {
        slot = slot_of[alive_list[i]];
        root = alias_of[slot];
        take = root != slot;
    }
    const unsigned mask = __ballot_sync(0xffffffffu, take);
    if (mask == 0u) return;
    const int lane = threadIdx.x & 31, leader = __ffs(mask) - 1;
    unsigned long long base = 0;
    if (lane == leader) {
        base = atomicAdd(&counters->n_mat, (unsigned long long)__popc(mask));
        atomicAdd(&counters->n_mat_leaders, (unsigned long long)__popc(mask));   // every item reads its source itself
    }
    base = __shfl_sync(0xffffffffu, base, leader);
    if (!take) return;
    CopyItem it;
    it.src = cells + (size_t)root * cells_per_grid;
    it.src_meta = meta + root;
    it.src_bands = bands + (size_t)root * n_bands;
    it.dst = cells + (size_t)slot * cells_per_grid;
    it.dst_meta = meta + slot;
    it.dst_bands = bands + (size_t)slot * n_bands;
    items[base + __popc(mask & ((1u << lane) - 1u))] = it;
    alias_of[slot] = slot;
}
void launch_materialize_alive(cudaStream_t stream, const uint32_t* alive_list, uint32_t n_local, const int32_t* slot_of,
                              int32_t* alias_of, uint32_t* cells, size_t cells_per_grid, SlotMeta* meta, uint32_t* bands,
                              uint32_t n_bands, CopyItem* items, StepCounters* counters) {
    k_materialize_alive<<<(n_local + 255) / 256, 256, 0, stream>>>(alive_list, slot_of, alias_of, cells, cells_per_grid, meta,
                                                                   bands, n_bands, items, counters);
}

void launch_materialize_list(cudaStream_t stream, const uint32_t* idx, uint32_t n_total, uint32_t first_particle,
                             uint32_t n_local, const int32_t* slot_of, int32_t* alias_of, uint32_t* cells,
                             size_t cells_per_grid, SlotMeta* meta, uint32_t* bands, uint32_t n_bands, CopyItem* items,
                             uint32_t* leaders, uint32_t* roots_scratch, StepCounters* counters) {
    k_materialize_list<<<1, 1024, 0, stream>>>(idx, n_total, first_particle, n_local, slot_of, alias_of, cells, cells_per_grid,
                                               meta, bands, n_bands, items, leaders, roots_scratch, counters);
}

bool plan_can_stage(uint32_t n_local, uint32_t n_spare_cap) {
    return n_local <= PLAN_STAGED_MAX_S && n_spare_cap <= n_local;
}

void launch_plan(cudaStream_t stream, const PlanArgs& a) {
    k_plan<<<1, 1024, a.staged ? plan_staged_bytes(a.n_local) : 0, stream>>>(a);
}

cudaError_t configure_resample_kernels() {
    return cudaFuncSetAttribute(k_plan, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan_staged_bytes(PLAN_STAGED_MAX_S));
}

}  // namespace slamrs
