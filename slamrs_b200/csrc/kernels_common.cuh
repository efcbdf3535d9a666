// Device helpers shared by the kernel translation units (kernels_*.cu): streaming 128/256-bit
// accesses and deterministic block-wide prefix sums.
#pragma once
#include "kernels.cuh"
#include "shared_stream.cuh"
#include <cooperative_groups.h>

namespace slamrs {

// Programmatic dependent launch (sm_90+): a kernel launched with launch_pdl() may become resident while the kernel in
// front of it in the stream still runs -- once that one has let its dependents go (pdl_launch_dependents) -- and must
// not touch anything the stream order protects before pdl_wait() returns (the prerequisite grid has completed and its
// writes are visible). Without the launch attribute both are no-ops. Takes the launch latency and the ramp of the
// short kernels of the step (motion -> likelihood -> weights -> indices) out of the critical path.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}


namespace cg = cooperative_groups;


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

}  // namespace slamrs
