// Bandwidth probe for the grid-copy kernel design (not part of the product):
// pure write streams, fan-out copies and plain copies with several store flavours on one B200.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/bw_probe tools/bw_probe.cu && /tmp/bw_probe
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__device__ __forceinline__ uint4 ld_nc(const uint4* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
template <int MODE> __device__ __forceinline__ void st(uint4* p, const uint4& v) {
    if (MODE == 0) asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
    if (MODE == 1) asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
    if (MODE == 2) *p = v;
    if (MODE == 3) asm volatile("st.global.wt.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
    if (MODE == 4) {
        uint64_t pol;
        asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
        asm volatile("st.global.L1::no_allocate.L2::cache_hint.v4.u32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "l"(pol) : "memory");
    }
}
__device__ __forceinline__ void st256(uint4* p, const uint4& a, const uint4& b) {
    asm volatile("st.global.L1::no_allocate.v8.u32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w) : "memory");
}
__device__ __forceinline__ void ld256(const uint4* p, uint4& a, uint4& b) {
    asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "l"(p));
}
// 256-bit variant: each thread moves 32 B per access, UNROLL accesses
template <int UNROLL, int THREADS>
__global__ void __launch_bounds__(THREADS) k_fan256(const uint4* __restrict__ src_base, uint4* __restrict__ dst_base, size_t v4_per_grid,
                                                    int n_src, int fan) {
    const uint32_t item = THREADS * UNROLL * 2;  // uint4 per item
    const uint32_t chunks = (uint32_t)((v4_per_grid + item - 1) / item);
    const size_t total = (size_t)n_src * chunks;
    for (size_t w = blockIdx.x; w < total; w += gridDim.x) {
        const size_t s = w / chunks; const uint32_t c = (uint32_t)(w - s * chunks);
        const uint4* src = src_base + s * v4_per_grid + (size_t)c * item;
        uint4 a[UNROLL], b[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) ld256(src + 2 * (threadIdx.x + u * THREADS), a[u], b[u]);
        for (int f = 0; f < fan; ++f) {
            uint4* dst = dst_base + ((size_t)s * fan + f) * v4_per_grid + (size_t)c * item;
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) st256(dst + 2 * (threadIdx.x + u * THREADS), a[u], b[u]);
        }
    }
}

// fan-out: read chunk of src once, write to FAN destinations (dst stride = grid size)
template <int MODE, int UNROLL>
__global__ void __launch_bounds__(256) k_fan(const uint4* __restrict__ src_base, uint4* __restrict__ dst_base, size_t v4_per_grid,
                                             int n_src, int fan) {
    const uint32_t item = 256 * UNROLL;
    const uint32_t chunks = (uint32_t)((v4_per_grid + item - 1) / item);
    const size_t total = (size_t)n_src * chunks;
    for (size_t w = blockIdx.x; w < total; w += gridDim.x) {
        const size_t s = w / chunks; const uint32_t c = (uint32_t)(w - s * chunks);
        const uint4* src = src_base + s * v4_per_grid;
        uint4 v[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) v[u] = ld_nc(src + (size_t)c * item + threadIdx.x + u * 256);
        for (int f = 0; f < fan; ++f) {
            uint4* dst = dst_base + ((size_t)s * fan + f) * v4_per_grid;
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) st<MODE>(dst + (size_t)c * item + threadIdx.x + u * 256, v[u]);
        }
    }
}

template <int MODE>
__global__ void __launch_bounds__(256) k_fill(uint4* __restrict__ dst, size_t n) {
    const uint4 v = make_uint4(1, 2, 3, 4);
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) st<MODE>(dst + i, v);
}

template <typename F> float time_ms(F f, int reps = 5) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) { cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms; }
    return best;
}

int main() {
    const size_t grid_bytes = 4u << 20, v4 = grid_bytes / 16;
    const int n_dst = 2048;             // 8 GiB of destinations
    const int n_src_max = 2048;
    uint4 *src, *dst;
    CK(cudaMalloc(&src, (size_t)n_src_max * grid_bytes));
    CK(cudaMalloc(&dst, (size_t)n_dst * grid_bytes));
    CK(cudaMemset(src, 1, (size_t)n_src_max * grid_bytes));
    int sms = 148;
    const double out_gb = (double)n_dst * grid_bytes / 1e9;
    printf("memset (driver)      : %7.1f GB/s written\n", out_gb / (time_ms([&] { cudaMemsetAsync(dst, 0, (size_t)n_dst * grid_bytes); }) * 1e-3));
    printf("memcpy D2D (driver)  : %7.1f GB/s r+w\n", 2 * out_gb / (time_ms([&] { cudaMemcpyAsync(dst, src, (size_t)n_dst * grid_bytes, cudaMemcpyDeviceToDevice); }) * 1e-3));
#define FILL(M, G) printf("fill mode %d grid %4d : %7.1f GB/s written\n", M, G, out_gb / (time_ms([&] { k_fill<M><<<G, 256>>>(dst, (size_t)n_dst * v4); }) * 1e-3));
    FILL(0, sms * 8) FILL(1, sms * 8) FILL(2, sms * 8) FILL(3, sms * 8) FILL(0, sms * 4) FILL(0, sms * 16) FILL(0, sms * 32)
#define FAN(M, U, G, F) { int ns = n_dst / (F); float ms = time_ms([&] { k_fan<M, U><<<G, 256>>>(src, dst, v4, ns, F); }); \
      printf("fan mode %d unroll %d grid %4d fan %2d : %6.3f ms  %7.1f GB/s moved (%.1f written)\n", M, U, G, F, ms, (out_gb + ns * grid_bytes / 1e9) / (ms * 1e-3), out_gb / (ms * 1e-3)); }
    for (int f : {1, 2, 4, 16}) { FAN(0, 4, sms * 8, f) }
    for (int f : {1, 16}) { FAN(1, 4, sms * 8, f) FAN(2, 4, sms * 8, f) FAN(0, 8, sms * 4, f) FAN(0, 2, sms * 8, f) FAN(0, 4, sms * 16, f) FAN(0, 4, sms * 4, f) FAN(0, 4, sms * 2, f) }
    printf("---- evict_first policy stores\n");
    FILL(4, sms * 8)
    for (int f : {1, 16}) { FAN(4, 4, sms * 8, f) FAN(4, 4, sms * 16, f) }
    printf("---- 256-bit accesses\n");
#define FAN256(U, T, G, F) { int ns = n_dst / (F); float ms = time_ms([&] { k_fan256<U, T><<<G, T>>>(src, dst, v4, ns, F); }); \
      printf("fan256 unroll %d threads %d grid %4d fan %2d : %6.3f ms  %7.1f GB/s moved (%.1f written)\n", U, T, G, F, ms, (out_gb + ns * grid_bytes / 1e9) / (ms * 1e-3), out_gb / (ms * 1e-3)); }
    for (int f : {1, 16}) { FAN256(2, 256, sms * 8, f) FAN256(4, 128, sms * 16, f) FAN256(2, 256, sms * 16, f) FAN256(2, 512, sms * 4, f) FAN256(4, 256, sms * 8, f) FAN256(2, 256, sms * 32, f) }
    for (int f : {16}) { FAN(0, 4, sms * 32, f) FAN(0, 4, sms * 64, f) FAN(0, 1, sms * 8, f) FAN(0, 1, sms * 64, f) }
    return 0;
}
