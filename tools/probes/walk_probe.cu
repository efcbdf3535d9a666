// Micro-benchmark for the ray walk's inner loop (tuning only; not part of the product library).
// Each thread walks one ray of a fan from the window's centre with the iterator's arithmetic; variants differ
// in how a visited cell is recorded. Prints cycles per CTA for 1..4 co-resident CTAs per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o walk_probe walk_probe.cu
#include <cstdio>
#include <cstdint>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

constexpr int W = 264, H = 132;          // half-disc window (upper half), 16-bit cells, row-major
constexpr int CX = 132, CY = 0;

template <int V>
__global__ void __launch_bounds__(192) k_walk(const float* __restrict__ ang, const float* __restrict__ len, int n_beams,
                                              unsigned long long* cycles, uint32_t* sink) {
    extern __shared__ __align__(16) uint32_t s_win[];
    for (int i = threadIdx.x; i < W * H / 2; i += blockDim.x) s_win[i] = 0u;
    __syncthreads();
    const long long t0 = clock64();
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(s_win);
    uint32_t acc_sink = 0;
    for (int b = threadIdx.x; b < n_beams; b += blockDim.x) {
        const float sx = CX + 0.37f, sy = CY + 0.41f;
        const float x1 = sx + len[b] * cosf(ang[b]), y1 = sy + len[b] * sinf(ang[b]);
        const float delta_x = fabsf(x1 - sx), delta_y = fabsf(y1 - sy);
        int x_inc = x1 > sx ? 1 : -1, y_inc = 1;
        float error = x1 > sx ? __fmul_rn(__fsub_rn(floorf(sx) + 1.0f, sx), delta_y) : __fmul_rn(__fsub_rn(sx, floorf(sx)), delta_y);
        error = __fsub_rn(error, __fmul_rn(__fsub_rn(floorf(sy) + 1.0f, sy), delta_x));
        int remaining = 3 + abs((int)floorf(x1) - CX) + abs((int)floorf(y1) - CY);
        const float x_step = (float)x_inc, y_step = (float)y_inc;
        float cxf = CX + 0.5f, cyf = CY + 0.5f;
        float dxs = __fsub_rn(sx, cxf), dys = __fsub_rn(sy, cyf);
        float dx2 = __fmul_rn(dxs, dxs), dy2 = __fmul_rn(dys, dys);
        const float free_below = (len[b] - 1.0f) * (len[b] - 1.0f);
        uint32_t x2 = 2u * CX;
        const uint32_t x_inc2 = (uint32_t)(2 * x_inc);
        uint32_t rb = base;     // row 0
        int ly = 0;
        if (V == 3) {
            // 2x2 blocks of 16-bit cells in one 64-bit word; pending adds are kept in a register pair
            uint32_t bx = CX >> 1, by = 0;
            unsigned long long pend = 0ull;
            int x = CX, y = 0;
            while (remaining > 0) {
                const float acc = __fadd_rn(dx2, dy2);
                if (!(acc < free_below)) break;
                pend += 1ull << (16 * ((x & 1) + 2 * (y & 1)));
                if (error > 0.0f) {
                    error = __fsub_rn(error, delta_x); cyf = __fadd_rn(cyf, y_step); dys = __fsub_rn(sy, cyf); dy2 = __fmul_rn(dys, dys);
                    y += 1;
                } else {
                    error = __fadd_rn(error, delta_y); cxf = __fadd_rn(cxf, x_step); dxs = __fsub_rn(sx, cxf); dx2 = __fmul_rn(dxs, dxs);
                    x += x_inc;
                }
                const uint32_t nbx = (uint32_t)x >> 1, nby = (uint32_t)y >> 1;
                if (nbx != bx || nby != by) {
                    const uint32_t a = base + 8u * (by * (W / 2) + bx);
                    asm volatile("red.shared.add.u64 [%0], %1;" ::"r"(a), "l"(pend) : "memory");
                    pend = 0ull; bx = nbx; by = nby;
                }
                remaining -= 1;
            }
            if (pend) { const uint32_t a = base + 8u * (by * (W / 2) + bx); asm volatile("red.shared.add.u64 [%0], %1;" ::"r"(a), "l"(pend) : "memory"); }
            continue;
        }
        while (remaining > 0) {
            const float acc = __fadd_rn(dx2, dy2);
            if (!(acc < free_below)) break;
            const uint32_t c2 = rb + x2;
            if (V == 0) asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(c2 & ~3u), "r"((c2 & 2u) ? 0x10000u : 1u) : "memory");
            if (V == 1) acc_sink += c2;
            if (V == 2) { unsigned short v; asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(c2) : "memory"); v += 1; asm volatile("st.shared.u16 [%0], %1;" ::"r"(c2), "h"(v) : "memory"); }
            if (V == 4) { asm volatile("st.shared.u16 [%0], %1;" ::"r"(c2), "h"((unsigned short)1) : "memory"); }
            if (error > 0.0f) {
                error = __fsub_rn(error, delta_x); cyf = __fadd_rn(cyf, y_step); dys = __fsub_rn(sy, cyf); dy2 = __fmul_rn(dys, dys);
                ly += y_inc; rb = base + (uint32_t)ly * (2u * W);
            } else {
                error = __fadd_rn(error, delta_y); cxf = __fadd_rn(cxf, x_step); dxs = __fsub_rn(sx, cxf); dx2 = __fmul_rn(dxs, dxs);
                x2 += x_inc2;
            }
            remaining -= 1;
        }
    }
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0) { atomicAdd(cycles, (unsigned long long)(t1 - t0)); }
    uint32_t s = acc_sink;
    for (int i = threadIdx.x; i < W * H / 2; i += blockDim.x) s += s_win[i];
    if (s == 0xdeadbeefu) sink[0] = s;
    if (blockIdx.x == 0 && threadIdx.x == 0) sink[1] = 0;
}

template <int V>
void run(const char* name, const float* d_ang, const float* d_len, int n_beams, int ctas_per_sm, unsigned long long* d_cyc, uint32_t* d_sink) {
    const size_t smem = (size_t)W * H * 2;
    cudaFuncSetAttribute(k_walk<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int grid = 148 * ctas_per_sm;
    for (int rep = 0; rep < 2; ++rep) {
        cudaMemset(d_cyc, 0, 8);
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        k_walk<V><<<grid, 192, smem>>>(d_ang, d_len, n_beams, d_cyc, d_sink);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        unsigned long long c; cudaMemcpy(&c, d_cyc, 8, cudaMemcpyDeviceToHost);
        if (rep) printf("%-28s ctas/sm %d: %8.0f cycles per CTA (walk), kernel %.1f us  [%s]\n", name, ctas_per_sm, (double)c / grid, ms * 1e3, cudaGetErrorString(cudaGetLastError()));
    }
}

int main() {
    const int n_beams = 180;
    std::vector<float> ang(n_beams), len(n_beams);
    // beams of a half scan (1 degree apart), shuffled as a sort by distance would, lengths 60..120 cells, longest first
    for (int i = 0; i < n_beams; ++i) { ang[i] = (float)(((i * 67) % 180) + 0.5f) * 3.14159265f / 180.0f; len[i] = 120.0f - 60.0f * i / n_beams; }
    float *d_ang, *d_len; unsigned long long* d_cyc; uint32_t* d_sink;
    cudaMalloc(&d_ang, n_beams * 4); cudaMalloc(&d_len, n_beams * 4); cudaMalloc(&d_cyc, 8); cudaMalloc(&d_sink, 8);
    cudaMemcpy(d_ang, ang.data(), n_beams * 4, cudaMemcpyHostToDevice); cudaMemcpy(d_len, len.data(), n_beams * 4, cudaMemcpyHostToDevice);
    double steps = 0; for (int i = 0; i < n_beams; ++i) steps += len[i] * (fabs(cos(ang[i])) + fabs(sin(ang[i])));
    printf("cell-steps per CTA ~ %.0f, longest ray ~ %.0f steps\n", steps, 120 * 1.414);
    for (int c = 1; c <= 3; c += 1) {
        run<0>("red.shared.u32 per step", d_ang, d_len, n_beams, c, d_cyc, d_sink);
        run<1>("no memory op", d_ang, d_len, n_beams, c, d_cyc, d_sink);
        run<2>("ld/st.shared.u16 (racy)", d_ang, d_len, n_beams, c, d_cyc, d_sink);
        run<4>("st.shared.u16 only", d_ang, d_len, n_beams, c, d_cyc, d_sink);
        run<3>("2x2 block, red.u64 on exit", d_ang, d_len, n_beams, c, d_cyc, d_sink);
    }
    return 0;
}
