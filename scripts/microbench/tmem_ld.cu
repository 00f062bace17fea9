// Micro-benchmark: how fast can the epilogue warps of one SM read accumulators out of tensor memory?
// tcgen05.ld.32x32b.xN issued by NW warps (warp w reads lane quarter w % 4), nothing else running on the SM.
// The fused distance + argmin kernel must look at every score once, so this read rate is a hard floor for small D
// (DESIGN.md section 7).  Build + run on a B200:
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/tmem_ld scripts/microbench/tmem_ld.cu && /tmp/tmem_ld
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define R8(r, o) "=r"(r[o + 0]), "=r"(r[o + 1]), "=r"(r[o + 2]), "=r"(r[o + 3]), "=r"(r[o + 4]), "=r"(r[o + 5]), "=r"(r[o + 6]), "=r"(r[o + 7])
#define U8(r, o) "r"(r[o + 0]), "r"(r[o + 1]), "r"(r[o + 2]), "r"(r[o + 3]), "r"(r[o + 4]), "r"(r[o + 5]), "r"(r[o + 6]), "r"(r[o + 7])

__device__ __forceinline__ void ld_x16(uint32_t (&r)[16], uint32_t taddr) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : R8(r, 0), R8(r, 8) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void ld_x32(uint32_t (&r)[32], uint32_t taddr) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
                 "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : R8(r, 0), R8(r, 8), R8(r, 16), R8(r, 24) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void ld_x64(uint32_t (&r)[64], uint32_t taddr) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x64.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
                 "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,"
                 "%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,"
                 "%48,%49,%50,%51,%52,%53,%54,%55,%56,%57,%58,%59,%60,%61,%62,%63}, [%64];"
                 : R8(r, 0), R8(r, 8), R8(r, 16), R8(r, 24), R8(r, 32), R8(r, 40), R8(r, 48), R8(r, 56) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
template <int N>
__device__ __forceinline__ float xor_of(uint32_t (&r)[N], float m) {   // cheapest possible use (one LOP3 per 2 scores): ptxas drops unused loads
    uint32_t x = __float_as_uint(m);
#pragma unroll
    for (int i = 0; i < N; i += 2) x ^= r[i] ^ r[i + 1];
    return __uint_as_float(x);
}
template <int N>
__device__ __forceinline__ float max_of(uint32_t (&r)[N], float m) {   // the real epilogue's minimum: a 3-input max per 2 scores
#pragma unroll
    for (int i = 0; i < N; i += 2) m = fmaxf(fmaxf(__uint_as_float(r[i]), __uint_as_float(r[i + 1])), m);
    return m;
}

// MODE 0: x32 + xor fold (LOP3)          MODE 1: two x32 per wait + xor fold      MODE 2: x64 + running maximum
// MODE 3: x32 + running maximum (FMNMX3) MODE 4: x16 + running maximum            MODE 5: two x32 per wait + running maximum
template <int MODE>
__global__ void __launch_bounds__(512) tmem_ld_kernel(int iters, long long* clocks, float* sink) {
    __shared__ uint32_t tmem_base;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&tmem_base)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = tmem_base + ((uint32_t)(warp & 3) * 32u << 16);   // lane quarter of this warp
    const uint32_t colw = (uint32_t)(warp >> 2) * 128u;                     // warps of one quarter start in different columns
    float m = -1e30f;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        const uint32_t c = (colw + (uint32_t)i * 64u) & 511u;
        if (MODE == 0 || MODE == 3) {
            uint32_t r[32];
            ld_x32(r, base + (c & 480u));
            wait_ld();
            if (MODE == 3) m = max_of(r, m); else m = xor_of(r, m);
        } else if (MODE == 1 || MODE == 5) {
            uint32_t a[32], b[32];
            ld_x32(a, base + (c & 448u));
            ld_x32(b, base + (c & 448u) + 32u);
            wait_ld();
            if (MODE == 5) { m = max_of(a, m); m = max_of(b, m); } else { m = xor_of(a, m); m = xor_of(b, m); }
        } else if (MODE == 2) {
            uint32_t r[64];
            ld_x64(r, base + (c & 448u));
            wait_ld();
            m = max_of(r, m);
        } else {
            uint32_t r[16];
            ld_x16(r, base + (c & 496u));
            wait_ld();
            m = max_of(r, m);
        }
    }
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0) clocks[blockIdx.x] = t1 - t0;
    if (m == 123.456f) sink[0] = m;   // never true in practice, keeps the folds alive
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}

template <int MODE>
static void run(const char* name, int cols_per_iter, int nw, int sms, long long* d_clk, float* d_sink) {
    const int iters = 4096;
    tmem_ld_kernel<MODE><<<sms, 32 * nw>>>(64, d_clk, d_sink);   // warm-up
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    tmem_ld_kernel<MODE><<<sms, 32 * nw>>>(iters, d_clk, d_sink);
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(err)); return; }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    long long clk[256];
    cudaMemcpy(clk, d_clk, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
    double mean = 0;
    for (int i = 0; i < sms; ++i) mean += (double)clk[i];
    mean /= sms;
    const double bytes = (double)nw * iters * 32.0 * cols_per_iter * 4.0;   // per SM
    printf("%-34s warps %2d: %8.1f B/clk/SM = %6.1f scores/clk/SM   (%.0f clk, %.3f ms, %.2f TB/s chip-wide)\n", name, nw, bytes / mean,
           bytes / mean / 4.0, mean, ms, bytes * sms / (ms * 1e-3) / 1e12);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    printf("%s, %d SMs, cc %d.%d\n", p.name, sms, p.major, p.minor);
    long long* d_clk; float* d_sink;
    cudaMalloc(&d_clk, sizeof(long long) * 256);
    cudaMalloc(&d_sink, 4);
    for (int nw : {4, 8, 16}) {
        run<4>("x16 + running max", 16, nw, sms, d_clk, d_sink);
        run<3>("x32 + running max", 32, nw, sms, d_clk, d_sink);
        run<2>("x64 + running max", 64, nw, sms, d_clk, d_sink);
        run<5>("2 x x32 per wait + running max", 64, nw, sms, d_clk, d_sink);
        run<0>("x32 + xor fold", 32, nw, sms, d_clk, d_sink);
        run<1>("2 x x32 per wait + xor fold", 64, nw, sms, d_clk, d_sink);
    }
    return 0;
}
