// CUDA-core kernels of the VQ bottleneck for sm_100a: codebook / latent preparation, the exact fp32 search
// (parity anchor + fallback of the tensor-core shortlist), the tail (fp32 rescoring in the reference's op order,
// codeword gather, SSE, straight-through value, per-code statistics; TMA-fed and register-staged forms) and the
// backward kernels.
//
// Reference semantics restated here (never copied): src/model/components/vector_quantizer.py:25-52.
// All of these are HBM-bound: they read latents straight from the reference's [B, D, W] layout (TMA boxes or
// frame-contiguous coalesced accesses), transpose through padded shared memory where a gather needs frame-major rows,
// and touch every latent byte once per kernel.
#include "vqb_internal.h"
#include "vqb_ptx.cuh"

#include <math.h>
#include <stdlib.h>

namespace vqb {

// ------------------------------------------------------------------------------------------------ small helpers
__device__ __forceinline__ float ld_stream(const float* p) {   // read-once data: keep it out of L1
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// better() / ref_distance() / red_add_v4(): vqb_internal.h (shared with the fused tail of the tensor-core kernel)

// Caller-supplied indices (BERT-predicted tokens, bert.py:72-78) are never trusted: a code outside [0, K) touches no memory.
// Its one-hot row stays all-zero, its gathered codeword / gradient is NaN (loud in the data); the Python wrappers also
// validate on the host and raise IndexError (the reference's scatter_ / one-hot matmul raise a device assert).
__device__ __forceinline__ bool code_ok(int64_t k, int K) { return (unsigned long long)k < (unsigned long long)K; }

// ------------------------------------------------------------------------------------------------ codebook prep
// One warp per code: |e_k|^2 in fp32 (vector_quantizer.py:33), bf16 copy for the tensor-core tiles, max |e|^2 for the
// guard band, non-finite flag.  Rows K..K_pad-1 get e2 = +inf (never shortlisted) and zero operands.
__global__ void __launch_bounds__(256) codebook_prep_kernel(const float* __restrict__ E, int K, int K_pad, int D,
                                                            float* __restrict__ e2, __nv_bfloat16* __restrict__ eb,
                                                            __nv_bfloat16* __restrict__ eh, WsMeta* meta, bool tf32,
                                                            float* __restrict__ ep, int ep_lpf) {
    const int lane = threadIdx.x & 31;
    const int k = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (k >= K_pad) return;
    float s = 0.f, st = 0.f, sd = 0.f;
    for (int d = lane; d < D; d += 32) {
        const float v = (k < K) ? E[(size_t)k * D + d] : 0.f;
        s = __fadd_rn(s, __fmul_rn(v, v));
        if (ep && k < K) ep[(size_t)k * D + tail3_perm_pos(d, ep_lpf)] = v;   // permuted fp32 copy for tail3_kernel
        if (tf32) {                               // what kind::tf32 reads: the low 13 mantissa bits dropped.  |e| bounds |tf32(e)| also
            const float vt = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u), dv = v - vt;   // for a rounding unit (factor below)
            st = fmaf(v, v, st);
            sd = fmaf(dv, dv, sd);
        } else if (eb) {
            const __nv_bfloat16 h = __float2bfloat16_rn(v);
            const float vt = __bfloat162float(h), dv = v - vt;
            st = fmaf(vt, vt, st);
            sd = fmaf(dv, dv, sd);
            eb[(size_t)k * D + d] = h;
        }
    }
    s = warp_sum(s);
    st = warp_sum(st);
    sd = warp_sum(sd);
    if (lane == 0) {
        if (k < K) {
            e2[k] = s;
            if (!isfinite(s)) atomicExch(&meta->cb_nonfinite, 1);
            else {
                atomicMax(&meta->emax2_bits, __float_as_uint(s));
                if (eb || tf32) {
                    atomicMax(&meta->etmax2_bits, __float_as_uint(tf32 ? st * 1.002f : st));   // |rn_tf32(e)| <= |e| (1 + 2^-11)
                    atomicMax(&meta->demax2_bits, __float_as_uint(sd));
                }
            }
        } else {
            e2[k] = INFINITY;
        }
        if (eh) {
            // three-term bf16 split of |e_k|^2 / 2 (24 bits): the bias operand of the tensor-core contraction.
            // Padding codes get a huge finite bias so their accumulator can never be the maximum.
            const float half = (k < K) ? 0.5f * s : 3.0e38f;
            const __nv_bfloat16 h1 = __float2bfloat16_rn(half);
            const float r1 = half - __bfloat162float(h1);
            const __nv_bfloat16 h2 = __float2bfloat16_rn(r1);
            const __nv_bfloat16 h3 = __float2bfloat16_rn(r1 - __bfloat162float(h2));
            __nv_bfloat16* o = eh + (size_t)k * 8;
            o[0] = h1; o[1] = (k < K) ? h2 : __float2bfloat16_rn(0.f); o[2] = (k < K) ? h3 : __float2bfloat16_rn(0.f);
            o[3] = o[4] = o[5] = o[6] = o[7] = __float2bfloat16_rn(0.f);
        }
    }
}

cudaError_t launch_codebook_prep(const float* codebook, int K, int K_pad, int D, float* e2, __nv_bfloat16* eb, __nv_bfloat16* eh,
                                 WsMeta* meta, cudaStream_t s, bool tf32, float* ep, int ep_lpf) {
    codebook_prep_kernel<<<(K_pad + 7) / 8, 256, 0, s>>>(codebook, K, K_pad, D, e2, eb, eh, meta, tf32, (D % 32 == 0) ? ep : nullptr, ep_lpf);
    note_launch();
    return cudaGetLastError();
}

static int tile_ldg_mode() {   // tile loads: register-staged LDG batches (default, measured faster) or cp.async (VQB_TILE_LDG=0)
    return env_get(ENV_TILE_LDG, 1) == 0 ? 0 : 1;
}

// ------------------------------------------------------------------------------------------------ frame tiles
// All streaming kernels work on tiles of 32 frames x D held frame-major in shared memory with rows of D+4 floats: the
// transposing global<->shared phases are coalesced over frames, the per-frame phase reads float4 rows, and LPF lanes
// cooperate on one frame (32 / LPF frames per warp at a time), each lane owning 4*J dims.
constexpr int TL_F = 32;              // frames per tile

template <int LPF>
__device__ __forceinline__ float group_sum(float v) {   // sum over the LPF lanes that share a frame
#pragma unroll
    for (int o = LPF / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// coalesced BCW -> frame-major shared tile: thread = (frame = tid & 31, dim group = tid >> 5), 4 dims per access.
// Scalar LDG (1.8 cycles per warp-instruction) + one conflict-free STS.128 per 4 dims is the cheapest transposition in
// pipe time; loads are issued in batches of 16 per thread (4 row groups) before any store so enough bytes are in flight.
template <int NW = 8>
__device__ __forceinline__ void tile_load(float* Xs, int ld, const float* __restrict__ z, size_t col, int64_t W, int D, bool valid) {
    const int f = threadIdx.x & 31;
    const float* p0 = z + (valid ? col : 0);
    for (int d0 = (threadIdx.x >> 5) * 4; d0 < D; d0 += 16 * NW) {
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int d = d0 + 4 * NW * u;
            v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (valid && d < D) {
                const float* p = p0 + (size_t)d * W;
                v[u].x = ld_stream(p);
                v[u].y = ld_stream(p + W);
                v[u].z = ld_stream(p + 2 * W);
                v[u].w = ld_stream(p + 3 * W);
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int d = d0 + 4 * NW * u;
            if (d < D) *reinterpret_cast<float4*>(Xs + f * ld + d) = v[u];
        }
    }
}
// Same transposition with cp.async (LDGSTS): no register staging, so every thread has its whole share of the tile
// (D/8 4-byte copies) in flight at once - the memory-level parallelism a latency-bound streaming kernel needs.
// Out-of-range frames are zero-filled (src-size 0).  Follow with tile_fetch_wait() + __syncthreads().
template <int NW = 8>
__device__ __forceinline__ void tile_fetch_async(float* Xs, int ld, const float* __restrict__ z, size_t col, int64_t W, int D, bool valid) {
    const int f = threadIdx.x & 31;
    const uint32_t dst0 = (uint32_t)__cvta_generic_to_shared(Xs + f * ld);
    const float* src0 = z + (valid ? col : 0);
    const uint32_t nbytes = valid ? 4u : 0u;
    for (int d = threadIdx.x >> 5; d < D; d += NW)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst0 + 4u * d), "l"(src0 + (size_t)d * W), "r"(nbytes) : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
}
__device__ __forceinline__ void tile_fetch_wait() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

template <int NW = 8>
__device__ __forceinline__ void tile_store(const float* Xs, int ld, float* __restrict__ out, size_t col, int64_t W, int D, bool valid) {
    const int f = threadIdx.x & 31;
    if (!valid) return;
    for (int d0 = (threadIdx.x >> 5) * 4; d0 < D; d0 += 4 * NW) {
        const float4 v = *reinterpret_cast<const float4*>(Xs + f * ld + d0);
        float* p = out + col + (size_t)d0 * W;
        st_stream(p, v.x);
        st_stream(p + W, v.y);
        st_stream(p + 2 * W, v.z);
        st_stream(p + 3 * W, v.w);
    }
}

// lanes per frame and float4 groups per lane for a given D: 8 lanes x (D/32) groups up to D = 64, 16 lanes x (D/64) beyond
#define VQB_DISPATCH_D(D, CALL)                    \
    do {                                           \
        if ((D) <= 32) { CALL(8, 1); }             \
        else if ((D) <= 64) { CALL(8, 2); }        \
        else if ((D) <= 128) { CALL(16, 2); }      \
        else if ((D) <= 192) { CALL(16, 3); }      \
        else if ((D) <= 256) { CALL(16, 4); }      \
        else if ((D) <= 384) { CALL(16, 6); }      \
        else { CALL(16, 8); }                      \
    } while (0)

// ------------------------------------------------------------------------------------------------ latent prep (bf16)
// z [B, D, W] fp32 -> xb [N_pad, D] bf16 (frame-major = K-major A operand for tcgen05) and the per-frame guard band.
// One block = 32 frames x all D, transposed through shared memory so both the read and the write are coalesced.
//
// Guard band (DESIGN.md "shortlist exactness"): with xt = bf16(x) = x + dx and et = bf16(e) = e + de,
//   xt.et - x.e = dx.et + x.de,  so  |score_bf16(k) - score(k)| <= 2 (|dx| |et_k| + |x| |de_k|) + accumulation error.
// Two codes are compared, hence the factor 4; the last term bounds fp32 accumulation in the tensor core and in the
// reference's own sgemm.  |dx|, |x| are measured per frame here; max|et|, max|de|, max|e| come from codebook_prep.
template <int LPF, int J>
__global__ void __launch_bounds__(256, (J >= 6) ? 2 : 3) latent_prep_bf16_kernel(const float* __restrict__ z, int D, int64_t W, int64_t N,
                                                                  int64_t N_pad, __nv_bfloat16* __restrict__ xb,
                                                                  float* __restrict__ band, const WsMeta* __restrict__ meta, int ldg) {
    extern __shared__ __align__(16) float Xbuf[];   // 2 x [32][D + 4]
    constexpr int FPW = 32 / LPF;
    constexpr int ITER = (TL_F / 8) / FPW;
    const int ld = D + 4;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = lane / LPF, sl = lane % LPF;
    const float etmax = sqrtf(__uint_as_float(meta->etmax2_bits)) * 1.0001f;
    const float demax = sqrtf(__uint_as_float(meta->demax2_bits)) * 1.0001f;
    const float emax = sqrtf(__uint_as_float(meta->emax2_bits)) * 1.0001f;
    auto tile_col = [&](int64_t tile, bool& valid) -> size_t {
        const int64_t nl = tile * TL_F + lane;
        valid = nl < N;
        int64_t b = 0, w = 0;
        if (valid) { b = nl / W; w = nl - b * W; }
        return (size_t)b * D * W + w;
    };
    int buf = 0;
    {
        bool v0;
        const size_t c0 = tile_col(blockIdx.x, v0);
        if (!ldg && (int64_t)blockIdx.x * TL_F < N_pad) tile_fetch_async(Xbuf, ld, z, c0, W, D, v0);
    }
    for (int64_t tile = blockIdx.x; tile * TL_F < N_pad; tile += gridDim.x, buf ^= 1) {
        float* Xs = Xbuf + (size_t)buf * TL_F * ld;
        __syncthreads();                          // every warp is done reading the other buffer
        const int64_t next = tile + gridDim.x;
        if (ldg) {
            bool vc;
            const size_t cc = tile_col(tile, vc);
            tile_load(Xs, ld, z, cc, W, D, vc);
        } else if (next * TL_F < N_pad) {
            bool vn;
            const size_t cn = tile_col(next, vn);
            tile_fetch_async(Xbuf + (size_t)(buf ^ 1) * TL_F * ld, ld, z, cn, W, D, vn);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            tile_fetch_wait();
        }
        __syncthreads();
#pragma unroll
        for (int it = 0; it < ITER; ++it) {
            const int f = warp * (TL_F / 8) + it * FPW + sub;
            const int64_t n = tile * TL_F + f;
            float s = 0.f, sd = 0.f;
#pragma unroll
            for (int j = 0; j < J; ++j) {
                const int d = 4 * sl + 4 * LPF * j;
                if (d < D) {
                    const float4 v = *reinterpret_cast<const float4*>(Xs + f * ld + d);
                    const __nv_bfloat162 h01 = __floats2bfloat162_rn(v.x, v.y), h23 = __floats2bfloat162_rn(v.z, v.w);
                    const float2 b01 = __bfloat1622float2(h01), b23 = __bfloat1622float2(h23);
                    const float d0 = v.x - b01.x, d1 = v.y - b01.y, d2 = v.z - b23.x, d3 = v.w - b23.y;
                    s = fmaf(v.x, v.x, s); s = fmaf(v.y, v.y, s); s = fmaf(v.z, v.z, s); s = fmaf(v.w, v.w, s);
                    sd = fmaf(d0, d0, sd); sd = fmaf(d1, d1, sd); sd = fmaf(d2, d2, sd); sd = fmaf(d3, d3, sd);
                    if (n < N_pad) {
                        uint2 pk;
                        pk.x = *reinterpret_cast<const uint32_t*>(&h01);
                        pk.y = *reinterpret_cast<const uint32_t*>(&h23);
                        *reinterpret_cast<uint2*>(xb + (size_t)n * D + d) = pk;     // frames past N are written as zeros
                    }
                }
            }
            s = group_sum<LPF>(s);
            sd = group_sum<LPF>(sd);
            if (sl == 0 && n < N) {
                const float xn = sqrtf(s) * 1.0001f, dxn = sqrtf(sd) * 1.0001f;
                // last terms: fp32 accumulation (tensor core and the reference's sgemm), the 3-term bf16 split of |e|^2/2
                band[n] = 4.0f * (dxn * etmax + xn * demax) * 1.001f + 8.0f * (float)(D + 16) * 2.3841858e-07f * xn * emax +
                          4.0e-7f * emax * emax + 1e-30f;
            }
        }
    }
}

cudaError_t launch_latent_prep_bf16(const float* z, int B, int D, int64_t W, int64_t N_pad, __nv_bfloat16* xb, float* band,
                                    const WsMeta* meta, cudaStream_t s) {
    const int64_t N = (int64_t)B * W;
    const size_t smem = (size_t)2 * TL_F * (D + 4) * 4;
    const int64_t tiles = (N_pad + TL_F - 1) / TL_F;
    int64_t grid = tiles < kTailGridMax ? tiles : kTailGridMax;
    if (grid < 1) grid = 1;
    cudaError_t e = cudaSuccess;
#define VQB_LP(LPF, J)                                                                                                            \
    do {                                                                                                                          \
        e = cudaFuncSetAttribute(latent_prep_bf16_kernel<LPF, J>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);       \
        if (e == cudaSuccess) latent_prep_bf16_kernel<LPF, J><<<(unsigned)grid, 256, smem, s>>>(z, D, W, N, N_pad, xb, band, meta, tile_ldg_mode()); \
    } while (0)
    VQB_DISPATCH_D(D, VQB_LP);
#undef VQB_LP
    if (e != cudaSuccess) return e;
    note_launch();
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ exact fp32 search
// Nearest code per frame with every distance evaluated in fp32 in the reference's association order.
// Block = 64 frames (resident in shared memory) x all K codes in tiles of 64, 4x4 register micro-tiles.
constexpr int XT_M = 64, XT_N = 64, XT_K = 16, XT_LD = XT_N + 4;

// torch.argmin order as one integer: NaN smallest, then by distance, then by index (atomicMin on it = the argmin)
__device__ __forceinline__ unsigned long long argmin_key(float d, int i) {
    const uint32_t b = __float_as_uint(d);
    const uint32_t hi = isnan(d) ? 0u : (b ^ ((b & 0x80000000u) ? 0xFFFFFFFFu : 0x80000000u));
    return ((unsigned long long)hi << 32) | (uint32_t)i;
}

// Fallback mode (rows != nullptr) splits the codebook over gridDim.y: a handful of frames must not be searched by a
// handful of blocks.  Slices meet in best64[list position] through atomicMin; fallback_commit_kernel publishes the codes.
__global__ void __launch_bounds__(256) exact_search_kernel(const float* __restrict__ z, const float* __restrict__ E,
                                                           const float* __restrict__ e2, int D, int64_t W, int64_t N, int K,
                                                           const int* __restrict__ rows, const int* __restrict__ row_count,
                                                           int* __restrict__ idx32, unsigned long long* __restrict__ best64) {
    extern __shared__ __align__(16) float smem_f[];
    float* Xs = smem_f;                                   // [D][64]   x tile, d-major
    float* Es = Xs + (size_t)D * XT_M;                    // [16][68]  codebook chunk, d-major
    float* x2s = Es + XT_K * XT_LD;                       // [64]
    int* rown = reinterpret_cast<int*>(x2s + XT_M);       // [64] frame id or -1
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int64_t total = rows ? (int64_t)*row_count : N;
    // many fallback frames (e.g. a NaN codebook sends all of them): re-reading every frame tile once per slice would
    // cost more than it saves, so slice 0 takes the whole codebook and the others retire
    const bool split = gridDim.y > 1 && total <= (int64_t)XT_M * gridDim.x * 4;
    if (!split && blockIdx.y > 0) return;
    const int tiles_k = (K + XT_N - 1) / XT_N;
    const int per_slice = split ? (tiles_k + (int)gridDim.y - 1) / (int)gridDim.y : tiles_k;
    const int k_lo = split ? (int)blockIdx.y * per_slice * XT_N : 0;
    const int k_hi = min(K, k_lo + per_slice * XT_N);

    for (int64_t tile = blockIdx.x; tile * XT_M < total; tile += gridDim.x) {
        __syncthreads();
        if (tid < XT_M) {
            const int64_t r = tile * XT_M + tid;
            rown[tid] = (r < total) ? (rows ? rows[r] : (int)r) : -1;
        }
        __syncthreads();
        {
            const int r = tid & 63;
            const int n = rown[r];
            int64_t b = 0, w = 0;
            if (n >= 0) { b = n / W; w = n - b * W; }
            const float* zp = z + (size_t)b * D * W + w;
            for (int d = tid >> 6; d < D; d += 4) Xs[d * XT_M + r] = (n >= 0) ? zp[(size_t)d * W] : 0.f;
        }
        __syncthreads();
        if (tid < XT_M) {   // |x|^2 = sum of rounded squares (vector_quantizer.py:32)
            float s = 0.f;
            for (int d = 0; d < D; ++d) { const float v = Xs[d * XT_M + tid]; s = __fadd_rn(s, __fmul_rn(v, v)); }
            x2s[tid] = s;
        }
        float bd[4];
        int bi[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { bd[i] = 0.f; bi[i] = -1; }

        for (int k0 = k_lo; k0 < k_hi; k0 += XT_N) {
            float acc[4][4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
            for (int d0 = 0; d0 < D; d0 += XT_K) {
                __syncthreads();
                {
                    const int dd = tid & 15;
                    for (int kk = tid >> 4; kk < XT_N; kk += 16) {
                        const int k = k0 + kk;
                        Es[dd * XT_LD + kk] = (k < K) ? E[(size_t)k * D + d0 + dd] : 0.f;
                    }
                }
                __syncthreads();
#pragma unroll
                for (int dd = 0; dd < XT_K; ++dd) {
                    const float4 a = *reinterpret_cast<const float4*>(&Xs[(d0 + dd) * XT_M + ty * 4]);
                    const float4 c = *reinterpret_cast<const float4*>(&Es[dd * XT_LD + tx * 4]);
                    const float av[4] = {a.x, a.y, a.z, a.w}, cv[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], cv[j], acc[i][j]);
                }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int k = k0 + tx * 4 + j;
                if (k < K) {
                    const float ek = e2[k];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float dist = ref_distance(x2s[ty * 4 + i], ek, acc[i][j]);
                        if (better(dist, k, bd[i], bi[i])) { bd[i] = dist; bi[i] = k; }
                    }
                }
            }
        }
        // combine the 16 threads (tx) that share a frame: lanes differ in their low 4 bits
#pragma unroll
        for (int i = 0; i < 4; ++i) {
#pragma unroll
            for (int o = 1; o < 16; o <<= 1) {
                const float od = __shfl_xor_sync(0xffffffffu, bd[i], o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi[i], o);
                if (oi >= 0 && better(od, oi, bd[i], bi[i])) { bd[i] = od; bi[i] = oi; }
            }
            if (tx == 0) {
                const int n = rown[ty * 4 + i];
                if (n >= 0) {
                    if (rows) { if (bi[i] >= 0) atomicMin(best64 + tile * XT_M + ty * 4 + i, argmin_key(bd[i], bi[i])); }
                    else idx32[n] = bi[i];
                }
            }
        }
    }
}

__global__ void __launch_bounds__(256) fallback_commit_kernel(const int* __restrict__ rows, const int* __restrict__ row_count,
                                                              const unsigned long long* __restrict__ best64,
                                                              uint8_t* __restrict__ cand_cnt, uint16_t* __restrict__ cand_idx) {
    const int total = *row_count;
    for (int r = blockIdx.x * 256 + threadIdx.x; r < total; r += gridDim.x * 256) {
        const int n = rows[r];
        cand_cnt[n] = kCandFinal;
        cand_idx[(size_t)n * kCandMax] = (uint16_t)(best64[r] & 0xFFFFu);
    }
}

cudaError_t launch_exact_search(const float* z, const float* codebook, const float* e2, int B, int D, int64_t W, int K,
                                const int* rows, const int* row_count, int* idx32, uint8_t* cand_cnt, uint16_t* cand_idx,
                                unsigned long long* best64, cudaStream_t s) {
    const int64_t N = (int64_t)B * W;
    const size_t smem = ((size_t)D * XT_M + XT_K * XT_LD + XT_M) * 4 + XT_M * 4;
    static bool attr_done_dev[64] = {};             // function attributes are per device
    int dev = 0;
    cudaGetDevice(&dev);
    if (!attr_done_dev[dev & 63]) {
        cudaError_t e = cudaFuncSetAttribute(exact_search_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) return e;
        attr_done_dev[dev & 63] = true;
    }
    const int64_t tiles = (N + XT_M - 1) / XT_M;
    const int per_sm = (int)(220 * 1024 / (smem + 1024)) > 0 ? (int)(220 * 1024 / (smem + 1024)) : 1;
    int64_t grid = 148LL * (per_sm > 4 ? 4 : per_sm);
    if (grid > tiles) grid = tiles;
    if (grid < 1) grid = 1;
    if (!rows) {
        exact_search_kernel<<<(unsigned)grid, 256, smem, s>>>(z, codebook, e2, D, W, N, K, nullptr, nullptr, idx32, nullptr);
        note_launch();
        return cudaGetLastError();
    }
    // one 64-code tile per slice up to K = 8192: the list is a handful of frames, so the kernel's duration is the serial chain of
    // one block (codebook chunk load -> barrier -> 16 FMA steps, per 16 dims) - 4 tiles per slice were 0.2 ms of every step
    int ksplit = (K + XT_N - 1) / XT_N;
    if (ksplit > 128) ksplit = 128;
    if (grid > 148) grid = 148;
    exact_search_kernel<<<dim3((unsigned)grid, (unsigned)ksplit), 256, smem, s>>>(z, codebook, e2, D, W, N, K, rows, row_count, nullptr,
                                                                                 best64);
    fallback_commit_kernel<<<148, 256, 0, s>>>(rows, row_count, best64, cand_cnt, cand_idx);
    note_launch(2);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ fallback tail
// Fused-tail mode (the tensor-core kernel finishes every frame whose shortlist held): the few frames that went to the
// exact search are finished here, one warp per frame of the list - codeword gather, straight-through value, SSE,
// histogram, residual sums, index.  Strided (uncoalesced) latent accesses are fine for a list this short.
__global__ void __launch_bounds__(256) fallback_tail_kernel(const float* __restrict__ z, const float* __restrict__ E, int D, int64_t W,
                                                            const int* __restrict__ rows, const int* __restrict__ row_count,
                                                            const unsigned long long* __restrict__ best64,
                                                            int64_t* __restrict__ idx_out, float* __restrict__ q_out,
                                                            int* __restrict__ counts, float* __restrict__ resid,
                                                            double* __restrict__ sse_partials) {
    __shared__ double red[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int total = *row_count;
    double sse = 0.0;
    for (int r = blockIdx.x * 8 + warp; r < total; r += gridDim.x * 8) {
        const int n = rows[r];
        const int k = (int)(uint32_t)(best64[r] & 0xFFFFFFFFull);     // argmin_key: the low word is the code
        const int64_t b = n / W, w = n - b * W;
        const float* xp = z + (size_t)b * D * W + w;
        float fs = 0.f;
        for (int d = lane; d < D; d += 32) {
            const float x = xp[(size_t)d * W];
            const float df = __fsub_rn(E[(size_t)k * D + d], x);
            fs = fmaf(df, df, fs);
            if (q_out) q_out[(size_t)b * D * W + (size_t)d * W + w] = __fadd_rn(x, df);   // straight-through VALUE (:48)
            if (resid) atomicAdd(resid + (size_t)k * D + d, -df);
        }
        sse += (double)warp_sum(fs);
        if (lane == 0) {
            atomicAdd(counts + k, 1);
            idx_out[n] = (int64_t)k;
        }
    }
    if (lane == 0) red[warp] = sse;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int i = 0; i < 8; ++i) s += red[i];
        sse_partials[blockIdx.x] = s;
    }
}

cudaError_t launch_fallback_tail(const float* z, const float* codebook, int B, int D, int64_t W, int K, const int* rows,
                                 const int* row_count, const unsigned long long* best64, int64_t* idx_out, float* q_out, int* counts,
                                 float* resid, double* sse_partials, cudaStream_t s) {
    (void)B; (void)K;
    fallback_tail_kernel<<<kFallbackTailGrid, 256, 0, s>>>(z, codebook, D, W, rows, row_count, best64, idx_out, q_out, counts, resid,
                                                           sse_partials);
    note_launch();
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ fused tail
// Per frame: (a) settle the index - given (exact search), unique shortlist entry, or by rescoring the shortlisted codes
// in fp32 in the reference's op order, ties -> lowest index; (b) gather the codeword (vector_quantizer.py:42 as a
// gather), (c) SSE for both MSE losses (:45-46), (d) straight-through value fl(x + fl(q - x)) written back in BCW
// (:48,:52), (e) code histogram (:49) and, for training, the per-code residual sums that give the codebook gradient.
// Latents are read exactly once.
//
// Tile = 32 frames x D in shared memory, frame-major with rows of D+4 floats: the transposing global<->shared phases
// move one float4 (4 consecutive dims of one frame) per thread, the per-frame phase reads float4 rows - both are
// bank-conflict free - and the codebook rows, residual atomics (red.v4) and shared accesses are all 16 bytes wide.
// LPF lanes cooperate on one frame (32 / LPF frames per warp at a time), each lane owning 4*J dims.
// Residual-sum replicas.  Every frame adds its D residuals to the row of its code with atomics; the L2 serialises atomics
// per address, so the most popular code bounds the whole pass (measured: 8 of 20 ms at BASELINE config 3 with a code that
// takes ~1.5 % of the frames).  The blocks of different SMs therefore add into one of n_rep copies (copy 0 is the caller's
// buffer, the others live in the workspace) and fold_resid_kernel sums the copies afterwards.
__global__ void __launch_bounds__(256) fold_resid_kernel(float* __restrict__ resid, const float* __restrict__ resid_rep, int n_rep,
                                                         size_t rep_stride) {
    // the caller's buffer (stats + K) is only 4-byte aligned when K % 4 != 0: plain scalar accesses, the arrays are small
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < rep_stride; i += (size_t)gridDim.x * 256) {
        float a = resid[i];
        for (int r = 0; r + 1 < n_rep; ++r) a += resid_rep[(size_t)r * rep_stride + i];
        resid[i] = a;
    }
}

// One frame of a frame-major shared tile (LPF lanes cooperate, lane `sl` of the group owns dims 4*sl + 4*LPF*j): settle the
// index (rescoring in fp32 in the reference's op order when more than one code is shortlisted), gather the codeword, write the
// straight-through value back into the tile, accumulate SSE / histogram / residual sums, publish the index.
// `given`: cl_lo.x already is the final code (exact search).  All 32 lanes of the warp must call this together.
struct TailAcc { float sse, sse_c; unsigned int n_resc, n_short; };

// kExact: D == 4 * LPF * J (D = 32, 64, 128, 256 ...): no dimension guards, no zero fill, constant strides.
template <int LPF, int J, bool kResid, bool kExact = false>
__device__ __forceinline__ void tail_frame(float* Xs, int ld, int f, int64_t n, bool live, int cnt_in, uint4 cl_lo, uint4 cl_hi, bool given,
                                           const float* __restrict__ E, const float* __restrict__ e2, int D,
                                           int64_t* __restrict__ idx_out, int* __restrict__ counts, float* __restrict__ resid,
                                           bool resid_v4, int sl, TailAcc& acc) {
    float& sse = acc.sse;
    float& sse_c = acc.sse_c;
    unsigned int& n_resc = acc.n_resc;
    unsigned int& n_short = acc.n_short;
    float4 xv[J];
#pragma unroll
    for (int j = 0; j < J; ++j) {
        const int d = 4 * sl + 4 * LPF * j;
        xv[j] = (kExact || d < D) ? *reinterpret_cast<const float4*>(Xs + f * ld + d) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const int cnt = cnt_in;
    int k = given ? (int)cl_lo.x : (int)(cl_lo.x & 0xFFFFu);
    const bool need = live && cnt != kCandFinal && cnt > 1;
    if (__any_sync(0xffffffffu, need)) {
        // fp32 rescoring of the shortlisted codes in the reference's op order (whole warp takes part in shuffles)
        float x2 = 0.f;
#pragma unroll
        for (int j = 0; j < J; ++j) {
            x2 = __fadd_rn(x2, __fmul_rn(xv[j].x, xv[j].x));
            x2 = __fadd_rn(x2, __fmul_rn(xv[j].y, xv[j].y));
            x2 = __fadd_rn(x2, __fmul_rn(xv[j].z, xv[j].z));
            x2 = __fadd_rn(x2, __fmul_rn(xv[j].w, xv[j].w));
        }
        x2 = group_sum<LPF>(x2);
        int cmax = need ? cnt : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) cmax = max(cmax, __shfl_xor_sync(0xffffffffu, cmax, o));
        // shortlist entry ci out of the prefetched row (registers only: no dependent global load per round)
        const uint32_t w8[8] = {cl_lo.x, cl_lo.y, cl_lo.z, cl_lo.w, cl_hi.x, cl_hi.y, cl_hi.z, cl_hi.w};
        auto entry = [&](int ci) -> int {
            const int wi = ci >> 1;
            const uint32_t a01 = (wi & 1) ? w8[1] : w8[0], a23 = (wi & 1) ? w8[3] : w8[2];
            const uint32_t a45 = (wi & 1) ? w8[5] : w8[4], a67 = (wi & 1) ? w8[7] : w8[6];
            const uint32_t lo = (wi & 2) ? a23 : a01, hi = (wi & 2) ? a67 : a45;
            const uint32_t wv = (wi & 4) ? hi : lo;
            return (int)((ci & 1) ? (wv >> 16) : (wv & 0xFFFFu));
        };
        float bd = 0.f;
        int bk = -1;
        // rounds are software-pipelined: the codeword of round ci+1 is in flight while round ci is reduced
        float4 ev[J], en[J];
        int kc = (need && 0 < cnt) ? entry(0) : 0;
        float e2c = e2[kc], e2n = 0.f;             // |e|^2 travels with its codeword: requested a round ahead as well
#pragma unroll
        for (int j = 0; j < J; ++j) {
            const int d = 4 * sl + 4 * LPF * j;
            ev[j] = (kExact || d < D) ? *reinterpret_cast<const float4*>(E + (size_t)kc * D + d) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        for (int ci = 0; ci < cmax; ++ci) {
            const bool act = need && ci < cnt;
            const int kn = (need && ci + 1 < cnt) ? entry(ci + 1) : 0;
            if (ci + 1 < cmax) {
                e2n = e2[kn];
#pragma unroll
                for (int j = 0; j < J; ++j) {
                    const int d = 4 * sl + 4 * LPF * j;
                    en[j] = (kExact || d < D) ? *reinterpret_cast<const float4*>(E + (size_t)kn * D + d) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
            float dot = 0.f;
#pragma unroll
            for (int j = 0; j < J; ++j) {
                dot = fmaf(xv[j].x, ev[j].x, dot);
                dot = fmaf(xv[j].y, ev[j].y, dot);
                dot = fmaf(xv[j].z, ev[j].z, dot);
                dot = fmaf(xv[j].w, ev[j].w, dot);
            }
            dot = group_sum<LPF>(dot);
            if (act) {
                const float dist = ref_distance(x2, e2c, dot);
                if (better(dist, kc, bd, bk)) { bd = dist; bk = kc; }
            }
            kc = kn;
            e2c = e2n;
#pragma unroll
            for (int j = 0; j < J; ++j) ev[j] = en[j];
        }
        if (need) {
            k = bk;
            if (sl == 0) { n_resc += 1; n_short += cnt; }
        }
    }
    if (live && !need && sl == 0) n_short += 1;
    if (live) {
        const float* er = E + (size_t)k * D;
        float4 qv[J];
#pragma unroll
        for (int j = 0; j < J; ++j) {
            const int d = 4 * sl + 4 * LPF * j;
            qv[j] = (kExact || d < D) ? *reinterpret_cast<const float4*>(er + d) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        float fs = 0.f;
#pragma unroll
        for (int j = 0; j < J; ++j) {
            const int d = 4 * sl + 4 * LPF * j;
            if (kExact || d < D) {
                // r = fl(x - e) is exactly -fl(e - x): the residual sums want r, and the straight-through VALUE
                // fl(x + fl(e - x)) (:48) equals fl(x - r) bit for bit - no negations
                float4 r, st;
                r.x = __fsub_rn(xv[j].x, qv[j].x); r.y = __fsub_rn(xv[j].y, qv[j].y);
                r.z = __fsub_rn(xv[j].z, qv[j].z); r.w = __fsub_rn(xv[j].w, qv[j].w);
                fs = fmaf(r.x, r.x, fs); fs = fmaf(r.y, r.y, fs); fs = fmaf(r.z, r.z, fs); fs = fmaf(r.w, r.w, fs);
                st.x = __fsub_rn(xv[j].x, r.x); st.y = __fsub_rn(xv[j].y, r.y);
                st.z = __fsub_rn(xv[j].z, r.z); st.w = __fsub_rn(xv[j].w, r.w);
                *reinterpret_cast<float4*>(Xs + f * ld + d) = st;
                if (kResid) {
                    float* rp = resid + (size_t)k * D + d;
                    if (resid_v4) red_add_v4(rp, r.x, r.y, r.z, r.w);
                    else { atomicAdd(rp, r.x); atomicAdd(rp + 1, r.y); atomicAdd(rp + 2, r.z); atomicAdd(rp + 3, r.w); }
                }
            }
        }
        {   // Kahan: sse += fs
            const float y = fs - sse_c, t = sse + y;
            sse_c = (t - sse) - y;
            sse = t;
        }
        if (sl == 0) {
            atomicAdd(counts + k, 1);
            idx_out[n] = (int64_t)k;
        }
    }
}

constexpr int TAIL_WARPS = 4;   // 128-thread blocks: more independent blocks per SM hide the load / barrier / gather latencies

template <int LPF, int J, bool kResid>
__global__ void __launch_bounds__(32 * TAIL_WARPS, (J >= 6) ? 4 : 6) tail_kernel(const float* __restrict__ z, const float* __restrict__ E,
                                                      const float* __restrict__ e2, int D, int64_t W, int64_t N, int K,
                                                      const int* __restrict__ idx32, const uint8_t* __restrict__ cand_cnt,
                                                      const uint16_t* __restrict__ cand_idx, int64_t* __restrict__ idx_out,
                                                      float* __restrict__ q_out, int* __restrict__ counts,
                                                      float* __restrict__ resid, double* __restrict__ sse_partials,
                                                      WsMeta* meta, float* resid_rep, int n_rep, size_t rep_stride, int ldg) {
    resid = pick_resid_replica(resid, resid_rep, n_rep, rep_stride);
    extern __shared__ __align__(16) float Xbuf[];   // 2 x [32][D + 4]: the next tile streams in while this one is worked on
    __shared__ double red[TAIL_WARPS];
    constexpr int FPW = 32 / LPF;                 // frames a warp works on at once
    constexpr int ITER = (TL_F / TAIL_WARPS) / FPW;   // rounds per warp and tile
    static_assert(ITER >= 1, "a warp must own at least FPW frames of the tile");
    const int ld = D + 4;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = lane / LPF, sl = lane % LPF;
    TailAcc acc{0.f, 0.f, 0u, 0u};                // Kahan-compensated per-thread SSE, diagnostics
    const bool resid_v4 = kResid && (reinterpret_cast<uintptr_t>(resid) & 15) == 0;   // stats + K is 16B aligned iff K % 4 == 0

    auto tile_col = [&](int64_t tile, bool& valid) -> size_t {
        const int64_t nl = tile * TL_F + lane;
        valid = nl < N;
        int64_t b = 0, w = 0;
        if (valid) { b = nl / W; w = nl - b * W; }
        return (size_t)b * D * W + w;
    };
    int buf = 0;
    {
        bool v0;
        const size_t c0 = tile_col(blockIdx.x, v0);
        if (!ldg && (int64_t)blockIdx.x * TL_F < N) tile_fetch_async<TAIL_WARPS>(Xbuf, ld, z, c0, W, D, v0);
    }
    for (int64_t tile = blockIdx.x; tile * TL_F < N; tile += gridDim.x, buf ^= 1) {
        float* Xs = Xbuf + (size_t)(ldg ? 0 : buf) * TL_F * ld;   // register-staged loads need one buffer only
        bool valid;
        const size_t col = tile_col(tile, valid);
        // shortlist headers of this warp's frames, requested early so their latency hides behind the tile fetch
        int cnt_r[ITER];
        uint4 cl_lo[ITER], cl_hi[ITER];           // the frame's whole shortlist row (16 x uint16), or the final code in .x
#pragma unroll
        for (int it = 0; it < ITER; ++it) {
            const int64_t n = tile * TL_F + warp * (TL_F / TAIL_WARPS) + it * FPW + sub;
            cnt_r[it] = 0;
            cl_lo[it] = make_uint4(0u, 0u, 0u, 0u);
            cl_hi[it] = make_uint4(0u, 0u, 0u, 0u);
            if (n < N) {
                if (idx32) { cnt_r[it] = kCandFinal; cl_lo[it].x = (uint32_t)idx32[n]; }
                else {
                    cnt_r[it] = cand_cnt[n];
                    const uint4* row = reinterpret_cast<const uint4*>(cand_idx + (size_t)n * kCandMax);
                    cl_lo[it] = row[0];
                    cl_hi[it] = row[1];
                }
            }
        }
        __syncthreads();                          // the other buffer's previous tile has been stored: it may be refilled
        const int64_t next = tile + gridDim.x;
        if (ldg) {
            tile_load<TAIL_WARPS>(Xs, ld, z, col, W, D, valid);
        } else if (next * TL_F < N) {
            bool vn;
            const size_t cn = tile_col(next, vn);
            tile_fetch_async<TAIL_WARPS>(Xbuf + (size_t)(buf ^ 1) * TL_F * ld, ld, z, cn, W, D, vn);
            asm volatile("cp.async.wait_group 1;" ::: "memory");   // this tile has landed; the next one stays in flight
        } else {
            tile_fetch_wait();
        }
        __syncthreads();
#pragma unroll
        for (int it = 0; it < ITER; ++it) {
            const int f = warp * (TL_F / TAIL_WARPS) + it * FPW + sub;
            const int64_t n = tile * TL_F + f;
            tail_frame<LPF, J, kResid>(Xs, ld, f, n, n < N, cnt_r[it], cl_lo[it], cl_hi[it], idx32 != nullptr, E, e2, D, idx_out, counts,
                                       resid, resid_v4, sl, acc);
        }
        if (q_out) {
            __syncthreads();
            tile_store<TAIL_WARPS>(Xs, ld, q_out, col, W, D, valid);
        }
    }
    double t = (double)acc.sse - (double)acc.sse_c;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (lane == 0) red[warp] = t;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int i = 0; i < TAIL_WARPS; ++i) s += red[i];
        sse_partials[blockIdx.x] = s;
    }
    if (acc.n_resc | acc.n_short) {
        atomicAdd(&meta->rescored, (unsigned long long)acc.n_resc);
        atomicAdd(&meta->shortlisted, (unsigned long long)acc.n_short);
    }
}

// ------------------------------------------------------------------------------------------------ fused tail, TMA-fed
// Same per-frame work as tail_kernel, but the latents arrive by TMA: persistent blocks walk tiles of 32 frames of one batch
// item; a 3-D TMA box [D dims][32 frames] of the NEXT tile is in flight (per block) while the current one is transposed from
// the box into the frame-major tile, worked on and written back.  No thread ever waits on a global latent load, which is what
// bounded tail_kernel (latency, 25 % occupancy).  Needs W % 4 == 0 (TMA global strides are multiples of 16 bytes).
// NW warps per block, NB box buffers (1: the next box is requested right after the transposition freed the buffer).
template <int LPF, int J, bool kResid, int NW, int NB, bool kExact>
__global__ void __launch_bounds__(32 * NW, (NW == 4) ? ((J <= 2) ? 6 : ((J <= 4) ? 4 : 3)) : ((J >= 6) ? 2 : 3))
tail_tma_kernel(const __grid_constant__ CUtensorMap tmap_z, const float* __restrict__ E, const float* __restrict__ e2, int D_arg, int64_t W,
                int tiles_per_item, int64_t num_tiles, int box_dims, const int* __restrict__ idx32,
                const uint8_t* __restrict__ cand_cnt, const uint16_t* __restrict__ cand_idx, int64_t* __restrict__ idx_out,
                float* __restrict__ q_out, int* __restrict__ counts, float* __restrict__ resid,
                double* __restrict__ sse_partials, WsMeta* meta, float* resid_rep, int n_rep, size_t rep_stride, int l2_once) {
    resid = pick_resid_replica(resid, resid_rep, n_rep, rep_stride);
    using namespace ptx;
    // kExact (D == 4 * LPF * J, the usual 64 / 128 / 256): D is a compile-time constant - row strides and codebook offsets fold
    // into immediates and the per-lane dimension guards disappear (ncu: half of the per-frame instructions were such overhead)
    const int D = kExact ? 4 * LPF * J : D_arg;
    constexpr int TT_WARPS = NW;
    extern __shared__ __align__(128) float tt_smem[];   // NB x [D][32] TMA boxes, then the frame-major tile [32][D + 4]
    __shared__ double red[TT_WARPS];
    __shared__ __align__(8) unsigned long long full[2];
    constexpr int FPW = 32 / LPF;                       // frames a warp works on at once
    constexpr int ITER = (TL_F / TT_WARPS) / FPW;       // rounds per warp and tile
    static_assert(ITER >= 1, "a warp must own at least FPW frames of the tile");
    const int ld = D + 4;
    float* box = tt_smem;
    float* Xs = tt_smem + NB * (size_t)D * TL_F;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = lane / LPF, sl = lane % LPF;
    TailAcc acc{0.f, 0.f, 0u, 0u};
    const bool resid_v4 = kResid && (reinterpret_cast<uintptr_t>(resid) & 15) == 0;
    const uint32_t box_bytes = (uint32_t)D * TL_F * 4;

    if (threadIdx.x == 0) {
        mbar_init(smem_u32(&full[0]), 1);
        mbar_init(smem_u32(&full[1]), 1);
        fence_barrier_init();
        prefetch_tmap(&tmap_z);
    }
    __syncthreads();
    auto issue = [&](int64_t tile, int buf) {           // one thread: the whole [D][32] box of `tile` (frames past W arrive as zeros)
        const int b = (int)(tile / tiles_per_item), w0 = (int)(tile - (int64_t)b * tiles_per_item) * TL_F;
        const uint32_t bar = smem_u32(&full[buf]);
        mbar_expect_tx(bar, box_bytes);
        for (int d0 = 0; d0 < D; d0 += box_dims) {
            if (l2_once) tma_load_3d_once(smem_u32(box + ((size_t)buf * D + d0) * TL_F), &tmap_z, bar, w0, d0, b);
            else tma_load_3d(smem_u32(box + ((size_t)buf * D + d0) * TL_F), &tmap_z, bar, w0, d0, b);
        }
    };
    int64_t tile = blockIdx.x;
    if (threadIdx.x == 0 && tile < num_tiles) issue(tile, 0);
    for (int it = 0; tile < num_tiles; tile += gridDim.x, ++it) {
        const int buf = NB == 2 ? (it & 1) : 0;
        const int b = (int)(tile / tiles_per_item), w0 = (int)(tile - (int64_t)b * tiles_per_item) * TL_F;
        const int64_t n0 = (int64_t)b * W + w0;           // global frame id of the tile's first frame
        const int wlim = (int)((W - w0) < TL_F ? (W - w0) : TL_F);   // frames of this tile that exist
        // shortlist headers of this warp's frames, requested early so their latency hides behind the box wait
        int cnt_r[ITER];
        uint4 cl_lo[ITER], cl_hi[ITER];
#pragma unroll
        for (int i = 0; i < ITER; ++i) {
            const int f = warp * (TL_F / TT_WARPS) + i * FPW + sub;
            cnt_r[i] = 0;
            cl_lo[i] = make_uint4(0u, 0u, 0u, 0u);
            cl_hi[i] = make_uint4(0u, 0u, 0u, 0u);
            if (f < wlim) {
                const int64_t n = n0 + f;
                if (idx32) { cnt_r[i] = kCandFinal; cl_lo[i].x = (uint32_t)idx32[n]; }
                else {
                    cnt_r[i] = cand_cnt[n];
                    const uint4* row = reinterpret_cast<const uint4*>(cand_idx + (size_t)n * kCandMax);
                    cl_lo[i] = row[0];
                    cl_hi[i] = row[1];
                }
            }
        }
        // the other box was transposed in the previous iteration (block-wide barriers since): refill it with the next tile
        const int64_t next = tile + gridDim.x;
        if (NB == 2 && threadIdx.x == 0 && next < num_tiles) issue(next, buf ^ 1);
        mbar_wait(smem_u32(&full[buf]), NB == 2 ? ((uint32_t)(it >> 1) & 1u) : ((uint32_t)it & 1u));
        {   // box [d][32 frames] -> frame-major tile: 4 conflict-free LDS.32 + one conflict-free STS.128 per 4 dims
            const float* bx = box + (size_t)buf * D * TL_F + lane;
            for (int d0 = warp * 4; d0 < D; d0 += 4 * TT_WARPS) {
                float4 v;
                v.x = bx[(d0 + 0) * TL_F];
                v.y = bx[(d0 + 1) * TL_F];
                v.z = bx[(d0 + 2) * TL_F];
                v.w = bx[(d0 + 3) * TL_F];
                *reinterpret_cast<float4*>(Xs + lane * ld + d0) = v;
            }
        }
        __syncthreads();
        if (NB == 1 && threadIdx.x == 0 && next < num_tiles) issue(next, 0);   // the box has been consumed: refill it behind the compute
#pragma unroll
        for (int i = 0; i < ITER; ++i) {
            const int f = warp * (TL_F / TT_WARPS) + i * FPW + sub;
            tail_frame<LPF, J, kResid, kExact>(Xs, ld, f, n0 + f, f < wlim, cnt_r[i], cl_lo[i], cl_hi[i], idx32 != nullptr, E, e2, D, idx_out, counts,
                                       resid, resid_v4, sl, acc);
        }
        if (q_out) {
            __syncthreads();
            if (lane < wlim) {
                // four running row pointers (one 64-bit add each per step) instead of a 64-bit multiply per store
                const size_t step = (size_t)(4 * TT_WARPS) * W;
                float* p0 = q_out + ((size_t)b * D + warp * 4) * W + w0 + lane;
                float* p1 = p0 + W;
                float* p2 = p1 + W;
                float* p3 = p2 + W;
                const float* xs = Xs + lane * ld + warp * 4;
#pragma unroll 4
                for (int d0 = warp * 4; d0 < D; d0 += 4 * TT_WARPS) {
                    const float4 v = *reinterpret_cast<const float4*>(xs);
                    st_stream(p0, v.x);
                    st_stream(p1, v.y);
                    st_stream(p2, v.z);
                    st_stream(p3, v.w);
                    p0 += step; p1 += step; p2 += step; p3 += step;
                    xs += 4 * TT_WARPS;
                }
            }
        }
        __syncthreads();                                  // the tile may be overwritten, the box refilled
    }
    double t = (double)acc.sse - (double)acc.sse_c;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (lane == 0) red[warp] = t;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int i = 0; i < TT_WARPS; ++i) s += red[i];
        sse_partials[blockIdx.x] = s;
    }
    if (acc.n_resc | acc.n_short) {
        atomicAdd(&meta->rescored, (unsigned long long)acc.n_resc);
        atomicAdd(&meta->shortlisted, (unsigned long long)acc.n_short);
    }
}

#define VQB_DISPATCH_D8(D, CALL)                   \
    do {                                           \
        if ((D) <= 32) { CALL(8, 1); }             \
        else if ((D) <= 64) { CALL(8, 2); }        \
        else if ((D) <= 96) { CALL(8, 3); }        \
        else if ((D) <= 128) { CALL(8, 4); }       \
        else if ((D) <= 192) { CALL(8, 6); }       \
        else if ((D) <= 256) { CALL(8, 8); }       \
        else if ((D) <= 384) { CALL(16, 6); }      \
        else { CALL(16, 8); }                      \
    } while (0)

template <int LPF, int J>
static cudaError_t launch_tail_t(const float* z, const float* codebook, const float* e2, int D, int64_t W, int64_t N, int K,
                                 const int* idx32, const uint8_t* cand_cnt, const uint16_t* cand_idx, int64_t* idx_out, float* q_out,
                                 int* counts, float* resid, double* part, int grid, WsMeta* meta, float* resid_rep, int n_rep,
                                 size_t rep_stride, cudaStream_t s) {
    const size_t smem = (size_t)(tile_ldg_mode() ? 1 : 2) * TL_F * (D + 4) * 4;   // the second buffer only serves the cp.async pipeline
    cudaError_t e;
    if (resid) {
        if ((e = cudaFuncSetAttribute(tail_kernel<LPF, J, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024)) != cudaSuccess) return e;
        tail_kernel<LPF, J, true><<<grid, 32 * TAIL_WARPS, smem, s>>>(z, codebook, e2, D, W, N, K, idx32, cand_cnt, cand_idx, idx_out, q_out, counts,
                                                          resid, part, meta, resid_rep, n_rep, rep_stride, tile_ldg_mode());
    } else {
        if ((e = cudaFuncSetAttribute(tail_kernel<LPF, J, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024)) != cudaSuccess) return e;
        tail_kernel<LPF, J, false><<<grid, 32 * TAIL_WARPS, smem, s>>>(z, codebook, e2, D, W, N, K, idx32, cand_cnt, cand_idx, idx_out, q_out, counts,
                                                           nullptr, part, meta, nullptr, 1, 0, tile_ldg_mode());
    }
    return cudaGetLastError();
}

// Copies of the residual sums: as many as keep the copies within 16 MiB (they should stay L2-resident next to the codebook),
// between 2 and kResidReplicasMax.  Measured at BASELINE config 3 (8 MiB per copy): 21.0 ms with one copy, 13.7 ms with two, no
// further gain from four or eight; small codebooks (the reference's K = 512) have hotter codes and get all eight.
int resid_replicas(int K, int D) {
    if (const int v = env_get(ENV_RESID_REPLICAS, 0); v >= 1 && v <= kResidReplicasMax) return v;
    const size_t per_copy = (size_t)K * D * 4;
    size_t n = (16u << 20) / (per_copy ? per_copy : 1);
    if (n < 2) n = 2;
    if (n > (size_t)kResidReplicasMax) n = kResidReplicasMax;
    return (int)n;
}
// VQB_TAIL_VARIANT (experiments): 1 (default) = 4 warps and one box per block (3 blocks per SM at D = 256: 13.7 ms in the
// BASELINE config 3 step), 0 = 8 warps and two boxes (2 blocks per SM: 14.6 ms).  A third form without block barriers (every
// warp transposing its own 8 frames out of a 128-byte-swizzled box, rescoring pairs dealt out over the lane groups) measured
// 14.7 ms and was dropped: after the residual replicas the pass is bound by its L2 / DRAM traffic, not by the barriers.
// Latents far larger than L2 are streamed with an evict-first policy (tma_load_3d_once); small batches keep the default, so
// that the tail still finds in L2 what the search kernel read (and the backward pass what the tail read).  VQB_L2_ONCE=0/1
// overrides (experiments).
bool latents_read_once(size_t latent_bytes) {
    if (const int v = env_get(ENV_L2_ONCE, -1); v >= 0) return v != 0;
    return latent_bytes > ((size_t)96 << 20);
}
static int tail_tma_variant() {
    return env_get(ENV_TAIL_VARIANT, 1);
}
static bool tail_tma_enabled() {   // VQB_TAIL_TMA=0 keeps the register-staged tail_kernel (experiments)
    return env_get(ENV_TAIL_TMA, 1) != 0;
}

template <int LPF, int J, int NW, int NB>
static cudaError_t launch_tail_tma_t(const CUtensorMap& map, const float* codebook, const float* e2, int D, int64_t W, int tiles_per_item,
                                     int64_t num_tiles, int box_dims, const int* idx32, const uint8_t* cand_cnt, const uint16_t* cand_idx,
                                     int64_t* idx_out, float* q_out, int* counts, float* resid, double* part, int n_partials, WsMeta* meta,
                                     float* resid_rep, int n_rep, size_t rep_stride, cudaStream_t s) {
    const size_t smem = (size_t)(NB * D * TL_F + TL_F * (D + 4)) * 4;
    auto go = [&](auto kernel) -> cudaError_t {
        // attribute + occupancy are looked up once per kernel instance and shared-memory size (this path is launch-bound for small
        // batches).  All instances share one signature, hence ONE instantiation of this lambda: the kernel pointer is part of the key.
        static thread_local size_t cached_smem = ~(size_t)0;
        static thread_local const void* cached_kernel = nullptr;
        static thread_local int cached_blocks = 0, cached_dev = -1;
        cudaError_t e = cudaSuccess;
        int dev = 0;
        cudaGetDevice(&dev);
        if (cached_smem != smem || cached_kernel != reinterpret_cast<const void*>(kernel) || cached_dev != dev) {
            if ((e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)) != cudaSuccess) return e;
            int per_sm = 1, sms = 148;
            if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, 32 * NW, smem)) != cudaSuccess) return e;
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
            cached_blocks = sms * (per_sm < 1 ? 1 : per_sm);
            cached_smem = smem;
            cached_kernel = reinterpret_cast<const void*>(kernel);
            cached_dev = dev;
        }
        int64_t grid = cached_blocks;
        if (grid > n_partials) grid = n_partials;
        if (grid > num_tiles) grid = num_tiles;
        if (grid < 1) grid = 1;
        kernel<<<(unsigned)grid, 32 * NW, smem, s>>>(map, codebook, e2, D, W, tiles_per_item, num_tiles, box_dims, idx32, cand_cnt, cand_idx,
                                                           idx_out, q_out, counts, resid, part, meta, resid_rep, n_rep, rep_stride,
                                                           latents_read_once((size_t)num_tiles * TL_F * D * 4) ? 1 : 0);
        return cudaGetLastError();
    };
    const bool exact_ok = env_get(ENV_TAIL_EXACT, 1) != 0;   // experiments
    // The default single-box form gets the constant-D specialisation where it measured faster: D = 192 (-2.7 %) and D = 256
    // (-3.3 %, 12.9 -> 12.5 ms at BASELINE config 3).  For D <= 128 the generic form is faster (D = 128: +6 %, D = 64: +24 %,
    // D = 32: +21 % with the constant-D code), so J <= 4 keeps it.
    if (D == 4 * LPF * J && NB == 1 && J >= 6 && exact_ok)
        return resid ? go(tail_tma_kernel<LPF, J, true, NW, NB, NB == 1>) : go(tail_tma_kernel<LPF, J, false, NW, NB, NB == 1>);
    return resid ? go(tail_tma_kernel<LPF, J, true, NW, NB, false>) : go(tail_tma_kernel<LPF, J, false, NW, NB, false>);
}

cudaError_t launch_tail(const float* z, const float* codebook, const float* e2, int B, int D, int64_t W, int K,
                        const int* idx32, const uint8_t* cand_cnt, const uint16_t* cand_idx, int64_t* idx_out, float* q_out,
                        int* counts, float* resid, float* sse_partials, int n_partials, WsMeta* meta, float* resid_rep, cudaStream_t s,
                        const float* ep) {
    // the caller (forward_impl) zeroed the SSE partials and resid_replicas(K, D) residual-sum replicas with the head of the workspace
    const int64_t N = (int64_t)B * W;
    cudaError_t e = cudaSuccess;
    double* part = reinterpret_cast<double*>(sse_partials);
    const size_t rep_stride = (size_t)K * D;
    int n_rep = (resid && resid_rep) ? resid_replicas(K, D) : 1;
    // 3 (default): tail3_kernel where it applies and pays; 30 / 300: tail3_kernel wherever it applies (300: red.v4 residual sums at every D);
    // 2 / 216 / 232: tail2_kernel; 0: round-1 kernels
    const int form = env_get(ENV_TAIL_FORM, 3);
    // tail3_kernel's blocks run a load -> compute -> store pipeline over many tiles; with fewer than a tile or two per block (small
    // batches: BASELINE config 1, N = 22 000) the pipeline never fills and tail2_kernel is faster (0.028 against 0.043 ms)
    const bool enough_tiles = (int64_t)B * ((W + 31) / 32) >= 1184;
    if (((form == 3 && tail3_preferred(D) && enough_tiles) || form == 30 || form == 300) && ep && tail_tma_enabled() && tail3_supports(D) && (W % 4) == 0 && (reinterpret_cast<uintptr_t>(z) & 15) == 0 &&
        (!q_out || (reinterpret_cast<uintptr_t>(q_out) & 15) == 0)) {
        // all n_rep residual-sum replicas live in the workspace (permuted layout); their un-permuted sum is added to `resid`
        if (resid && !resid_rep) return cudaErrorInvalidValue;
        return launch_tail3(z, ep, e2, B, D, W, K, idx32, cand_cnt, cand_idx, idx_out, q_out, counts, resid, part, n_partials, meta, resid_rep,
                            n_rep, s);
    }
    auto fold = [&]() -> cudaError_t {            // sum the residual replicas into the caller's buffer
        if (n_rep <= 1) return cudaSuccess;
        const size_t blocks = (rep_stride + 255) / 256;
        fold_resid_kernel<<<(unsigned)(blocks < 1184 ? blocks : 1184), 256, 0, s>>>(resid, resid_rep, n_rep, rep_stride);
        note_launch();
        return cudaGetLastError();
    };
    if (form != 0 && tail_tma_enabled() && tail2_supports(D) && (W % 4) == 0 && (reinterpret_cast<uintptr_t>(z) & 15) == 0) {
        e = launch_tail2(z, codebook, e2, B, D, W, K, idx32, cand_cnt, cand_idx, idx_out, q_out, counts, resid, part, n_partials, meta,
                         resid_rep, n_rep, rep_stride, form, s);
        note_launch();
        return e != cudaSuccess ? e : fold();
    }
    if (tail_tma_enabled() && (W % 4) == 0 && (reinterpret_cast<uintptr_t>(z) & 15) == 0) {
        // TMA-fed tail: tiles of 32 frames that never straddle a batch item
        const int box_dims = D <= 256 ? D : D / 2;
        CUtensorMap map;
        if (make_latent_map(&map, z, (uint64_t)B, (uint64_t)D, (uint64_t)W, TL_F, (uint32_t)box_dims) != 0) return cudaErrorInvalidValue;
        const int tiles_per_item = (int)((W + TL_F - 1) / TL_F);
        const int64_t num_tiles = (int64_t)B * tiles_per_item;
#define VQB_TAIL_TMA(LPF, J) e = launch_tail_tma_t<LPF, J, 8, 2>(map, codebook, e2, D, W, tiles_per_item, num_tiles, box_dims, idx32, cand_cnt, cand_idx, idx_out, q_out, counts, resid, part, n_partials, meta, resid_rep, n_rep, rep_stride, s)
#define VQB_TAIL_TMA41(LPF, J) e = launch_tail_tma_t<LPF, J, 4, 1>(map, codebook, e2, D, W, tiles_per_item, num_tiles, box_dims, idx32, cand_cnt, cand_idx, idx_out, q_out, counts, resid, part, n_partials, meta, resid_rep, n_rep, rep_stride, s)
        if (tail_tma_variant() == 1) VQB_DISPATCH_D8(D, VQB_TAIL_TMA41);
        else VQB_DISPATCH_D8(D, VQB_TAIL_TMA);
#undef VQB_TAIL_TMA
#undef VQB_TAIL_TMA41
        note_launch();
        return e != cudaSuccess ? e : fold();
    }
    const int64_t tiles = (N + TL_F - 1) / TL_F;
    int64_t grid = n_partials < tiles ? n_partials : tiles;
    if (grid < 1) grid = 1;
    const int g = (int)grid;
#define VQB_TAIL(LPF, J) e = launch_tail_t<LPF, J>(z, codebook, e2, D, W, N, K, idx32, cand_cnt, cand_idx, idx_out, q_out, counts, resid, part, g, meta, resid_rep, n_rep, rep_stride, s)
    VQB_DISPATCH_D8(D, VQB_TAIL);   // measured: 8 lanes per frame beat 16 for the tail at D = 256 (0.91 vs 1.09 ms per 2^20 frames)
#undef VQB_TAIL
    note_launch();
    return e != cudaSuccess ? e : fold();
}

// counts (int) and SSE partials (double) -> the fp32 statistics buffer [counts | resid | SSE | N]
__global__ void __launch_bounds__(256) pack_stats_kernel(const int* __restrict__ counts, const double* __restrict__ part,
                                                         int n_partials, int64_t N, int K, int D, float* __restrict__ stats,
                                                         bool accumulate) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i < K) stats[i] = (float)counts[i];
    if (blockIdx.x == 0) {
        __shared__ double red[256];
        double s = 0.0;
        for (int j = threadIdx.x; j < n_partials; j += 256) s += part[j];
        red[threadIdx.x] = s;
        __syncthreads();
        for (int o = 128; o > 0; o >>= 1) {
            if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            const size_t base = (size_t)K * ((size_t)D + 1);
            stats[base] = (float)(red[0] + (accumulate ? (double)stats[base] : 0.0));
            stats[base + 1] = (float)((double)N + (accumulate ? (double)stats[base + 1] : 0.0));
        }
    }
}

cudaError_t launch_pack_stats(const int* counts, const float* sse_partials, int n_partials, int64_t N, int K, int D,
                              float* stats, bool accumulate, cudaStream_t s) {
    pack_stats_kernel<<<(K + 255) / 256, 256, 0, s>>>(counts, reinterpret_cast<const double*>(sse_partials), n_partials, N, K, D,
                                                      stats, accumulate);
    note_launch();
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ finalize
__global__ void __launch_bounds__(1024) finalize_kernel(const float* __restrict__ stats, int K, int D, float beta,
                                                        float* __restrict__ losses) {
    __shared__ double red[32];
    const size_t base = (size_t)K * ((size_t)D + 1);
    const float n = stats[base + 1];
    double s = 0.0;
    for (int k = threadIdx.x; k < K; k += 1024) {
        const float p = stats[k] / n;                         // avg_probs (vector_quantizer.py:49)
        s += (double)(p * logf(p + 1e-10f));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < 32; ++i) t += red[i];
        const float mse = (float)((double)stats[base] / ((double)n * (double)D));
        losses[0] = mse;                                     // embedding_loss (:46)
        losses[1] = beta * mse;                              // commitment_loss (:45)
        losses[2] = expf((float)(-t));                       // perplexity (:50)
    }
}

cudaError_t launch_finalize(const float* stats, int K, int D, float beta, float* losses, cudaStream_t s) {
    finalize_kernel<<<1, 1024, 0, s>>>(stats, K, D, beta, losses);
    note_launch();
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ backward
// dX[b,:,w] = Gq[b,:,w] + g_c * beta * 2 (x - q) / (N D); same float4 frame-major tile as the tail kernel: the commitment
// term is formed per frame (8 or 16 lanes each) in shared memory, the upstream gradient is added on the way out.
template <int LPF, int J>
__global__ void __launch_bounds__(256, (J >= 6) ? 2 : 3) backward_dx_kernel(const float* __restrict__ z, const float* __restrict__ E,
                                                             const int64_t* __restrict__ idx, const float* __restrict__ Gq,
                                                             const float* __restrict__ g_c, float beta, int D, int64_t W,
                                                             int64_t N, int K, float* __restrict__ dX, int ldg) {
    extern __shared__ __align__(16) float Xs[];   // [32][D + 4] latents, then [32][D + 4] upstream gradient
    constexpr int FPW = 32 / LPF;
    constexpr int ITER = (TL_F / 8) / FPW;
    const int ld = D + 4;
    float* Gs = Xs + TL_F * ld;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = lane / LPF, sl = lane % LPF;
    const float gc = g_c ? *g_c : 0.f;
    const float coef = gc * beta * (2.0f / ((float)N * (float)D));
    for (int64_t tile = blockIdx.x; tile * TL_F < N; tile += gridDim.x) {
        const int64_t nl = tile * TL_F + lane;
        const bool valid = nl < N;
        int64_t b = 0, w = 0;
        if (valid) { b = nl / W; w = nl - b * W; }
        const size_t col = (size_t)b * D * W + w;
        int64_t k_r[ITER];
#pragma unroll
        for (int it = 0; it < ITER; ++it) {
            const int64_t n = tile * TL_F + warp * (TL_F / 8) + it * FPW + sub;
            k_r[it] = (n < N) ? idx[n] : 0;
        }
        __syncthreads();
        if (ldg) {
            tile_load(Xs, ld, z, col, W, D, valid);
            if (Gq) tile_load(Gs, ld, Gq, col, W, D, valid);
        } else {
            tile_fetch_async(Xs, ld, z, col, W, D, valid);
            if (Gq) tile_fetch_async(Gs, ld, Gq, col, W, D, valid);
            tile_fetch_wait();
        }
        __syncthreads();
#pragma unroll
        for (int it = 0; it < ITER; ++it) {
            const int f = warp * (TL_F / 8) + it * FPW + sub;
            const bool ok = code_ok(k_r[it], K);
            const float* er = E + (size_t)(ok ? k_r[it] : 0) * D;
            const float nanv = __int_as_float(0x7fc00000);
#pragma unroll
            for (int j = 0; j < J; ++j) {
                const int d = 4 * sl + 4 * LPF * j;
                if (d < D) {
                    const float4 x = *reinterpret_cast<const float4*>(Xs + f * ld + d);
                    const float4 q = ok ? *reinterpret_cast<const float4*>(er + d) : make_float4(nanv, nanv, nanv, nanv);
                    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (Gq) g = *reinterpret_cast<const float4*>(Gs + f * ld + d);
                    float4 o;
                    o.x = g.x + coef * __fsub_rn(x.x, q.x); o.y = g.y + coef * __fsub_rn(x.y, q.y);
                    o.z = g.z + coef * __fsub_rn(x.z, q.z); o.w = g.w + coef * __fsub_rn(x.w, q.w);
                    *reinterpret_cast<float4*>(Xs + f * ld + d) = o;
                }
            }
        }
        __syncthreads();
        tile_store(Xs, ld, dX, col, W, D, valid);
    }
}

// Second form of dX: only the codeword term needs a gather, so only the codewords are transposed.  Per tile of 32 frames of one
// batch item: the 32 selected codebook rows are fetched with coalesced 16-byte loads (8 lanes per row) and written transposed
// into Es[d][33] (conflict-free), then every warp streams whole latent rows in the tensors' own [B, D, W] layout -
// dX[b, d, w0 + lane] = Gq + coef * (x - Es[d][lane]) - with 2 x DX_UNROLL independent coalesced loads in flight per thread and
// no shared-memory round trip for the latents or the upstream gradient.
constexpr int DX_UNROLL = 8;

template <int J>
__global__ void __launch_bounds__(256) backward_dx_rows_kernel(const float* __restrict__ z, const float* __restrict__ E,
                                                               const int64_t* __restrict__ idx, const float* __restrict__ Gq,
                                                               const float* __restrict__ g_c, float beta, int D, int64_t W, int64_t N,
                                                               int K, int tiles_per_item, int64_t num_tiles, float* __restrict__ dX) {
    extern __shared__ __align__(16) float Es[];   // [D][33] codeword components of the tile's frames, dim-major
    constexpr int LPF = 8, EP = TL_F + 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = lane / LPF, sl = lane % LPF;
    const float gc = g_c ? *g_c : 0.f;
    const float coef = gc * beta * (2.0f / ((float)N * (float)D));
    for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int b = (int)(tile / tiles_per_item), w0 = (int)(tile - (int64_t)b * tiles_per_item) * TL_F;
        {   // this warp's four frames: one codebook row each, 8 lanes x 16 bytes per request
            const int f = warp * 4 + sub;
            const bool live = w0 + f < W;
            const int64_t k = live ? idx[(int64_t)b * W + w0 + f] : 0;
            const bool ok = code_ok(k, K);
            const float* er = E + (size_t)(ok ? k : 0) * D;
            const float nanv = __int_as_float(0x7fc00000);
            float4 ev[J];
#pragma unroll
            for (int j = 0; j < J; ++j) {
                const int d = 4 * sl + 4 * LPF * j;
                ev[j] = (live && d < D) ? (ok ? *reinterpret_cast<const float4*>(er + d) : make_float4(nanv, nanv, nanv, nanv))
                                        : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int j = 0; j < J; ++j) {
                const int d = 4 * sl + 4 * LPF * j;
                if (d < D) {
                    Es[(d + 0) * EP + f] = ev[j].x;
                    Es[(d + 1) * EP + f] = ev[j].y;
                    Es[(d + 2) * EP + f] = ev[j].z;
                    Es[(d + 3) * EP + f] = ev[j].w;
                }
            }
        }
        __syncthreads();
        if (w0 + lane < W) {
            const size_t at = (size_t)b * D * W + w0 + lane;
            for (int d0 = warp * DX_UNROLL; d0 < D; d0 += 8 * DX_UNROLL) {
                float x[DX_UNROLL], g[DX_UNROLL];
#pragma unroll
                for (int u = 0; u < DX_UNROLL; ++u) {
                    const int d = d0 + u;
                    x[u] = (d < D) ? ld_stream(z + at + (size_t)d * W) : 0.f;
                    g[u] = (d < D && Gq) ? ld_stream(Gq + at + (size_t)d * W) : 0.f;
                }
#pragma unroll
                for (int u = 0; u < DX_UNROLL; ++u) {
                    const int d = d0 + u;
                    if (d < D) st_stream(dX + at + (size_t)d * W, g[u] + coef * __fsub_rn(x[u], Es[d * EP + lane]));
                }
            }
        }
        __syncthreads();                           // Es is rewritten for the next tile
    }
}

cudaError_t launch_backward_dx(const float* z, const float* codebook, const int64_t* idx, const float* Gq, const float* g_c,
                               float beta, int B, int D, int64_t W, int K, float* dX, cudaStream_t s) {
    const int64_t N = (int64_t)B * W;
    cudaError_t e = cudaSuccess;
    if (D <= 256 && env_get(ENV_DX_TILES, 0) != 1) {
        const int tiles_per_item = (int)((W + TL_F - 1) / TL_F);
        const int64_t num_tiles = (int64_t)B * tiles_per_item;
        const size_t smem = (size_t)D * (TL_F + 1) * 4;
        auto go = [&](auto kernel) -> cudaError_t {
            static thread_local size_t cached_smem = ~(size_t)0;   // see launch_tail_tma_t
            static thread_local const void* cached_kernel = nullptr;
            static thread_local int cached_blocks = 0, cached_dev = -1;
            cudaError_t e2 = cudaSuccess;
            int dev = 0;
            cudaGetDevice(&dev);
            if (cached_smem != smem || cached_kernel != reinterpret_cast<const void*>(kernel) || cached_dev != dev) {
                if ((e2 = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024)) != cudaSuccess) return e2;
                int per_sm = 1, sms = 148;
                if ((e2 = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, 256, smem)) != cudaSuccess) return e2;
                cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
                cached_blocks = sms * (per_sm < 1 ? 1 : per_sm);
                cached_smem = smem;
                cached_kernel = reinterpret_cast<const void*>(kernel);
                cached_dev = dev;
            }
            int64_t grid = cached_blocks;
            if (grid > num_tiles) grid = num_tiles;
            if (grid < 1) grid = 1;
            kernel<<<(unsigned)grid, 256, smem, s>>>(z, codebook, idx, Gq, g_c, beta, D, W, N, K, tiles_per_item, num_tiles, dX);
            return cudaGetLastError();
        };
        if (D <= 32) e = go(backward_dx_rows_kernel<1>);
        else if (D <= 64) e = go(backward_dx_rows_kernel<2>);
        else if (D <= 96) e = go(backward_dx_rows_kernel<3>);
        else if (D <= 128) e = go(backward_dx_rows_kernel<4>);
        else if (D <= 192) e = go(backward_dx_rows_kernel<6>);
        else e = go(backward_dx_rows_kernel<8>);
        if (e != cudaSuccess) return e;
        note_launch();
        return cudaSuccess;
    }
    const size_t smem = (size_t)2 * TL_F * (D + 4) * 4;
    const int64_t tiles = (N + TL_F - 1) / TL_F;
    int64_t grid = tiles < kTailGridMax ? tiles : kTailGridMax;
    if (grid < 1) grid = 1;
#define VQB_DX(LPF, J)                                                                                                          \
    do {                                                                                                                        \
        e = cudaFuncSetAttribute(backward_dx_kernel<LPF, J>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);          \
        if (e == cudaSuccess) backward_dx_kernel<LPF, J><<<(unsigned)grid, 256, smem, s>>>(z, codebook, idx, Gq, g_c, beta, D, W, N, K, dX, tile_ldg_mode()); \
    } while (0)
    VQB_DISPATCH_D(D, VQB_DX);
#undef VQB_DX
    if (e != cudaSuccess) return e;
    note_launch();
    return cudaGetLastError();
}

// dE[k,:] = -g_e * (2 / (N D)) * resid[k,:]; rows never selected have resid == 0 exactly -> dE == 0 exactly.
__global__ void __launch_bounds__(256) backward_de_kernel(const float* __restrict__ stats, const float* __restrict__ g_e,
                                                          int K, int D, float* __restrict__ dE) {
    const size_t total = (size_t)K * D;
    const float n = stats[total + K + 1];
    const float coef = -(g_e ? *g_e : 0.f) * (2.0f / (n * (float)D));
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (size_t)gridDim.x * 256)
        dE[i] = coef * stats[K + i];
}

cudaError_t launch_backward_de(const float* stats, const float* g_e, int K, int D, float* dE, cudaStream_t s) {
    const size_t total = (size_t)K * D;
    size_t grid = (total + 255) / 256;
    if (grid > 148 * 16) grid = 148 * 16;
    backward_de_kernel<<<(unsigned)grid, 256, 0, s>>>(stats, g_e, K, D, dE);
    note_launch();
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ EMA codebook (extension)
// Exponential-moving-average codebook update from the SAME statistics buffer (not in the reference, which trains the codebook
// with Adam; off by default - SURVEY.md row f4):
//   cluster_size <- g * cluster_size + (1-g) * counts ;  embed_sum <- g * embed_sum + (1-g) * (resid + counts * e)
//   e_k <- embed_sum_k / ((cluster_size_k + eps) / (n + K eps) * n),  n = sum_k cluster_size_k
__global__ void __launch_bounds__(1024) ema_cluster_kernel(const float* __restrict__ stats, float* __restrict__ cluster_size, int K,
                                                           float decay) {
    __shared__ double red[32];
    double s = 0.0;
    for (int k = threadIdx.x; k < K; k += 1024) {
        const float c = decay * cluster_size[k] + (1.f - decay) * stats[k];
        cluster_size[k] = c;
        s += (double)c;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < 32; ++i) t += red[i];
        cluster_size[K] = (float)t;
    }
}
__global__ void __launch_bounds__(256) ema_codebook_kernel(const float* __restrict__ stats, const float* __restrict__ cluster_size,
                                                           float* __restrict__ embed_sum, float* __restrict__ E, int K, int D,
                                                           float decay, float eps) {
    const size_t total = (size_t)K * D;
    const float n = cluster_size[K];
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (size_t)gridDim.x * 256) {
        const int k = (int)(i / D);
        const float sumx = stats[K + i] + stats[k] * E[i];                 // sum of the frames assigned to code k
        const float m = decay * embed_sum[i] + (1.f - decay) * sumx;
        embed_sum[i] = m;
        E[i] = m / ((cluster_size[k] + eps) / (n + (float)K * eps) * n);
    }
}

cudaError_t launch_ema_update(const float* stats, float* codebook, float* cluster_size, float* embed_sum, int K, int D, float decay,
                              float eps, cudaStream_t s) {
    ema_cluster_kernel<<<1, 1024, 0, s>>>(stats, cluster_size, K, decay);
    const size_t total = (size_t)K * D;
    size_t grid = (total + 255) / 256;
    if (grid > 148 * 8) grid = 148 * 8;
    ema_codebook_kernel<<<(unsigned)grid, 256, 0, s>>>(stats, cluster_size, embed_sum, codebook, K, D, decay, eps);
    note_launch(2);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ one-hot / gather / windows
__global__ void __launch_bounds__(256) onehot_kernel(const int64_t* __restrict__ idx, int64_t N, int K, float* __restrict__ out) {
    const int64_t n = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (n < N) {
        const int64_t k = idx[n];
        if (code_ok(k, K)) out[(size_t)n * K + k] = 1.0f;
    }
}

cudaError_t launch_onehot(const int64_t* idx, int64_t N, int K, float* out, cudaStream_t s) {
    cudaError_t e = cudaMemsetAsync(out, 0, (size_t)N * K * sizeof(float), s);
    if (e != cudaSuccess) return e;
    onehot_kernel<<<(unsigned)((N + 255) / 256), 256, 0, s>>>(idx, N, K, out);
    note_launch();
    return cudaGetLastError();
}

template <int LPF, int J>
__global__ void __launch_bounds__(256, (J >= 6) ? 2 : 3) gather_kernel(const float* __restrict__ E, const int64_t* __restrict__ idx, int D,
                                                        int64_t W, int64_t N, int K, float* __restrict__ out) {
    extern __shared__ __align__(16) float Xs[];   // [32][D + 4]
    constexpr int FPW = 32 / LPF;
    constexpr int ITER = (TL_F / 8) / FPW;
    const int ld = D + 4;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = lane / LPF, sl = lane % LPF;
    for (int64_t tile = blockIdx.x; tile * TL_F < N; tile += gridDim.x) {
        const int64_t nl = tile * TL_F + lane;
        const bool valid = nl < N;
        int64_t b = 0, w = 0;
        if (valid) { b = nl / W; w = nl - b * W; }
        const size_t col = (size_t)b * D * W + w;
        __syncthreads();
#pragma unroll
        for (int it = 0; it < ITER; ++it) {
            const int f = warp * (TL_F / 8) + it * FPW + sub;
            const int64_t n = tile * TL_F + f;
            if (n < N) {
                const int64_t k = idx[n];
                const bool ok = code_ok(k, K);
                const float* er = E + (size_t)(ok ? k : 0) * D;
                const float nanv = __int_as_float(0x7fc00000);
#pragma unroll
                for (int j = 0; j < J; ++j) {
                    const int d = 4 * sl + 4 * LPF * j;
                    if (d < D) *reinterpret_cast<float4*>(Xs + f * ld + d) = ok ? *reinterpret_cast<const float4*>(er + d) : make_float4(nanv, nanv, nanv, nanv);
                }
            }
        }
        __syncthreads();
        tile_store(Xs, ld, out, col, W, D, valid);
    }
}

cudaError_t launch_gather(const float* codebook, const int64_t* idx, int B, int D, int64_t W, int K, float* out, cudaStream_t s) {
    const int64_t N = (int64_t)B * W;
    const size_t smem = (size_t)TL_F * (D + 4) * 4;
    const int64_t tiles = (N + TL_F - 1) / TL_F;
    int64_t grid = tiles < kTailGridMax ? tiles : kTailGridMax;
    if (grid < 1) grid = 1;
    cudaError_t e = cudaSuccess;
#define VQB_GA(LPF, J)                                                                                                 \
    do {                                                                                                               \
        e = cudaFuncSetAttribute(gather_kernel<LPF, J>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);       \
        if (e == cudaSuccess) gather_kernel<LPF, J><<<(unsigned)grid, 256, smem, s>>>(codebook, idx, D, W, N, K, out); \
    } while (0)
    VQB_DISPATCH_D(D, VQB_GA);
#undef VQB_GA
    if (e != cudaSuccess) return e;
    note_launch();
    return cudaGetLastError();
}

// idx [B, L] -> tokens [B, n_win, window] (pad_id past L) + mask (bert.py:50-69)
__global__ void __launch_bounds__(256) window_kernel(const int64_t* __restrict__ idx, int64_t L, int window, int64_t n_win,
                                                     int64_t total, int64_t pad_id, int64_t* __restrict__ tokens,
                                                     float* __restrict__ mask) {
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
        const int64_t per_b = n_win * window;
        const int64_t b = i / per_b, pos = i - b * per_b;
        const bool real = pos < L;
        tokens[i] = real ? idx[b * L + pos] : pad_id;
        mask[i] = real ? 1.0f : 0.0f;
    }
}

cudaError_t launch_window(const int64_t* idx, int B, int64_t L, int window, int64_t pad_id, int64_t* tokens, float* mask,
                          cudaStream_t s) {
    const int64_t n_win = (L + window - 1) / window;
    const int64_t total = (int64_t)B * n_win * window;
    int64_t grid = (total + 255) / 256;
    if (grid > 148 * 16) grid = 148 * 16;
    if (grid < 1) grid = 1;
    window_kernel<<<(unsigned)grid, 256, 0, s>>>(idx, L, window, n_win, total, pad_id, tokens, mask);
    note_launch();
    return cudaGetLastError();
}

}  // namespace vqb
