// The tail of the VQ bottleneck, second form (round 2): per frame settle the index (fp32 rescoring of the shortlisted codes in
// the reference's op order, ties -> lowest index), gather the codeword, write the straight-through value fl(x + fl(e - x)) back
// in [B, D, W], accumulate SSE, the code histogram and the per-code residual sums.  Replaces vector_quantizer.py:37-52 after
// the search; same results as tail_tma_kernel (vqb_kernels.cu), which stays as the generic-D form.
//
// What changed against tail_tma_kernel, each answering a line of its ncu profile (profiles/r01_tail_cfg3_ncu_full.txt:
// 18.6 % warps active, 16 % of the stall samples at the block barrier, 12 + 7 % on codebook gathers, 3.9 G warp instructions):
//  * D = 32 J is a compile-time constant for every instance and the per-frame phase streams - it never holds the latent, the
//    codeword of this round, the codeword of the next round and the winner at once.
//  * Rescoring is DEALT: the (frame, code) pairs of a warp's frames go one per 8-lane group and round (software-pipelined two
//    deep), so a warp runs ceil(pairs / 4) rounds instead of the longest shortlist among its concurrent frames, and the four
//    warps of a block reach the write-back barrier at about the same time.
//  * The tile's shortlists are staged in shared memory before the box wait: a candidate code is never a global load that the
//    codeword gather has to wait for.
//  * |x|^2 is evaluated once per frame that has pairs, not once per pair.
//  * Runs of frames with the same code (collapsed codebooks early in training: BASELINE config 4 starts at perplexity 1.75;
//    real audio: neighbouring frames) keep their histogram count and, for D <= 128, their residual sum in registers and issue
//    ONE set of atomics per run instead of one per frame - same-address atomics were the whole tail at K = 512 (0.81 -> 0.10 ms).
// Measured (profiles/r02_exp_tail_forms.jsonl): D = 256 equal to tail_tma_kernel within 3 % at any occupancy - the pass is bound
// by the SM's L1TEX data pipe (every phase goes through shared memory), see DESIGN.md section 3.3.
#include "vqb_internal.h"
#include "vqb_ptx.cuh"

namespace vqb {

namespace t2 {

constexpr int LPF = 8;            // lanes per frame: lane `sl` of a group owns dims 4 sl + 32 j, j < J
constexpr int NWARP = 4;
constexpr int kPairMax = 12;      // shortlist entries the search publishes per frame (kCandFill in vqb_tc.cu)

__device__ __forceinline__ float group_sum8(float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v;
}

// J: D = 32 J.  TF: frames per tile (16 or 32).  kRun: keep the residual sum of a run of equal codes in registers (J <= 4).
template <int J, int TF, bool kResid, bool kRun>
__global__ void __launch_bounds__(32 * NWARP, (J >= 6) ? (TF == 16 ? 5 : 3) : ((J >= 3) ? 6 : 8))
tail2_kernel(const __grid_constant__ CUtensorMap tmap_z, const float* __restrict__ E, const float* __restrict__ e2, int64_t W,
             int tiles_per_item, int64_t num_tiles, const int* __restrict__ idx32, const uint8_t* __restrict__ cand_cnt,
             const uint16_t* __restrict__ cand_idx, int64_t* __restrict__ idx_out, float* __restrict__ q_out, int* __restrict__ counts,
             float* __restrict__ resid, double* __restrict__ sse_partials, WsMeta* meta, float* resid_rep, int n_rep, size_t rep_stride,
             int l2_once) {
    using namespace ptx;
    constexpr int D = 32 * J, LD = D + 4;
    constexpr int FW = TF / NWARP;                 // frames a warp owns per tile: 4 or 8
    constexpr int ITER = FW / 4;                   // four frames at a time (one per lane group)
    static_assert(TF == 16 || TF == 32, "tile = 16 or 32 frames");
    resid = pick_resid_replica(resid, resid_rep, n_rep, rep_stride);
    extern __shared__ __align__(128) float t2_smem[];
    float* box = t2_smem;                          // [D][TF]: the TMA box, dim-major
    float* Xs = box + D * TF;                      // [TF][LD]: the same tile, frame-major
    float2* pairres = reinterpret_cast<float2*>(Xs + TF * LD);       // [NWARP][FW * kPairMax] (distance, code) of the dealt pairs
    float* x2s = reinterpret_cast<float*>(pairres + NWARP * FW * kPairMax);   // [TF] |x|^2 of the frames that have pairs
    uint16_t* sC = reinterpret_cast<uint16_t*>(x2s + TF);            // [TF][kCandMax] the tile's shortlists (or final codes), staged early
    __shared__ double red[NWARP];
    __shared__ __align__(8) unsigned long long full;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 3, sl = lane & 7;
    const bool resid_v4 = kResid && (reinterpret_cast<uintptr_t>(resid) & 15) == 0;   // stats + K is 16-byte aligned iff K % 4 == 0
    float sse = 0.f, sse_c = 0.f;                  // Kahan-compensated per-thread SSE
    unsigned int n_resc = 0, n_short = 0;
    // run of equal codes of this lane group: count and (kRun) residual sum, flushed when the code changes / at the end
    int run_k = -1, run_n = 0;
    float4 run_r[kRun ? J : 1];

    auto flush_run = [&]() {
        if (run_k >= 0) {
            if (sl == 0) atomicAdd(counts + run_k, run_n);
            if (kResid && kRun) {
#pragma unroll
                for (int j = 0; j < J; ++j) {
                    float* rp = resid + (size_t)run_k * D + 4 * sl + 32 * j;
                    if (resid_v4) red_add_v4(rp, run_r[j].x, run_r[j].y, run_r[j].z, run_r[j].w);
                    else { atomicAdd(rp, run_r[j].x); atomicAdd(rp + 1, run_r[j].y); atomicAdd(rp + 2, run_r[j].z); atomicAdd(rp + 3, run_r[j].w); }
                }
            }
        }
    };

    if (threadIdx.x == 0) {
        mbar_init(smem_u32(&full), 1);
        fence_barrier_init();
        prefetch_tmap(&tmap_z);
    }
    __syncthreads();
    auto issue = [&](int64_t tile) {               // one thread: the whole [D][TF] box of `tile` (frames past W arrive as zeros)
        const int b = (int)(tile / tiles_per_item), w0 = (int)(tile - (int64_t)b * tiles_per_item) * TF;
        const uint32_t bar = smem_u32(&full);
        mbar_expect_tx(bar, (uint32_t)D * TF * 4);
        if (l2_once) tma_load_3d_once(smem_u32(box), &tmap_z, bar, w0, 0, b);
        else tma_load_3d(smem_u32(box), &tmap_z, bar, w0, 0, b);
    };
    int64_t tile = blockIdx.x;
    if (threadIdx.x == 0 && tile < num_tiles) issue(tile);
    for (uint32_t it = 0; tile < num_tiles; tile += gridDim.x, ++it) {
        const int b = (int)(tile / tiles_per_item), w0 = (int)(tile - (int64_t)b * tiles_per_item) * TF;
        const int64_t n0 = (int64_t)b * W + w0;     // global frame id of the tile's first frame
        const int wlim = (int)((W - w0) < TF ? (W - w0) : TF);   // frames of this tile that exist
        const int fbase = warp * FW;                // this warp's frames: fbase .. fbase + FW - 1
        // ---- shortlist lengths of the warp's frames (lane u < FW holds frame fbase + u), requested before the box wait
        int cnt_l = 0;
        if (lane < FW && fbase + lane < wlim) cnt_l = idx32 ? kCandFinal : (int)cand_cnt[n0 + fbase + lane];
        // ---- and the shortlists themselves (32 bytes per frame, two lanes per frame) into shared memory: every later use of a
        // candidate code is then a shared-memory read instead of a global load the codeword gather has to wait for (ncu: 22 % of
        // the stall samples sat on exactly that dependent pair of loads)
        if (lane < 2 * FW) {
            const int u = lane >> 1, half = lane & 1;
            uint4 v = make_uint4(0u, 0u, 0u, 0u);
            if (fbase + u < wlim) {
                const int64_t n = n0 + fbase + u;
                if (idx32) { if (half == 0) v.x = (uint32_t)idx32[n]; }
                else v = reinterpret_cast<const uint4*>(cand_idx + (size_t)n * kCandMax)[half];
            }
            reinterpret_cast<uint4*>(sC + (fbase + u) * kCandMax)[half] = v;
        }
        mbar_wait(smem_u32(&full), it & 1u);
        {   // box [d][TF frames] -> frame-major tile: conflict-free LDS.32 x 4 + one STS.128 per 4 dims
            if (TF == 32) {
                const float* bx = box + lane;
#pragma unroll 2
                for (int d0 = warp * 4; d0 < D; d0 += 4 * NWARP) {
                    float4 v;
                    v.x = bx[(d0 + 0) * TF]; v.y = bx[(d0 + 1) * TF]; v.z = bx[(d0 + 2) * TF]; v.w = bx[(d0 + 3) * TF];
                    *reinterpret_cast<float4*>(Xs + lane * LD + d0) = v;
                }
            } else {
                // 16 frames: a half-warp per dim group; the upper half reads its four rows in the order 1,0,3,2 so that the two
                // halves never meet in a bank (rows are 16 words long)
                const int f = lane & 15, hi = lane >> 4;
                const float* bx = box + f;
#pragma unroll 2
                for (int d0 = (warp * 2 + hi) * 4; d0 < D; d0 += 8 * NWARP) {
                    float a0 = bx[(d0 + (0 ^ hi)) * TF], a1 = bx[(d0 + (1 ^ hi)) * TF], a2 = bx[(d0 + (2 ^ hi)) * TF], a3 = bx[(d0 + (3 ^ hi)) * TF];
                    float4 v;
                    v.x = hi ? a1 : a0; v.y = hi ? a0 : a1; v.z = hi ? a3 : a2; v.w = hi ? a2 : a3;
                    *reinterpret_cast<float4*>(Xs + f * LD + d0) = v;
                }
            }
        }
        __syncthreads();
        {
            const int64_t next = tile + gridDim.x;
            if (threadIdx.x == 0 && next < num_tiles) issue(next);      // the box has been consumed: refill it behind the compute
        }
        // ---- pairs of the warp: frame u contributes np_u = cnt_u (> 1) pairs; inclusive prefix over the FW frames
        const int np_l = (cnt_l != kCandFinal && cnt_l > 1) ? cnt_l : 0;
        int incl_l = np_l;
#pragma unroll
        for (int o = 1; o < FW; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl_l, o);
            if (lane >= o) incl_l += t;
        }
        const int total = __shfl_sync(0xffffffffu, incl_l, FW - 1);
        if (total > 0) {
            // |x|^2 of the frames that have pairs: sum of rounded squares per lane, then over the group (as tail_frame does)
#pragma unroll
            for (int i = 0; i < ITER; ++i) {
                // (every group computes: the group sums are full-warp shuffles and must not sit in a divergent branch)
                const int fi = i * 4 + g;
                const float* xr = Xs + (fbase + fi) * LD + 4 * sl;
                float x2 = 0.f;
#pragma unroll
                for (int j = 0; j < J; ++j) {
                    const float4 x = *reinterpret_cast<const float4*>(xr + 32 * j);
                    x2 = __fadd_rn(x2, __fmul_rn(x.x, x.x)); x2 = __fadd_rn(x2, __fmul_rn(x.y, x.y));
                    x2 = __fadd_rn(x2, __fmul_rn(x.z, x.z)); x2 = __fadd_rn(x2, __fmul_rn(x.w, x.w));
                }
                x2 = group_sum8(x2);
                if (sl == 0) x2s[fbase + fi] = x2;
            }
            __syncwarp();
            float2* pr = pairres + warp * (FW * kPairMax);
            const int rounds = (total + 3) >> 2;
            // pair p = 4 r + g of round r belongs to the frame whose [excl, incl) holds it; inactive groups rescore code 0 of frame 0
            auto pair_of = [&](int r, int& fi, int& kc) {
                const int p = 4 * r + g;
                int ex = 0;
                fi = 0;
#pragma unroll
                for (int u = 0; u < FW; ++u) {
                    const int iu = __shfl_sync(0xffffffffu, incl_l, u);
                    if (iu <= p) { fi = u + 1; ex = iu; }
                }
                const bool act = p < total;
                if (!act) { fi = 0; ex = p; }
                kc = act ? (int)sC[(fbase + fi) * kCandMax + (p - ex)] : 0;
                return act;
            };
            auto fetch_row = [&](int kc, float4 (&ev)[J], float& e2c) {
                const float* er = E + (size_t)kc * D + 4 * sl;
#pragma unroll
                for (int j = 0; j < J; ++j) ev[j] = *reinterpret_cast<const float4*>(er + 32 * j);
                e2c = e2[kc];
            };
            auto score = [&](int r, bool act, int fi, int kc, const float4 (&ev)[J], float e2c) {
                const float* xr = Xs + (fbase + fi) * LD + 4 * sl;
                float dot = 0.f;
#pragma unroll
                for (int j = 0; j < J; ++j) {
                    const float4 x = *reinterpret_cast<const float4*>(xr + 32 * j);
                    dot = fmaf(x.x, ev[j].x, dot); dot = fmaf(x.y, ev[j].y, dot);
                    dot = fmaf(x.z, ev[j].z, dot); dot = fmaf(x.w, ev[j].w, dot);
                }
                dot = group_sum8(dot);
                if (act && sl == 0) pr[4 * r + g] = make_float2(ref_distance(x2s[fbase + fi], e2c, dot), __int_as_float(kc));
            };
            // rounds are software-pipelined two deep (ping-pong registers): the codeword of round r + 1 is in flight while
            // round r is reduced
            float4 eva[J], evb[J];
            float e2a = 0.f, e2b = 0.f;
            int fia = 0, kca = 0, fib = 0, kcb = 0;
            bool acta = pair_of(0, fia, kca), actb = false;
            fetch_row(kca, eva, e2a);
            for (int r = 0; r < rounds; r += 2) {
                if (r + 1 < rounds) { actb = pair_of(r + 1, fib, kcb); fetch_row(kcb, evb, e2b); }
                score(r, acta, fia, kca, eva, e2a);
                if (r + 1 < rounds) {
                    if (r + 2 < rounds) { acta = pair_of(r + 2, fia, kca); fetch_row(kca, eva, e2a); }
                    score(r + 1, actb, fib, kcb, evb, e2b);
                }
            }
            __syncwarp();
        }
        // ---- own frames: final code, codeword gather, straight-through value, SSE, histogram, residual sums
#pragma unroll
        for (int i = 0; i < ITER; ++i) {
            const int fi = i * 4 + g;
            const int f = fbase + fi;
            const int in_o = __shfl_sync(0xffffffffu, incl_l, fi), np_o = __shfl_sync(0xffffffffu, np_l, fi);
            const bool live = f < wlim;
            if (live) {
                const int64_t n = n0 + f;
                int k;
                if (np_o > 0) {                     // settle among the dealt pairs: torch.argmin order, ties -> lowest index
                    const float2* pr = pairres + warp * (FW * kPairMax);
                    float bd = 0.f;
                    int bk = -1;
                    for (int pp = in_o - np_o; pp < in_o; ++pp) {
                        const float2 v = pr[pp];
                        const int kk = __float_as_int(v.y);
                        if (better(v.x, kk, bd, bk)) { bd = v.x; bk = kk; }
                    }
                    k = bk;
                    if (sl == 0) { n_resc += 1; n_short += np_o; }
                } else {
                    k = (int)sC[f * kCandMax];           // the only shortlisted code / the exact search's answer
                    if (sl == 0) n_short += 1;
                }
                float* xr = Xs + f * LD + 4 * sl;
                const float* er = E + (size_t)k * D + 4 * sl;
                float4 qv[J];
#pragma unroll
                for (int j = 0; j < J; ++j) qv[j] = *reinterpret_cast<const float4*>(er + 32 * j);
                if (k != run_k) {                   // a new run of this lane group
                    flush_run();
                    run_k = k;
                    run_n = 0;
                    if (kRun) {
#pragma unroll
                        for (int j = 0; j < J; ++j) run_r[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                }
                run_n += 1;
                float fs = 0.f;
#pragma unroll
                for (int j = 0; j < J; ++j) {
                    const float4 x = *reinterpret_cast<const float4*>(xr + 32 * j);
                    // r = fl(x - e) is exactly -fl(e - x): the residual sums want r, and the straight-through VALUE
                    // fl(x + fl(e - x)) (vector_quantizer.py:48) equals fl(x - r) bit for bit
                    float4 r, st;
                    r.x = __fsub_rn(x.x, qv[j].x); r.y = __fsub_rn(x.y, qv[j].y); r.z = __fsub_rn(x.z, qv[j].z); r.w = __fsub_rn(x.w, qv[j].w);
                    fs = fmaf(r.x, r.x, fs); fs = fmaf(r.y, r.y, fs); fs = fmaf(r.z, r.z, fs); fs = fmaf(r.w, r.w, fs);
                    st.x = __fsub_rn(x.x, r.x); st.y = __fsub_rn(x.y, r.y); st.z = __fsub_rn(x.z, r.z); st.w = __fsub_rn(x.w, r.w);
                    *reinterpret_cast<float4*>(xr + 32 * j) = st;
                    if (kResid) {
                        if (kRun) {
                            run_r[j].x += r.x; run_r[j].y += r.y; run_r[j].z += r.z; run_r[j].w += r.w;
                        } else {
                            float* rp = resid + (size_t)k * D + 4 * sl + 32 * j;
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
                if (sl == 0) idx_out[n] = (int64_t)k;
            }
        }
        if (q_out) {
            __syncthreads();
            if (TF == 32) {
                if (lane < wlim) {
                    const size_t step = (size_t)(4 * NWARP) * W;
                    float* p0 = q_out + ((size_t)b * D + warp * 4) * W + w0 + lane;
                    const float* xs = Xs + lane * LD + warp * 4;
#pragma unroll 2
                    for (int d0 = warp * 4; d0 < D; d0 += 4 * NWARP) {
                        const float4 v = *reinterpret_cast<const float4*>(xs);
                        st_stream(p0, v.x); st_stream(p0 + W, v.y); st_stream(p0 + 2 * W, v.z); st_stream(p0 + 3 * W, v.w);
                        p0 += step;
                        xs += 4 * NWARP;
                    }
                }
            } else {
                const int f = lane & 15, hi = lane >> 4;
                if (f < wlim) {
                    const size_t step = (size_t)(8 * NWARP) * W;
                    float* p0 = q_out + ((size_t)b * D + (warp * 2 + hi) * 4) * W + w0 + f;
                    const float* xs = Xs + f * LD + (warp * 2 + hi) * 4;
#pragma unroll 2
                    for (int d0 = (warp * 2 + hi) * 4; d0 < D; d0 += 8 * NWARP) {
                        const float4 v = *reinterpret_cast<const float4*>(xs);
                        st_stream(p0, v.x); st_stream(p0 + W, v.y); st_stream(p0 + 2 * W, v.z); st_stream(p0 + 3 * W, v.w);
                        p0 += step;
                        xs += 8 * NWARP;
                    }
                }
            }
        }
        __syncthreads();                            // the tile may be overwritten
    }
    flush_run();
    double t = (double)sse - (double)sse_c;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (lane == 0) red[warp] = t;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int i = 0; i < NWARP; ++i) s += red[i];
        sse_partials[blockIdx.x] = s;
    }
    if (n_resc | n_short) {
        atomicAdd(&meta->rescored, (unsigned long long)n_resc);
        atomicAdd(&meta->shortlisted, (unsigned long long)n_short);
    }
}

template <int J, int TF>
static cudaError_t launch_t(const CUtensorMap& map, const float* codebook, const float* e2, int64_t W, int tiles_per_item, int64_t num_tiles,
                            const int* idx32, const uint8_t* cand_cnt, const uint16_t* cand_idx, int64_t* idx_out, float* q_out, int* counts,
                            float* resid, double* part, int n_partials, WsMeta* meta, float* resid_rep, int n_rep, size_t rep_stride,
                            bool read_once, cudaStream_t s) {
    constexpr int D = 32 * J, FW = TF / NWARP;
    const size_t smem = (size_t)(D * TF + TF * (D + 4)) * 4 + (size_t)NWARP * FW * kPairMax * 8 + TF * 4 + (size_t)TF * kCandMax * 2;
    auto go = [&](auto kernel) -> cudaError_t {
        // attribute + occupancy are looked up once per kernel instance and device (this path is launch-bound for small batches)
        static thread_local const void* cached_kernel = nullptr;
        static thread_local int cached_blocks = 0, cached_dev = -1;
        cudaError_t e = cudaSuccess;
        int dev = 0;
        cudaGetDevice(&dev);
        if (cached_kernel != reinterpret_cast<const void*>(kernel) || cached_dev != dev) {
            if ((e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)) != cudaSuccess) return e;
            int per_sm = 1, sms = 148;
            if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, 32 * NWARP, smem)) != cudaSuccess) return e;
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
            cached_blocks = sms * (per_sm < 1 ? 1 : per_sm);
            cached_kernel = reinterpret_cast<const void*>(kernel);
            cached_dev = dev;
        }
        int64_t grid = cached_blocks;
        if (grid > n_partials) grid = n_partials;
        if (grid > num_tiles) grid = num_tiles;
        if (grid < 1) grid = 1;
        kernel<<<(unsigned)grid, 32 * NWARP, smem, s>>>(map, codebook, e2, W, tiles_per_item, num_tiles, idx32, cand_cnt, cand_idx, idx_out, q_out,
                                                        counts, resid, part, meta, resid_rep, n_rep, rep_stride, read_once ? 1 : 0);
        return cudaGetLastError();
    };
    constexpr bool kRun = J <= 4;
    return resid ? go(tail2_kernel<J, TF, true, kRun>) : go(tail2_kernel<J, TF, false, false>);
}

}  // namespace t2

bool tail2_supports(int D) { return D % 32 == 0 && (D / 32 <= 4 || D == 192 || D == 256); }

// Tile size: 32 frames (VQB_TAIL_FORM=216 selects 16-frame tiles: 5 blocks per SM instead of 3 at D = 256, measured 8 % SLOWER -
// the pass is bound by the L1TEX data pipe, not by latency).
cudaError_t launch_tail2(const float* z, const float* codebook, const float* e2, int B, int D, int64_t W, int K, const int* idx32,
                         const uint8_t* cand_cnt, const uint16_t* cand_idx, int64_t* idx_out, float* q_out, int* counts, float* resid,
                         double* part, int n_partials, WsMeta* meta, float* resid_rep, int n_rep, size_t rep_stride, int form,
                         cudaStream_t s) {
    (void)K;
    const int tf = form == 216 ? 16 : 32;
    CUtensorMap map;
    if (make_latent_map(&map, z, (uint64_t)B, (uint64_t)D, (uint64_t)W, (uint32_t)tf, (uint32_t)D) != 0) return cudaErrorInvalidValue;
    const int tiles_per_item = (int)((W + tf - 1) / tf);
    const int64_t num_tiles = (int64_t)B * tiles_per_item;
    const bool once = latents_read_once((size_t)B * D * W * 4);
    cudaError_t e = cudaErrorInvalidValue;
#define VQB_T2(J)                                                                                                                          \
    e = tf == 16 ? t2::launch_t<J, 16>(map, codebook, e2, W, tiles_per_item, num_tiles, idx32, cand_cnt, cand_idx, idx_out, q_out, counts, \
                                       resid, part, n_partials, meta, resid_rep, n_rep, rep_stride, once, s)                               \
                 : t2::launch_t<J, 32>(map, codebook, e2, W, tiles_per_item, num_tiles, idx32, cand_cnt, cand_idx, idx_out, q_out, counts, \
                                       resid, part, n_partials, meta, resid_rep, n_rep, rep_stride, once, s)
    switch (D / 32) {
        case 1: VQB_T2(1); break;
        case 2: VQB_T2(2); break;
        case 3: VQB_T2(3); break;
        case 4: VQB_T2(4); break;
        case 6: VQB_T2(6); break;
        case 8: VQB_T2(8); break;
        default: break;
    }
#undef VQB_T2
    return e;
}

}  // namespace vqb
