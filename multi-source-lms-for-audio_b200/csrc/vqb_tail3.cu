// The tail of the VQ bottleneck, third form (round 2, second half): the same per-frame work as tail2_kernel - settle the index
// (fp32 rescoring of the shortlisted codes in the reference's op order, ties -> lowest index), gather the codeword, write the
// straight-through value fl(x + fl(e - x)) back in [B, D, W], accumulate SSE, the code histogram and the per-code residual sums
// (vector_quantizer.py:37-52) - WITHOUT the two shared-memory transpositions that made tail2_kernel / tail_tma_kernel L1TEX-bound
// (~106 L1TEX wavefronts per frame at D = 256 against ~200 SM cycles per frame).
//
//  * The tile's latents arrive as ONE 128-byte-swizzled TMA box [D dims][32 frames] and are read IN PLACE.  LPF lanes (8, 4 or 2)
//    share a frame; lane `sl` owns the dims whose d % 8 lies in [S sl, S sl + S), S = 8 / LPF.  With the 128-byte swizzle the
//    16-byte chunk (4 frames) of row d sits at chunk position (f / 4) ^ (d % 8), so the 32 lanes of a warp (32 / LPF consecutive
//    frames x LPF lanes) always hit 32 different banks - no frame-major copy.
//  * The codebook is read from a PERMUTED fp32 copy (workspace, written by codebook_prep_kernel; tail3_perm_pos): the m-th dim a
//    lane owns sits at position (m / 4) 4 LPF + 4 sl + m % 4, so a lane's dims are 16-byte vectors and the lanes of a frame read
//    contiguous 16 LPF bytes.  The residual sums are accumulated in the same permuted layout (all replicas live in the workspace)
//    and un-permuted by fold_resid_perm_kernel when the replicas are summed into the caller's statistics.
//  * Small D wants few lanes per frame: at D = 64 the frame-major kernels spend 116 warp instructions per frame (ncu, issue slots
//    57 % busy) on per-frame bookkeeping that 8 lanes with 8 dims each cannot amortise.
//  * The straight-through value is written back into the box at the latent's own position and leaves as ONE TMA store
//    (cp.async.bulk.tensor ... global <- shared): no write-back loop, full 128-byte rows.
//  * Two boxes per block, four compute warps that never meet at a block barrier (each owns 8 frames of every tile) and a fifth
//    warp that only moves data: it waits for the four warps on an mbarrier, stores the box, waits until the store has read it and
//    refills it with the tile after next.
#include "vqb_internal.h"
#include "vqb_ptx.cuh"

namespace vqb {

namespace t3 {

constexpr int TF = 32;            // frames per tile = one 128-byte swizzle row
constexpr int kPairMax = 12;      // shortlist entries the search publishes per frame (kCandFill in vqb_tc.cu)

template <int LPF>
__device__ __forceinline__ float group_sum(float v) {
    if (LPF >= 8) v += __shfl_xor_sync(0xffffffffu, v, 4);
    if (LPF >= 4) v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v;
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, bool once) {
    if (once)   // written once, never read by this pass: evict first (same fixed policy encoding as tma_load_3d_once)
        asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%2, %3, %4}], [%1], %5;" ::"l"(map), "r"(src),
                     "r"(c0), "r"(c1), "r"(c2), "l"(0x12F0000000000000ull)
                     : "memory");
    else
        asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(map), "r"(src), "r"(c0), "r"(c1),
                     "r"(c2)
                     : "memory");
}
// Residual sums of one frame as ONE bulk reduction (the TMA unit adds `bytes` of shared memory into global memory with L2 atomics):
// the SM issues one instruction per frame instead of D / 4 red.v4 lane operations - at D = 256 the LSU's reduction issue rate
// (~1.3 cycles per lane operation) was what bounded the pass (1.48 ms with, 1.02 ms without the residual sums per 2^21 frames).
__device__ __forceinline__ void bulk_reduce_add_f32(float* dst, uint32_t src, uint32_t bytes) {
    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

struct Bars { unsigned long long full[2], done[2]; };

// J: D = 32 J.  LPF: lanes per frame (8, 4, 2; D / LPF <= 32 dims per lane).  kRun: keep the residual sum of a run of equal codes in
// registers (at most 16 dims per lane).  kAhead: the next tile's shortlists are fetched one tile ahead and |x|^2 rides along with the pair
// scoring (chosen per shape by measurement, see launch_tail3).  kBulk: residual sums leave as one bulk reduction per frame from a per-warp staging row (needs kResid, !kRun).
template <int J, int LPF, bool kResid, bool kRun, bool kBulk, bool kAhead>
__global__ void __launch_bounds__(32 * ((LPF == 2 ? 2 : 4) + 1), (32 * J / LPF >= 24) ? 3 : ((32 * J / LPF >= 12) ? (LPF <= 4 ? 5 : 4) : 6))
tail3_kernel(const __grid_constant__ CUtensorMap tmap_z, const __grid_constant__ CUtensorMap tmap_q, const float* __restrict__ Ep,
             const float* __restrict__ e2, int64_t W, int tiles_per_item, int num_tiles, const int* __restrict__ idx32,
             const uint8_t* __restrict__ cand_cnt, const uint16_t* __restrict__ cand_idx, int64_t* __restrict__ idx_out, int has_q,
             int* __restrict__ counts, float* __restrict__ resid_rep, int n_rep, size_t rep_stride, double* __restrict__ sse_partials,
             WsMeta* meta, int l2_once) {
    using namespace ptx;
    static_assert(!kBulk || (kResid && !kRun), "bulk reductions replace the per-lane red.v4 of the run-less form");
    constexpr int D = 32 * J;
    constexpr int G = 32 / LPF;                    // frames a warp works on at a time (one per lane group)
    constexpr int S = 8 / LPF;                     // dims a lane owns in every block of 8
    constexpr int M = D / LPF, T = M / 4;          // dims per lane; as float4s of the permuted codebook / residual rows
    constexpr int NWARP = LPF == 2 ? 2 : 4;        // compute warps; warp NWARP is the data mover
    constexpr int FW = TF / NWARP;                 // frames a compute warp owns per tile
    constexpr int ITER = FW / G;
    static_assert(M % 4 == 0 && M <= 32 && FW % G == 0, "unsupported (D, LPF)");
    constexpr uint32_t BOX_BYTES = (uint32_t)D * TF * 4;
    extern __shared__ __align__(1024) unsigned char t3_smem[];
    unsigned char* box0 = t3_smem;                                   // 2 x [D][32 frames] fp32, 128-byte swizzled rows
    float2* pairres = reinterpret_cast<float2*>(t3_smem + 2 * BOX_BYTES);              // [NWARP][FW * kPairMax] (distance, code)
    float* x2s = reinterpret_cast<float*>(pairres + NWARP * FW * kPairMax);            // [TF] |x|^2 of the frames that have pairs (D <= 128)
    uint16_t* sC = reinterpret_cast<uint16_t*>(x2s + TF);                              // [TF][kCandMax] the tile's shortlists
    double* red = reinterpret_cast<double*>(sC + TF * kCandMax);                       // [NWARP]
    Bars* bars = reinterpret_cast<Bars*>(red + NWARP);
    float* stage = reinterpret_cast<float*>(bars + 1);                                 // kBulk: [NWARP][G frames][D] residual rows (16-byte aligned)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane / LPF, sl = lane % LPF;
    const unsigned gmask = (LPF == 32 ? 0xffffffffu : ((1u << LPF) - 1u)) << (LPF * g);   // the lanes of this lane group
    float* resid = nullptr;
    if (kResid) {                                   // all replicas live in the workspace (permuted layout); chosen by SM id
        unsigned int smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        resid = resid_rep + (size_t)(smid % (unsigned int)n_rep) * rep_stride;
    }
    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) { mbar_init(smem_u32(&bars->full[i]), 1); mbar_init(smem_u32(&bars->done[i]), NWARP); }
        fence_barrier_init();
        prefetch_tmap(&tmap_z);
        if (has_q) prefetch_tmap(&tmap_q);
    }
    __syncthreads();
    const int tile0 = (int)blockIdx.x, tstep = (int)gridDim.x;

    if (warp == NWARP) {
        // ================================================================ data mover (one lane)
        if (lane == 0) {
            auto load = [&](int tile, int buf) {        // the whole [D][32] box of `tile` (frames past W arrive as zeros)
                const int b = tile / tiles_per_item, w0 = (tile - b * tiles_per_item) * TF;
                const uint32_t bar = smem_u32(&bars->full[buf]);
                mbar_expect_tx(bar, BOX_BYTES);
                if (l2_once) tma_load_3d_once(smem_u32(box0 + buf * BOX_BYTES), &tmap_z, bar, w0, 0, b);
                else tma_load_3d(smem_u32(box0 + buf * BOX_BYTES), &tmap_z, bar, w0, 0, b);
            };
            if (tile0 < num_tiles) load(tile0, 0);
            if (tile0 + tstep < num_tiles) load(tile0 + tstep, 1);
            uint32_t it = 0;
            for (int tile = tile0; tile < num_tiles; tile += tstep, ++it) {
                const int buf = (int)(it & 1u);
                mbar_wait(smem_u32(&bars->done[buf]), (it >> 1) & 1u);          // the compute warps are through with this box
                if (has_q) {
                    const int b = tile / tiles_per_item, w0 = (tile - b * tiles_per_item) * TF;
                    tma_store_3d(&tmap_q, smem_u32(box0 + buf * BOX_BYTES), w0, 0, b, l2_once != 0);   // frames past W are clipped
                    bulk_commit();
                }
                const int next = tile + 2 * tstep;
                if (next < num_tiles) {
                    if (has_q) bulk_wait_read0();                               // the store has read the box: it may be overwritten
                    load(next, buf);
                }
            }
            if (has_q) bulk_wait0();
        }
    } else {
        // ================================================================ compute warps: frames fbase .. fbase + FW - 1 of every tile
        const int fbase = warp * FW;
        float sse = 0.f, sse_c = 0.f;                  // Kahan-compensated per-thread SSE
        unsigned int n_resc = 0, n_short = 0;
        int run_k = -1, run_n = 0;                     // run of equal codes of this lane group
        float4 run_r[kRun ? T : 1];
        auto flush_run = [&]() {
            if (run_k >= 0) {
                if (sl == 0) atomicAdd(counts + run_k, run_n);
                if (kResid && kRun) {
#pragma unroll
                    for (int t = 0; t < T; ++t)
                        red_add_v4(resid + (size_t)run_k * D + 4 * LPF * t + 4 * sl, run_r[t].x, run_r[t].y, run_r[t].z, run_r[t].w);
                }
            }
        };
        float2* pr = pairres + warp * (FW * kPairMax);
        // (batch item, tile inside the item) of this warp's current tile, advanced by the grid stride without a division per tile
        // (three divisions per tile were 8 % of the instructions at D = 64)
        const int qs = tstep / tiles_per_item, rs = tstep - qs * tiles_per_item;
        int pb = tile0 / tiles_per_item, pt = tile0 - pb * tiles_per_item;
        auto advance = [&](int& b_, int& t_) {
            b_ += qs;
            t_ += rs;
            if (t_ >= tiles_per_item) { t_ -= tiles_per_item; ++b_; }
        };
        // shortlist length (lane u < FW holds frame fbase + u) and shortlist (two lanes per frame) of a tile, into registers; with kAhead
        // they are fetched ONE TILE AHEAD (at the top of a tile these loads were the longest single stall of the pass at D = 256)
        auto fetch_lists = [&](int tile, int b, int t_in, int& cnt, uint4& v) {
            cnt = 0;
            v = make_uint4(0u, 0u, 0u, 0u);
            if (tile >= num_tiles) return;
            const int w0 = t_in * TF;
            const int64_t n0 = (int64_t)b * W + w0;
            const int wlim = (int)((W - w0) < TF ? (W - w0) : TF);
            if (lane < FW && fbase + lane < wlim) cnt = idx32 ? kCandFinal : (int)cand_cnt[n0 + fbase + lane];
            if (lane < 2 * FW && fbase + (lane >> 1) < wlim) {
                const int64_t n = n0 + fbase + (lane >> 1);
                if (idx32) { if ((lane & 1) == 0) v.x = (uint32_t)idx32[n]; }
                else v = reinterpret_cast<const uint4*>(cand_idx + (size_t)n * kCandMax)[lane & 1];
            }
        };
        constexpr bool pf = kAhead, foldx2 = kAhead;
        int cnt_nx;
        uint4 v_nx;
        if (pf) fetch_lists(tile0, pb, pt, cnt_nx, v_nx);
        uint32_t it = 0;
        for (int tile = tile0; tile < num_tiles; tile += tstep, ++it) {
            const int buf = (int)(it & 1u);
            const int b = pb, w0 = pt * TF;
            const int64_t n0 = (int64_t)b * W + w0;    // global frame id of the tile's first frame
            const int wlim = (int)((W - w0) < TF ? (W - w0) : TF);
            advance(pb, pt);                           // (pb, pt) now is the position of tile + tstep
            if (!pf) fetch_lists(tile, b, w0 / TF, cnt_nx, v_nx);
            const int cnt_l = cnt_nx;
            if (lane < 2 * FW) reinterpret_cast<uint4*>(sC + (fbase + (lane >> 1)) * kCandMax)[lane & 1] = v_nx;
            if (pf) fetch_lists(tile + tstep, pb, pt, cnt_nx, v_nx);
            const int np_l = (cnt_l != kCandFinal && cnt_l > 1) ? cnt_l : 0;
            int incl_l = np_l;
#pragma unroll
            for (int o = 1; o < FW; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl_l, o);
                if (lane >= o) incl_l += t;
            }
            const int total = __shfl_sync(0xffffffffu, incl_l, FW - 1);
            __syncwarp();                               // sC of this tile is in place
            mbar_wait(smem_u32(&bars->full[buf]), (it >> 1) & 1u);
            unsigned char* box = box0 + buf * BOX_BYTES;
            // The m-th dim of this lane is d = 8 (m / S) + S sl + m % S.  Element (d, frame f): row d at d * 128 bytes, the 16-byte
            // chunk f / 4 at chunk position (f / 4) ^ (d % 8).  xrow(f, c): address of dims with m % S = c; dim m at + (m / S) * 1024.
            auto xrow = [&](int f, int c) -> unsigned char* {
                const unsigned r = (unsigned)(S * sl + c);
                return box + r * 128 + ((((unsigned)f >> 2) ^ r) << 4) + (f & 3) * 4;
            };
            if (total > 0) {
                if (!foldx2) {                          // |x|^2 of the frames that have pairs in a pass of its own
                    const unsigned has_pairs = __ballot_sync(0xffffffffu, np_l > 0);   // bit u: frame fbase + u has pairs
#pragma unroll
                    for (int i = 0; i < ITER; ++i) {
                        if ((has_pairs >> (G * i)) & ((1u << G) - 1u)) {               // warp-uniform
                            const int fi = i * G + g;
                            unsigned char* xp[S];
#pragma unroll
                            for (int c = 0; c < S; ++c) xp[c] = xrow(fbase + fi, c);
                            float x2 = 0.f;
#pragma unroll
                            for (int m = 0; m < M; ++m) {
                                const float x = *reinterpret_cast<const float*>(xp[m % S] + (m / S) * 1024);
                                x2 = __fadd_rn(x2, __fmul_rn(x, x));
                            }
                            x2 = group_sum<LPF>(x2);
                            if (sl == 0) x2s[fbase + fi] = x2;
                        }
                    }
                    __syncwarp();
                }
                const int rounds = (total + G - 1) / G;
                // pair p = G r + g of round r belongs to the frame whose [excl, incl) holds it; inactive groups rescore code 0 of frame 0
                auto pair_of = [&](int r, int& fi, int& kc) {
                    const int p = G * r + g;
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
                auto fetch_row = [&](int kc, float4 (&ev)[T], float& e2c) {
                    const float* er = Ep + (size_t)kc * D + 4 * sl;
#pragma unroll
                    for (int t = 0; t < T; ++t) ev[t] = *reinterpret_cast<const float4*>(er + 4 * LPF * t);
                    e2c = e2[kc];
                };
                auto score_plain = [&](int r, bool act, int fi, int kc, const float4 (&ev)[T], float e2c) {
                    unsigned char* xp[S];
#pragma unroll
                    for (int c = 0; c < S; ++c) xp[c] = xrow(fbase + fi, c);
                    float dot = 0.f;
#pragma unroll
                    for (int t = 0; t < T; ++t) {
                        const float ee[4] = {ev[t].x, ev[t].y, ev[t].z, ev[t].w};
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int m = 4 * t + j;
                            dot = fmaf(*reinterpret_cast<const float*>(xp[m % S] + (m / S) * 1024), ee[j], dot);
                        }
                    }
                    dot = group_sum<LPF>(dot);
                    if (act && sl == 0) pr[G * r + g] = make_float2(ref_distance(x2s[fbase + fi], e2c, dot), __int_as_float(kc));
                };
                auto score_fold = [&](int r, bool act, int fi, int kc, const float4 (&ev)[T], float e2c) {
                    // |x|^2 rides along (sum of rounded squares per lane, then over the group: every pair of a frame gets the same
                    // value) - the latent is being read anyway, a separate pass would read it once more
                    unsigned char* xp[S];
#pragma unroll
                    for (int c = 0; c < S; ++c) xp[c] = xrow(fbase + fi, c);
                    float dot = 0.f, x2 = 0.f;
#pragma unroll
                    for (int t = 0; t < T; ++t) {
                        const float ee[4] = {ev[t].x, ev[t].y, ev[t].z, ev[t].w};
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int m = 4 * t + j;
                            const float x = *reinterpret_cast<const float*>(xp[m % S] + (m / S) * 1024);
                            dot = fmaf(x, ee[j], dot);
                            x2 = __fadd_rn(x2, __fmul_rn(x, x));
                        }
                    }
                    dot = group_sum<LPF>(dot);
                    x2 = group_sum<LPF>(x2);
                    if (act && sl == 0) pr[G * r + g] = make_float2(ref_distance(x2, e2c, dot), __int_as_float(kc));
                };
                auto score = [&](int r, bool act, int fi, int kc, const float4 (&ev)[T], float e2c) {
                    if (foldx2) score_fold(r, act, fi, kc, ev, e2c);
                    else score_plain(r, act, fi, kc, ev, e2c);
                };
                if (G >= 8) {
                    // 8 or 16 pairs per round: a warp's frames rarely need more than one round, and the second row buffer of the
                    // pipelined form below costs the registers of a fifth block per SM
                    for (int r = 0; r < rounds; ++r) {
                        float4 ev[T];
                        float e2c;
                        int fi, kc;
                        const bool act = pair_of(r, fi, kc);
                        fetch_row(kc, ev, e2c);
                        score(r, act, fi, kc, ev, e2c);
                    }
                } else {
                    // rounds are software-pipelined two deep (ping-pong registers)
                    float4 eva[T], evb[T];
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
                }
            }
            __syncwarp();                               // pairres written by other lanes of this warp
            // ---- own frames: final code, codeword gather, straight-through value (in place), SSE, histogram, residual sums
#pragma unroll
            for (int i = 0; i < ITER; ++i) {
                const int fi = i * G + g;
                const int f = fbase + fi;
                const int in_o = __shfl_sync(0xffffffffu, incl_l, fi), np_o = __shfl_sync(0xffffffffu, np_l, fi);
                if (f < wlim) {
                    const int64_t n = n0 + f;
                    int k;
                    if (np_o > 0) {                     // settle among the dealt pairs: torch.argmin order, ties -> lowest index
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
                        k = (int)sC[f * kCandMax];      // the only shortlisted code / the exact search's answer
                        if (sl == 0) n_short += 1;
                    }
                    const float* er = Ep + (size_t)k * D + 4 * sl;
                    float4 qv[T];
#pragma unroll
                    for (int t = 0; t < T; ++t) qv[t] = *reinterpret_cast<const float4*>(er + 4 * LPF * t);
                    if (k != run_k) {                   // a new run of this lane group
                        flush_run();
                        run_k = k;
                        run_n = 0;
                        if (kRun) {
#pragma unroll
                            for (int t = 0; t < T; ++t) run_r[t] = make_float4(0.f, 0.f, 0.f, 0.f);
                        }
                    }
                    run_n += 1;
                    unsigned char* xp[S];
#pragma unroll
                    for (int c = 0; c < S; ++c) xp[c] = xrow(f, c);
                    float* srow = stage + (size_t)(warp * G + g) * D;
                    if (kBulk) {                                   // the previous reduction out of this staging row has read it
                        if (sl == 0) bulk_wait_read0();
                        __syncwarp(gmask);
                    }
                    float fs = 0.f;
#pragma unroll
                    for (int t = 0; t < T; ++t) {
                        const float ee[4] = {qv[t].x, qv[t].y, qv[t].z, qv[t].w};
                        float rr[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int m = 4 * t + j;
                            float* px = reinterpret_cast<float*>(xp[m % S] + (m / S) * 1024);
                            const float x = *px;
                            // r = fl(x - e) is exactly -fl(e - x): the residual sums want r, and the straight-through VALUE
                            // fl(x + fl(e - x)) (vector_quantizer.py:48) equals fl(x - r) bit for bit
                            rr[j] = __fsub_rn(x, ee[j]);
                            fs = fmaf(rr[j], rr[j], fs);
                            if (has_q) *px = __fsub_rn(x, rr[j]);
                        }
                        if (kResid) {
                            if (kRun) {
                                run_r[t].x += rr[0]; run_r[t].y += rr[1]; run_r[t].z += rr[2]; run_r[t].w += rr[3];
                            } else if (kBulk) {
                                *reinterpret_cast<float4*>(srow + 4 * LPF * t + 4 * sl) = make_float4(rr[0], rr[1], rr[2], rr[3]);
                            } else {
                                red_add_v4(resid + (size_t)k * D + 4 * LPF * t + 4 * sl, rr[0], rr[1], rr[2], rr[3]);
                            }
                        }
                    }
                    if (kBulk) {                                   // the frame's residual row -> its code's row of this SM's replica
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                        __syncwarp(gmask);
                        if (sl == 0) { bulk_reduce_add_f32(resid + (size_t)k * D, smem_u32(srow), (uint32_t)D * 4u); bulk_commit(); }
                    }
                    {   // Kahan: sse += fs
                        const float y = fs - sse_c, tt = sse + y;
                        sse_c = (tt - sse) - y;
                        sse = tt;
                    }
                    if (sl == 0) idx_out[n] = (int64_t)k;
                }
            }
            // this warp is through with the box: its writes become visible to the async proxy (TMA store), then it reports
            if (has_q) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&bars->done[buf]));
        }
        flush_run();
        if (kBulk && sl == 0) bulk_wait0();             // shared memory must outlive the reductions that read it
        double t = (double)sse - (double)sse_c;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (lane == 0) red[warp] = t;
        if (n_resc | n_short) {
            atomicAdd(&meta->rescored, (unsigned long long)n_resc);
            atomicAdd(&meta->shortlisted, (unsigned long long)n_short);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int i = 0; i < NWARP; ++i) s += red[i];
        sse_partials[blockIdx.x] = s;
    }
}

template <int J, int LPF>
static cudaError_t launch_t(const CUtensorMap& mz, const CUtensorMap& mq, const float* ep, const float* e2, int64_t W, int tiles_per_item,
                            int num_tiles, const int* idx32, const uint8_t* cand_cnt, const uint16_t* cand_idx, int64_t* idx_out,
                            bool has_q, int* counts, float* resid_rep, int n_rep, size_t rep_stride, double* part, int n_partials, WsMeta* meta,
                            bool read_once, bool bulk, bool ahead, cudaStream_t s) {
    constexpr int D = 32 * J, NWARP = LPF == 2 ? 2 : 4, FW = TF / NWARP, G = 32 / LPF, NTHREADS = 32 * (NWARP + 1);
    constexpr bool kRun = D / LPF <= 16;
    bulk = bulk && resid_rep && !kRun;
    const size_t smem = (size_t)2 * D * TF * 4 + (size_t)NWARP * FW * kPairMax * 8 + TF * 4 + (size_t)TF * kCandMax * 2 + NWARP * 8 + sizeof(Bars) +
                        (bulk ? (size_t)NWARP * G * D * 4 : 0);
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
            if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, NTHREADS, smem)) != cudaSuccess) return e;
            if (per_sm > kTailGridMax / 148) per_sm = kTailGridMax / 148;
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
            cached_blocks = sms * (per_sm < 1 ? 1 : per_sm);
            cached_kernel = reinterpret_cast<const void*>(kernel);
            cached_dev = dev;
        }
        int grid = cached_blocks;
        if (grid > n_partials) grid = n_partials;
        if (grid > num_tiles) grid = num_tiles;
        if (grid < 1) grid = 1;
        kernel<<<(unsigned)grid, NTHREADS, smem, s>>>(mz, mq, ep, e2, W, tiles_per_item, num_tiles, idx32, cand_cnt, cand_idx, idx_out, has_q ? 1 : 0,
                                                      counts, resid_rep, n_rep, rep_stride, part, meta, read_once ? 1 : 0);
        return cudaGetLastError();
    };
    if (ahead) {
        if (!resid_rep) return go(tail3_kernel<J, LPF, false, false, false, true>);
        if constexpr (!kRun) { if (bulk) return go(tail3_kernel<J, LPF, true, false, true, true>); }
        return go(tail3_kernel<J, LPF, true, kRun, false, true>);
    }
    if (!resid_rep) return go(tail3_kernel<J, LPF, false, false, false, false>);
    if constexpr (!kRun) { if (bulk) return go(tail3_kernel<J, LPF, true, false, true, false>); }
    return go(tail3_kernel<J, LPF, true, kRun, false, false>);
}

// sums the (permuted) replicas into the caller's residual sums, un-permuting
__global__ void __launch_bounds__(256) fold_resid_perm_kernel(float* __restrict__ resid, const float* __restrict__ resid_rep, int n_rep,
                                                              size_t rep_stride, int D, int lpf) {
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < rep_stride; i += (size_t)gridDim.x * 256) {
        const size_t k = i / (size_t)D;
        const int d = (int)(i - k * (size_t)D);
        const size_t p = k * (size_t)D + (size_t)tail3_perm_pos(d, lpf);
        float a = resid[i];
        for (int r = 0; r < n_rep; ++r) a += resid_rep[(size_t)r * rep_stride + p];
        resid[i] = a;
    }
}

}  // namespace t3

bool tail3_supports(int D) { return D % 32 == 0 && (D / 32 <= 4 || D == 192 || D == 256); }
// Lanes per frame (decides the permutation of the codebook copy and of the residual replicas).  VQB_TAIL_LPF (experiments) overrides
// where the shape allows it.
int tail3_lpf(int D) {
    int lpf = D >= 96 ? 8 : 4;   // measured: D = 64 / 32 want fewer lanes per frame (per-frame bookkeeping), D = 128 is equal with 4 and 8
    const int v = env_get(ENV_TAIL_LPF, 0);
    if ((v == 2 || v == 4 || v == 8) && D % (4 * v) == 0 && D / v <= 32 && (v != 2 || D <= 64) && (v != 4 || D <= 128)) lpf = v;
    return lpf;
}
// Measured against tail2_kernel (profiles/r03_exp_tail_forms.jsonl, r03_exp_tail3_ahead*.jsonl): D = 256 -10 %, D = 128 -10 %; at D = 64 the two
// are within the +-5 % run-to-run spread of each other (r03_exp_tail_address_alias.jsonl), so every supported D takes this kernel.
bool tail3_preferred(int D) { return tail3_supports(D); }

// `ep`: the permuted fp32 codebook (codebook_prep_kernel, tail3_lpf(D)); `resid_rep`: n_rep zeroed copies of [K, D] in the workspace
// (or null); `resid`: the caller's residual sums, to which the folded replicas are ADDED.
cudaError_t launch_tail3(const float* z, const float* ep, const float* e2, int B, int D, int64_t W, int K, const int* idx32,
                         const uint8_t* cand_cnt, const uint16_t* cand_idx, int64_t* idx_out, float* q_out, int* counts, float* resid,
                         double* part, int n_partials, WsMeta* meta, float* resid_rep, int n_rep, cudaStream_t s) {
    using namespace t3;
    CUtensorMap mz, mq;
    if (make_latent_map(&mz, z, (uint64_t)B, (uint64_t)D, (uint64_t)W, (uint32_t)TF, (uint32_t)D, true) != 0) return cudaErrorInvalidValue;
    if (q_out) { if (make_latent_map(&mq, q_out, (uint64_t)B, (uint64_t)D, (uint64_t)W, (uint32_t)TF, (uint32_t)D, true) != 0) return cudaErrorInvalidValue; }
    else mq = mz;
    const int tiles_per_item = (int)((W + TF - 1) / TF);
    const int64_t tiles64 = (int64_t)B * tiles_per_item;
    if (tiles64 >= (1LL << 30)) return cudaErrorInvalidValue;
    const int num_tiles = (int)tiles64;
    const bool once = latents_read_once((size_t)B * D * W * 4);
    const size_t rep_stride = (size_t)K * D;
    float* rep = resid ? resid_rep : nullptr;
    // bulk reductions where there is no run merging (D >= 192; cfg-3 slice of 2^21 frames: 1.33 ms against 1.48 ms with red.v4);
    // VQB_TAIL_FORM=300 (experiments): red.v4 everywhere
    const bool bulk = env_get(ENV_TAIL_FORM, 3) != 300;
    const int lpf = tail3_lpf(D);
    // shortlists one tile ahead + |x|^2 folded into the pair scoring: measured per shape (profiles/r03_exp_tail3_ahead*.jsonl): a gain
    // at D = 256 (1.34 -> 1.31 - 1.33 ms per 2^21 frames in most runs), none at D = 64, a loss at D = 128 with 8 lanes per frame
    bool ahead = D >= 192;
    if (const int v = env_get(ENV_TAIL_AHEAD, -1); v >= 0) ahead = v != 0;   // experiments
    cudaError_t e = cudaErrorInvalidValue;
#define VQB_T3(J, LPF) e = launch_t<J, LPF>(mz, mq, ep, e2, W, tiles_per_item, num_tiles, idx32, cand_cnt, cand_idx, idx_out, q_out != nullptr, counts, \
                                            rep, n_rep, rep_stride, part, n_partials, meta, once, bulk, ahead, s)
    switch ((D / 32) * 16 + lpf) {
        case 1 * 16 + 2: VQB_T3(1, 2); break;
        case 1 * 16 + 4: VQB_T3(1, 4); break;
        case 1 * 16 + 8: VQB_T3(1, 8); break;
        case 2 * 16 + 2: VQB_T3(2, 2); break;
        case 2 * 16 + 4: VQB_T3(2, 4); break;
        case 2 * 16 + 8: VQB_T3(2, 8); break;
        case 3 * 16 + 4: VQB_T3(3, 4); break;
        case 3 * 16 + 8: VQB_T3(3, 8); break;
        case 4 * 16 + 4: VQB_T3(4, 4); break;
        case 4 * 16 + 8: VQB_T3(4, 8); break;
        case 6 * 16 + 8: VQB_T3(6, 8); break;
        case 8 * 16 + 8: VQB_T3(8, 8); break;
        default: break;
    }
#undef VQB_T3
    note_launch();
    if (e != cudaSuccess || !rep) return e;
    const size_t blocks = (rep_stride + 255) / 256;
    fold_resid_perm_kernel<<<(unsigned)(blocks < 1184 ? blocks : 1184), 256, 0, s>>>(resid, rep, n_rep, rep_stride, D, lpf);
    note_launch();
    return cudaGetLastError();
}

}  // namespace vqb
