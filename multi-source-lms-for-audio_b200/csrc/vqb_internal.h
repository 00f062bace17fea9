// Internal declarations shared by the translation units of libvqb_b200.so (not part of the C ABI).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stddef.h>
#include <stdint.h>
#include "../../include/vqb.h"

namespace vqb {

constexpr int kCandMax   = 16;    // uint16 slots per frame in cand_idx (32 B rows); the tensor-core search fills <= 12
constexpr int kCandFinal = 255;   // cand_cnt marker: slot 0 holds the final code (written by the exact fallback search)
constexpr int kTileCodes = 256;   // codes per tcgen05 N tile; |e|^2 and the bf16 codebook copy are padded to this
constexpr int kTileRows  = 128;   // frames per tcgen05 M tile; the bf16 latent copy is padded to this

// Small device-resident header at the start of every workspace.
struct WsMeta {
    unsigned int       emax2_bits;      // max_k |e_k|^2 as ordered uint bits (atomicMax on non-negative floats)
    int                cb_nonfinite;    // != 0 if any |e_k|^2 is inf/NaN -> every frame takes the exact path
    int                fallback_count;  // frames appended to fallback_rows by the tensor-core search
    unsigned int       etmax2_bits;     // max_k |bf16(e_k)|^2           } rounding-residual norms that make the
    unsigned int       demax2_bits;     // max_k |e_k - bf16(e_k)|^2     } shortlist guard band a rigorous bound
    int                pad0[3];
    unsigned long long rescored;        // diagnostics (vqb_debug_counters)
    unsigned long long shortlisted;
    unsigned long long fallback_total;
    unsigned long long pad1;
};

struct WsLayout {
    size_t meta, e2, counts, sse_partials, resid_rep, ep, idx32, cand_cnt, cand_idx, fallback_rows, best64, x2, eb, eh, xb, ev, total;
    int    k_pad;
    int64_t n_pad;
    int    n_partials;
};
WsLayout ws_layout(int64_t N, int K, int D, int flags, bool need_xb, int ev_ctas);
WsLayout ws_layout_for(const float* z, int B, int D, int64_t W, int K, int flags);   // z may be null (= 16-byte aligned)

constexpr int kTcMaxCtas = 160;         // persistent grid bound of the tensor-core search (event scratch is sized for it)
size_t tc_event_scratch_bytes(int ctas);
constexpr int kTailGridMax = 148 * 8;   // persistent grid of the fused tail kernel (sse partial slots)
constexpr int kResidReplicasMax = 8;     // residual-sum replicas against same-address atomic serialisation (see pick_resid_replica)
constexpr int kFallbackTailGrid = 148;  // fused-tail mode: sse slots [kTcMaxCtas, kTcMaxCtas + kFallbackTailGrid) belong to fallback_tail_kernel

// CUDA-event pair around one stage of vqb_forward when vqb_debug_kernel_timing(1) is on (else begin returns null, end is a no-op)
void* stage_timing_begin(cudaStream_t s, int stage);
void  stage_timing_end(void* slot, cudaStream_t s);

// Experiment switches.  They are read ONCE (first use, or vqb_debug_reload_env()) and honoured only when
// VQB_EXPERIMENTS=1 is set: a stray variable can never change production behaviour and the launch path never calls getenv.
enum EnvKey {
    ENV_TC_MODE, ENV_TC_CLUSTER, ENV_TC_FUSE, ENV_TC_STAGES, ENV_TC_ASLOTS, ENV_TC_EVSM, ENV_TC_EHSLOTS, ENV_TC_TAIL, ENV_TMA_PROMO,
    ENV_TILE_LDG, ENV_RESID_REPLICAS, ENV_L2_ONCE, ENV_TAIL_VARIANT, ENV_TAIL_TMA, ENV_TAIL_EXACT, ENV_DX_TILES, ENV_TC_EPI, ENV_TAIL_FORM,
    ENV_TAIL_LPF, ENV_TAIL_AHEAD, ENV_TC_SLEEP,
    ENV_COUNT
};
int  env_get(EnvKey key, int unset_value);         // value of the switch, or unset_value when it is not set / not enabled
void env_reload();

void set_error(const char* fmt, ...);
void note_launch(int n = 1);                       // counts this library's kernel launches (vqb_debug_launch_count)
int  cuda_fail(cudaError_t e, const char* what);   // records message, returns (int)e

#ifdef __CUDACC__
// torch.argmin ordering (vector_quantizer.py:37): NaN beats everything, otherwise smaller distance, ties -> lower index.
__device__ __forceinline__ bool better(float d, int i, float bd, int bi) {
    if (bi < 0) return true;
    const bool dn = isnan(d), bn = isnan(bd);
    if (dn) return !bn || i < bi;
    if (bn) return false;
    return d < bd || (d == bd && i < bi);
}
// The reference's distance in its association order (vector_quantizer.py:32-33):
//   fl(|x|^2 + fl(|e|^2 - fl(2 * dot)));  2*dot is exact, so fma(-2, dot, e2) is the same single rounding.
__device__ __forceinline__ float ref_distance(float x2, float e2, float dot) {
    return __fadd_rn(x2, __fmaf_rn(-2.0f, dot, e2));
}
__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void st_stream(float* p, float v) {   // written once, not read again by this kernel: keep it out of L1
    asm volatile("st.global.L1::no_allocate.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}
// Residual-sum replicas (see launch_tail): the blocks of different SMs add into one of n_rep copies, chosen by SM id
__device__ __forceinline__ float* pick_resid_replica(float* resid, float* resid_rep, int n_rep, size_t rep_stride) {
    if (!resid || n_rep <= 1) return resid;
    unsigned int smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    const unsigned int r = smid % (unsigned int)n_rep;
    return r ? resid_rep + (size_t)(r - 1) * rep_stride : resid;
}
#endif

// What the tensor-core search needs to finish frames itself (fused tail, see vqb_tc.cu): everything the stand-alone
// tail kernel would have read or written.  q_out / resid may be null.
struct TailArgs {
    const float* z;           // [B, D, W] fp32 latents
    const float* codebook;    // [K, D] fp32
    const float* e2;          // [K_pad] |e_k|^2
    int64_t*     idx_out;     // [N]
    float*       q_out;       // [B, D, W] straight-through value, or null
    int*         counts;      // [K]
    float*       resid;       // [K, D] sum of (x - e_k) over the frames of code k, or null
    double*      sse_partials;   // one slot per CTA of the search kernel
};

// ---- launchers (vqb_kernels.cu) -------------------------------------------------------------------------------
// tf32: the rounding-residual norms of the guard band are those of the tf32 operand (low 13 mantissa bits dropped); no bf16 copy
// ep != nullptr (D % 32 == 0): also the permuted fp32 copy tail3_kernel gathers from (vqb_tail3.cu)
cudaError_t launch_codebook_prep(const float* codebook, int K, int K_pad, int D, float* e2, __nv_bfloat16* eb,
                                 __nv_bfloat16* eh, WsMeta* meta, cudaStream_t s, bool tf32 = false, float* ep = nullptr, int ep_lpf = 8);
cudaError_t launch_latent_prep_bf16(const float* z, int B, int D, int64_t W, int64_t N_pad, __nv_bfloat16* xb, float* band,
                                    const WsMeta* meta, cudaStream_t s);
// rows == nullptr: all N frames -> idx32[n]; else the frames listed in rows[0..*row_count) -> cand_cnt/cand_idx (count 1)
cudaError_t launch_exact_search(const float* z, const float* codebook, const float* e2, int B, int D, int64_t W, int K,
                                const int* rows, const int* row_count, int* idx32, uint8_t* cand_cnt, uint16_t* cand_idx,
                                unsigned long long* best64, cudaStream_t s);
cudaError_t launch_tail(const float* z, const float* codebook, const float* e2, int B, int D, int64_t W, int K,
                        const int* idx32, const uint8_t* cand_cnt, const uint16_t* cand_idx, int64_t* idx_out, float* q_out,
                        int* counts, float* resid, float* sse_partials, int n_partials, WsMeta* meta, float* resid_rep, cudaStream_t s,
                        const float* ep = nullptr);   // ep: permuted codebook copy (enables tail3_kernel)
// third form of the tail (vqb_tail3.cu): swizzled TMA box read and written in place, TMA store of `quantized`, no block barriers;
// gathers from the permuted codebook copy `ep`, accumulates the residual sums in n_rep permuted workspace replicas and ADDS their
// un-permuted sum to `resid`
bool tail3_supports(int D);
bool tail3_preferred(int D);   // the shapes where it measured faster than tail2_kernel
int  tail3_lpf(int D);         // lanes per frame tail3_kernel uses at this D: decides the permutation below
// Position, inside a codebook / residual row, of dim d in the layout tail3_kernel reads with `lpf` lanes per frame: lane sl owns the
// dims whose d % 8 lies in [S sl, S sl + S), S = 8 / lpf; its m-th dim (m = S (d / 8) + d % 8 % S) sits at (m / 4) 4 lpf + 4 sl + m % 4
__host__ __device__ __forceinline__ int tail3_perm_pos(int d, int lpf) {
    const int S = 8 / lpf, r = d & 7, sl = r / S, m = (d >> 3) * S + r % S;
    return (m >> 2) * 4 * lpf + sl * 4 + (m & 3);
}
cudaError_t launch_tail3(const float* z, const float* ep, const float* e2, int B, int D, int64_t W, int K, const int* idx32,
                         const uint8_t* cand_cnt, const uint16_t* cand_idx, int64_t* idx_out, float* q_out, int* counts, float* resid,
                         double* part, int n_partials, WsMeta* meta, float* resid_rep, int n_rep, cudaStream_t s);
// second form of the tail (vqb_tail2.cu): D = 32 J compile-time, dealt rescoring, 16-frame tiles for large D, run-length atomics
bool tail2_supports(int D);
cudaError_t launch_tail2(const float* z, const float* codebook, const float* e2, int B, int D, int64_t W, int K, const int* idx32,
                         const uint8_t* cand_cnt, const uint16_t* cand_idx, int64_t* idx_out, float* q_out, int* counts, float* resid,
                         double* part, int n_partials, WsMeta* meta, float* resid_rep, int n_rep, size_t rep_stride, int form,
                         cudaStream_t s);
int resid_replicas(int K, int D);   // copies of the residual sums the tail kernels spread their atomics over (workspace holds kResidReplicasMax: tail3_kernel keeps all of them there)
// fused-tail mode: finishes the (rare) frames the exact fallback search decided - one warp per frame of the list
cudaError_t launch_fallback_tail(const float* z, const float* codebook, int B, int D, int64_t W, int K, const int* rows,
                                 const int* row_count, const unsigned long long* best64, int64_t* idx_out, float* q_out, int* counts,
                                 float* resid, double* sse_partials, cudaStream_t s);
cudaError_t launch_pack_stats(const int* counts, const float* sse_partials, int n_partials, int64_t N, int K, int D,
                              float* stats, bool accumulate, cudaStream_t s);
cudaError_t launch_finalize(const float* stats, int K, int D, float beta, float* losses, cudaStream_t s);
cudaError_t launch_backward_dx(const float* z, const float* codebook, const int64_t* idx, const float* Gq, const float* g_c,
                               float beta, int B, int D, int64_t W, int K, float* dX, cudaStream_t s);
cudaError_t launch_backward_de(const float* stats, const float* g_e, int K, int D, float* dE, cudaStream_t s);
cudaError_t launch_ema_update(const float* stats, float* codebook, float* cluster_size, float* embed_sum, int K, int D, float decay,
                              float eps, cudaStream_t s);
cudaError_t launch_onehot(const int64_t* idx, int64_t N, int K, float* out, cudaStream_t s);
cudaError_t launch_gather(const float* codebook, const int64_t* idx, int B, int D, int64_t W, int K, float* out, cudaStream_t s);
cudaError_t launch_window(const int64_t* idx, int B, int64_t L, int window, int64_t pad_id, int64_t* tokens, float* mask,
                          cudaStream_t s);

// 3-D TMA view of a [B, D, W] fp32 tensor (vqb_tc.cu): box = box_frames x box_dims of one batch item; swizzle128 needs
// box_frames == 32 (128-byte rows) and a 1024-byte aligned destination
bool latents_read_once(size_t latent_bytes);   // stream the latents with an L2 evict-first policy? (vqb_kernels.cu)
int make_latent_map(CUtensorMap* map, const float* z, uint64_t B, uint64_t D, uint64_t W, uint32_t box_frames, uint32_t box_dims,
                    bool swizzle128 = false, bool atom32 = false);   // atom32: 128B swizzle with 32-byte atoms (tf32 MN-major operand)

// ---- tensor-core search (vqb_tc.cu) ---------------------------------------------------------------------------
// Shortlist per frame from bf16 tcgen05 scores: cand_cnt/cand_idx, overflow frames appended to fallback_rows.
bool tc_can_fuse(const float* z, int B, int D, int64_t W, int prec = VQB_PREC_BF16);   // can the tensor-core kernel read the fp32 [B, D, W] latents itself?
bool tc_fused_tail_fits(int B, int D, int64_t W);            // ... and its shared memory leaves room for the tile pipeline
bool tc_fused_tail_enabled();                               // VQB_TC_TAIL=0 keeps the stand-alone tail kernel (experiments)
// z_fused != nullptr: fused operand preparation (xb / band unused); else xb / band from latent_prep_bf16_kernel
int launch_tc_search(const float* z_fused, int B, int64_t W, const __nv_bfloat16* xb, const __nv_bfloat16* eb, const __nv_bfloat16* eh,
                     const float* band, int64_t N, int64_t N_pad,
                     int K, int K_pad, int D, uint8_t* cand_cnt, uint16_t* cand_idx, int* fallback_rows, WsMeta* meta,
                     unsigned long long* best64, float* scores_dbg, void* ev_scratch, const TailArgs* tail, const float* codebook_f32,
                     cudaStream_t s);   // codebook_f32 != nullptr: kind::tf32 from the fp32 operands (needs z_fused, no fused tail)

}  // namespace vqb
