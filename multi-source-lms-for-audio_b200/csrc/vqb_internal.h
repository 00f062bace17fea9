// Internal declarations shared by the translation units of libvqb_b200.so (not part of the C ABI).
#pragma once
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
    size_t meta, e2, counts, sse_partials, idx32, cand_cnt, cand_idx, fallback_rows, best64, x2, eb, eh, xb, ev, total;
    int    k_pad;
    int64_t n_pad;
    int    n_partials;
};
WsLayout ws_layout(int64_t N, int K, int D, int flags);

constexpr int kTcMaxCtas = 160;         // persistent grid bound of the tensor-core search (event scratch is sized for it)
size_t tc_event_scratch_bytes();
constexpr int kTailGridMax = 148 * 8;   // persistent grid of the fused tail kernel (sse partial slots)

void set_error(const char* fmt, ...);
void note_launch(int n = 1);                       // counts this library's kernel launches (vqb_debug_launch_count)
int  cuda_fail(cudaError_t e, const char* what);   // records message, returns (int)e

// ---- launchers (vqb_kernels.cu) -------------------------------------------------------------------------------
cudaError_t launch_codebook_prep(const float* codebook, int K, int K_pad, int D, float* e2, __nv_bfloat16* eb,
                                 __nv_bfloat16* eh, WsMeta* meta, cudaStream_t s);
cudaError_t launch_latent_prep_bf16(const float* z, int B, int D, int64_t W, int64_t N_pad, __nv_bfloat16* xb, float* band,
                                    const WsMeta* meta, cudaStream_t s);
// rows == nullptr: all N frames -> idx32[n]; else the frames listed in rows[0..*row_count) -> cand_cnt/cand_idx (count 1)
cudaError_t launch_exact_search(const float* z, const float* codebook, const float* e2, int B, int D, int64_t W, int K,
                                const int* rows, const int* row_count, int* idx32, uint8_t* cand_cnt, uint16_t* cand_idx,
                                unsigned long long* best64, cudaStream_t s);
cudaError_t launch_tail(const float* z, const float* codebook, const float* e2, int B, int D, int64_t W, int K,
                        const int* idx32, const uint8_t* cand_cnt, const uint16_t* cand_idx, int64_t* idx_out, float* q_out,
                        int* counts, float* resid, float* sse_partials, int n_partials, WsMeta* meta, cudaStream_t s);
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

// ---- tensor-core search (vqb_tc.cu) ---------------------------------------------------------------------------
// Shortlist per frame from bf16 tcgen05 scores: cand_cnt/cand_idx, overflow frames appended to fallback_rows.
bool tc_can_fuse(const float* z, int B, int D, int64_t W);   // can the tensor-core kernel read the fp32 [B, D, W] latents itself?
// z_fused != nullptr: fused operand preparation (xb / band unused); else xb / band from latent_prep_bf16_kernel
int launch_tc_search(const float* z_fused, int B, int64_t W, const __nv_bfloat16* xb, const __nv_bfloat16* eb, const __nv_bfloat16* eh,
                     const float* band, int64_t N, int64_t N_pad,
                     int K, int K_pad, int D, uint8_t* cand_cnt, uint16_t* cand_idx, int* fallback_rows, WsMeta* meta,
                     unsigned long long* best64, float* scores_dbg, void* ev_scratch, cudaStream_t s);

}  // namespace vqb
