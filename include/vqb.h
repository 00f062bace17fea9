/* vqb.h - C ABI of libvqb_b200.so, the B200 (sm_100a) vector-quantiser bottleneck.
 *
 * The reference (deborahdore/multi-source-lms-for-audio) has no FFI layer: its boundary for this path is the Python
 * class `VectorQuantizer` (src/model/components/vector_quantizer.py:6-54), constructed at src/model/vqvae.py:46-48 and
 * called at src/model/vqvae.py:84,92.  This library is what a maintainer binds underneath that class (ctypes stub in
 * INTEGRATION.md); every entry point below names the reference lines it replaces.
 *
 * Conventions
 *  - extern "C", plain pointers and sizes, no C++ or torch types.  `stream` is a cudaStream_t passed as void*.
 *  - All data pointers are DEVICE pointers owned by the caller unless the name ends in `_host`.
 *  - Return value: 0 = OK; negative = argument / capability error detected before any launch (VQB_E_*);
 *    positive = the cudaError_t / ncclResult_t (+1000) of a failed runtime call.  vqb_last_error() returns a
 *    thread-local human-readable message for the last non-zero return on this thread.
 *  - Calls only enqueue work on `stream`; they never synchronise the host and use no hidden streams, so a sequence of
 *    calls is CUDA-graph capturable.  Scalars (losses, perplexity) are written to device memory.  (Exceptions, named
 *    below: vqb_forward_host and the vqb_debug_* readers synchronise.)
 *  - Threading: the data-path calls keep no state between calls and may be issued concurrently from several threads,
 *    devices and streams as long as each call has its own workspace.  NOT re-entrant: vqb_forward_host on one device
 *    (serialised internally), and the vqb_debug_* timing / launch counters (process-global, for single-threaded benches).
 *  - Index arguments are validated on the device: a code outside [0, K) touches no memory; its one-hot row stays zero,
 *    its gathered codeword / gradient is NaN.
 *  - There is no CPU fallback.  A device that is not compute capability 10.x is a hard error (VQB_E_DEVICE).
 *  - Latents are fp32 in the reference's [B, D, W] ("BCW") layout; frame n = b*W + w (vector_quantizer.py:25-29).
 *    The codebook is fp32 [K, D] row-major (nn.Embedding weight, vector_quantizer.py:18).
 */
#ifndef VQB_H_
#define VQB_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VQB_VERSION 200
#if defined(__GNUC__)
#define VQB_API __attribute__((visibility("default")))
#else
#define VQB_API
#endif

/* ---- flags for vqb_forward / vqb_workspace_bytes ------------------------------------------------------------- */
#define VQB_PREC_MASK   0x0F
#define VQB_PREC_FP32   0x00  /* exact: fp32 CUDA-core search in the reference's op order (parity anchor / fallback)    */
#define VQB_PREC_BF16   0x01  /* tcgen05 bf16 shortlist + guard band, fp32 rescoring in the reference's op order        */
#define VQB_PREC_TF32   0x02  /* tcgen05 kind::tf32 shortlist straight from the fp32 operands (no converted copies, tighter
                                 band, half the tensor rate of bf16) + the same fp32 rescoring                           */
#define VQB_WANT_Q      0x10  /* write the straight-through value fl(x + fl(q - x)) to q_bcw_out (v_q.py:42,48,52)      */
#define VQB_WANT_RESID  0x20  /* accumulate per-code residual sums into stats (needed for the codebook gradient)        */

/* ---- error codes ---------------------------------------------------------------------------------------------- */
#define VQB_E_NULL      (-1)  /* a required pointer is NULL                                                             */
#define VQB_E_SHAPE     (-2)  /* unsupported B / D / W / K (see vqb_forward)                                            */
#define VQB_E_ALIGN     (-3)  /* a pointer is not 16-byte aligned                                                       */
#define VQB_E_WORKSPACE (-4)  /* workspace smaller than vqb_workspace_bytes reports                                     */
#define VQB_E_DEVICE    (-5)  /* current device is not sm_100 (no fallback path exists)                                 */
#define VQB_E_FLAGS     (-6)  /* unknown precision / flag combination                                                   */
#define VQB_E_NCCL      (-7)  /* libnccl.so.2 could not be loaded                                                       */

/* statistics buffer layout (floats): [counts[K] | resid[K*D] | SSE | N]                                               */
#define VQB_STATS_LEN(K, D) ((size_t)(K) * ((size_t)(D) + 1) + 2)

VQB_API int         vqb_version(void);
VQB_API const char* vqb_last_error(void);

/* Bytes of scratch vqb_forward needs.  (Replaces the N x K fp32 `distances` and `encodings` temporaries of
 * vector_quantizer.py:32-39, which are never materialised here.)
 * vqb_workspace_bytes:    upper bound for ANY [B, W] with B * W = N (includes a bf16 latent copy that only shapes the
 *                         tensor-core kernel cannot read directly need: W % 4 != 0, short clips).
 * vqb_workspace_bytes_bw: exact for one [B, D, W] shape (16-byte aligned latents, which vqb_forward requires anyway). */
VQB_API int vqb_workspace_bytes(int64_t N, int K, int D, int flags, size_t* bytes_out);
VQB_API int vqb_workspace_bytes_bw(int B, int D, int64_t W, int K, int flags, size_t* bytes_out);

/* Forward of the bottleneck: vector_quantizer.py:25-52 minus the dense one-hot.
 *   z_bcw      [B, D, W] fp32        codebook  [K, D] fp32
 *   idx_out    [N] int64             nearest code per frame (vector_quantizer.py:37; lowest index on ties, NaN wins)
 *   q_bcw_out  [B, D, W] fp32        straight-through value, only with VQB_WANT_Q (may be NULL otherwise)
 *   stats_out  [VQB_STATS_LEN] fp32  counts, residual sums sum_{n in k}(x_n - e_k) (only with VQB_WANT_RESID, else
 *                                    left untouched), SSE = sum (q - x)^2, N.  This is the buffer a multi-GPU job
 *                                    all-reduces (vqb_allreduce_stats) before vqb_finalize / vqb_backward.
 * Supported: 1 <= K <= 65536, D % 16 == 0, 16 <= D <= 512, N < 2^31.  */
VQB_API int vqb_forward(const float* z_bcw, const float* codebook, int B, int D, int64_t W, int K, int flags,
                int64_t* idx_out, float* q_bcw_out, float* stats_out,
                void* workspace, size_t workspace_bytes, void* stream);

/* losses_out[0] = embedding_loss = SSE / (N D)        (vector_quantizer.py:46)
 * losses_out[1] = commitment_loss = beta * SSE / (N D) (vector_quantizer.py:45)
 * losses_out[2] = perplexity = exp(-sum p log(p + 1e-10)), p = counts / N   (vector_quantizer.py:49-50) */
VQB_API int vqb_finalize(const float* stats, int K, int D, float beta, float* losses_out, void* stream);

/* Backward (what autograd derives from vector_quantizer.py:42-52):
 *   dX[b,:,w] = Gq[b,:,w] + g_c * beta * 2 (x - q) / (N_local D)          N_local = B*W of this call
 *   dE[k,:]   = - g_e * (2 / (N_stats D)) * resid[k,:]                     N_stats = stats[N slot] (global after allreduce)
 * g_e_dev / g_c_dev are DEVICE pointers to the upstream scalar gradients of embedding_loss / commitment_loss (NULL = 0),
 * Gq_bcw the upstream gradient of `quantized` (NULL = 0).  dX_bcw or dE may be NULL to skip that output. */
VQB_API int vqb_backward(const float* z_bcw, const float* codebook, const int64_t* idx, const float* stats,
                 const float* Gq_bcw, const float* g_e_dev, const float* g_c_dev, float beta,
                 int B, int D, int64_t W, int K, float* dX_bcw, float* dE, void* stream);

/* Dense one-hot `encodings` [N, K] fp32 (vector_quantizer.py:38-39), for callers that really want it. */
VQB_API int vqb_onehot(const int64_t* idx, int64_t N, int K, float* encodings_out, void* stream);

/* Codeword gather ("de-quantise"): out[b,:,w] = codebook[idx[b*W+w], :]  (the one-hot matmul of
 * vector_quantizer.py:42 and bert.py:75-78 as a gather, written straight in BCW). */
VQB_API int vqb_gather(const float* codebook, const int64_t* idx, int B, int D, int64_t W, int K, float* out_bcw, void* stream);

/* Index-stream export for the BERT stage (bert.py:50-69): idx [B*L] -> tokens [B, n_win, window] int64 (last window
 * padded with pad_id) and attention mask [B, n_win, window] fp32 (0 on padding), n_win = ceil(L / window). */
VQB_API int vqb_window_indices(const int64_t* idx, int B, int64_t L, int window, int64_t pad_id,
                       int64_t* tokens_out, float* mask_out, void* stream);

/* EXTENSION, not in the reference (its codebook is trained by Adam, vqvae.py:168-171): exponential-moving-average codebook
 * update from the same statistics buffer (needs VQB_WANT_RESID).  cluster_size is [K + 1] (slot K receives the running
 * total), embed_sum is [K, D]; both are state owned by the caller, codebook is updated in place. */
VQB_API int vqb_ema_update(const float* stats, float* codebook, float* cluster_size, float* embed_sum, int K, int D,
                           float decay, float eps, void* stream);

/* Host-buffer convenience used for end-to-end timing and by non-torch callers: copies z (pinned or pageable HOST
 * memory) to the device in chunks overlapped with compute, runs vqb_forward per chunk of whole batch items and copies
 * the indices, with VQB_WANT_Q the straight-through output (what a call of the reference returns, vector_quantizer.py:54;
 * full duplex with the uploads) and the statistics back.  `comm` (may be NULL) all-reduces the statistics over the ranks
 * before they are copied back.  SYNCHRONOUS, on private streams of the CURRENT device; its device scratch is cached per
 * device (vqb_host_release frees the current device's) and calls on the same device serialise on a mutex. */
VQB_API int vqb_forward_host(const float* z_bcw_host, const float* codebook_host, int B, int D, int64_t W, int K, int flags,
                     int64_t* idx_out_host, float* q_bcw_out_host, float* stats_out_host, int chunk_batches, void* comm);
VQB_API int vqb_host_release(void);

/* ---- multi-GPU: the one exchange on the path (implicit DDP all-reduce of codebook.weight.grad in the reference,
 *      configs/trainer/default.yaml:9-10) --------------------------------------------------------------------------- */
#define VQB_UNIQUE_ID_BYTES 128
VQB_API int vqb_comm_unique_id(void* id_out_host /* 128 bytes */);
VQB_API int vqb_comm_init(const void* id_host, int rank, int world, void** comm_out);
VQB_API int vqb_allreduce_stats(void* comm, float* stats, size_t n_floats, void* stream);   /* in-place ncclSum */
VQB_API int vqb_comm_destroy(void* comm);

/* ---- diagnostics (used by tests and bench.py) ------------------------------------------------------------------- */
/* counters_out_host[0] = frames whose shortlist held > 1 code (rescored), [1] = frames sent to the exact fallback,
 * [2] = total shortlisted codes; valid after the stream of the last vqb_forward using `workspace` has been synchronised. */
VQB_API int vqb_debug_counters(const void* workspace, int64_t* counters_out_host);
/* Raw tcgen05 scores for testing the tensor-core tile: out[n, k] = |e_k|^2 - 2 bf16(x_n).bf16(e_k), n < N, k < K. */
VQB_API int vqb_debug_tc_scores(const float* z_bcw, const float* codebook, int B, int D, int64_t W, int K, int flags,
                        float* scores_out, void* workspace, size_t workspace_bytes, void* stream);

/* Number of this library's kernel launches since the last reset (bench.py's `gpu_launches`). */
VQB_API long long vqb_debug_launch_count(int reset);
/* The experiment switches of DESIGN.md section 8 (VQB_* environment variables) are read once and honoured only when
 * VQB_EXPERIMENTS=1 is set; this re-reads them (tests flip them inside one process). */
VQB_API int vqb_debug_reload_env(void);
/* Layout contract of tail3_kernel (host-only, no device needed): the lanes-per-frame it uses at this D, and the position inside a
 * permuted codebook / residual row that holds dim d for a given lanes-per-frame (negative: bad arguments). */
VQB_API int vqb_debug_tail3_lanes(int D);
VQB_API int vqb_debug_tail3_perm_pos(int d, int lanes_per_frame);
/* CUDA-event timing of the stages of vqb_forward on its launching stream: enable, run, then read the summed duration and
 * launch count of a stage (bench.py's roofline leg).  At most 2048 stage launches are recorded per enable.
 * vqb_debug_kernel_time_ms reads VQB_STAGE_SEARCH, the dominant kernel (tc_search_kernel). */
#define VQB_STAGE_SEARCH   0   /* tc_search_kernel (bf16) or exact_search_kernel over all frames (fp32) */
#define VQB_STAGE_PREP     1   /* memsets, codebook_prep, latent_prep (unfused operand preparation only) */
#define VQB_STAGE_FALLBACK 2   /* exact search of the frames whose shortlist overflowed + commit / list tail */
#define VQB_STAGE_TAIL     3   /* rescoring, gather, straight-through value, statistics */
#define VQB_STAGE_PACK     4   /* pack_stats */
#define VQB_N_STAGES       5
VQB_API int vqb_debug_kernel_timing(int enable);
VQB_API int vqb_debug_kernel_time_ms(double* total_ms, int* launches);
VQB_API int vqb_debug_stage_time_ms(int stage, double* total_ms, int* launches);

#ifdef __cplusplus
}
#endif
#endif /* VQB_H_ */
