// C ABI of libvqb_b200.so (see include/vqb.h).  Argument checking, workspace carving and the launch sequence of
// the forward / backward of the VQ bottleneck.  No host synchronisation, no hidden streams, no CPU fallback.
#include "vqb_internal.h"

#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <mutex>

namespace vqb {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
int cuda_fail(cudaError_t e, const char* what) {
    set_error("%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
    return (int)e;
}

// ---- experiment switches: one table, filled once -----------------------------------------------------------------------
static const char* const kEnvNames[ENV_COUNT] = {
    "VQB_TC_MODE", "VQB_TC_CLUSTER", "VQB_TC_FUSE", "VQB_TC_STAGES", "VQB_TC_ASLOTS", "VQB_TC_EVSM", "VQB_TC_EHSLOTS", "VQB_TC_TAIL",
    "VQB_TMA_PROMO", "VQB_TILE_LDG", "VQB_RESID_REPLICAS", "VQB_L2_ONCE", "VQB_TAIL_VARIANT", "VQB_TAIL_TMA", "VQB_TAIL_EXACT",
    "VQB_DX_TILES", "VQB_TC_EPI", "VQB_TAIL_FORM", "VQB_TAIL_LPF", "VQB_TAIL_AHEAD", "VQB_TC_SLEEP"};
static int g_env_val[ENV_COUNT];
static bool g_env_set[ENV_COUNT];
static std::atomic<bool> g_env_loaded{false};
static std::mutex g_env_mutex;

void env_reload() {
    std::lock_guard<std::mutex> lock(g_env_mutex);
    const char* gate = getenv("VQB_EXPERIMENTS");
    const bool on = gate && gate[0] == '1';
    for (int i = 0; i < ENV_COUNT; ++i) {
        const char* v = on ? getenv(kEnvNames[i]) : nullptr;
        g_env_set[i] = v != nullptr && v[0] != 0;
        g_env_val[i] = g_env_set[i] ? atoi(v) : 0;
    }
    g_env_loaded.store(true, std::memory_order_release);
}
int env_get(EnvKey key, int unset_value) {
    if (!g_env_loaded.load(std::memory_order_acquire)) env_reload();
    return g_env_set[key] ? g_env_val[key] : unset_value;
}

static std::atomic<long long> g_launches{0};
void note_launch(int n) { g_launches += n; }

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// need_xb: reserve the bf16 latent copy (only the unfused operand preparation touches it: 2 N D bytes, 8.6 GB at BASELINE
// config 3); ev_ctas: CTAs of the search kernel the event scratch must serve (<= kTcMaxCtas).
WsLayout ws_layout(int64_t N, int K, int D, int flags, bool need_xb, int ev_ctas) {
    WsLayout L{};
    const int prec = flags & VQB_PREC_MASK;
    L.k_pad = (K + kTileCodes - 1) / kTileCodes * kTileCodes;
    L.n_pad = (N + kTileRows - 1) / kTileRows * kTileRows;
    L.n_partials = kTailGridMax;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 1024); return o; };
    L.meta = take(sizeof(WsMeta));
    L.e2 = take((size_t)L.k_pad * 4);
    L.counts = take((size_t)K * 4);
    L.sse_partials = take((size_t)L.n_partials * 8);
    L.resid_rep = (flags & VQB_WANT_RESID) ? take((size_t)kResidReplicasMax * K * D * 4) : 0;
    L.ep = (D % 32 == 0) ? take((size_t)K * D * 4) : 0;   // permuted fp32 codebook copy for tail3_kernel
    if (prec == VQB_PREC_FP32) {
        L.idx32 = take((size_t)N * 4);
        L.cand_cnt = L.cand_idx = L.fallback_rows = L.best64 = L.x2 = L.eb = L.eh = L.xb = L.ev = 0;
    } else {
        L.idx32 = 0;
        L.cand_cnt = take((size_t)L.n_pad);
        L.cand_idx = take((size_t)L.n_pad * kCandMax * 2);
        L.fallback_rows = take((size_t)L.n_pad * 4);
        L.best64 = take((size_t)L.n_pad * 8);
        L.x2 = take((size_t)L.n_pad * 4);
        L.eb = take((size_t)L.k_pad * D * 2);
        L.eh = take((size_t)L.k_pad * 16);
        L.xb = need_xb ? take((size_t)L.n_pad * D * 2) : 0;
        L.ev = take(tc_event_scratch_bytes(ev_ctas));
    }
    L.total = off;
    return L;
}

static int check_device() {
    static thread_local int ok_device = -1;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return cuda_fail(e, "cudaGetDevice");
    if (dev == ok_device) return 0;
    int major = 0;
    e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    if (e != cudaSuccess) return cuda_fail(e, "cudaDeviceGetAttribute");
    if (major != 10) {
        set_error("device %d has compute capability %d.x; libvqb_b200 only runs on sm_100 (B200) and has no fallback", dev, major);
        return VQB_E_DEVICE;
    }
    ok_device = dev;
    return 0;
}

static int check_shape(int B, int D, int64_t W, int K) {
    if (B < 1 || W < 1 || K < 1 || K > 65536 || D < 16 || D > 512 || (D % 16) != 0) {
        set_error("unsupported shape B=%d D=%d W=%lld K=%d (need 1<=K<=65536, 16<=D<=512, D%%16==0)", B, D, (long long)W, K);
        return VQB_E_SHAPE;
    }
    if ((int64_t)B * W >= (1LL << 31) - kTileRows) {
        set_error("N = B*W = %lld does not fit the 31-bit frame index; split the batch", (long long)((int64_t)B * W));
        return VQB_E_SHAPE;
    }
    return 0;
}

#define VQB_CUDA(call, what)                                   \
    do {                                                       \
        cudaError_t e__ = (call);                              \
        if (e__ != cudaSuccess) return cuda_fail(e__, what);   \
    } while (0)

// Exact layout for one [B, D, W] call.  z == nullptr: a 16-byte aligned tensor (vqb_forward insists on that anyway).
WsLayout ws_layout_for(const float* z, int B, int D, int64_t W, int K, int flags) {
    const int prec = flags & VQB_PREC_MASK;
    const int64_t N = (int64_t)B * W;
    if (prec == VQB_PREC_FP32) return ws_layout(N, K, D, flags, false, 0);
    bool fuse = tc_can_fuse(z, B, D, W, prec);
    if (prec == VQB_PREC_TF32 && !fuse) fuse = tc_can_fuse(z, B, D, W, VQB_PREC_BF16);   // falls back to the bf16 shortlist
    const int64_t tiles = fuse ? (int64_t)B * ((W + kTileRows - 1) / kTileRows) : (N + kTileRows - 1) / kTileRows;
    return ws_layout(N, K, D, flags, !fuse, (int)(tiles < kTcMaxCtas ? tiles : kTcMaxCtas));
}

// accumulate: keep counts / resid / SSE / N from earlier chunks (used by the host-buffer path)
int forward_impl(const float* z, const float* codebook, int B, int D, int64_t W, int K, int flags, int64_t* idx_out,
                 float* q_out, float* stats_out, void* workspace, size_t ws_bytes, cudaStream_t s, bool accumulate,
                 float* scores_dbg) {
    int rc;
    if ((rc = check_device()) != 0) return rc;
    if (!z || !codebook || !workspace || (!scores_dbg && (!idx_out || !stats_out))) {
        set_error("vqb_forward: NULL pointer argument");
        return VQB_E_NULL;
    }
    if ((rc = check_shape(B, D, W, K)) != 0) return rc;
    int prec = flags & VQB_PREC_MASK;
    if (prec != VQB_PREC_FP32 && prec != VQB_PREC_BF16 && prec != VQB_PREC_TF32) {
        set_error("vqb_forward: unknown precision %d (VQB_PREC_FP32, VQB_PREC_BF16 or VQB_PREC_TF32)", prec);
        return VQB_E_FLAGS;
    }
    // tf32 reads the fp32 latents in place; shapes the tensor-core kernel cannot read that way (W % 4 != 0, short clips, D > 256)
    // take the bf16 shortlist instead - the fp32 rescoring makes the result the same either way
    if (prec == VQB_PREC_TF32 && !tc_can_fuse(z, B, D, W, VQB_PREC_TF32)) prec = VQB_PREC_BF16;
    flags = (flags & ~VQB_PREC_MASK) | prec;
    if ((flags & VQB_WANT_Q) && !q_out) {
        set_error("vqb_forward: VQB_WANT_Q set but q_bcw_out is NULL");
        return VQB_E_NULL;
    }
    if (((uintptr_t)z | (uintptr_t)codebook | (uintptr_t)workspace) & 15) {
        set_error("vqb_forward: z, codebook and workspace must be 16-byte aligned");
        return VQB_E_ALIGN;
    }
    const int64_t N = (int64_t)B * W;
    const WsLayout L = ws_layout_for(z, B, D, W, K, flags);
    if (ws_bytes < L.total) {
        set_error("vqb_forward: workspace has %zu bytes, %zu needed", ws_bytes, L.total);
        return VQB_E_WORKSPACE;
    }
    char* ws = static_cast<char*>(workspace);
    WsMeta* meta = reinterpret_cast<WsMeta*>(ws + L.meta);
    float* e2 = reinterpret_cast<float*>(ws + L.e2);
    int* counts = reinterpret_cast<int*>(ws + L.counts);
    float* part = reinterpret_cast<float*>(ws + L.sse_partials);
    float* resid = (flags & VQB_WANT_RESID) ? stats_out + K : nullptr;
    float* resid_rep = resid ? reinterpret_cast<float*>(ws + L.resid_rep) : nullptr;
    float* ep = (D % 32 == 0) ? reinterpret_cast<float*>(ws + L.ep) : nullptr;

    void* tprep = stage_timing_begin(s, VQB_STAGE_PREP);
    // ONE memset over the workspace regions that start a call at zero - meta | (e2, rewritten by codebook_prep) | counts | SSE partials |
    // residual-sum replicas lie back to back (ws_layout) - instead of four: small batches are launch-bound
    const int n_rep = resid ? resid_replicas(K, D) : 0;
    const size_t zero_end = resid_rep ? L.resid_rep + (size_t)n_rep * K * D * 4 : L.sse_partials + (size_t)L.n_partials * sizeof(double);
    if (!accumulate) {
        VQB_CUDA(cudaMemsetAsync(ws + L.meta, 0, zero_end - L.meta, s), "memset workspace head");
        if (resid) VQB_CUDA(cudaMemsetAsync(resid, 0, (size_t)K * D * 4, s), "memset resid");
    } else {                                      // the host-buffer path accumulates counts / statistics over its chunks
        VQB_CUDA(cudaMemsetAsync(&meta->fallback_count, 0, sizeof(int), s), "memset fallback_count");
        VQB_CUDA(cudaMemsetAsync(ws + L.sse_partials, 0, zero_end - L.sse_partials, s), "memset partials + replicas");
    }

    if (prec == VQB_PREC_FP32) {
        int* idx32 = reinterpret_cast<int*>(ws + L.idx32);
        VQB_CUDA(launch_codebook_prep(codebook, K, L.k_pad, D, e2, nullptr, nullptr, meta, s, false, ep, tail3_lpf(D)), "codebook_prep");
        stage_timing_end(tprep, s);
        void* t0 = stage_timing_begin(s, VQB_STAGE_SEARCH);
        VQB_CUDA(launch_exact_search(z, codebook, e2, B, D, W, K, nullptr, nullptr, idx32, nullptr, nullptr, nullptr, s), "exact_search");
        stage_timing_end(t0, s);
        void* t1 = stage_timing_begin(s, VQB_STAGE_TAIL);
        VQB_CUDA(launch_tail(z, codebook, e2, B, D, W, K, idx32, nullptr, nullptr, idx_out, (flags & VQB_WANT_Q) ? q_out : nullptr,
                             counts, resid, part, L.n_partials, meta, resid_rep, s, ep), "tail");
        stage_timing_end(t1, s);
    } else {
        uint8_t* cand_cnt = reinterpret_cast<uint8_t*>(ws + L.cand_cnt);
        uint16_t* cand_idx = reinterpret_cast<uint16_t*>(ws + L.cand_idx);
        int* fb_rows = reinterpret_cast<int*>(ws + L.fallback_rows);
        unsigned long long* best64 = reinterpret_cast<unsigned long long*>(ws + L.best64);
        float* x2 = reinterpret_cast<float*>(ws + L.x2);
        __nv_bfloat16* eb = reinterpret_cast<__nv_bfloat16*>(ws + L.eb);
        __nv_bfloat16* xb = reinterpret_cast<__nv_bfloat16*>(ws + L.xb);
        __nv_bfloat16* eh = reinterpret_cast<__nv_bfloat16*>(ws + L.eh);
        const bool tf32 = prec == VQB_PREC_TF32;
        VQB_CUDA(launch_codebook_prep(codebook, K, L.k_pad, D, e2, tf32 ? nullptr : eb, eh, meta, s, tf32, ep, tail3_lpf(D)), "codebook_prep");
        const bool fuse = tc_can_fuse(z, B, D, W, prec);
        // fused tail: the search kernel finishes the frames itself (index, quantized, statistics); only the frames it
        // sends to the exact search are finished by a small list kernel.  Otherwise the stand-alone tail kernel runs.
        const bool fused_tail = fuse && !tf32 && !scores_dbg && tc_fused_tail_enabled() && (reinterpret_cast<uintptr_t>(codebook) & 31) == 0 &&
                                tc_fused_tail_fits(B, D, W);
        double* part_d = reinterpret_cast<double*>(part);
        TailArgs targs{z, codebook, e2, idx_out, (flags & VQB_WANT_Q) ? q_out : nullptr, counts, resid, part_d};
        if (!fuse) VQB_CUDA(launch_latent_prep_bf16(z, B, D, W, L.n_pad, xb, x2, meta, s), "latent_prep");
        stage_timing_end(tprep, s);
        rc = launch_tc_search(fuse ? z : nullptr, B, W, xb, eb, eh, x2, N, L.n_pad, K, L.k_pad, D, cand_cnt, cand_idx, fb_rows, meta, best64,
                              scores_dbg, ws + L.ev, fused_tail ? &targs : nullptr, tf32 ? codebook : nullptr, s);
        if (rc != 0) return rc;
        if (scores_dbg) return 0;
        void* tfb = stage_timing_begin(s, VQB_STAGE_FALLBACK);
        VQB_CUDA(launch_exact_search(z, codebook, e2, B, D, W, K, fb_rows, &meta->fallback_count, nullptr, cand_cnt, cand_idx, best64, s),
                 "exact_search(fallback)");
        if (fused_tail) {
            VQB_CUDA(launch_fallback_tail(z, codebook, B, D, W, K, fb_rows, &meta->fallback_count, best64, idx_out, targs.q_out, counts, resid,
                                          part_d + kTcMaxCtas, s), "fallback_tail");
            stage_timing_end(tfb, s);
        } else {
            stage_timing_end(tfb, s);
            void* tt = stage_timing_begin(s, VQB_STAGE_TAIL);
            VQB_CUDA(launch_tail(z, codebook, e2, B, D, W, K, nullptr, cand_cnt, cand_idx, idx_out, (flags & VQB_WANT_Q) ? q_out : nullptr,
                                 counts, resid, part, L.n_partials, meta, resid_rep, s, ep), "tail");
            stage_timing_end(tt, s);
        }
    }
    void* tpk = stage_timing_begin(s, VQB_STAGE_PACK);
    VQB_CUDA(launch_pack_stats(counts, part, L.n_partials, N, K, D, stats_out, accumulate, s), "pack_stats");
    stage_timing_end(tpk, s);
    return 0;
}

}  // namespace vqb

using namespace vqb;

extern "C" {

int vqb_version(void) { return VQB_VERSION; }
const char* vqb_last_error(void) { return g_err; }

int vqb_workspace_bytes(int64_t N, int K, int D, int flags, size_t* bytes_out) {
    if (!bytes_out) { set_error("vqb_workspace_bytes: NULL output"); return VQB_E_NULL; }
    if (N < 1 || K < 1 || K > 65536 || D < 16 || D > 512 || D % 16) {
        set_error("vqb_workspace_bytes: unsupported N=%lld K=%d D=%d", (long long)N, K, D);
        return VQB_E_SHAPE;
    }
    const int prec = flags & VQB_PREC_MASK;
    if (prec != VQB_PREC_FP32 && prec != VQB_PREC_BF16 && prec != VQB_PREC_TF32) { set_error("vqb_workspace_bytes: unknown precision %d", prec); return VQB_E_FLAGS; }
    *bytes_out = ws_layout(N, K, D, flags, true, kTcMaxCtas).total;   // upper bound over every [B, W] with B * W = N
    return 0;
}

int vqb_workspace_bytes_bw(int B, int D, int64_t W, int K, int flags, size_t* bytes_out) {
    if (!bytes_out) { set_error("vqb_workspace_bytes_bw: NULL output"); return VQB_E_NULL; }
    int rc;
    if ((rc = check_shape(B, D, W, K)) != 0) return rc;
    const int prec = flags & VQB_PREC_MASK;
    if (prec != VQB_PREC_FP32 && prec != VQB_PREC_BF16 && prec != VQB_PREC_TF32) { set_error("vqb_workspace_bytes_bw: unknown precision %d", prec); return VQB_E_FLAGS; }
    *bytes_out = ws_layout_for(nullptr, B, D, W, K, flags).total;
    return 0;
}

int vqb_forward(const float* z_bcw, const float* codebook, int B, int D, int64_t W, int K, int flags, int64_t* idx_out,
                float* q_bcw_out, float* stats_out, void* workspace, size_t workspace_bytes, void* stream) {
    return forward_impl(z_bcw, codebook, B, D, W, K, flags, idx_out, q_bcw_out, stats_out, workspace, workspace_bytes,
                        static_cast<cudaStream_t>(stream), false, nullptr);
}

int vqb_finalize(const float* stats, int K, int D, float beta, float* losses_out, void* stream) {
    int rc;
    if ((rc = check_device()) != 0) return rc;
    if (!stats || !losses_out) { set_error("vqb_finalize: NULL pointer argument"); return VQB_E_NULL; }
    if (K < 1 || D < 1) { set_error("vqb_finalize: bad K/D"); return VQB_E_SHAPE; }
    VQB_CUDA(launch_finalize(stats, K, D, beta, losses_out, static_cast<cudaStream_t>(stream)), "finalize");
    return 0;
}

int vqb_backward(const float* z_bcw, const float* codebook, const int64_t* idx, const float* stats, const float* Gq_bcw,
                 const float* g_e_dev, const float* g_c_dev, float beta, int B, int D, int64_t W, int K, float* dX_bcw,
                 float* dE, void* stream) {
    int rc;
    if ((rc = check_device()) != 0) return rc;
    if ((rc = check_shape(B, D, W, K)) != 0) return rc;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (dX_bcw) {
        if (!z_bcw || !codebook || !idx) { set_error("vqb_backward: dX needs z, codebook and idx"); return VQB_E_NULL; }
        VQB_CUDA(launch_backward_dx(z_bcw, codebook, idx, Gq_bcw, g_c_dev, beta, B, D, W, K, dX_bcw, s), "backward_dx");
    }
    if (dE) {
        if (!stats) { set_error("vqb_backward: dE needs stats"); return VQB_E_NULL; }
        VQB_CUDA(launch_backward_de(stats, g_e_dev, K, D, dE, s), "backward_de");
    }
    return 0;
}

int vqb_ema_update(const float* stats, float* codebook, float* cluster_size, float* embed_sum, int K, int D, float decay, float eps,
                   void* stream) {
    int rc;
    if ((rc = check_device()) != 0) return rc;
    if (!stats || !codebook || !cluster_size || !embed_sum) { set_error("vqb_ema_update: NULL pointer argument"); return VQB_E_NULL; }
    if (K < 1 || D < 1 || !(decay >= 0.f && decay <= 1.f)) { set_error("vqb_ema_update: bad K/D/decay"); return VQB_E_SHAPE; }
    VQB_CUDA(launch_ema_update(stats, codebook, cluster_size, embed_sum, K, D, decay, eps, static_cast<cudaStream_t>(stream)), "ema_update");
    return 0;
}

int vqb_onehot(const int64_t* idx, int64_t N, int K, float* encodings_out, void* stream) {
    int rc;
    if ((rc = check_device()) != 0) return rc;
    if (!idx || !encodings_out) { set_error("vqb_onehot: NULL pointer argument"); return VQB_E_NULL; }
    if (N < 1 || K < 1) { set_error("vqb_onehot: bad N/K"); return VQB_E_SHAPE; }
    VQB_CUDA(launch_onehot(idx, N, K, encodings_out, static_cast<cudaStream_t>(stream)), "onehot");
    return 0;
}

int vqb_gather(const float* codebook, const int64_t* idx, int B, int D, int64_t W, int K, float* out_bcw, void* stream) {
    int rc;
    if ((rc = check_device()) != 0) return rc;
    if (!codebook || !idx || !out_bcw) { set_error("vqb_gather: NULL pointer argument"); return VQB_E_NULL; }
    if ((rc = check_shape(B, D, W, K)) != 0) return rc;
    VQB_CUDA(launch_gather(codebook, idx, B, D, W, K, out_bcw, static_cast<cudaStream_t>(stream)), "gather");
    return 0;
}

int vqb_window_indices(const int64_t* idx, int B, int64_t L, int window, int64_t pad_id, int64_t* tokens_out, float* mask_out,
                       void* stream) {
    int rc;
    if ((rc = check_device()) != 0) return rc;
    if (!idx || !tokens_out || !mask_out) { set_error("vqb_window_indices: NULL pointer argument"); return VQB_E_NULL; }
    if (B < 1 || L < 1 || window < 1) { set_error("vqb_window_indices: bad B/L/window"); return VQB_E_SHAPE; }
    VQB_CUDA(launch_window(idx, B, L, window, pad_id, tokens_out, mask_out, static_cast<cudaStream_t>(stream)), "window");
    return 0;
}

long long vqb_debug_launch_count(int reset) {
    return reset ? g_launches.exchange(0) : g_launches.load();
}

int vqb_debug_reload_env(void) {
    env_reload();
    return 0;
}

int vqb_debug_tail3_lanes(int D) { return (D >= 32 && D % 32 == 0) ? tail3_lpf(D) : VQB_E_SHAPE; }
int vqb_debug_tail3_perm_pos(int d, int lanes_per_frame) {
    if (d < 0 || (lanes_per_frame != 2 && lanes_per_frame != 4 && lanes_per_frame != 8)) return VQB_E_SHAPE;
    return tail3_perm_pos(d, lanes_per_frame);
}

int vqb_debug_counters(const void* workspace, int64_t* counters_out_host) {
    if (!workspace || !counters_out_host) { set_error("vqb_debug_counters: NULL pointer argument"); return VQB_E_NULL; }
    WsMeta m;
    VQB_CUDA(cudaMemcpy(&m, workspace, sizeof(WsMeta), cudaMemcpyDeviceToHost), "memcpy meta");
    counters_out_host[0] = (int64_t)m.rescored;
    counters_out_host[1] = (int64_t)m.fallback_total;
    counters_out_host[2] = (int64_t)m.shortlisted;
    return 0;
}

int vqb_debug_tc_scores(const float* z_bcw, const float* codebook, int B, int D, int64_t W, int K, int flags, float* scores_out,
                        void* workspace, size_t workspace_bytes, void* stream) {
    if (!scores_out) { set_error("vqb_debug_tc_scores: NULL output"); return VQB_E_NULL; }
    const int prec = (flags & VQB_PREC_MASK) == VQB_PREC_TF32 ? VQB_PREC_TF32 : VQB_PREC_BF16;
    return forward_impl(z_bcw, codebook, B, D, W, K, (flags & ~VQB_PREC_MASK) | prec, nullptr, nullptr, nullptr, workspace,
                        workspace_bytes, static_cast<cudaStream_t>(stream), false, scores_out);
}

}  // extern "C"
