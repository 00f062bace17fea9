// Host-buffer entry point: the reference-facing call for callers whose latents live in HOST memory (and the path the
// end-to-end benchmark times).  Latents stream to the device in chunks of whole batch items over one copy stream while
// the previous chunk is quantised on a compute stream; indices stream back on a third.  Device scratch is cached.
#include "vqb_internal.h"

namespace vqb {

int forward_impl(const float* z, const float* codebook, int B, int D, int64_t W, int K, int flags, int64_t* idx_out,
                 float* q_out, float* stats_out, void* workspace, size_t ws_bytes, cudaStream_t s, bool accumulate,
                 float* scores_dbg);

struct HostCtx {
    cudaStream_t s_in = nullptr, s_run = nullptr, s_out = nullptr;
    cudaEvent_t in_done[2] = {nullptr, nullptr}, run_done[2] = {nullptr, nullptr}, out_done[2] = {nullptr, nullptr};
    float* z[2] = {nullptr, nullptr};
    int64_t* idx[2] = {nullptr, nullptr};
    float *codebook = nullptr, *stats = nullptr;
    void* ws = nullptr;
    size_t z_bytes = 0, idx_bytes = 0, cb_bytes = 0, stats_bytes = 0, ws_bytes = 0;
    bool init = false;
};
static HostCtx g_ctx;

#define VQB_CUDA(call, what)                                   \
    do {                                                       \
        cudaError_t e__ = (call);                              \
        if (e__ != cudaSuccess) return cuda_fail(e__, what);   \
    } while (0)

static int grow(void** p, size_t* have, size_t need) {
    if (*have >= need) return 0;
    if (*p) VQB_CUDA(cudaFree(*p), "cudaFree");
    *p = nullptr;
    *have = 0;
    VQB_CUDA(cudaMalloc(p, need), "cudaMalloc");
    *have = need;
    return 0;
}

static int ctx_init() {
    if (g_ctx.init) return 0;
    VQB_CUDA(cudaStreamCreateWithFlags(&g_ctx.s_in, cudaStreamNonBlocking), "stream");
    VQB_CUDA(cudaStreamCreateWithFlags(&g_ctx.s_run, cudaStreamNonBlocking), "stream");
    VQB_CUDA(cudaStreamCreateWithFlags(&g_ctx.s_out, cudaStreamNonBlocking), "stream");
    for (int i = 0; i < 2; ++i) {
        VQB_CUDA(cudaEventCreateWithFlags(&g_ctx.in_done[i], cudaEventDisableTiming), "event");
        VQB_CUDA(cudaEventCreateWithFlags(&g_ctx.run_done[i], cudaEventDisableTiming), "event");
        VQB_CUDA(cudaEventCreateWithFlags(&g_ctx.out_done[i], cudaEventDisableTiming), "event");
    }
    g_ctx.init = true;
    return 0;
}

}  // namespace vqb

using namespace vqb;

extern "C" {

int vqb_forward_host(const float* z_host, const float* codebook_host, int B, int D, int64_t W, int K, int flags,
                     int64_t* idx_out_host, float* stats_out_host, int chunk_batches) {
    if (!z_host || !codebook_host || !idx_out_host) { set_error("vqb_forward_host: NULL pointer argument"); return VQB_E_NULL; }
    if (B < 1 || W < 1) { set_error("vqb_forward_host: bad B/W"); return VQB_E_SHAPE; }
    if (flags & VQB_WANT_Q) { set_error("vqb_forward_host: VQB_WANT_Q is not supported on the host-buffer path"); return VQB_E_FLAGS; }
    int rc;
    if ((rc = ctx_init()) != 0) return rc;
    if (chunk_batches < 1) {   // default: ~256 MiB of latents per chunk
        const int64_t per_item = (int64_t)D * W * 4;
        chunk_batches = (int)((256LL << 20) / (per_item > 0 ? per_item : 1));
        if (chunk_batches < 1) chunk_batches = 1;
    }
    if (chunk_batches > B) chunk_batches = B;
    const int64_t n_chunk = (int64_t)chunk_batches * W;
    size_t ws_need = 0;
    if ((rc = vqb_workspace_bytes(n_chunk, K, D, flags, &ws_need)) != 0) return rc;
    const size_t z_need = (size_t)chunk_batches * D * W * 4, idx_need = (size_t)n_chunk * 8;
    const size_t cb_need = (size_t)K * D * 4, st_need = VQB_STATS_LEN(K, D) * 4;
    for (int i = 0; i < 2; ++i) {
        size_t zb = g_ctx.z_bytes, ib = g_ctx.idx_bytes;
        if ((rc = grow(reinterpret_cast<void**>(&g_ctx.z[i]), &zb, z_need)) != 0) return rc;
        if ((rc = grow(reinterpret_cast<void**>(&g_ctx.idx[i]), &ib, idx_need)) != 0) return rc;
        if (i == 1) { g_ctx.z_bytes = zb; g_ctx.idx_bytes = ib; }
    }
    if ((rc = grow(reinterpret_cast<void**>(&g_ctx.codebook), &g_ctx.cb_bytes, cb_need)) != 0) return rc;
    if ((rc = grow(reinterpret_cast<void**>(&g_ctx.stats), &g_ctx.stats_bytes, st_need)) != 0) return rc;
    if ((rc = grow(&g_ctx.ws, &g_ctx.ws_bytes, ws_need)) != 0) return rc;

    VQB_CUDA(cudaMemcpyAsync(g_ctx.codebook, codebook_host, cb_need, cudaMemcpyHostToDevice, g_ctx.s_run), "H2D codebook");
    int it = 0;
    for (int b0 = 0; b0 < B; b0 += chunk_batches, ++it) {
        const int nb = (B - b0 < chunk_batches) ? (B - b0) : chunk_batches;
        const int buf = it & 1;
        const size_t zbytes = (size_t)nb * D * W * 4;
        if (it >= 2) {   // the buffers of chunk it-2 must have been consumed / drained
            VQB_CUDA(cudaStreamWaitEvent(g_ctx.s_in, g_ctx.run_done[buf], 0), "wait");
            VQB_CUDA(cudaStreamWaitEvent(g_ctx.s_run, g_ctx.out_done[buf], 0), "wait");
        }
        VQB_CUDA(cudaMemcpyAsync(g_ctx.z[buf], z_host + (size_t)b0 * D * W, zbytes, cudaMemcpyHostToDevice, g_ctx.s_in), "H2D z");
        VQB_CUDA(cudaEventRecord(g_ctx.in_done[buf], g_ctx.s_in), "record");
        VQB_CUDA(cudaStreamWaitEvent(g_ctx.s_run, g_ctx.in_done[buf], 0), "wait");
        rc = forward_impl(g_ctx.z[buf], g_ctx.codebook, nb, D, W, K, flags, g_ctx.idx[buf], nullptr, g_ctx.stats, g_ctx.ws,
                          g_ctx.ws_bytes, g_ctx.s_run, it > 0, nullptr);
        if (rc != 0) return rc;
        VQB_CUDA(cudaEventRecord(g_ctx.run_done[buf], g_ctx.s_run), "record");
        VQB_CUDA(cudaStreamWaitEvent(g_ctx.s_out, g_ctx.run_done[buf], 0), "wait");
        VQB_CUDA(cudaMemcpyAsync(idx_out_host + (size_t)b0 * W, g_ctx.idx[buf], (size_t)nb * W * 8, cudaMemcpyDeviceToHost, g_ctx.s_out),
                 "D2H idx");
        VQB_CUDA(cudaEventRecord(g_ctx.out_done[buf], g_ctx.s_out), "record");
    }
    if (stats_out_host)
        VQB_CUDA(cudaMemcpyAsync(stats_out_host, g_ctx.stats, st_need, cudaMemcpyDeviceToHost, g_ctx.s_run), "D2H stats");
    VQB_CUDA(cudaStreamSynchronize(g_ctx.s_run), "sync run");
    VQB_CUDA(cudaStreamSynchronize(g_ctx.s_out), "sync out");
    return 0;
}

int vqb_host_release(void) {
    if (!g_ctx.init) return 0;
    cudaDeviceSynchronize();
    for (int i = 0; i < 2; ++i) {
        if (g_ctx.z[i]) cudaFree(g_ctx.z[i]);
        if (g_ctx.idx[i]) cudaFree(g_ctx.idx[i]);
        cudaEventDestroy(g_ctx.in_done[i]);
        cudaEventDestroy(g_ctx.run_done[i]);
        cudaEventDestroy(g_ctx.out_done[i]);
    }
    if (g_ctx.codebook) cudaFree(g_ctx.codebook);
    if (g_ctx.stats) cudaFree(g_ctx.stats);
    if (g_ctx.ws) cudaFree(g_ctx.ws);
    cudaStreamDestroy(g_ctx.s_in);
    cudaStreamDestroy(g_ctx.s_run);
    cudaStreamDestroy(g_ctx.s_out);
    g_ctx = HostCtx{};
    return 0;
}

}  // extern "C"
