// Host-buffer entry point: the reference-facing call for callers whose latents live in HOST memory (and the path the
// end-to-end benchmark times).  Latents stream to the device in chunks of whole batch items over one copy stream while
// the previous chunk is quantised on a compute stream; indices and (with VQB_WANT_Q) the straight-through output stream
// back on a third, full duplex with the next chunk's upload.  Device scratch is cached PER DEVICE and guarded by a
// mutex: concurrent calls on one device serialise, calls on different devices run side by side.
#include "vqb_internal.h"

#include <mutex>

namespace vqb {

int forward_impl(const float* z, const float* codebook, int B, int D, int64_t W, int K, int flags, int64_t* idx_out,
                 float* q_out, float* stats_out, void* workspace, size_t ws_bytes, cudaStream_t s, bool accumulate,
                 float* scores_dbg);

constexpr int kMaxDevices = 64;

struct HostCtx {
    std::mutex mu;
    cudaStream_t s_in = nullptr, s_run = nullptr, s_out = nullptr;
    cudaEvent_t in_done[2] = {nullptr, nullptr}, run_done[2] = {nullptr, nullptr}, out_done[2] = {nullptr, nullptr};
    void* z[2] = {nullptr, nullptr};
    void* q[2] = {nullptr, nullptr};
    void* idx[2] = {nullptr, nullptr};
    void *codebook = nullptr, *stats = nullptr, *ws = nullptr;
    size_t z_bytes[2] = {0, 0}, q_bytes[2] = {0, 0}, idx_bytes[2] = {0, 0}, cb_bytes = 0, stats_bytes = 0, ws_bytes = 0;
    bool init = false;
};
static HostCtx g_ctx[kMaxDevices];

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

static int ctx_init(HostCtx& c) {
    if (c.init) return 0;
    VQB_CUDA(cudaStreamCreateWithFlags(&c.s_in, cudaStreamNonBlocking), "stream");
    VQB_CUDA(cudaStreamCreateWithFlags(&c.s_run, cudaStreamNonBlocking), "stream");
    VQB_CUDA(cudaStreamCreateWithFlags(&c.s_out, cudaStreamNonBlocking), "stream");
    for (int i = 0; i < 2; ++i) {
        VQB_CUDA(cudaEventCreateWithFlags(&c.in_done[i], cudaEventDisableTiming), "event");
        VQB_CUDA(cudaEventCreateWithFlags(&c.run_done[i], cudaEventDisableTiming), "event");
        VQB_CUDA(cudaEventCreateWithFlags(&c.out_done[i], cudaEventDisableTiming), "event");
    }
    c.init = true;
    return 0;
}

static void ctx_release(HostCtx& c) {
    if (!c.init) return;
    cudaStreamSynchronize(c.s_in);
    cudaStreamSynchronize(c.s_run);
    cudaStreamSynchronize(c.s_out);
    for (int i = 0; i < 2; ++i) {
        if (c.z[i]) cudaFree(c.z[i]);
        if (c.q[i]) cudaFree(c.q[i]);
        if (c.idx[i]) cudaFree(c.idx[i]);
        c.z[i] = c.q[i] = c.idx[i] = nullptr;
        c.z_bytes[i] = c.q_bytes[i] = c.idx_bytes[i] = 0;
        cudaEventDestroy(c.in_done[i]);
        cudaEventDestroy(c.run_done[i]);
        cudaEventDestroy(c.out_done[i]);
    }
    if (c.codebook) cudaFree(c.codebook);
    if (c.stats) cudaFree(c.stats);
    if (c.ws) cudaFree(c.ws);
    c.codebook = c.stats = c.ws = nullptr;
    c.cb_bytes = c.stats_bytes = c.ws_bytes = 0;
    cudaStreamDestroy(c.s_in);
    cudaStreamDestroy(c.s_run);
    cudaStreamDestroy(c.s_out);
    c.init = false;
}

// the pipeline proper; on any error the caller drains the three streams before it returns
static int run_chunks(HostCtx& c, const float* z_host, const float* codebook_host, int B, int D, int64_t W, int K, int flags,
                      int64_t* idx_out_host, float* q_out_host, float* stats_out_host, int chunk_batches, void* comm) {
    int rc;
    const int64_t n_chunk = (int64_t)chunk_batches * W;
    size_t ws_need = 0;
    // upper bound over the chunk shapes of this call (the last chunk may be shorter and plan differently)
    if ((rc = vqb_workspace_bytes(n_chunk, K, D, flags, &ws_need)) != 0) return rc;
    const size_t z_need = (size_t)chunk_batches * D * W * 4, idx_need = (size_t)n_chunk * 8;
    const size_t cb_need = (size_t)K * D * 4, st_need = VQB_STATS_LEN(K, D) * 4;
    for (int i = 0; i < 2; ++i) {
        if ((rc = grow(&c.z[i], &c.z_bytes[i], z_need)) != 0) return rc;
        if ((rc = grow(&c.idx[i], &c.idx_bytes[i], idx_need)) != 0) return rc;
        if (q_out_host && (rc = grow(&c.q[i], &c.q_bytes[i], z_need)) != 0) return rc;
    }
    if ((rc = grow(&c.codebook, &c.cb_bytes, cb_need)) != 0) return rc;
    if ((rc = grow(&c.stats, &c.stats_bytes, st_need)) != 0) return rc;
    if ((rc = grow(&c.ws, &c.ws_bytes, ws_need)) != 0) return rc;
    float* stats = static_cast<float*>(c.stats);

    VQB_CUDA(cudaMemcpyAsync(c.codebook, codebook_host, cb_need, cudaMemcpyHostToDevice, c.s_run), "H2D codebook");
    int it = 0;
    for (int b0 = 0; b0 < B; b0 += chunk_batches, ++it) {
        const int nb = (B - b0 < chunk_batches) ? (B - b0) : chunk_batches;
        const int buf = it & 1;
        const size_t zbytes = (size_t)nb * D * W * 4;
        if (it >= 2) {   // the buffers of chunk it-2 must have been consumed / drained
            VQB_CUDA(cudaStreamWaitEvent(c.s_in, c.run_done[buf], 0), "wait");
            VQB_CUDA(cudaStreamWaitEvent(c.s_run, c.out_done[buf], 0), "wait");
        }
        VQB_CUDA(cudaMemcpyAsync(c.z[buf], z_host + (size_t)b0 * D * W, zbytes, cudaMemcpyHostToDevice, c.s_in), "H2D z");
        VQB_CUDA(cudaEventRecord(c.in_done[buf], c.s_in), "record");
        VQB_CUDA(cudaStreamWaitEvent(c.s_run, c.in_done[buf], 0), "wait");
        rc = forward_impl(static_cast<const float*>(c.z[buf]), static_cast<const float*>(c.codebook), nb, D, W, K, flags,
                          static_cast<int64_t*>(c.idx[buf]), q_out_host ? static_cast<float*>(c.q[buf]) : nullptr, stats, c.ws, c.ws_bytes,
                          c.s_run, it > 0, nullptr);
        if (rc != 0) return rc;
        VQB_CUDA(cudaEventRecord(c.run_done[buf], c.s_run), "record");
        VQB_CUDA(cudaStreamWaitEvent(c.s_out, c.run_done[buf], 0), "wait");
        VQB_CUDA(cudaMemcpyAsync(idx_out_host + (size_t)b0 * W, c.idx[buf], (size_t)nb * W * 8, cudaMemcpyDeviceToHost, c.s_out), "D2H idx");
        if (q_out_host)
            VQB_CUDA(cudaMemcpyAsync(q_out_host + (size_t)b0 * D * W, c.q[buf], zbytes, cudaMemcpyDeviceToHost, c.s_out), "D2H quantized");
        VQB_CUDA(cudaEventRecord(c.out_done[buf], c.s_out), "record");
    }
    if (comm && (rc = vqb_allreduce_stats(comm, stats, VQB_STATS_LEN(K, D), c.s_run)) != 0) return rc;
    if (stats_out_host) VQB_CUDA(cudaMemcpyAsync(stats_out_host, stats, st_need, cudaMemcpyDeviceToHost, c.s_run), "D2H stats");
    return 0;
}

}  // namespace vqb

using namespace vqb;

extern "C" {

int vqb_forward_host(const float* z_host, const float* codebook_host, int B, int D, int64_t W, int K, int flags,
                     int64_t* idx_out_host, float* q_out_host, float* stats_out_host, int chunk_batches, void* comm) {
    if (!z_host || !codebook_host || !idx_out_host) { set_error("vqb_forward_host: NULL pointer argument"); return VQB_E_NULL; }
    if (B < 1 || W < 1) { set_error("vqb_forward_host: bad B/W"); return VQB_E_SHAPE; }
    if ((flags & VQB_WANT_Q) && !q_out_host) { set_error("vqb_forward_host: VQB_WANT_Q set but q_out_host is NULL"); return VQB_E_NULL; }
    if (!(flags & VQB_WANT_Q)) q_out_host = nullptr;
    int dev = 0;
    VQB_CUDA(cudaGetDevice(&dev), "cudaGetDevice");
    if (dev < 0 || dev >= kMaxDevices) { set_error("vqb_forward_host: device ordinal %d out of range", dev); return VQB_E_DEVICE; }
    HostCtx& c = g_ctx[dev];
    std::lock_guard<std::mutex> lock(c.mu);
    int rc;
    if ((rc = ctx_init(c)) != 0) return rc;
    if (chunk_batches < 1) {   // default: ~256 MiB of latents per chunk
        const int64_t per_item = (int64_t)D * W * 4;
        chunk_batches = (int)((256LL << 20) / (per_item > 0 ? per_item : 1));
        if (chunk_batches < 1) chunk_batches = 1;
    }
    if (chunk_batches > B) chunk_batches = B;
    rc = run_chunks(c, z_host, codebook_host, B, D, W, K, flags, idx_out_host, q_out_host, stats_out_host, chunk_batches, comm);
    // success or not: nothing of this call may still be in flight on the private streams when it returns
    cudaError_t e0 = cudaStreamSynchronize(c.s_in), e1 = cudaStreamSynchronize(c.s_run), e2 = cudaStreamSynchronize(c.s_out);
    if (rc != 0) return rc;
    if (e0 != cudaSuccess) return cuda_fail(e0, "sync copy-in stream");
    if (e1 != cudaSuccess) return cuda_fail(e1, "sync compute stream");
    if (e2 != cudaSuccess) return cuda_fail(e2, "sync copy-out stream");
    return 0;
}

int vqb_host_release(void) {   // frees the cached scratch of the CURRENT device
    int dev = 0;
    VQB_CUDA(cudaGetDevice(&dev), "cudaGetDevice");
    if (dev < 0 || dev >= kMaxDevices) return 0;
    HostCtx& c = g_ctx[dev];
    std::lock_guard<std::mutex> lock(c.mu);
    ctx_release(c);
    return 0;
}

}  // extern "C"
