// PTX wrappers shared by the translation units of libvqb_b200.so: mbarrier, elect.sync and TMA basics (sm_100a).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace vqb {
namespace ptx {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Bounded wait: a pipeline bug must become a trapped launch, never a hung GPU.  try_wait carries a suspend-time hint, so
// a waiting warp sleeps in hardware (and is woken by the completing arrive) instead of burning issue slots that the
// single-thread producer / MMA loops on the same scheduler need.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    long long t0 = 0;
    for (uint32_t it = 0;; ++it) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity), "r"(0x989680u)
            : "memory");
        if (ok) return;
        if ((it & 63u) == 63u) {
            const long long now = clock64();
            if (t0 == 0) t0 = now;
            else if (now - t0 > 8000000000LL) __trap();   // seconds: only a broken pipeline gets here
        }
    }
}
// The same wait for roles that are far off the critical path (the TMA producer waiting for a free stage, the MMA issuer waiting for
// the epilogue, converters waiting for a free operand slot): between two looks the warp sleeps `ns` nanoseconds.  On this part a
// failed try_wait comes back after a few tens of cycles whatever its suspend hint says, so a waiting helper warp spins - at K = 1024,
// D = 64 those loops were 23 % of ALL executed instructions of the search kernel (ncu), issued on the same schedulers as the
// issue-bound epilogue warps.
__device__ __forceinline__ void mbar_wait_sleep(uint32_t bar, uint32_t parity, uint32_t ns) {
    uint32_t ok = 0;
    long long t0 = 0;
    for (uint32_t it = 0;; ++it) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity), "r"(0x989680u)
            : "memory");
        if (ok) return;
        if (ns) __nanosleep(ns);
        if ((it & 63u) == 63u) {
            const long long now = clock64();
            if (t0 == 0) t0 = now;
            else if (now - t0 > 8000000000LL) __trap();   // seconds: only a broken pipeline gets here
        }
    }
}
// One non-blocking look at a barrier phase (acquire on success): issued EARLY, so that its latency hides behind other work and the
// blocking wait can be skipped when the phase has long completed (a try_wait on a completed phase still costs ~270 cycles here)
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// Polling wait (no hardware suspend): for the two barriers of the accumulator hand-shake, where the wake-up latency of a
// suspended warp sits on the critical path when one codebook tile is only a few hundred tensor-core cycles (small D).
__device__ __forceinline__ void mbar_wait_poll(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    long long t0 = 0;
    for (uint32_t it = 0;; ++it) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
        if (ok) return;
        if ((it & 1023u) == 1023u) {
            const long long now = clock64();
            if (t0 == 0) t0 = now;
            else if (now - t0 > 8000000000LL) __trap();
        }
    }
}
// One lane of a converged warp; ptxas recognises the elect.sync predicate and issues the following tcgen05 / TMA
// instructions directly instead of wrapping each one in an elect-and-retry loop.
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
        "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
// The same load with an L2 evict-first policy: latents are read once per pass, and as evict-normal lines they push the hot
// working set (codebook, |e|^2, residual-sum replicas) out of L2 (ncu, tail at BASELINE config 3: 53 % of the residual
// reductions missed L2).  0x12F0... is the fixed encoding of createpolicy.fractional.L2::evict_first with fraction 1.0.
__device__ __forceinline__ void tma_load_3d_once(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, %5}], [%2], %6;" ::"r"(dst),
        "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "l"(0x12F0000000000000ull)
        : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

}  // namespace ptx
}  // namespace vqb
