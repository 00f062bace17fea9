// Tensor-core nearest-code shortlist for sm_100a: the fused distance + running-argmin kernel.
//
//   score[n, k] = |e_k|^2 - 2 * bf16(x_n) . bf16(e_k)        (the |x_n|^2 term is constant per frame)
//               = -2 * acc[n, k],   acc = bf16(x_n) . bf16(e_k) - |e_k|^2 / 2
//
// The bias is folded INTO the tensor-core contraction: one extra K step per tile multiplies a constant A operand
// (-1, -1, -1, 0...) with (h1, h2, h3, 0...), the three-term bf16 split of |e_k|^2 / 2, so the accumulator already is the
// (negated, halved) score and the epilogue is a pure running arg-max - no bias loads, no FMAs.
//
// replaces `distances` + `argmin` of src/model/components/vector_quantizer.py:32-37.  The N x K score matrix never
// leaves the SM: tcgen05.mma accumulates a 128-frame x 256-code tile in TMEM, epilogue warps pull it back with
// tcgen05.ld, add |e_k|^2 and keep, per frame, the codes whose score is within a rigorous guard band of the running
// minimum (at most 12 per frame).  The tail kernel (vqb_kernels.cu) then rescores those few codes in fp32 in the reference's
// operation order, which decides the index; frames whose shortlist overflowed go to the exact fp32 search.
//
// Structure (one persistent CTA per SM, 640 threads, warp-specialised; CTA pairs with cta_group::2 MMAs by default):
//   warps 0-15  epilogue: warp w owns TMEM lanes 32*(w%4).., column quarter w/4 (64 of the 256 columns)
//   warp 16     TMA producer (one thread): codebook tiles B (256 codes x 64 dims per stage; half tiles per CTA of a pair)
//               as 128B-swizzled K-major boxes, bias-operand slices; in the unfused mode also the bf16 latent tile A
//   warp 17     MMA issuer (one thread): tcgen05.mma kind::f16, M128 (M256 over a pair) N256 K16, fp32 accumulate, two
//               TMEM accumulator stages (2 x 256 columns) so the epilogue of tile j overlaps the MMAs of tile j+1
//   warps 18-19 fused operand preparation: fp32 [B, D, W] latents -> bf16 swizzled A chunks (warp 19 also issues the 3-D
//               TMA loads of the staging ring; warp 18 also allocates TMEM)
//   warps 20-23 only in the opt-in fused-tail variant (kTail): finish the frames one tile behind the epilogue
#include "vqb_internal.h"
#include "vqb_ptx.cuh"

#include <cuda.h>
#include <stdlib.h>

namespace vqb {

namespace tc {

constexpr int BM = kTileRows;      // 128 frames per tile (= TMEM lanes)
constexpr int BN = kTileCodes;     // 256 codes per tile (= TMEM columns of one accumulator stage)
constexpr int BK = 64;             // bf16 elements per 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int A_CHUNK_BYTES = BM * BK * 2;   // 16 KiB
constexpr int B_STAGE_BYTES = BN * BK * 2;   // 32 KiB
constexpr int EH_SLICE_BYTES = BN * 16;      // 4 KiB: (h1,h2,h3,0,0,0,0,0) bf16 per code = K-half 0 of the bias operand
constexpr int EH_SLOTS = 2, MAX_EH_SLOTS = 8;   // default / maximum depth of the bias-operand ring (runtime: eh_slots)
constexpr int AX_BYTES = BM * 16;            // 2 KiB: K-half 0 of the constant A operand
constexpr int ZERO_BYTES = EH_SLICE_BYTES;   // shared all-zero K-half 1 of both bias operands
// fused operand preparation: fp32 latents arrive straight from the reference's [B, D, W] layout as 3-D TMA boxes of
// 16 dims x 128 frames (8 KiB) and are converted to the bf16 K-major swizzled A tile inside the kernel
constexpr int SUB_DIMS = 16, STG_SLOTS = 4, STG_BYTES = SUB_DIMS * BM * 4;
constexpr int CONVERT_WARP = 18, ALOAD_WARP = 19;
constexpr int MAX_A_SLOTS = 8, MAX_B_STAGES = 8;   // 4 whole-tile stages (1-CTA) or 8 half-tile stages (2-CTA)
constexpr int EPI_WARP0 = 0, EPI_WARPS = 16, EPI_THREADS = EPI_WARPS * 32;   // 4 warps per TMEM lane quarter: 64 columns each
// The single-thread producer / MMA loops sit in the HIGHEST warp ids: the scheduler favours them over waiting epilogue warps.
constexpr int PRODUCER_WARP = 16, MMA_WARP = 17, ALLOC_WARP = 18;
constexpr int NUM_THREADS = EPI_THREADS + 4 * 32;                              // 640
// fused tail (kTail): four more warps finish the frames of the PREVIOUS tile while the tensor core sweeps the next one
// (warps 20-23; lane = frame).  The tile's fp32 latents stream back in as 3-D TMA boxes of TAIL_CHUNK dims x 128 frames
// through a TX_SLOTS-deep shared-memory ring (issued by the first tail warp, a few boxes ahead of its own consumption):
// DRAM latency is hidden by the ring, not by registers.  A fifth warp would cost every thread 8 registers (warps are
// allocated four at a time).
constexpr int TAIL_WARP0 = 20, TAIL_WARPS = 4;
constexpr int NUM_THREADS_TAIL = NUM_THREADS + TAIL_WARPS * 32;   // 768
constexpr int TAIL_CHUNK = 8;                // dims per box
constexpr int TX_SLOTS = 6, TX_BYTES = TAIL_CHUNK * BM * 4;   // 6 x 4 KiB
constexpr int TAIL_EQ = 2;                   // codebook-row chunks a lane keeps in flight (registers; these gathers hit L2)
constexpr int COLS_PER_WARP = BN / (EPI_WARPS / 4);                            // 64
constexpr int kCandFill = 12;                // shortlist entries published per frame (cand_idx rows hold kCandMax = 16)

struct Barriers {
    unsigned long long b_full[MAX_B_STAGES], b_empty[MAX_B_STAGES];
    unsigned long long a_full[MAX_A_SLOTS], a_empty[MAX_A_SLOTS];
    unsigned long long eh_full[MAX_EH_SLOTS], eh_empty[MAX_EH_SLOTS], tmem_full[2], tmem_empty[2];
    unsigned long long stg_full[STG_SLOTS], stg_empty[STG_SLOTS];
    unsigned long long a_land[2];     // tf32 mode: the TMA boxes of a whole fp32 A tile have landed (local; a_full follows the norm pass)
    unsigned int tmem_base;
    unsigned int pad;
};

struct TailBarriers {   // fused tail only (kept out of Barriers: the default kernel has 120 bytes of shared memory to spare)
    unsigned long long tail_full[2], tail_empty[2];
    unsigned long long tx_full[TX_SLOTS], tx_empty[TX_SLOTS];
};

// ---------------------------------------------------------------------------------------------- PTX wrappers
// (mbarrier / elect / TMA basics shared with the stand-alone tail kernel: vqb_ptx.cuh)
using namespace ptx;
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
        "l"(map), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
// same box delivered to the same shared-memory offset of every CTA in `mask`; each destination's barrier gets the bytes
__device__ __forceinline__ void tma_load_2d_mc(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(dst),
        "l"(map), "r"(bar), "r"(c0), "r"(c1), "h"(mask)
        : "memory");
}
// 2-CTA (cta_group::2) variants.  The pair's leader (cluster rank 0) owns the "full" barriers: both CTAs' TMA loads
// complete their transaction bytes on the leader's barrier (peer bit 24 of the shared-window address cleared).
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
        "l"(map), "r"(bar & kPeerBitMask), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar, uint32_t cta) {   // arrive on the same barrier of CTA `cta`
    asm volatile(
        "{\n\t.reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}" ::"r"(bar),
        "r"(cta)
        : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(bar)
                 : "memory");
}

__device__ __forceinline__ void tmem_alloc_2sm(uint32_t smem_dst, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
// D[tmem of both CTAs] (+)= [A0; A1] * [B0; B1]^T over the CTA pair: M = 256 (128 frames per CTA), N = 256 (128 codes per CTA)
__device__ __forceinline__ void umma_bf16_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint32_t bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask)
                 : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, bf16 inputs, fp32 accumulate
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, fp32 operands read as tf32 (kind::tf32, K = 8 per instruction), fp32 accumulate
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_tf32_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// arrive on the barrier at this offset in every CTA of `mask` once the MMAs issued so far have completed
__device__ __forceinline__ void umma_commit_mc(uint32_t bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask)
                 : "memory");
}

// K-major operand, 128-byte swizzle: rows of 128 B, 8-row groups 1024 B apart (SBO), LBO unused (=1), version 1 (sm_100)
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// K-major operand without swizzle, one K step (16 bf16) wide: 8x8 core matrices of 128 contiguous bytes; `sbo` bytes between
// 8-row groups, `lbo` bytes from K-half 0 to K-half 1
__device__ __forceinline__ uint64_t make_desc_noswz(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
// MN-major tf32 operand (the fp32 [dims][frames] latent boxes as they lie in the reference's [B, D, W] layout).  For 32-bit
// MN-major operands the only shared-memory layout the tensor core accepts is "128-byte swizzle with 32-byte atoms" (layout type 1;
// TMA writes it with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B): rows of 128 B = 32 fp32 along MN, atoms of 4 K-rows (512 B) in which the
// 32-byte chunk index is XORed with the row number.  `lbo` bytes between consecutive 32-element groups along MN, `sbo` bytes
// between consecutive 4-row atoms along K (cute/atom/mma_traits_sm100.hpp, Layout_MN_SW128_32B_Atom).
__device__ __forceinline__ uint64_t make_desc_mn_sw128_32b(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46) |
           (1ull << 61);
}
// kind::tf32: a/b format TF32 (2<<7, 2<<10), A MN-major (bit 15), B K-major
constexpr uint32_t kIdescTf32 = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
constexpr uint32_t kIdescTf32_2 = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((2 * BM) >> 4) << 24);
// c_format F32 (1<<4), a/b format BF16 (1<<7, 1<<10), K-major both, N>>3 at bit 17, M>>4 at bit 24
constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
constexpr uint32_t kIdesc2 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((2 * BM) >> 4) << 24);

#define VQB_R32(r) \
    "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), \
    "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),   \
    "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),  \
    "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
#define VQB_RW32(r) \
    "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]), \
    "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]), "+r"(r[17]), "+r"(r[18]),   \
    "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]),  \
    "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])

// 32 lanes x 32 consecutive columns: thread i of the warp receives columns [col, col+32) of TMEM lane (lane_base + i)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, "
        "%25, %26, %27, %28, %29, %30, %31}, [%32];"
        : VQB_R32(r)
        : "r"(taddr)
        : "memory");
}
// The wait names the destination registers as read-write operands so that no use of them can be scheduled above it.
__device__ __forceinline__ void tmem_ld_wait(uint32_t (&r)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;" : VQB_RW32(r)::"memory");
}

// ---------------------------------------------------------------------------------------------- shortlist
// Per-thread event stack in global (L2-resident) scratch.  Whenever an 8-code chunk's minimum score comes within the
// guard band of the thread's running minimum, the chunk is appended RAW: its 8 accumulators straight from the TMEM
// registers, its minimum and its id (48 bytes, three predicated 16-byte stores, no arithmetic, no divergent code).
// Which codes of the surviving chunks are really inside the band is worked out once per frame tile, after the sweep,
// against the frame's final minimum.  EV_CAP chunks per thread and frame tile; running out (adversarial orderings) only
// sends that frame to the exact search.
constexpr int EV_CAP = 32;
constexpr int EV_WORDS = 12;   // 8 accumulators, chunk minimum, chunk id, 2 pad

struct EventStack {
    uint32_t* base;   // third level: this thread's EV_CAP x EV_WORDS words in global (L2-resident) scratch
    uint32_t* sbase;  // first level: the thread's own `sm` entries in shared memory (when the tile pipeline leaves room: small D)
    int       sm;
    int       n;      // entries in the thread's own shared-memory slots (<= sm)
    int       ng;     // entries in the global stack (may exceed EV_CAP: overflow -> the frame goes to the exact search)
    // second level: a pool of the WARP in shared memory (entries carry their lane).  With a small codebook there are few codebook
    // tiles per frame tile, and the end-of-round resolution used to wait for two dependent L2 round trips whenever ONE thread of the
    // block had more events than own slots (timeline at K = 1024: 2 300 + 1 600 cycles of a 12 100-cycle round); the threads of a warp
    // rarely need more than a handful of extra entries between them.
    uint32_t  pool_u, pool_cnt_u, pool_cap, lane;
    __device__ __forceinline__ const uint32_t* own(int i) const { return sbase + i * EV_WORDS; }
    __device__ __forceinline__ const uint32_t* glob(int i) const { return base + i * EV_WORDS; }
    __device__ __forceinline__ void push_if(bool p, float tmin, int chunk, const uint32_t* a) {
        const bool q = p && n < sm;
        const uint32_t dst = smem_u32(sbase) + (uint32_t)n * (EV_WORDS * 4);
        asm volatile(
            "{\n\t.reg .pred q;\n\t"
            "setp.ne.u32 q, %0, 0;\n\t"
            "@q st.shared.v4.u32 [%1], {%2, %3, %4, %5};\n\t"
            "@q st.shared.v4.u32 [%1+16], {%6, %7, %8, %9};\n\t"
            "@q st.shared.v2.u32 [%1+32], {%10, %11};\n\t}" ::"r"((uint32_t)q),
            "r"(dst), "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]), "r"(__float_as_uint(tmin)),
            "r"((uint32_t)chunk)
            : "memory");
        n += q ? 1 : 0;
        if (p && !q) {                             // rare: own slots are full
            uint32_t pos = pool_cap;
            if (pool_cap) asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(pos) : "r"(pool_cnt_u) : "memory");
            if (pos < pool_cap) {
                const uint32_t d = pool_u + pos * (EV_WORDS * 4);
                asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(d), "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]) : "memory");
                asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(d + 16), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]) : "memory");
                asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(d + 32), "r"(__float_as_uint(tmin)), "r"((uint32_t)chunk), "r"(lane), "r"(0u)
                             : "memory");
            } else {
                if (ng < EV_CAP) {
                    uint32_t* g = base + ng * EV_WORDS;
                    *reinterpret_cast<uint4*>(g) = make_uint4(a[0], a[1], a[2], a[3]);
                    *reinterpret_cast<uint4*>(g + 4) = make_uint4(a[4], a[5], a[6], a[7]);
                    *reinterpret_cast<uint2*>(g + 8) = make_uint2(__float_as_uint(tmin), (uint32_t)chunk);
                }
                ng += 1;
            }
        }
    }
};

// monotone float <-> int mapping so that a shared running maximum can be kept with an integer atomicMax
__device__ __forceinline__ int f2ord(float f) { const int b = __float_as_int(f); return b >= 0 ? b : b ^ 0x7FFFFFFF; }
__device__ __forceinline__ float ord2f(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7FFFFFFF); }

// One 32-column slab of accumulators (= -score/2) for this thread's frame: a pure running arg-max.
// Fast path per slab: 16 FMNMX3 + one compare.  When the slab maximum is within the band of the running maximum, the
// four 8-code chunks are appended under predicates (straight-line code) and the threshold tightens.
__device__ __forceinline__ float slab_max32(const uint32_t (&r)[32]) {
    float m = fmaxf(fmaxf(__uint_as_float(r[0]), __uint_as_float(r[1])), __uint_as_float(r[2]));
#pragma unroll
    for (int j = 3; j < 31; j += 2) m = fmaxf(fmaxf(m, __uint_as_float(r[j])), __uint_as_float(r[j + 1]));
    return fmaxf(m, __uint_as_float(r[31]));
}

// The four threads that share a frame (one per column quarter, in four different warps) pool their running maximum in
// shared memory (`smax`): every thread's threshold tracks the best score ANY quarter has seen, which cuts the number of
// appended chunks per frame from 4 x ln(K/32) to ln(K/8)-ish.
__device__ __forceinline__ void scan_slab(const uint32_t (&r)[32], int chunk0, float hband, float& thr, EventStack& ev, int* smax) {
    thr = fmaxf(thr, ord2f(*reinterpret_cast<volatile int*>(smax)) - hband);
    float t[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        float m = fmaxf(fmaxf(__uint_as_float(r[g * 8 + 0]), __uint_as_float(r[g * 8 + 1])), __uint_as_float(r[g * 8 + 2]));
        m = fmaxf(fmaxf(m, __uint_as_float(r[g * 8 + 3])), __uint_as_float(r[g * 8 + 4]));
        m = fmaxf(fmaxf(m, __uint_as_float(r[g * 8 + 5])), __uint_as_float(r[g * 8 + 6]));
        t[g] = fmaxf(m, __uint_as_float(r[g * 8 + 7]));
    }
    const float slab_max = fmaxf(fmaxf(fmaxf(t[0], t[1]), t[2]), t[3]);
    if (slab_max > thr) {
#pragma unroll
        for (int g = 0; g < 4; ++g) ev.push_if(t[g] > thr, t[g], chunk0 + g, &r[g * 8]);
        thr = fmaxf(thr, slab_max - hband);
        atomicMax(smax, f2ord(slab_max));
    }
}

// ---- small codebooks (K <= 1024): per-WARP slab queues in shared memory ------------------------------------------------------------
// What the hand-off timeline (scripts/trace_tc.py, profiles/r03_trace_cfg2_*.txt) showed for the per-thread stacks at K = 1024, D = 64:
// a frame-tile round of 12 000 cycles = 4 tiles x ~1 600 (the scan phases are issue-bound: ~100 instructions per 32-column slab and
// warp, four warps per scheduler in lock-step, because with a small codebook a 32-frame x 32-code slab nearly always holds an event
// and the four predicated chunk appends run for every slab) + ~3 500 cycles of end-of-round resolution; the tensor core needs 640
// cycles per tile.  This form keeps the lock-step of all 16 warps (the thread-level parallelism the scan needs; a split into two
// groups of 8 warps made the scan latency-bound at the same throughput) and makes both parts cheaper:
//  * a lane whose slab maximum reaches its threshold stores the slab RAW (32 accumulators + maximum + first code | lane) as ONE entry
//    of its warp's queue - slot by ballot prefix, the fill is a warp-uniform register, no atomics, no per-chunk bookkeeping;
//  * the pooled running maximum of a frame's four threads is read once per tile, not once per slab;
//  * at the end of the round the queue's entries are dealt one per lane, whatever frame they belong to: 32 compares into a bit mask,
//    ONE shared-memory atomic for the entry's shortlist positions; barriers of the 128 threads that share 32 frames;
//  * a full queue spills into the warp's global scratch (never a fall-back of 32 frames at once); a lane that floods only loses itself.
constexpr int Q_ENTRY = 48;        // overflow-pool entries of the per-thread stacks: 8 raw accumulators, (chunk maximum, chunk id, lane, 0)
constexpr int SQ_ENTRY = 144;      // slab-queue entries: 32 raw accumulators, (slab maximum, first code | lane << 16), 8 pad (16-byte reads stay conflict-free)
constexpr int SQ_LANE_CAP = 16;    // entries one lane may append per round: a frame that floods (adversarial code order) only loses itself
constexpr int SQ_SPILL = 256;      // entries per warp in global scratch behind the shared-memory queue
__device__ __forceinline__ void scan_slab_sq(const uint32_t (&r)[32], int code0, float hband, float& thr, float& mymax, int* smax, uint32_t q_u,
                                             uint32_t q_cap, unsigned char* spill, uint32_t& qn, int& mine, int lane) {
    float t[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) {      // depth-3 trees (the chain form of slab_max32 is 16 dependent instructions)
        const float m0 = fmaxf(fmaxf(__uint_as_float(r[g * 8 + 0]), __uint_as_float(r[g * 8 + 1])), __uint_as_float(r[g * 8 + 2]));
        const float m1 = fmaxf(fmaxf(__uint_as_float(r[g * 8 + 3]), __uint_as_float(r[g * 8 + 4])), __uint_as_float(r[g * 8 + 5]));
        t[g] = fmaxf(fmaxf(fmaxf(m0, m1), __uint_as_float(r[g * 8 + 6])), __uint_as_float(r[g * 8 + 7]));
    }
    const float m = fmaxf(fmaxf(t[0], t[1]), fmaxf(t[2], t[3]));
    const bool want = m >= thr;
    if (__any_sync(0xffffffffu, want)) {
        const bool hit = want && mine < SQ_LANE_CAP;
        if (want && !hit) mine = SQ_LANE_CAP + 1;        // flooded: this frame goes to the exact search, it takes no more slots
        const unsigned mk = __ballot_sync(0xffffffffu, hit);
        if (hit) {
            const uint32_t pos = qn + (uint32_t)__popc(mk & ((1u << lane) - 1u));
            const uint32_t hd1 = (uint32_t)code0 | ((uint32_t)lane << 16);
            if (pos < q_cap) {
                const uint32_t a = q_u + pos * SQ_ENTRY;
#pragma unroll
                for (int q = 0; q < 8; ++q)
                    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a + 16 * q), "r"(r[4 * q]), "r"(r[4 * q + 1]), "r"(r[4 * q + 2]),
                                 "r"(r[4 * q + 3])
                                 : "memory");
                asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(a + 128), "r"(__float_as_uint(m)), "r"(hd1) : "memory");
            } else if (pos < q_cap + (uint32_t)SQ_SPILL) {
                uint4* gdst = reinterpret_cast<uint4*>(spill + (size_t)(pos - q_cap) * SQ_ENTRY);
#pragma unroll
                for (int q = 0; q < 8; ++q) gdst[q] = make_uint4(r[4 * q], r[4 * q + 1], r[4 * q + 2], r[4 * q + 3]);
                *reinterpret_cast<uint2*>(gdst + 8) = make_uint2(__float_as_uint(m), hd1);
            } else {
                mine = SQ_LANE_CAP + 1;                  // no room anywhere
            }
            mine += 1;
        }
        qn += (uint32_t)__popc(mk);
        if (want) {
            thr = fmaxf(thr, m - hband);
            if (m > mymax) {
                mymax = m;
                atomicMax(smax, f2ord(m));
            }
        }
    }
}

__device__ __forceinline__ void dump_slab(const uint32_t (&r)[32], int code0, int K, float* row_out) {
#pragma unroll
    for (int j = 0; j < 32; ++j)
        if (code0 + j < K) row_out[code0 + j] = -2.f * __uint_as_float(r[j]);
}

// ---------------------------------------------------------------------------------------------- fused tail helpers
// L2 policy for the straight-through output (written once, never read by this kernel): evict first, so that it does
// not push the codebook, the event stacks or the residual sums out of L2
__device__ __forceinline__ uint64_t policy_evict_first() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void stg_once(float* p, float v, uint64_t) {   // written once, never read by this kernel
    asm volatile("st.global.cs.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}
// 32 bytes of a codebook row per lane and request: each lane gathers from its own row, so every request costs one L1 tag
// lookup per lane whatever its width - 256-bit loads halve that cost against 128-bit ones (the pointer is 32-byte aligned)
__device__ __forceinline__ void ldg256(const float* p, float (&v)[8]) {
    asm volatile("ld.global.nc.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
                 : "l"(p));
}
// How many pair sweeps the slowest tail warp needs for this tile: every tail warp deals the (frame, code) pairs of its 32
// frames out one per lane, 32 per sweep.  The tail warps and their loader all evaluate this on the same counts, so they
// agree on the number of latent streams without talking to each other.
__device__ __forceinline__ int tail_max_sweeps(const uint8_t* cnts, int lane) {
    const uint32_t c4 = reinterpret_cast<const uint32_t*>(cnts)[lane];   // frames 4*lane .. 4*lane+3: tail warp lane / 8
    int np = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int c = (int)((c4 >> (8 * j)) & 0xFFu);
        np += c > 1 ? c : 0;
    }
    np += __shfl_xor_sync(0xffffffffu, np, 1);
    np += __shfl_xor_sync(0xffffffffu, np, 2);
    np += __shfl_xor_sync(0xffffffffu, np, 4);
    int sw = (np + 31) >> 5;
    sw = max(sw, __shfl_xor_sync(0xffffffffu, sw, 8));
    sw = max(sw, __shfl_xor_sync(0xffffffffu, sw, 16));
    return sw;
}
// One pass of a tail warp over the tile's latents as they stream through the ring: box g holds dims [8g, 8g+8) of all 128
// frames, [dim][frame] fp32.  Each active lane reads column `xcol` of every box and the same dims of codebook row `ep`
// (two 16-byte gathers per box, kept TAIL_EQ boxes ahead in registers).
//   kFinal = false: accumulates x.e and |x|^2 (fp32 rescoring of one (frame, code) pair)
//   kFinal = true : codeword gather, straight-through value fl(x + fl(e - x)) -> qp, SSE -> fs, residual sums -> rp
// All four tail warps consume every box in the same order; `it` counts boxes since the kernel started.
// Position of a tail warp in the ring: the slot it consumes next and the fill parity of that slot.
struct RingPos {
    uint32_t slot, ph;
    __device__ __forceinline__ void advance() { if (++slot == TX_SLOTS) { slot = 0; ph ^= 1u; } }
};
constexpr int TX_AHEAD = TX_SLOTS - 2;   // boxes the loader keeps in flight ahead of its own consumption
// The first tail warp doubles as the loader of the ring and keeps no state for it: the box that is TX_AHEAD positions ahead of
// the one it consumes next lands TX_AHEAD slots further on, and its slot is free once all four warps released the box that was
// there before (so the other warps may lag one box behind without stalling the loader).
__device__ __forceinline__ void tail_issue(const RingPos& pos, int ahead, int dim0, int b, int w0, uint32_t sTx_u, uint32_t bar_full,
                                           uint32_t bar_empty, const CUtensorMap* map) {
    uint32_t sl = pos.slot + (uint32_t)ahead, ph = pos.ph;
    if (sl >= (uint32_t)TX_SLOTS) { sl -= TX_SLOTS; ph ^= 1u; }
    mbar_wait(bar_empty + sl * 8, ph ^ 1u);
    if (elect_one()) {
        mbar_expect_tx(bar_full + sl * 8, TX_BYTES);
        tma_load_3d(sTx_u + sl * TX_BYTES, map, bar_full + sl * 8, w0, dim0, b);
    }
    __syncwarp();
}

// `more` = boxes of this tile that follow this pass (the loader runs ahead across pass boundaries, never across tiles).
template <bool kFinal>
__device__ __forceinline__ void tail_stream(const unsigned char* sTx, uint32_t bar_full, uint32_t bar_empty, RingPos& pos, int nbox,
                                            bool act, int xcol, const float* __restrict__ ep, int lane, float& acc0, float& acc1,
                                            int64_t W, float* qp, float* rp, bool resid_v4, uint64_t pol, bool loader, int more, int b,
                                            int w0, const CUtensorMap* map) {
    float eq[TAIL_EQ][TAIL_CHUNK];
#pragma unroll
    for (int u = 0; u < TAIL_EQ; ++u)
        if (act && u < nbox) ldg256(ep + u * TAIL_CHUNK, eq[u]);
    for (int g0 = 0; g0 < nbox; g0 += TAIL_EQ) {
#pragma unroll
        for (int u = 0; u < TAIL_EQ; ++u) {
            const int g = g0 + u;
            if (g < nbox) {
                if (loader && g + TX_AHEAD < nbox + more) {
                    int gi = g + TX_AHEAD;
                    while (gi >= nbox) gi -= nbox;
                    tail_issue(pos, TX_AHEAD, gi * TAIL_CHUNK, b, w0, smem_u32(sTx), bar_full, bar_empty, map);
                }
                mbar_wait(bar_full + pos.slot * 8, pos.ph);
                if (act) {
                    const float* bx = reinterpret_cast<const float*>(sTx + (size_t)pos.slot * TX_BYTES) + xcol;
                    float x[TAIL_CHUNK];
#pragma unroll
                    for (int i = 0; i < TAIL_CHUNK; ++i) x[i] = bx[i * BM];
                    const float (&ev)[TAIL_CHUNK] = eq[u];
                    if (!kFinal) {
                        float cd = 0.f, cx = 0.f;         // blocked summation: box sums first, then the running totals
#pragma unroll
                        for (int i = 0; i < TAIL_CHUNK; ++i) { cd = fmaf(x[i], ev[i], cd); cx = fmaf(x[i], x[i], cx); }
                        acc0 += cd;
                        acc1 += cx;
                    } else {
                        float df[TAIL_CHUNK], cs = 0.f;
#pragma unroll
                        for (int i = 0; i < TAIL_CHUNK; ++i) { df[i] = __fsub_rn(ev[i], x[i]); cs = fmaf(df[i], df[i], cs); }
                        acc0 += cs;
                        if (qp) {
                            float* q = qp + (size_t)g * TAIL_CHUNK * W;
#pragma unroll
                            for (int i = 0; i < TAIL_CHUNK; ++i) stg_once(q + (size_t)i * W, __fadd_rn(x[i], df[i]), pol);   // vector_quantizer.py:48
                        }
                        if (rp) {
                            float* r = rp + g * TAIL_CHUNK;
                            if (resid_v4) {
                                red_add_v4(r, -df[0], -df[1], -df[2], -df[3]);
                                red_add_v4(r + 4, -df[4], -df[5], -df[6], -df[7]);
                            } else {
#pragma unroll
                                for (int i = 0; i < TAIL_CHUNK; ++i) atomicAdd(r + i, -df[i]);
                            }
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_empty + pos.slot * 8);
                if (act && g + TAIL_EQ < nbox) ldg256(ep + (g + TAIL_EQ) * TAIL_CHUNK, eq[u]);
                pos.advance();
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------- trace (debug builds only)
// -DVQB_TC_TRACE (scripts/build_trace.py -> libvqb_b200_trace.so, never the shipped library): one lane per role of CTA 0 / 1 writes
// SM-clock timestamps of its hand-offs into a global buffer (vqb_debug_set_trace): the timeline that the per-line stall profile of
// a warp-specialised kernel cannot show.  Record slot = [CTA][role][event][tile counter], value = bit 63 | clock low 32 bits.
#ifdef VQB_TC_TRACE
__device__ unsigned long long* g_trace_buf = nullptr;
__device__ unsigned int g_trace_cap = 0;      // records per CTA: 8 roles x 8 events x 1024 counters
// the record base of this CTA is read ONCE per thread (VQB_TRACE_INIT): a device-global read per event would cost a memory round trip
__device__ __forceinline__ void trace_ev(unsigned long long* base, int role, int event, unsigned int counter) {
    if (base && counter < 1024u)                  // fire-and-forget store into the record's own slot: no atomics, no waiting
        base[(size_t)(((role << 3) | event) << 10) + counter] = (1ull << 63) | (unsigned long long)(unsigned int)clock64();
}
#define VQB_TRACE_INIT() unsigned long long* const trace_base__ = (blockIdx.x < 2 && g_trace_buf) ? g_trace_buf + (size_t)blockIdx.x * g_trace_cap : nullptr
#define VQB_TRACE(role, event, counter) trace_ev(trace_base__, role, event, counter)
#else
#define VQB_TRACE_INIT() ((void)0)
#define VQB_TRACE(role, event, counter) ((void)0)
#endif

// ---------------------------------------------------------------------------------------------- the kernel
// kTwo = false: cta_group::1 MMAs, every CTA holds whole codebook tiles (optionally multicast inside a cluster).
// kTwo = true : CTA pairs with cta_group::2 MMAs (M = 256 over the pair): each CTA holds its own 128 frames and HALF of every
//               codebook tile (128 codes), which halves the per-SM operand ingest and doubles the ring depth (8 stages).
// kFuse = true : the A operand is built in the kernel from the fp32 [B, D, W] latents (tmap_x is then a 3-D fp32 map):
//               warp 19 streams 16-dim x 128-frame boxes through a small ring, warp 18 converts them to bf16 into the
//               swizzled A chunks, measures |x| and |x - bf16(x)| per frame and publishes the guard band in shared memory.
//               Frame tiles then never straddle a batch item (tile = (b, w0 .. w0+127)); no bf16 copy of the latents exists.
// kTail = true (needs kFuse): warps 20-23 finish every frame whose shortlist held - fp32 rescoring in the reference's op order,
//               codeword gather, straight-through value, SSE, histogram, residual sums, index - one tile behind the tensor
//               core.  The tile's latents are read a second time while they are still in L2, so the forward pass touches
//               HBM once for the latents and once for `quantized`; no stand-alone tail kernel runs.
// kTf32 = true (needs kFuse, no kTail): kind::tf32 MMAs straight from the fp32 operands.  The A tile is the tile's fp32 latents as four
//               128-byte-swizzled TMA boxes [D dims][32 frames] (an MN-major operand: no conversion, no copy); the codebook tiles are
//               fp32 boxes [codes][32 dims].  Warp 19 only loads A tiles, warp 18 measures |x| and |x - tf32(x)| of the landed tile
//               (guard band) and then releases it to the MMA issuer.  tf32 runs at half the bf16 tensor rate, the band is ~3x tighter.
// kGrouped = true (needs kFuse, no kTail): the slab-queue epilogue (scan_slab_sq: per-warp queues of raw slabs) instead of per-thread stacks;
//               chosen by tc_plan for small codebooks (K <= 1024).
template <bool kTwo, bool kFuse, bool kTail, bool kTf32 = false, bool kGrouped = false>
__global__ void __launch_bounds__(kTail ? NUM_THREADS_TAIL : NUM_THREADS, 1)
tc_search_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_e,
                 const __grid_constant__ CUtensorMap tmap_eh, const __grid_constant__ CUtensorMap tmap_xt, const __nv_bfloat16* __restrict__ eh, int64_t W, int tiles_per_item, int D,
                 const WsMeta* __restrict__ meta_ro, const float* __restrict__ band_g, int64_t N, int num_m_tiles, int num_n_tiles,
                 int num_kb, int a_slots, int b_stages, int cs, int K, uint8_t* __restrict__ cand_cnt,
                 uint16_t* __restrict__ cand_idx, int* __restrict__ fallback_rows, WsMeta* meta,
                 unsigned long long* __restrict__ best64, float* __restrict__ scores_dbg, uint32_t* __restrict__ ev_scratch,
                 const TailArgs tail, const int ev_sm, const int l2_once, const int eh_slots, const int wq_cap, const uint32_t spin_ns) {
    VQB_TRACE_INIT();
    static_assert(!kTail || kFuse, "the fused tail needs frame tiles that never straddle a batch item");
    static_assert(!kTf32 || (kFuse && !kTail), "tf32 reads the fp32 latents in place; no fused tail");
    static_assert(!kGrouped || (kFuse && !kTail), "the slab-queue epilogue reads the guard bands from shared memory; no fused tail");
    extern __shared__ __align__(1024) unsigned char smem[];
    const uint32_t a_slot_bytes = kTf32 ? (uint32_t)BM * (uint32_t)D * 4u : (uint32_t)A_CHUNK_BYTES;   // tf32: a_slots whole fp32 tiles
    unsigned char* sA = smem;                                            // a_slots x 16 KiB (tf32: a_slots x 128 D x 4 bytes)
    unsigned char* sB = sA + (size_t)a_slots * a_slot_bytes;             // b_stages x 32 KiB (16 KiB half tiles in 2-CTA mode)
    unsigned char* sEH = sB + (size_t)b_stages * (kTwo ? B_STAGE_BYTES / 2 : B_STAGE_BYTES);   // eh_slots x 4 KiB bias operand B (K-half 0)
    unsigned char* sAX = sEH + eh_slots * EH_SLICE_BYTES;                // 2 KiB constant bias operand A (K-half 0)
    unsigned char* sZero = sAX + AX_BYTES;                               // 4 KiB of zeros: K-half 1 of both bias operands
    unsigned char* sStg = sZero + ZERO_BYTES;                            // fused mode: STG_SLOTS x 8 KiB fp32 boxes [16 dims][128 frames]
    float* sBand = reinterpret_cast<float*>(sStg + ((kFuse && !kTf32) ? STG_SLOTS * STG_BYTES : 0));   // fused mode: [4][128] guard bands
    // ([4][128]: the bands of round rd are written while the epilogue may still be up to two accumulator stages behind the tensor
    // core - with one or two codebook tiles per frame tile that is up to three rounds back)
    float* sMin = sBand + (kFuse ? 4 * BM : 0);                          // [4][128] running maxima of the four column quarters
    int* sCnt = reinterpret_cast<int*>(sMin + 4 * BM);                   // [128] shortlist fill per frame, [128] overflow flags,
    uint32_t* sEv = reinterpret_cast<uint32_t*>(sCnt + 3 * BM);          // [512][ev_sm] shared-memory part of the event stacks
    // fused tail: the shortlists of two frame tiles (the epilogue fills one while the tail warps consume the other)
    // slab-queue epilogue: [16 warps][wq_cap] queue entries of 144 bytes, then the tile's shortlists [128][kCandMax] (written out as 32-byte rows)
    // per-thread stacks: [16 warps][wq_cap] overflow pool entries of the warps (same 48-byte entries)
    unsigned char* sWarpQ = reinterpret_cast<unsigned char*>(sEv + (size_t)EPI_THREADS * ev_sm * EV_WORDS);
    uint16_t* sList = reinterpret_cast<uint16_t*>(sWarpQ + (size_t)EPI_WARPS * wq_cap * (kGrouped ? SQ_ENTRY : Q_ENTRY));
    uint32_t* sPoolCnt = reinterpret_cast<uint32_t*>(reinterpret_cast<unsigned char*>(sList) + (kGrouped ? BM * kCandMax * 2 : 0));   // [16] pool fills
    uint16_t* sCand = reinterpret_cast<uint16_t*>(sPoolCnt + 2 * EPI_WARPS);   // (128 bytes: what follows keeps its alignment) [2][128][kCandFill] codes
    uint8_t* sCandCnt = reinterpret_cast<uint8_t*>(sCand + (kTail ? 2 * BM * kCandFill : 0));   // [2][128] 0 = not for the tail
    float2* sPair = reinterpret_cast<float2*>(sCandCnt + (kTail ? 2 * BM : 0));                  // [4][32] (distance, code) hand-back
    unsigned char* sTx = reinterpret_cast<unsigned char*>(sPair + (kTail ? TAIL_WARPS * 32 + 16 : 0));   // (+128 B of per-warp totals) TX_SLOTS x 4 KiB latent boxes
    TailBarriers* tbars = reinterpret_cast<TailBarriers*>(sTx + (kTail ? TX_SLOTS * TX_BYTES : 0));
    Barriers* bars = reinterpret_cast<Barriers*>(reinterpret_cast<unsigned char*>(tbars) + (kTail ? sizeof(TailBarriers) : 0));   // sCnt: [128] pooled running maximum per frame (ordered int)

    // Logical warp id = role.  With the fused tail the roles are rotated so that the tail warps are the four LOWEST hardware
    // warps: the scheduler favours high warp ids, and the tail must never win an issue slot against the MMA / TMA / epilogue warps.
    // (Hardware warp = logical warp + 4 mod 24: the TMEM lane quarter warp % 4 of the epilogue warps is unchanged.)
    const int lane = threadIdx.x & 31;
    const int warp = kTail ? (int)(((threadIdx.x >> 5) + (unsigned)TAIL_WARP0) % (unsigned)(TAIL_WARP0 + TAIL_WARPS)) : (int)(threadIdx.x >> 5);
    // Cluster of `cs` CTAs: every CTA quantises its own 128-frame tile, but each codebook tile is fetched from L2 only
    // once per cluster - CTA r loads rows [r*256/cs, (r+1)*256/cs) and multicasts them into all cs shared memories.
    // All CTAs of a cluster therefore walk the same (tile round, codebook tile) sequence; a CTA whose frame tile lies
    // past the end works on zero-filled rows and publishes nothing.
    const uint32_t crank = cs > 1 ? cluster_ctarank() : 0u;
    const uint16_t cmask = (uint16_t)((1u << cs) - 1u);
    constexpr uint32_t kStageBytes = kTwo ? B_STAGE_BYTES / 2 : B_STAGE_BYTES;   // per-CTA bytes of one codebook stage
    constexpr uint32_t kEhBytes = kTwo ? EH_SLICE_BYTES / 2 : EH_SLICE_BYTES;
    constexpr int kBK = kTf32 ? 32 : BK;         // elements per 128-byte swizzle row of a codebook stage: 32 fp32 or 64 bf16
    const bool leader = !kTwo || crank == 0;
    const int n_clusters = (int)gridDim.x / cs;
    const int cluster_id = (int)blockIdx.x / cs;
    const int rounds = (num_m_tiles + n_clusters * cs - 1) / (n_clusters * cs);

    if (warp == PRODUCER_WARP && lane == 0) {
        prefetch_tmap(&tmap_x);
        prefetch_tmap(&tmap_e);
    }
    if (warp == MMA_WARP && lane == 0) {
        for (int i = 0; i < b_stages; ++i) { mbar_init(smem_u32(&bars->b_full[i]), 1); mbar_init(smem_u32(&bars->b_empty[i]), kTwo ? 1 : cs); }
        for (int i = 0; i < a_slots; ++i) {
            // fused: two converter warps per CTA (both CTAs in 2-CTA mode); tf32: one norm warp per CTA
            mbar_init(smem_u32(&bars->a_full[i]), kTf32 ? (kTwo ? 2 : 1) : (kFuse ? (kTwo ? 4 : 2) : 1));
            mbar_init(smem_u32(&bars->a_empty[i]), 1);
        }
        for (int i = 0; i < STG_SLOTS; ++i) { mbar_init(smem_u32(&bars->stg_full[i]), 1); mbar_init(smem_u32(&bars->stg_empty[i]), 2); }
        mbar_init(smem_u32(&bars->a_land[0]), 1);
        mbar_init(smem_u32(&bars->a_land[1]), 1);
        for (int i = 0; i < eh_slots; ++i) {
            mbar_init(smem_u32(&bars->eh_full[i]), 1);
            mbar_init(smem_u32(&bars->eh_empty[i]), 1);
        }
        if (kTail) {
            for (int i = 0; i < 2; ++i) {
                mbar_init(smem_u32(&tbars->tail_full[i]), 4);                // the four epilogue warps that publish the counts
                mbar_init(smem_u32(&tbars->tail_empty[i]), TAIL_WARPS);
            }
            for (int i = 0; i < TX_SLOTS; ++i) {
                mbar_init(smem_u32(&tbars->tx_full[i]), 1);
                mbar_init(smem_u32(&tbars->tx_empty[i]), TAIL_WARPS);
            }
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(smem_u32(&bars->tmem_full[i]), 1);
            mbar_init(smem_u32(&bars->tmem_empty[i]), (kTwo ? 2 : 1) * EPI_WARPS);   // 2-CTA: both CTAs' epilogues report to the leader
        }
        fence_barrier_init();
    }
    if (warp == ALLOC_WARP) {
        if (kTwo) tmem_alloc_2sm(smem_u32(&bars->tmem_base), 512);
        else tmem_alloc(smem_u32(&bars->tmem_base), 512);
    }
    {   // constant bias operand A: every frame row is (-1, -1, -1, 0, 0, 0, 0, 0) in K-half 0; K-half 1 is the zero block
        uint4* ax = reinterpret_cast<uint4*>(sAX);
        for (int i = threadIdx.x; i < AX_BYTES / 16; i += (int)blockDim.x) ax[i] = make_uint4(0xBF80BF80u, 0x0000BF80u, 0u, 0u);
        uint4* zz = reinterpret_cast<uint4*>(sZero);
        for (int i = threadIdx.x; i < ZERO_BYTES / 16; i += (int)blockDim.x) zz[i] = make_uint4(0u, 0u, 0u, 0u);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core
    }
    tc_fence_before();
    __syncthreads();
    if (cs > 1) cluster_sync_all();   // peers must see initialised barriers before the first multicast / remote arrive
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;

    if (warp == PRODUCER_WARP) {
        // ================================================================ TMA producer (converged warp, one elected lane issues)
        uint32_t a_slot0 = 0, a_ph = 0, b_st = 0, b_ph = 0, es = 0, e_ph = 0;
        const uint32_t slice = (uint32_t)B_STAGE_BYTES / (uint32_t)cs;
        const int b_row_off = kTwo ? (int)crank * (BN / 2) : (int)crank * (BN / cs);
        const uint32_t sA_u = smem_u32(sA), sB_u = smem_u32(sB) + (kTwo ? 0u : crank * slice), sE_u = smem_u32(sEH);
        const uint32_t bar_afull = smem_u32(&bars->a_full[0]), bar_aempty = smem_u32(&bars->a_empty[0]);
        const uint32_t bar_bfull = smem_u32(&bars->b_full[0]), bar_bempty = smem_u32(&bars->b_empty[0]);
        const uint32_t bar_efull = smem_u32(&bars->eh_full[0]), bar_eempty = smem_u32(&bars->eh_empty[0]);
        for (int rd = 0; rd < rounds; ++rd) {
            const int mt = (rd * n_clusters + cluster_id) * cs + (int)crank;   // may be >= num_m_tiles: dummy tile
            const int a_row = (mt < num_m_tiles ? mt : 0) * BM;               // dummy tiles re-read tile 0, nothing is published
            for (int nt = 0; nt < num_n_tiles; ++nt) {
                mbar_wait_sleep(bar_eempty + es * 8, e_ph ^ 1, spin_ns);
                if (elect_one()) {
                    if (kTwo) {   // own half of the bias operand; both halves complete on the leader's barrier
                        if (leader) mbar_expect_tx(bar_efull + es * 8, EH_SLICE_BYTES);
                        tma_load_2d_2sm(sE_u + es * kEhBytes, &tmap_eh, bar_efull + es * 8, 0, nt * BN + b_row_off);
                    } else {
                        mbar_expect_tx(bar_efull + es * 8, EH_SLICE_BYTES);
                        bulk_load_1d(sE_u + es * EH_SLICE_BYTES, eh + (size_t)nt * BN * 8, EH_SLICE_BYTES, bar_efull + es * 8);
                    }
                }
                __syncwarp();
                if (++es == (uint32_t)eh_slots) { es = 0; e_ph ^= 1; }
                for (int kb = 0; kb < num_kb; ++kb) {
                    uint32_t slot = a_slot0 + kb, a_phk = a_ph;          // ring position of chunk kb of this tile
                    if (slot >= (uint32_t)a_slots) { slot -= a_slots; a_phk ^= 1; }
                    if (nt == 0 && !kFuse) mbar_wait_sleep(bar_aempty + slot * 8, a_phk ^ 1, spin_ns);
                    mbar_wait_sleep(bar_bempty + b_st * 8, b_ph ^ 1, spin_ns);
                    if (elect_one()) {
                        if (nt == 0 && !kFuse) {
                            if (kTwo) {
                                if (leader) mbar_expect_tx(bar_afull + slot * 8, 2 * A_CHUNK_BYTES);
                                tma_load_2d_2sm(sA_u + slot * A_CHUNK_BYTES, &tmap_x, bar_afull + slot * 8, kb * BK, a_row);
                            } else {
                                mbar_expect_tx(bar_afull + slot * 8, A_CHUNK_BYTES);
                                tma_load_2d(sA_u + slot * A_CHUNK_BYTES, &tmap_x, bar_afull + slot * 8, kb * BK, a_row);
                            }
                        }
                        if (kTwo) {
                            if (leader) mbar_expect_tx(bar_bfull + b_st * 8, B_STAGE_BYTES);   // both halves
                            tma_load_2d_2sm(sB_u + b_st * kStageBytes, &tmap_e, bar_bfull + b_st * 8, kb * kBK, nt * BN + b_row_off);
                        } else {
                            mbar_expect_tx(bar_bfull + b_st * 8, B_STAGE_BYTES);   // own slice + the peers' slices
                            if (cs == 1)
                                tma_load_2d(sB_u + b_st * B_STAGE_BYTES, &tmap_e, bar_bfull + b_st * 8, kb * kBK, nt * BN);
                            else
                                tma_load_2d_mc(sB_u + b_st * B_STAGE_BYTES, &tmap_e, bar_bfull + b_st * 8, kb * kBK, nt * BN + b_row_off, cmask);
                        }
                    }
                    __syncwarp();
                    if (++b_st == (uint32_t)b_stages) { b_st = 0; b_ph ^= 1; }
                }
            }
            a_slot0 += num_kb;
            if (a_slot0 >= (uint32_t)a_slots) { a_slot0 -= a_slots; a_ph ^= 1; }
        }
    } else if (warp == MMA_WARP) {
        // ================================================================ MMA issuer (converged warp, one elected lane issues;
        //                                                                   in 2-CTA mode only the pair's leader)
        if (leader) {
        uint32_t a_slot0 = 0, a_ph = 0, b_st = 0, b_ph = 0, as = 0, t_ph = 0, es = 0, e_ph = 0;
        const uint64_t dA0 = kTf32 ? make_desc_mn_sw128_32b(smem_u32(sA), (uint32_t)D * 128u, 512u) : make_desc_sw128(smem_u32(sA));
        const uint64_t dB0 = make_desc_sw128(smem_u32(sB));
        // bias operands: 8-row groups 128 B apart, K-half 1 = the shared zero block
        const uint64_t dAX = make_desc_noswz(smem_u32(sAX), smem_u32(sZero) - smem_u32(sAX), 128);
        const uint32_t sEH_u = smem_u32(sEH), sZero_u = smem_u32(sZero);
        const uint32_t bar_efull = smem_u32(&bars->eh_full[0]), bar_eempty = smem_u32(&bars->eh_empty[0]);
        const uint32_t bar_afull = smem_u32(&bars->a_full[0]), bar_aempty = smem_u32(&bars->a_empty[0]);
        const uint32_t bar_bfull = smem_u32(&bars->b_full[0]), bar_bempty = smem_u32(&bars->b_empty[0]);
        const uint32_t bar_tfull = smem_u32(&bars->tmem_full[0]), bar_tempty = smem_u32(&bars->tmem_empty[0]);
        for (int rd = 0; rd < rounds; ++rd) {
            for (int nt = 0; nt < num_n_tiles; ++nt) {
                if (lane == 0) VQB_TRACE(1, 0, rd * num_n_tiles + nt);        // MMA: about to wait for the accumulator stage
                mbar_wait_sleep(bar_tempty + as * 8, t_ph ^ 1, spin_ns >> 1);   // (the epilogue is waiting for what this wait gates: half the nap)
                if (lane == 0) VQB_TRACE(1, 1, rd * num_n_tiles + nt);        // MMA: stage free
                const uint32_t tmem_d = tmem_base + as * BN;
                const bool last_nt = nt == num_n_tiles - 1;
                for (int kb = 0; kb < num_kb; ++kb) {
                    if (kTf32) {
                        // the A operand is one whole fp32 tile per round (slot = rd mod a_slots, released by the norm warp); a codebook
                        // stage holds 32 dims = four K steps of 8 (fewer in the last stage when D % 32 != 0)
                        const uint32_t tslot = (uint32_t)rd % (uint32_t)a_slots, t_phase = ((uint32_t)rd / (uint32_t)a_slots) & 1u;
                        if (nt == 0 && kb == 0) mbar_wait(bar_afull + tslot * 8, t_phase);
                        mbar_wait(bar_bfull + b_st * 8, b_ph);
                        tc_fence_after();
                        if (elect_one()) {
                            // A: +64 (1024 bytes) per K step = the next group of 8 dims; B: +2 (32 bytes) per K step inside the 128-byte row
                            const uint64_t da = dA0 + (uint64_t)(tslot * (a_slot_bytes >> 4)) + (uint64_t)(kb * 4 * 64);
                            const uint64_t db = dB0 + (uint64_t)(b_st * (kStageBytes >> 4));
                            const int steps = (D - kb * 32) >= 32 ? 4 : (D - kb * 32) / 8;
                            for (int st = 0; st < steps; ++st) {
                                if (kTwo) umma_tf32_2sm(tmem_d, da + (uint64_t)(st * 64), db + (uint64_t)(st * 2), kIdescTf32_2, (kb | st) ? 1u : 0u);
                                else umma_tf32(tmem_d, da + (uint64_t)(st * 64), db + (uint64_t)(st * 2), kIdescTf32, (kb | st) ? 1u : 0u);
                            }
                            if (kTwo) {
                                umma_commit_2sm(bar_bempty + b_st * 8, 3);
                                if (last_nt && kb == num_kb - 1) umma_commit_2sm(bar_aempty + tslot * 8, 3);
                            } else {
                                if (cs == 1) umma_commit(bar_bempty + b_st * 8);
                                else umma_commit_mc(bar_bempty + b_st * 8, cmask);
                                if (last_nt && kb == num_kb - 1) umma_commit(bar_aempty + tslot * 8);
                            }
                        }
                        __syncwarp();
                        if (++b_st == (uint32_t)b_stages) { b_st = 0; b_ph ^= 1; }
                        continue;
                    }
                    uint32_t slot = a_slot0 + kb, a_phk = a_ph;
                    if (slot >= (uint32_t)a_slots) { slot -= a_slots; a_phk ^= 1; }
                    if (nt == 0) mbar_wait(bar_afull + slot * 8, a_phk);
                    if (lane == 0 && kb == 0) VQB_TRACE(1, 2, rd * num_n_tiles + nt);   // MMA: A chunk 0 ready
                    mbar_wait(bar_bfull + b_st * 8, b_ph);
                    if (lane == 0 && kb == 0) VQB_TRACE(1, 3, rd * num_n_tiles + nt);   // MMA: codebook stage 0 ready
                    tc_fence_after();
                    if (elect_one()) {
                        // descriptors address shared memory in 16-byte units: +1024 per 16 KiB A chunk, +2048 per 32 KiB B stage,
                        // +2 per K step of 16 bf16 (32 bytes) inside the 128-byte swizzle row
                        const uint64_t da = dA0 + (uint64_t)(slot * (A_CHUNK_BYTES >> 4));
                        const uint64_t db = dB0 + (uint64_t)(b_st * (kStageBytes >> 4));
                        if (kTwo) {
                            umma_bf16_2sm(tmem_d, da, db, kIdesc2, kb ? 1u : 0u);
                            umma_bf16_2sm(tmem_d, da + 2, db + 2, kIdesc2, 1u);
                            umma_bf16_2sm(tmem_d, da + 4, db + 4, kIdesc2, 1u);
                            umma_bf16_2sm(tmem_d, da + 6, db + 6, kIdesc2, 1u);
                            umma_commit_2sm(bar_bempty + b_st * 8, 3);
                            if (last_nt) umma_commit_2sm(bar_aempty + slot * 8, 3);
                        } else {
                            umma_bf16(tmem_d, da, db, kIdesc, kb ? 1u : 0u);
                            umma_bf16(tmem_d, da + 2, db + 2, kIdesc, 1u);
                            umma_bf16(tmem_d, da + 4, db + 4, kIdesc, 1u);
                            umma_bf16(tmem_d, da + 6, db + 6, kIdesc, 1u);
                            if (cs == 1) umma_commit(bar_bempty + b_st * 8);
                            else umma_commit_mc(bar_bempty + b_st * 8, cmask);
                            if (last_nt) umma_commit(bar_aempty + slot * 8);
                        }
                    }
                    __syncwarp();
                    if (++b_st == (uint32_t)b_stages) { b_st = 0; b_ph ^= 1; }
                }
                // the bias K step: acc -= |e_k|^2 / 2, then hand the accumulator to the epilogue
                mbar_wait(bar_efull + es * 8, e_ph);
                if (lane == 0) VQB_TRACE(1, 4, rd * num_n_tiles + nt);        // MMA: bias operand ready, tile about to be committed
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t eh_u = sEH_u + es * kEhBytes;
                    if (kTwo) {
                        umma_bf16_2sm(tmem_d, dAX, make_desc_noswz(eh_u, sZero_u - eh_u, 128), kIdesc2, 1u);
                        umma_commit_2sm(bar_eempty + es * 8, 3);
                        umma_commit_2sm(bar_tfull + as * 8, 3);
                    } else {
                        umma_bf16(tmem_d, dAX, make_desc_noswz(eh_u, sZero_u - eh_u, 128), kIdesc, 1u);
                        umma_commit(bar_eempty + es * 8);
                        umma_commit(bar_tfull + as * 8);
                    }
                }
                __syncwarp();
                if (++es == (uint32_t)eh_slots) { es = 0; e_ph ^= 1; }
                as ^= 1;
                if (as == 0) t_ph ^= 1;
            }
            a_slot0 += num_kb;
            if (a_slot0 >= (uint32_t)a_slots) { a_slot0 -= a_slots; a_ph ^= 1; }
        }
        }
    } else if (kTf32 && warp == ALOAD_WARP) {
        // ================================================================ tf32: loader of the fp32 A tiles.  One tile = four boxes
        // [D dims][32 frames] (128-byte swizzle with 32-byte atoms), frame group fg at fg * D * 128 bytes of the slot: the MN-major
        // canonical layout of a 32-bit operand.
        const uint32_t bar_aempty = smem_u32(&bars->a_empty[0]), bar_land = smem_u32(&bars->a_land[0]), sA_u = smem_u32(sA);
        for (int rd = 0; rd < rounds; ++rd) {
            const uint32_t tslot = (uint32_t)rd % (uint32_t)a_slots, t_phase = ((uint32_t)rd / (uint32_t)a_slots) & 1u;
            int mt = (rd * n_clusters + cluster_id) * cs + (int)crank;
            if (mt >= num_m_tiles) mt = 0;                                   // dummy tile: re-read tile 0, nothing is published
            const int b_ = mt / tiles_per_item, w0_ = (mt - b_ * tiles_per_item) * BM;
            mbar_wait_sleep(bar_aempty + tslot * 8, t_phase ^ 1, spin_ns);   // the MMAs that read this slot last are done
            if (elect_one()) {
                mbar_expect_tx(bar_land + tslot * 8, a_slot_bytes);
#pragma unroll
                for (int fg = 0; fg < BM / 32; ++fg) {                       // frames past W arrive as zeros
                    const uint32_t dst = sA_u + tslot * a_slot_bytes + (uint32_t)fg * (uint32_t)D * 128u;
                    if (l2_once) tma_load_3d_once(dst, &tmap_x, bar_land + tslot * 8, w0_ + fg * 32, 0, b_);
                    else tma_load_3d(dst, &tmap_x, bar_land + tslot * 8, w0_ + fg * 32, 0, b_);
                }
            }
            __syncwarp();
        }
    } else if (kTf32 && warp == CONVERT_WARP) {
        // ================================================================ tf32: norm warp.  Lane l measures frames l, l + 32, l + 64,
        // l + 96 of the landed tile (one per frame group: a row of the swizzled box holds 32 frames in a permuted order, so the 32
        // lanes of a load never meet in a bank), publishes the guard bands and only then releases the tile to the MMA issuer - the
        // epilogue therefore finds the bands of a tile as soon as it sees the tile's first accumulator.
        const uint32_t bar_afull = smem_u32(&bars->a_full[0]), bar_land = smem_u32(&bars->a_land[0]);
        const float etmax = sqrtf(__uint_as_float(meta_ro->etmax2_bits)) * 1.0001f;
        const float demax = sqrtf(__uint_as_float(meta_ro->demax2_bits)) * 1.0001f;
        const float emax = sqrtf(__uint_as_float(meta_ro->emax2_bits)) * 1.0001f;
        for (int rd = 0; rd < rounds; ++rd) {
            const uint32_t tslot = (uint32_t)rd % (uint32_t)a_slots, t_phase = ((uint32_t)rd / (uint32_t)a_slots) & 1u;
            mbar_wait_sleep(bar_land + tslot * 8, t_phase, spin_ns >> 1);
            const unsigned char* tile = sA + (size_t)tslot * a_slot_bytes;
            float s2[4] = {0.f, 0.f, 0.f, 0.f}, sd2[4] = {0.f, 0.f, 0.f, 0.f};
            for (int d = 0; d < D; ++d) {
                // 32-byte chunk (8 frames) of row d sits at chunk position (lane / 8) ^ (d % 4)
                const uint32_t off = (uint32_t)d * 128u + ((((uint32_t)lane >> 3) ^ ((uint32_t)d & 3u)) << 5) + (((uint32_t)lane & 7u) << 2);
#pragma unroll
                for (int fg = 0; fg < 4; ++fg) {
                    const float x = *reinterpret_cast<const float*>(tile + (size_t)fg * D * 128 + off);
                    // what the tensor core drops: the low 13 mantissa bits (a round-to-nearest unit would drop no more than that)
                    const float dx = x - __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
                    s2[fg] = fmaf(x, x, s2[fg]);
                    sd2[fg] = fmaf(dx, dx, sd2[fg]);
                }
            }
#pragma unroll
            for (int fg = 0; fg < 4; ++fg) {
                const float xn = sqrtf(s2[fg]) * 1.0001f, dxn = sqrtf(sd2[fg]) * 1.0001f;
                sBand[(rd & 3) * BM + fg * 32 + lane] = 4.0f * (dxn * etmax + xn * demax) * 1.001f +
                                                        8.0f * (float)(D + 16) * 2.3841858e-07f * xn * emax + 4.0e-7f * emax * emax + 1e-30f;
            }
            __syncwarp();
            if (lane == 0) {
                if (kTwo && crank != 0) mbar_arrive_remote(bar_afull + tslot * 8, 0);
                else mbar_arrive(bar_afull + tslot * 8);
            }
        }
    } else if (!kTf32 && kFuse && (warp == CONVERT_WARP || warp == ALOAD_WARP)) {
        // ================================================================ fused operand preparation: two converter warps.
        // Warp 18 owns frames 0..63 of the tile, warp 19 frames 64..127 (two frames per lane).  Warp 19 is also the loader:
        // before it touches sub-chunk g it makes sure the TMA loads up to g + STG_SLOTS - 1 are issued.  A load waits only
        // for BOTH converters to have released its ring slot, so the loader never waits on its own future work.
        const int num_sub = D / SUB_DIMS;
        const int half_row0 = (warp == CONVERT_WARP) ? 0 : BM / 2;
        const bool is_loader = warp == ALOAD_WARP;
        const long long total_sub = (long long)rounds * num_sub;
        long long issued = 0, g = 0;                                             // sub-chunk counters over all tiles of this CTA
        uint32_t a_slot0 = 0, a_ph = 0;
        const uint32_t bar_sfull = smem_u32(&bars->stg_full[0]), bar_sempty = smem_u32(&bars->stg_empty[0]), sS_u = smem_u32(sStg);
        const uint32_t bar_afull = smem_u32(&bars->a_full[0]), bar_aempty = smem_u32(&bars->a_empty[0]);
        const float etmax = sqrtf(__uint_as_float(meta_ro->etmax2_bits)) * 1.0001f;
        const float demax = sqrtf(__uint_as_float(meta_ro->demax2_bits)) * 1.0001f;
        const float emax = sqrtf(__uint_as_float(meta_ro->emax2_bits)) * 1.0001f;
        auto ensure_issued = [&](long long upto) {
            while (issued <= upto && issued < total_sub) {
                const uint32_t sl_ = (uint32_t)(issued % STG_SLOTS), ph_ = (uint32_t)((issued / STG_SLOTS) & 1);
                mbar_wait_sleep(bar_sempty + sl_ * 8, ph_ ^ 1, spin_ns);
                const int rd_ = (int)(issued / num_sub), sub_ = (int)(issued - (long long)rd_ * num_sub);
                int mt_ = (rd_ * n_clusters + cluster_id) * cs + (int)crank;
                if (mt_ >= num_m_tiles) mt_ = 0;                                 // dummy tile: re-read tile 0, nothing is published
                const int b_ = mt_ / tiles_per_item, w0_ = (mt_ - b_ * tiles_per_item) * BM;
                if (elect_one()) {
                    mbar_expect_tx(bar_sfull + sl_ * 8, STG_BYTES);
                    // frames past W arrive as zeros
                    if (l2_once) tma_load_3d_once(sS_u + sl_ * STG_BYTES, &tmap_x, bar_sfull + sl_ * 8, w0_, sub_ * SUB_DIMS, b_);
                    else tma_load_3d(sS_u + sl_ * STG_BYTES, &tmap_x, bar_sfull + sl_ * 8, w0_, sub_ * SUB_DIMS, b_);
                }
                __syncwarp();
                ++issued;
            }
        };
        for (int rd = 0; rd < rounds; ++rd) {
            float s2[2] = {0.f, 0.f}, sd2[2] = {0.f, 0.f};                       // |x|^2 and |x - bf16(x)|^2 of this lane's 2 frames
            const int r0 = half_row0 + 2 * lane;
            int sub = 0;
            for (int kb = 0; kb < num_kb; ++kb) {
                uint32_t slot = a_slot0 + kb, a_phk = a_ph;
                if (slot >= (uint32_t)a_slots) { slot -= a_slots; a_phk ^= 1; }
                if (is_loader) ensure_issued(g + STG_SLOTS - 1);
                mbar_wait_sleep(bar_aempty + slot * 8, a_phk ^ 1, spin_ns);      // the MMAs that read this ring slot last are done
                unsigned char* chunk = sA + (size_t)slot * A_CHUNK_BYTES;
                for (int q = 0; q < BK / SUB_DIMS; ++q, ++sub) {
                    if (sub < num_sub) {
                        if (is_loader) ensure_issued(g + STG_SLOTS - 1);
                        const uint32_t sg = (uint32_t)(g % STG_SLOTS), sg_ph = (uint32_t)((g / STG_SLOTS) & 1);
                        mbar_wait_sleep(bar_sfull + sg * 8, sg_ph, spin_ns >> 1);
                        const float* stg = reinterpret_cast<const float*>(sStg + (size_t)sg * STG_BYTES) + r0;
#pragma unroll
                        for (int h = 0; h < 2; ++h) {                             // two 16-byte units (8 dims) per sub-chunk
                            float2 v[8];
#pragma unroll
                            for (int d = 0; d < 8; ++d) v[d] = *reinterpret_cast<const float2*>(stg + (h * 8 + d) * BM);
#pragma unroll
                            for (int i = 0; i < 2; ++i) {
                                uint32_t pk[4];
#pragma unroll
                                for (int d = 0; d < 4; ++d) {
                                    const float xa = i ? v[2 * d].y : v[2 * d].x, xb_ = i ? v[2 * d + 1].y : v[2 * d + 1].x;
                                    const __nv_bfloat162 hb = __floats2bfloat162_rn(xa, xb_);
                                    const float2 bk = __bfloat1622float2(hb);
                                    const float e0 = xa - bk.x, e1 = xb_ - bk.y;
                                    s2[i] = fmaf(xa, xa, s2[i]);
                                    s2[i] = fmaf(xb_, xb_, s2[i]);
                                    sd2[i] = fmaf(e0, e0, sd2[i]);
                                    sd2[i] = fmaf(e1, e1, sd2[i]);
                                    pk[d] = *reinterpret_cast<const uint32_t*>(&hb);
                                }
                                const int r = r0 + i, u = 2 * q + h;
                                *reinterpret_cast<uint4*>(chunk + (r >> 3) * 1024 + (r & 7) * 128 + ((u ^ (r & 7)) << 4)) =
                                    make_uint4(pk[0], pk[1], pk[2], pk[3]);
                            }
                        }
                        __syncwarp();
                        if (lane == 0) mbar_arrive(bar_sempty + sg * 8);
                        ++g;
                    } else {                                                      // D % 64 != 0: the rest of the last chunk is zero
#pragma unroll
                        for (int h = 0; h < 2; ++h)
#pragma unroll
                            for (int i = 0; i < 2; ++i) {
                                const int r = r0 + i, u = 2 * q + h;
                                *reinterpret_cast<uint4*>(chunk + (r >> 3) * 1024 + (r & 7) * 128 + ((u ^ (r & 7)) << 4)) = make_uint4(0u, 0u, 0u, 0u);
                            }
                    }
                }
                if (kb == num_kb - 1) {                                           // guard band of this tile's frames (see latent_prep_bf16_kernel)
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        const float xn = sqrtf(s2[i]) * 1.0001f, dxn = sqrtf(sd2[i]) * 1.0001f;
                        sBand[(rd & 3) * BM + r0 + i] = 4.0f * (dxn * etmax + xn * demax) * 1.001f +
                                                         8.0f * (float)(D + 16) * 2.3841858e-07f * xn * emax + 4.0e-7f * emax * emax + 1e-30f;
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes -> visible to the tensor core
                __syncwarp();
                if (lane == 0) {
                    if (kTwo && crank != 0) mbar_arrive_remote(bar_afull + slot * 8, 0);
                    else mbar_arrive(bar_afull + slot * 8);
                    VQB_TRACE(warp == CONVERT_WARP ? 2 : 3, 0, rd * num_kb + kb);   // converter: A chunk published
                }
            }
            a_slot0 += num_kb;
            if (a_slot0 >= (uint32_t)a_slots) { a_slot0 -= a_slots; a_ph ^= 1; }
        }
    } else if (kTail && warp >= TAIL_WARP0) {
        // ================================================================ fused tail: lane = frame, one tile behind the epilogue.
        // Pair sweeps settle the index of the frames with more than one shortlisted code: the warp's (frame, code) pairs are
        // dealt out one per lane (no divergence over shortlist lengths), each lane evaluates the reference's fp32 distance
        // fl(|x|^2 + fl(|e|^2 - 2 x.e)) of its pair and hands it back through shared memory; ties -> lowest index.
        // The final pass gathers the codeword, writes the straight-through value and accumulates SSE / histogram /
        // residual sums.  Every pass reads the tile's latents from the ring (see tail_stream).
        const int tw = warp - TAIL_WARP0;
        const int f = tw * 32 + lane;
        const int nbox = D / TAIL_CHUNK;
        const bool resid_v4 = tail.resid && (reinterpret_cast<uintptr_t>(tail.resid) & 15) == 0;
        float2* pair = sPair + tw * 32;
        double* acc_sse = reinterpret_cast<double*>(sPair + TAIL_WARPS * 32) + tw;          // per-warp running SSE (lane 0)
        unsigned int* acc_cnt = reinterpret_cast<unsigned int*>(sPair + TAIL_WARPS * 32 + 4) + 2 * tw;   // rescored / shortlisted
        if (lane == 0) { *acc_sse = 0.0; acc_cnt[0] = 0u; acc_cnt[1] = 0u; }
        const uint64_t pol_once = 0;
        const uint32_t bar_full = smem_u32(&tbars->tx_full[0]), bar_empty = smem_u32(&tbars->tx_empty[0]);
        const bool loader = tw == 0;
        RingPos pos{0u, 0u};
        if (loader && lane == 0) prefetch_tmap(&tmap_xt);
        const int mt_step = n_clusters * cs;
        int mt = cluster_id * cs + (int)crank;
        for (int rd = 0; rd < rounds; ++rd, mt += mt_step) {
            const int tbuf = rd & 1;
            mbar_wait(smem_u32(&tbars->tail_full[tbuf]), (rd >> 1) & 1);
            if (mt < num_m_tiles) {
                const int b = mt / tiles_per_item, w0 = (mt - b * tiles_per_item) * BM;
                const int cnt = sCandCnt[tbuf * BM + f];
                const uint16_t* cl = sCand + (tbuf * BM + tw * 32) * kCandFill;   // this warp's 32 shortlists
                const int sweeps = tail_max_sweeps(sCandCnt + tbuf * BM, lane);
                if (loader) {                                // prime the ring: the first TX_AHEAD boxes of this tile
                    const int total_boxes = (sweeps + 1) * nbox;
                    for (int a = 0; a < TX_AHEAD && a < total_boxes; ++a) {
                        int gi = a;
                        while (gi >= nbox) gi -= nbox;
                        tail_issue(pos, a, gi * TAIL_CHUNK, b, w0, smem_u32(sTx), bar_full, bar_empty, &tmap_xt);
                    }
                }
                int k = cnt ? (int)cl[lane * kCandFill] : 0;
                const bool need = cnt > 1 && sweeps > 0;
                const int np = need ? cnt : 0;
                int incl = np;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int t = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += t;
                }
                const int excl = incl - np, total = __shfl_sync(0xffffffffu, incl, 31);
                float bd = 0.f;
                int bk = -1;
                for (int sw = 0; sw < sweeps; ++sw) {
                    const int s0 = sw * 32, p = s0 + lane;
                    const bool act = p < total;
                    int fl = 0;                              // pair p belongs to the frame whose [excl, incl) holds it
#pragma unroll
                    for (int st = 16; st > 0; st >>= 1) {
                        const int t = __shfl_sync(0xffffffffu, incl, fl + st - 1);
                        if (t <= p) fl += st;
                    }
                    const int ci = p - __shfl_sync(0xffffffffu, excl, fl);
                    const int kc = act ? (int)cl[fl * kCandFill + ci] : 0;
                    float dot = 0.f, x2 = 0.f;
                    tail_stream<false>(sTx, bar_full, bar_empty, pos, nbox, act, tw * 32 + fl,
                                       tail.codebook + (size_t)kc * D, lane, dot, x2, W, nullptr, nullptr,
                                       false, pol_once, loader, (sweeps - sw) * nbox, b, w0, &tmap_xt);
                    const float dist = act ? ref_distance(x2, tail.e2[kc], dot) : 0.f;
                    pair[lane] = make_float2(dist, __int_as_float(kc));
                    __syncwarp();
                    if (need) {
                        const int lo = max(excl, s0), hi = min(incl, s0 + 32);
                        for (int pp = lo; pp < hi; ++pp) {
                            const float2 v = pair[pp - s0];
                            const int kk = __float_as_int(v.y);
                            if (better(v.x, kk, bd, bk)) { bd = v.x; bk = kk; }
                        }
                    }
                    __syncwarp();
                }
                if (need) k = bk;
                const size_t at = (size_t)b * D * W + (size_t)(w0 + f);   // element (b, 0, w) of the [B, D, W] tensors
                if (cnt) {
                    tail.idx_out[(int64_t)b * W + w0 + f] = (int64_t)k;
                    atomicAdd(tail.counts + k, 1);
                }
                float fs = 0.f, unused = 0.f;
                tail_stream<true>(sTx, bar_full, bar_empty, pos, nbox, cnt != 0, f,
                                  tail.codebook + (size_t)k * D, lane, fs, unused, W,
                                  tail.q_out ? tail.q_out + at : nullptr,
                                  tail.resid ? tail.resid + (size_t)k * D : nullptr, resid_v4, pol_once, loader, 0, b,
                                  w0, &tmap_xt);
                // per-warp running totals live in shared memory (lane 0 only): registers are scarce in this kernel
                fs = cnt ? fs : 0.f;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) fs += __shfl_xor_sync(0xffffffffu, fs, o);
                const unsigned int n_resc = __popc(__ballot_sync(0xffffffffu, need));
                const unsigned int n_short = __reduce_add_sync(0xffffffffu, (unsigned int)cnt);
                if (lane == 0) { *acc_sse += (double)fs; acc_cnt[0] += n_resc; acc_cnt[1] += n_short; }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&tbars->tail_empty[tbuf]));
        }
        asm volatile("bar.sync 3, 128;" ::: "memory");
        if (tw == 0 && lane == 0) {
            const double* a = reinterpret_cast<const double*>(sPair + TAIL_WARPS * 32);
            tail.sse_partials[blockIdx.x] = (a[0] + a[1]) + (a[2] + a[3]);
            const unsigned int* c = reinterpret_cast<const unsigned int*>(sPair + TAIL_WARPS * 32 + 4);
            atomicAdd(&meta->rescored, (unsigned long long)(c[0] + c[2] + c[4] + c[6]));
            atomicAdd(&meta->shortlisted, (unsigned long long)(c[1] + c[3] + c[5] + c[7]));
        }
    } else if (kGrouped && warp < EPI_WARPS) {
        // ================================================================ slab-queue epilogue (small codebooks)
        const int ew = warp - EPI_WARP0;
        const int quarter = warp & 3;             // TMEM lanes this warp may touch: 32*quarter .. +31
        const int colq = ew >> 2;                 // which 64-column quarter of every tile: two 32-column slabs
        const int row_in_tile = quarter * 32 + lane;
        const uint32_t t_lane = (uint32_t)(quarter * 32) << 16;
        const bool force_fallback = meta->cb_nonfinite != 0;
        int* smax = sCnt + 2 * BM + row_in_tile;  // pooled running maximum of the frame (ordered int)
        const uint32_t q_u = smem_u32(sWarpQ + (size_t)ew * wq_cap * SQ_ENTRY);
        unsigned char* spill = reinterpret_cast<unsigned char*>(ev_scratch + ((size_t)blockIdx.x * EPI_THREADS + (size_t)ew * 32) * (EV_CAP * EV_WORDS));
        if (colq == 0) { sCnt[row_in_tile] = 0; sCnt[BM + row_in_tile] = 0; *smax = f2ord(-INFINITY); }
        asm volatile("bar.sync 1, 512;" ::: "memory");
        const bool tr = lane == 0 && (ew == 0 || ew == 13);
        const int trole = 4 + (ew != 0);
        (void)tr; (void)trole;
        uint32_t n_it = 0;
        bool pre_ok = false;
        for (int rd = 0; rd < rounds; ++rd) {
            const int mt = (rd * n_clusters + cluster_id) * cs + (int)crank;   // may be >= num_m_tiles: dummy tile
            const int b = mt / tiles_per_item;
            const int64_t w = (int64_t)(mt - b * tiles_per_item) * BM + row_in_tile;
            const int64_t row = (mt < num_m_tiles && w < W) ? (int64_t)b * W + w : N;   // global frame index, or N: no frame
            float band = 0.f, hband = 0.f, thr = -INFINITY, mymax = -INFINITY;
            uint32_t qn = 0;                      // fill of this warp's queue (warp-uniform)
            int mine = 0;                         // entries this lane appended
            for (int nt = 0; nt < num_n_tiles; ++nt, ++n_it) {
                const uint32_t as = n_it & 1u, ph = (n_it >> 1) & 1u;
                if (tr) VQB_TRACE(trole, 0, n_it);
                if (!pre_ok) mbar_wait(smem_u32(&bars->tmem_full[as]), ph);
                tc_fence_after();
                if (tr) VQB_TRACE(trole, 1, n_it);
                const uint32_t taddr = tmem_base + t_lane + as * BN + (uint32_t)colq * COLS_PER_WARP;
                const int code0 = nt * BN + colq * COLS_PER_WARP;
                uint32_t ra[32], rb[32];
                if (nt == 0) {                    // the converter published this tile's bands before the first MMA could start
                    band = sBand[(rd & 3) * BM + row_in_tile];
                    hband = 0.5f * band;          // the band in accumulator units (acc = -score / 2)
                    // first codebook tile of the round: its maximum first (the accumulator stays in TMEM), then the scan
                    tmem_ld32(taddr, ra);
                    tmem_ld_wait(ra);
                    tmem_ld32(taddr + 32, rb);
                    float pre = slab_max32(ra);
                    tmem_ld_wait(rb);
                    pre = fmaxf(pre, slab_max32(rb));
                    mymax = pre;
                    thr = pre - hband;
                    atomicMax(smax, f2ord(pre));
                }
                thr = fmaxf(thr, ord2f(*reinterpret_cast<volatile int*>(smax)) - hband);   // what the frame's other threads have seen
                tmem_ld32(taddr, ra);
                tmem_ld_wait(ra);
                tmem_ld32(taddr + 32, rb);        // in flight behind the scan of the first slab
                if (tr && ew == 0) VQB_TRACE(6, 0, n_it);
                scan_slab_sq(ra, code0, hband, thr, mymax, smax, q_u, (uint32_t)wq_cap, spill, qn, mine, lane);
                if (tr && ew == 0) VQB_TRACE(6, 1, n_it);
                tmem_ld_wait(rb);
                // the last slab of this accumulator stage is in registers: hand the stage back before scanning it
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if (kTwo && crank != 0) mbar_arrive_remote(smem_u32(&bars->tmem_empty[as]), 0);   // the leader issues the pair's MMAs
                    else mbar_arrive(smem_u32(&bars->tmem_empty[as]));
                }
                if (tr) VQB_TRACE(trole, 2, n_it);
                pre_ok = mbar_test(smem_u32(&bars->tmem_full[(n_it + 1u) & 1u]), ((n_it + 1u) >> 1) & 1u);
                scan_slab_sq(rb, code0 + 32, hband, thr, mymax, smax, q_u, (uint32_t)wq_cap, spill, qn, mine, lane);
                if (tr && ew == 0) VQB_TRACE(6, 3, n_it);
            }
            // ---- end of the round: the four warps that share these 32 frames meet (the other lane quarters run on), every warp deals
            // ITS queue's entries one per lane and filters them against the final maximum of the frame each entry belongs to
            if (tr) VQB_TRACE(trole, 3, n_it);
            sMin[colq * BM + row_in_tile] = mymax;
            if (!(band < INFINITY) || mine > SQ_LANE_CAP) sCnt[BM + row_in_tile] = 1;   // non-finite latent / flooded: exact search
            asm volatile("bar.sync %0, 128;" ::"r"(4 + quarter) : "memory");
            if (tr) VQB_TRACE(trole, 6, n_it);
            {
                if (qn > (uint32_t)wq_cap + (uint32_t)SQ_SPILL) {   // (only when no lane cap applies: unreachable with 32 lanes x SQ_LANE_CAP <= capacity)
                    sCnt[BM + row_in_tile] = 1;
                    qn = (uint32_t)wq_cap + (uint32_t)SQ_SPILL;
                }
                const unsigned char* qb = sWarpQ + (size_t)ew * wq_cap * SQ_ENTRY;
                for (uint32_t e = (uint32_t)lane; e < qn; e += 32u) {
                    const bool in_sm = e < (uint32_t)wq_cap;
                    const unsigned char* en = in_sm ? qb + (size_t)e * SQ_ENTRY : spill + (size_t)(e - (uint32_t)wq_cap) * SQ_ENTRY;
                    const uint2 hd = *reinterpret_cast<const uint2*>(en + 128);
                    const int r = quarter * 32 + (int)(hd.y >> 16);
                    const float gmax = fmaxf(fmaxf(sMin[r], sMin[BM + r]), fmaxf(sMin[2 * BM + r], sMin[3 * BM + r]));
                    const float cutoff = gmax - 0.5f * sBand[(rd & 3) * BM + r];
                    if (__uint_as_float(hd.x) >= cutoff) {
                        const int k0 = (int)(hd.y & 0xFFFFu);
                        unsigned pm = 0;          // which of the 32 codes pass: ONE shared-memory atomic per entry
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            const uint4 a = *reinterpret_cast<const uint4*>(en + 16 * q);
                            pm |= (__uint_as_float(a.x) >= cutoff ? 1u : 0u) << (4 * q);
                            pm |= (__uint_as_float(a.y) >= cutoff ? 2u : 0u) << (4 * q);
                            pm |= (__uint_as_float(a.z) >= cutoff ? 4u : 0u) << (4 * q);
                            pm |= (__uint_as_float(a.w) >= cutoff ? 8u : 0u) << (4 * q);
                        }
                        int pos = atomicAdd(&sCnt[r], __popc(pm));
                        while (pm) {
                            const int j = __ffs((int)pm) - 1;
                            pm &= pm - 1u;
                            if (pos < kCandFill) sList[r * kCandMax + pos] = (uint16_t)(k0 + j);
                            else sCnt[BM + r] = 1;
                            ++pos;
                        }
                    }
                }
            }
            if (colq == 0) *smax = f2ord(-INFINITY);      // nobody reads the pooled maximum between the two barriers
            if (tr) VQB_TRACE(trole, 4, n_it);
            asm volatile("bar.sync %0, 128;" ::"r"(4 + quarter) : "memory");
            if (tr) VQB_TRACE(trole, 5, n_it);
            if (colq == 0) {
                if (row < N) {
                    const int cnt = sCnt[row_in_tile];
                    if (force_fallback || cnt == 0 || sCnt[BM + row_in_tile]) {
                        cand_cnt[row] = kCandFinal;          // the exact search fills in the final code
                        const int fp = atomicAdd(&meta->fallback_count, 1);
                        fallback_rows[fp] = (int)row;
                        best64[fp] = ~0ull;
                        atomicAdd(&meta->fallback_total, 1ull);
                    } else {
                        cand_cnt[row] = (uint8_t)(cnt < kCandFill ? cnt : kCandFill);
                        const uint4* src = reinterpret_cast<const uint4*>(sList + row_in_tile * kCandMax);
                        uint4* dst = reinterpret_cast<uint4*>(cand_idx + (size_t)row * kCandMax);
                        dst[0] = src[0];
                        dst[1] = src[1];
                    }
                }
                sCnt[row_in_tile] = 0;
                sCnt[BM + row_in_tile] = 0;
            }
        }
    } else if (warp < EPI_WARPS) {
        // ================================================================ epilogue
        const int ew = warp - EPI_WARP0;
        const int quarter = warp & 3;            // TMEM lanes this warp may touch: 32*quarter .. +31
        const int colq = ew >> 2;                // which 64-column quarter of every tile
        const int row_in_tile = quarter * 32 + lane;
        const uint32_t t_lane = (uint32_t)(quarter * 32) << 16;
        const bool force_fallback = meta->cb_nonfinite != 0;
        const int et = (warp - EPI_WARP0) * 32 + lane;           // 0..511
        EventStack ev;
        ev.base = ev_scratch + ((size_t)blockIdx.x * EPI_THREADS + et) * (EV_CAP * EV_WORDS);
        ev.sbase = sEv + (size_t)et * ev_sm * EV_WORDS;
        ev.sm = ev_sm;
        ev.n = 0;
        ev.ng = 0;
        ev.pool_u = smem_u32(sWarpQ + (size_t)ew * wq_cap * Q_ENTRY);
        ev.pool_cnt_u = smem_u32(sPoolCnt + ew);
        ev.pool_cap = (kFuse && !kTail) ? (uint32_t)wq_cap : 0u;   // (the pool's resolution reads the bands and writes global shortlists)
        ev.lane = (uint32_t)lane;
        if (lane == 0) sPoolCnt[ew] = 0u;
        if (colq == 0) { sCnt[row_in_tile] = 0; sCnt[BM + row_in_tile] = 0; sCnt[2 * BM + row_in_tile] = f2ord(-INFINITY); }
        int* smax = sCnt + 2 * BM + row_in_tile;
        asm volatile("bar.sync 1, 512;" ::: "memory");
        uint32_t n_it = 0;
        bool pre_ok = false;
        for (int rd = 0; rd < rounds; ++rd) {
            const int mt = (rd * n_clusters + cluster_id) * cs + (int)crank;   // may be >= num_m_tiles: dummy tile
            int64_t row;                                  // global frame index n = b*W + w, or N when this lane has no frame
            if (kFuse) {
                const int b = mt / tiles_per_item;
                const int64_t w = (int64_t)(mt - b * tiles_per_item) * BM + row_in_tile;
                row = (mt < num_m_tiles && w < W) ? (int64_t)b * W + w : N;
            } else {
                row = (int64_t)mt * BM + row_in_tile;
            }
            float band = (!kFuse && row < N) ? band_g[row] : 0.f;
            float hband = 0.5f * band;                    // the band in accumulator units (acc = -score / 2)
            float thr = -INFINITY;
            ev.n = 0;
            ev.ng = 0;
            for (int nt = 0; nt < num_n_tiles; ++nt, ++n_it) {
                const uint32_t as = n_it & 1, ph = (n_it >> 1) & 1;
                if (lane == 0 && (ew == 0 || ew == 13)) VQB_TRACE(4 + (ew != 0), 0, n_it);   // epilogue: about to wait for the accumulator
                if (!pre_ok) mbar_wait(smem_u32(&bars->tmem_full[as]), ph);   // (pre_ok: seen complete before the last slab of the previous tile was scanned)
                if (lane == 0 && (ew == 0 || ew == 13)) VQB_TRACE(4 + (ew != 0), 1, n_it);   // epilogue: accumulator visible
                tc_fence_after();
                if (kFuse && nt == 0) {                   // the converter published this tile's bands before the first MMA could start
                    band = sBand[(rd & 3) * BM + row_in_tile];
                    hband = 0.5f * band;
                }
                const uint32_t taddr = tmem_base + t_lane + as * BN + colq * COLS_PER_WARP;
                const int code0 = nt * BN + colq * COLS_PER_WARP;
                uint32_t ra[32];
                if (nt == 0 && !scores_dbg) {
                    // First codebook tile of a frame tile: the threshold is still -inf and everything would be appended.  Take
                    // the tile's maximum first (the accumulator stays in TMEM), then scan it with a tight threshold.
                    float pre = -INFINITY;
#pragma unroll
                    for (int sb = 0; sb < COLS_PER_WARP / 32; ++sb) {
                        tmem_ld32(taddr + sb * 32, ra);
                        tmem_ld_wait(ra);
                        pre = fmaxf(pre, slab_max32(ra));
                    }
                    thr = fmaxf(thr, pre - hband);
                    atomicMax(smax, f2ord(pre));
                }
#pragma unroll
                for (int sb = 0; sb < COLS_PER_WARP / 32; ++sb) {
                    tmem_ld32(taddr + sb * 32, ra);
                    tmem_ld_wait(ra);
                    if (lane == 0 && ew == 0) VQB_TRACE(6, 2 * sb, n_it);                        // epilogue: slab in registers
                    if (sb == COLS_PER_WARP / 32 - 1) {
                        // the last slab of this accumulator stage is in registers: hand the stage back BEFORE scanning it - the
                        // hand-off round trip (release -> MMA issue -> commit -> wake-up -> read-out), not the scan, is what a
                        // 700-cycle tile at D = 64 waits for
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) {
                            if (kTwo && crank != 0) mbar_arrive_remote(smem_u32(&bars->tmem_empty[as]), 0);   // the leader issues the pair's MMAs
                            else mbar_arrive(smem_u32(&bars->tmem_empty[as]));
                            if (ew == 0 || ew == 13) VQB_TRACE(4 + (ew != 0), 2, n_it);                   // epilogue: stage released
                        }
                        // a first look at the NEXT accumulator before this slab is scanned: with a small codebook the epilogue is the
                        // slower side, the next accumulator is nearly always complete, and the look's latency hides behind the scan
                        pre_ok = mbar_test(smem_u32(&bars->tmem_full[(n_it + 1u) & 1u]), ((n_it + 1u) >> 1) & 1u);
                    }
                    if (scores_dbg) {
                        if (row < N) dump_slab(ra, code0 + sb * 32, K, scores_dbg + (size_t)row * K);
                    } else {
                        scan_slab(ra, (code0 + sb * 32) >> 3, hband, thr, ev, smax);
                    }
                    if (lane == 0 && ew == 0) VQB_TRACE(6, 2 * sb + 1, n_it);                    // epilogue: slab scanned
                }
            }
            // ---- resolve this thread's chunks against the frame's final maximum and publish the shortlist
            if (lane == 0 && (ew == 0 || ew == 13)) VQB_TRACE(4 + (ew != 0), 3, n_it);   // epilogue: sweep scanned, resolution starts
            sMin[colq * BM + row_in_tile] = thr + hband;                     // running maximum of this column quarter
            const int tbuf = rd & 1;
            if (kTail) mbar_wait(smem_u32(&tbars->tail_empty[tbuf]), ((rd >> 1) & 1) ^ 1);   // the tail is done with the tile before last
            // the 128 threads that share these 32 frames (four column quarters of one TMEM lane quarter) meet; the other lane quarters
            // run on - a 512-thread barrier here made every round wait for the slowest of 16 warps (timeline: 700 + 900 cycles of 8 300 at K = 512)
            asm volatile("bar.sync %0, 128;" ::"r"(4 + quarter) : "memory");
            if (lane == 0 && (ew == 0 || ew == 13)) VQB_TRACE(4 + (ew != 0), 6, n_it);   // epilogue: past the first barrier
            if (row < N && !scores_dbg) {
                const float gmax = fmaxf(fmaxf(sMin[row_in_tile], sMin[BM + row_in_tile]),
                                         fmaxf(sMin[2 * BM + row_in_tile], sMin[3 * BM + row_in_tile]));
                const float cutoff = gmax - hband;
                uint16_t* dst = kTail ? sCand + (tbuf * BM + row_in_tile) * kCandFill : cand_idx + (size_t)row * kCandMax;
                bool lost = ev.ng > EV_CAP || !(band < INFINITY);
                // one event: which of its 8 codes pass, ONE shared-memory atomic for them (unrolled over the codes, the returning atomics
                // of the lanes of a warp ran one after the other: 2 700 cycles of a 12 700-cycle round at K = 1024)
                auto take = [&](const uint32_t* en, uint32_t chunk) {
                    const uint4 a0 = *reinterpret_cast<const uint4*>(en), a1 = *reinterpret_cast<const uint4*>(en + 4);
                    const int k0 = (int)chunk * 8;
                    const uint32_t av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                    unsigned pm = 0;
#pragma unroll
                    for (int j = 0; j < 8; ++j) pm |= (__uint_as_float(av[j]) >= cutoff) ? (1u << j) : 0u;
                    int pos = atomicAdd(&sCnt[row_in_tile], __popc(pm));
                    while (pm) {
                        const int j = __ffs((int)pm) - 1;
                        pm &= pm - 1u;
                        if (pos < kCandFill) dst[pos] = (uint16_t)(k0 + j);
                        else lost = true;
                        ++pos;
                    }
                };
                if (lane == 0 && ew == 0) VQB_TRACE(7, 0, n_it);
                {   // own shared-memory slots (at most 3): every entry is read up front (one shared-memory latency, not one per entry
                    // and per dependent step) and the thread asks for all its shortlist positions with ONE atomic
                    unsigned pm[3];
                    uint32_t ch[3];
                    int total = 0;
#pragma unroll
                    for (int i = 0; i < 3; ++i) {
                        pm[i] = 0u;
                        ch[i] = 0u;
                        if (i < ev.n) {
                            const uint32_t* en = ev.own(i);
                            const uint2 hd = *reinterpret_cast<const uint2*>(en + 8);
                            const uint4 a0 = *reinterpret_cast<const uint4*>(en), a1 = *reinterpret_cast<const uint4*>(en + 4);
                            const uint32_t av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
                            for (int j = 0; j < 8; ++j) pm[i] |= (__uint_as_float(av[j]) >= cutoff) ? (1u << j) : 0u;
                            ch[i] = hd.y;
                            total += __popc(pm[i]);
                        }
                    }
                    if (total) {
                        int pos = atomicAdd(&sCnt[row_in_tile], total);
#pragma unroll
                        for (int i = 0; i < 3; ++i) {
                            unsigned m = pm[i];
                            while (m) {
                                const int j = __ffs((int)m) - 1;
                                m &= m - 1u;
                                if (pos < kCandFill) dst[pos] = (uint16_t)(ch[i] * 8u + (uint32_t)j);
                                else lost = true;
                                ++pos;
                            }
                        }
                    }
                }
                if (lane == 0 && ew == 0) VQB_TRACE(7, 1, n_it);
                const int n_g = ev.ng < EV_CAP ? ev.ng : EV_CAP;
                for (int e0 = 0; e0 < n_g; e0 += 8) {        // global stack (rare): headers of 8 events fetched together
                    uint2 hd[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u)
                        hd[u] = (e0 + u < n_g) ? *reinterpret_cast<const uint2*>(ev.glob(e0 + u) + 8) : make_uint2(0xff800000u, 0u);
#pragma unroll
                    for (int u = 0; u < 8; ++u)
                        if (__uint_as_float(hd[u].x) >= cutoff) take(ev.glob(e0 + u), hd[u].y);
                }
                if (lost) sCnt[BM + row_in_tile] = 1;
            }
            if (lane == 0 && ew == 0) VQB_TRACE(7, 2, n_it);
            if (kFuse && !kTail && wq_cap > 0 && !scores_dbg) {
                // the warp's overflow pool: entries dealt one per lane; an entry belongs to the frame of the lane that appended it
                uint32_t pn;
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(pn) : "r"(ev.pool_cnt_u) : "memory");
                if (pn > (uint32_t)wq_cap) pn = (uint32_t)wq_cap;
                for (uint32_t e0 = 0; e0 < pn; e0 += 32u) {
                    const uint32_t e = e0 + (uint32_t)lane;
                    const bool act = e < pn;
                    const uint32_t* en = reinterpret_cast<const uint32_t*>(sWarpQ + ((size_t)ew * wq_cap + (act ? e : 0u)) * Q_ENTRY);
                    const uint4 hd = *reinterpret_cast<const uint4*>(en + 8);            // chunk maximum, chunk id, lane, 0
                    const int src = act ? (int)hd.z : 0;
                    const int64_t row_e = __shfl_sync(0xffffffffu, row, src);            // global frame of that lane (N: none)
                    const int r = quarter * 32 + src;
                    if (act && row_e < N) {
                        const float gmax = fmaxf(fmaxf(sMin[r], sMin[BM + r]), fmaxf(sMin[2 * BM + r], sMin[3 * BM + r]));
                        const float cutoff = gmax - 0.5f * sBand[(rd & 3) * BM + r];
                        if (__uint_as_float(hd.x) >= cutoff) {
                            const uint4 a0 = *reinterpret_cast<const uint4*>(en), a1 = *reinterpret_cast<const uint4*>(en + 4);
                            const int k0 = (int)hd.y * 8;
                            const uint32_t av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                            unsigned pm = 0;
#pragma unroll
                            for (int j = 0; j < 8; ++j) pm |= (__uint_as_float(av[j]) >= cutoff) ? (1u << j) : 0u;
                            int pos = atomicAdd(&sCnt[r], __popc(pm));
                            uint16_t* dst = cand_idx + (size_t)row_e * kCandMax;
                            while (pm) {
                                const int j = __ffs((int)pm) - 1;
                                pm &= pm - 1u;
                                if (pos < kCandFill) dst[pos] = (uint16_t)(k0 + j);
                                else sCnt[BM + r] = 1;
                                ++pos;
                            }
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) sPoolCnt[ew] = 0u;
            }
            if (colq == 0) *smax = f2ord(-INFINITY);       // nobody reads the pooled maximum between the two barriers
            if (lane == 0 && (ew == 0 || ew == 13)) VQB_TRACE(4 + (ew != 0), 4, n_it);   // epilogue: own stack resolved
            asm volatile("bar.sync %0, 128;" ::"r"(4 + quarter) : "memory");
            if (lane == 0 && (ew == 0 || ew == 13)) VQB_TRACE(4 + (ew != 0), 5, n_it);   // epilogue: past the second barrier
            if (colq == 0) {
                int tcnt = 0;                                    // what the tail warps get: 0 = frame is not theirs
                if (row < N && !scores_dbg) {
                    const int cnt = sCnt[row_in_tile];
                    if (force_fallback || cnt == 0 || sCnt[BM + row_in_tile]) {
                        if (!kTail) cand_cnt[row] = kCandFinal;   // the exact search fills in the final code
                        const int fp = atomicAdd(&meta->fallback_count, 1);
                        fallback_rows[fp] = (int)row;
                        best64[fp] = ~0ull;                      // the sliced exact search meets here through atomicMin
                        atomicAdd(&meta->fallback_total, 1ull);
                    } else {
                        tcnt = cnt < kCandFill ? cnt : kCandFill;
                        if (!kTail) cand_cnt[row] = (uint8_t)tcnt;
                    }
                }
                sCnt[row_in_tile] = 0;
                sCnt[BM + row_in_tile] = 0;
                if (kTail) {
                    sCandCnt[tbuf * BM + row_in_tile] = (uint8_t)tcnt;
                    __syncwarp();
                    if (lane == 0) mbar_arrive(smem_u32(&tbars->tail_full[tbuf]));   // release: the shortlists were written before bar 2
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (cs > 1) cluster_sync_all();   // no CTA may retire while a peer can still multicast into it or arrive on its barriers
    if (warp == ALLOC_WARP) {
        tc_fence_after();
        if (kTwo) tmem_dealloc_2sm(tmem_base, 512);
        else tmem_dealloc(tmem_base, 512);
    }
}

// ---------------------------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn) return fn;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !p)
        return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(p);
    return fn;
}

// rows x D bf16, row-major; box = 64 columns (128 B) x box_rows, 128B swizzle, out-of-bounds reads return zeros
static int make_map(CUtensorMap* map, const void* base, uint64_t rows, uint64_t D, uint32_t box_rows) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return VQB_E_DEVICE; }
    const cuuint64_t dims[2] = {D, rows};
    const cuuint64_t strides[1] = {D * 2};
    const cuuint32_t box[2] = {(cuuint32_t)BK, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r); return 1000 + (int)r; }
    return 0;
}

// fp32 codebook [K, D] row-major as the tf32 B operand: box = 32 dims (128 B) x box_rows codes, 128B swizzle; rows past K and
// dims past D read as zeros (their bias operand keeps padded codes from ever winning)
static int make_map_f32(CUtensorMap* map, const void* base, uint64_t rows, uint64_t D, uint32_t box_rows) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return VQB_E_DEVICE; }
    const cuuint64_t dims[2] = {D, rows};
    const cuuint64_t strides[1] = {D * 4};
    const cuuint32_t box[2] = {32, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(codebook fp32) failed with CUresult %d", (int)r); return 1000 + (int)r; }
    return 0;
}

// bias operand [K_pad, 8] bf16 (16-byte rows, no swizzle): box = 128 codes = one CTA's half of a codebook tile
static int make_map_eh(CUtensorMap* map, const void* base, uint64_t rows) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return VQB_E_DEVICE; }
    const cuuint64_t dims[2] = {8, rows};
    const cuuint64_t strides[1] = {16};
    const cuuint32_t box[2] = {8, (cuuint32_t)(BN / 2)};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(eh) failed with CUresult %d", (int)r); return 1000 + (int)r; }
    return 0;
}

// fp32 latents in the reference's [B, D, W] layout: box = 128 frames x box_dims dims of one batch item, zero fill past W
static int make_map_z(CUtensorMap* map, const float* z, uint64_t B, uint64_t D, uint64_t W, uint32_t box_dims = SUB_DIMS) {
    return make_latent_map(map, z, B, D, W, (uint32_t)BM, box_dims);
}

}  // namespace tc

// 3-D TMA view of the [B, D, W] fp32 latents (or of `quantized`): box = box_frames x box_dims x 1, out-of-range frames read as
// zeros and are clipped on stores.  Needs 16-byte global strides: W % 4 == 0 and a 16-byte aligned base.
int make_latent_map(CUtensorMap* map, const float* z, uint64_t B, uint64_t D, uint64_t W, uint32_t box_frames, uint32_t box_dims,
                    bool swizzle128, bool atom32) {
    tc::EncodeTiledFn fn = tc::get_encode_fn();
    if (!fn) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return VQB_E_DEVICE; }
    const cuuint64_t dims[3] = {W, D, B};
    const cuuint64_t strides[2] = {W * 4, D * W * 4};
    const cuuint32_t box[3] = {box_frames, box_dims, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    // L2 promotion no wider than a box row: with 128-byte rows (the tail's 32-frame boxes) a 256-byte promotion also fetches the
    // neighbouring tile's sectors, which another block wants at another time - under the evict-first policy they were gone
    // by then and came from DRAM twice (ncu: +3 GB of evict-first misses at BASELINE config 3)
    CUtensorMapL2promotion promo = box_frames * 4 >= 256 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B
                                 : box_frames * 4 >= 128 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_L2_64B;
    if (const int v = env_get(ENV_TMA_PROMO, -1); v >= 0) {   // experiments: 0 none, 64, 128, 256
        promo = v == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : v == 64 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
              : v == 128 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
    }
    const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(z), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE,
                          swizzle128 ? (atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B) : CU_TENSOR_MAP_SWIZZLE_NONE,
                          promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(z) failed with CUresult %d", (int)r); return 1000 + (int)r; }
    return 0;
}

// Shared-memory plan of one launch: cluster mode, A ring, codebook stages, shared-memory event-stack entries.
// One function decides for the launcher AND for the callers that must know beforehand whether a variant fits
// (tc_can_fuse, tc_fused_tail_fits).
constexpr int kHelperNapNs = 0;        // default nap of waiting helper warps in ns (set by measurement)
constexpr int kSlabQueueDefault = 0;   // 1: the slab-queue epilogue is the default for K <= 1024 (set by measurement)
struct TcPlan {
    bool ok, two, grouped;   // grouped: grouped epilogue (wq_cap = queue entries per warp); else wq_cap = overflow-pool entries per warp
    int cs, a_slots, b_stages, ev_sm, eh_slots, wq_cap;
    size_t smem;
};

// K_pad = 0: the caller only asks whether a plan exists (the ring epilogue is an optimisation, never a requirement)
static TcPlan tc_plan(bool fuse, bool with_tail, bool dbg, int num_m_tiles, int D, bool tf32 = false, int K_pad = 0) {
    using namespace tc;
    TcPlan p{};
    const int num_kb = tf32 ? (D + 31) / 32 : (D + BK - 1) / BK;
    // default: CTA pairs with cta_group::2 MMAs; VQB_TC_MODE=1: cta_group::1 with VQB_TC_CLUSTER-way codebook multicast
    p.two = true;
    if (env_get(ENV_TC_MODE, 0) == 1) p.two = false;
    p.cs = 2;
    p.cs = env_get(ENV_TC_CLUSTER, 2);
    if (p.cs != 1 && p.cs != 2 && p.cs != 4) p.cs = 2;
    if (p.two) p.cs = 2;
    if (num_m_tiles < 2 * p.cs) { p.cs = 1; p.two = false; }
    const size_t stage_bytes = p.two ? B_STAGE_BYTES / 2 : B_STAGE_BYTES;
    // A ring.  Unfused: double-buffered for small D, else one tile.  Fused operand preparation: a full second tile when it fits
    // (the converter then works a whole frame tile ahead of the tensor core: no bubble at the tile boundary, where otherwise
    // the last chunks of the next tile can only be converted inside the last codebook tile), else two spare chunks.
    p.a_slots = num_kb <= 2 ? 2 * num_kb : num_kb;
    if (fuse && num_kb > 2 && num_kb + 2 <= 6) p.a_slots = num_kb + 2;
    // tf32: the A ring holds whole fp32 tiles of 128 frames x D x 4 bytes - two when they leave three codebook stages, else one
    const size_t a_unit = tf32 ? (size_t)BM * D * 4 : (size_t)A_CHUNK_BYTES;
    if (tf32) p.a_slots = 2;
    // Bias-operand ring.  The producer walks the codebook tiles in order and waits for the bias slot of tile j before it
    // loads anything of tile j, so this ring bounds how far ALL operand loads run ahead of the tensor core.  Measured at
    // K = 1024, D = 64 (where a codebook tile is only ~640 tensor-core cycles): 2, 4 and 8 slots give the same kernel time -
    // the operand loads are not what bounds small shapes.
    p.eh_slots = EH_SLOTS;
    if (const int v = env_get(ENV_TC_EHSLOTS, 0); v >= 2 && v <= MAX_EH_SLOTS) p.eh_slots = v;   // experiments
    const size_t fixed_no_a = (size_t)p.eh_slots * EH_SLICE_BYTES + AX_BYTES + ZERO_BYTES + 4 * BM * 4 + 3 * BM * 4 +
                              (fuse ? (tf32 ? 0 : STG_SLOTS * STG_BYTES) + 4 * BM * 4 : 0) +
                              (with_tail ? 2 * BM * kCandFill * 2 + 2 * BM + TAIL_WARPS * 32 * 8 + 128 + TX_SLOTS * TX_BYTES + sizeof(TailBarriers) : 0) +
                              2 * EPI_WARPS * 4 + sizeof(Barriers);   // (+ 128 bytes for the pool fills) no slack: the dynamic segment starts 1024-byte aligned (no static shared memory in this kernel)
    if (!tf32 && fuse && !with_tail && num_kb > 2 && 2 * num_kb <= MAX_A_SLOTS) {
        // the full second tile must leave three codebook stages (it does for CTA pairs at D = 256: 3 x 16 KiB half-tile stages)
        bool full_second_tile = fixed_no_a + (size_t)2 * num_kb * A_CHUNK_BYTES + 3 * stage_bytes <= 227 * 1024;
        if (const int v = env_get(ENV_TC_ASLOTS, -1); v >= 0) full_second_tile = full_second_tile && v >= 2 * num_kb;   // experiments
        if (full_second_tile) p.a_slots = 2 * num_kb;   // pays with one codebook stage (3 instead of 4: measured equal)
    }
    if (tf32 && fixed_no_a + 2 * a_unit + 3 * stage_bytes > 227 * 1024) p.a_slots = 1;
    const size_t fixed = (size_t)p.a_slots * a_unit + fixed_no_a;
    // Shared-memory part of the event stacks, from what four codebook stages leave: own slots of every epilogue thread (24 KiB per
    // level) and an overflow pool per warp (EventStack).  With room for 72 KiB (D = 64): two own slots + 32 pool entries per warp.
    const size_t ev_entry_bytes = (size_t)EPI_THREADS * EV_WORDS * 4;   // one entry for every epilogue thread: 24 KiB
    const size_t pool_entry_bytes = (size_t)EPI_WARPS * Q_ENTRY;        // one pool entry for every epilogue warp: 768 bytes
    p.ev_sm = 0;
    p.wq_cap = 0;
    size_t wq_bytes = 0;
    if (!with_tail && !dbg && 227 * 1024 > fixed + 4 * stage_bytes) {
        const size_t room = 227 * 1024 - fixed - 4 * stage_bytes;
        const bool pool_ok = fuse && env_get(ENV_TC_EVSM, -1) != -2;        // VQB_TC_EVSM=-2 (experiments): no pool
        if (pool_ok && room >= 2 * ev_entry_bytes + 32 * pool_entry_bytes) { p.ev_sm = 2; p.wq_cap = 32; }
        else if (pool_ok && room >= 2 * ev_entry_bytes + 16 * pool_entry_bytes) { p.ev_sm = 2; p.wq_cap = 16; }
        else if (pool_ok && room >= ev_entry_bytes + 32 * pool_entry_bytes) { p.ev_sm = 1; p.wq_cap = 32; }
        else { p.ev_sm = (int)(room / ev_entry_bytes); if (p.ev_sm > 3) p.ev_sm = 3; }
        wq_bytes = (size_t)p.wq_cap * pool_entry_bytes;
    }
    if (const int v = env_get(ENV_TC_EVSM, -1); v >= 0 && v < p.ev_sm) p.ev_sm = v;   // experiments
    // Slab-queue epilogue (small codebooks, K <= 1024; see scan_slab_sq): per-warp queues of raw slabs + the tile's shortlists instead of
    // the per-thread stacks.  About 22 entries per warp and frame-tile round at K = 1024 (fewer at K = 512); what does not fit the
    // shared-memory queue spills into the warp's global scratch.  VQB_TC_EPI=0 / 1 (experiments) forces the per-thread stacks / this form.
    if (fuse && !with_tail && !dbg && K_pad > 0 && K_pad <= 1024 && env_get(ENV_TC_EPI, kSlabQueueDefault) == 1) {
        const size_t list_bytes = (size_t)BM * kCandMax * 2;
        const size_t base = fixed + 4 * stage_bytes + list_bytes;
        if (227 * 1024 > base) {
            int cap = (int)((227 * 1024 - base) / ((size_t)EPI_WARPS * SQ_ENTRY));
            if (cap > 48) cap = 48;
            if (cap >= 24) { p.grouped = true; p.wq_cap = cap; p.ev_sm = 0; wq_bytes = (size_t)EPI_WARPS * cap * SQ_ENTRY + list_bytes; }
        }
    }
    const size_t fixed_ev = fixed + (size_t)p.ev_sm * ev_entry_bytes + wq_bytes;
    p.b_stages = fixed_ev < 227 * 1024 ? (int)((227 * 1024 - fixed_ev) / stage_bytes) : 0;
    if (p.b_stages > (p.two ? 8 : 4)) p.b_stages = p.two ? 8 : 4;
    if (const int v = env_get(ENV_TC_STAGES, 0); v >= 2 && v < p.b_stages) p.b_stages = v;   // experiments
    p.ok = p.b_stages >= 2 && p.a_slots <= MAX_A_SLOTS && !(tf32 && (with_tail || !fuse));
    p.smem = fixed_ev + (size_t)p.b_stages * stage_bytes;
    return p;
}

bool tc_can_fuse(const float* z, int B, int D, int64_t W, int prec) {
    if (env_get(ENV_TC_FUSE, 1) == 0) return false;
    if (prec == VQB_PREC_TF32) {
        // the fp32 tile is the A operand itself: a box of D dims (TMA boxes hold at most 256 rows) per 32 frames
        const int64_t t = (int64_t)B * ((W + tc::BM - 1) / tc::BM);
        return D <= 256 && (D % 8) == 0 && (W % 4) == 0 && (reinterpret_cast<uintptr_t>(z) & 15) == 0 && (W % 128 == 0 || W >= 1024) &&
               tc_plan(true, false, false, (int)(t < (1 << 30) ? t : (1 << 30)), D, true).ok;
    }
    // the staging ring of the fused preparation must leave room for the pipeline: it does not with an eight-chunk A tile
    // (D > 448) outside the CTA-pair mode (whole-tile codebook stages: fewer than four frame tiles, or VQB_TC_MODE=1)
    const int64_t tiles = (int64_t)B * ((W + tc::BM - 1) / tc::BM);
    if (!tc_plan(true, false, false, (int)(tiles < (1 << 30) ? tiles : (1 << 30)), D).ok) return false;
    // 3-D TMA needs 16-byte global strides; short clips would waste most of every 128-frame tile on padding
    return (W % 4) == 0 && (reinterpret_cast<uintptr_t>(z) & 15) == 0 && (D % tc::SUB_DIMS) == 0 && (W % 128 == 0 || W >= 1024);
}

// Opt-in (VQB_TC_TAIL=1).  Measured on B200 at BASELINE config 3: the fused tail makes the whole forward 68.5 ms instead of
// 74.3 ms with the round-1 stand-alone tail, but the search kernel itself slows from 50.9 to 68.5 ms - the epilogue already
// keeps the SM's non-tensor resources busy, so the tail's work costs about as much inside the kernel as outside it.
// A stand-alone tail at HBM speed beats it, hence off by default.
bool tc_fused_tail_fits(int B, int D, int64_t W) {   // does the kTail variant's extra shared memory leave a pipeline?
    const int64_t tiles = (int64_t)B * ((W + tc::BM - 1) / tc::BM);
    return tc_plan(true, true, false, (int)(tiles < (1 << 30) ? tiles : (1 << 30)), D).ok;
}
bool tc_fused_tail_enabled() {
    return env_get(ENV_TC_TAIL, 0) == 1;
}

size_t tc_event_scratch_bytes(int ctas) {
    if (ctas < 2) ctas = 2;
    if (ctas > kTcMaxCtas) ctas = kTcMaxCtas;
    return (size_t)ctas * tc::EPI_THREADS * tc::EV_CAP * tc::EV_WORDS * 4;
}

// ---- optional CUDA-event timing of the stages of vqb_forward, on the launching stream (bench.py's roofline leg) -------
struct TimingSlot { cudaEvent_t start, stop; int stage; };
static bool g_timing = false;
static TimingSlot g_slots[2048];
static int g_slots_used = 0, g_slots_made = 0;

void* stage_timing_begin(cudaStream_t s, int stage) {
    if (!g_timing || g_slots_used >= 2048) return nullptr;
    if (g_slots_used >= g_slots_made) {
        if (cudaEventCreate(&g_slots[g_slots_made].start) != cudaSuccess || cudaEventCreate(&g_slots[g_slots_made].stop) != cudaSuccess)
            return nullptr;
        ++g_slots_made;
    }
    TimingSlot* t = &g_slots[g_slots_used++];
    t->stage = stage;
    cudaEventRecord(t->start, s);
    return t;
}
void stage_timing_end(void* slot, cudaStream_t s) {
    if (slot) cudaEventRecord(static_cast<TimingSlot*>(slot)->stop, s);
}
static TimingSlot* timing_begin(cudaStream_t s) { return static_cast<TimingSlot*>(stage_timing_begin(s, VQB_STAGE_SEARCH)); }
static void timing_end(TimingSlot* t, cudaStream_t s) { stage_timing_end(t, s); }

int launch_tc_search(const float* z_fused, int B, int64_t W, const __nv_bfloat16* xb, const __nv_bfloat16* eb, const __nv_bfloat16* eh,
                     const float* band, int64_t N, int64_t N_pad, int K, int K_pad, int D, uint8_t* cand_cnt, uint16_t* cand_idx,
                     int* fallback_rows, WsMeta* meta, unsigned long long* best64, float* scores_dbg, void* ev_scratch,
                     const TailArgs* tail_args, const float* codebook_f32, cudaStream_t s) {
    using namespace tc;
    const bool fuse = z_fused != nullptr;          // the caller decided with tc_can_fuse(): A operand built in-kernel from fp32 BCW
    const bool with_tail = tail_args != nullptr;   // the kernel also finishes the frames (needs the fused operand preparation)
    const bool tf32 = codebook_f32 != nullptr;     // kind::tf32 straight from the fp32 latents and the fp32 codebook
    if (tf32 && (!fuse || with_tail)) { set_error("tc_search: tf32 needs the in-place fp32 operand path"); return VQB_E_FLAGS; }
    if (with_tail && (!fuse || scores_dbg)) { set_error("tc_search: the fused tail needs the fused operand preparation"); return VQB_E_FLAGS; }
    if (with_tail && ((reinterpret_cast<uintptr_t>(tail_args->codebook) & 31) || (D % 8))) {
        set_error("tc_search: the fused tail gathers codebook rows with 32-byte loads; the codebook must be 32-byte aligned");
        return VQB_E_ALIGN;
    }
    CUtensorMap mx;
    int rc;
    if (tf32) rc = make_latent_map(&mx, z_fused, (uint64_t)B, (uint64_t)D, (uint64_t)W, 32, (uint32_t)D, true, true);
    else if (fuse) rc = make_map_z(&mx, z_fused, (uint64_t)B, (uint64_t)D, (uint64_t)W);
    else rc = make_map(&mx, xb, (uint64_t)N_pad, (uint64_t)D, BM);
    if (rc != 0) return rc;
    const int tiles_per_item = (int)((W + BM - 1) / BM);
    const int num_kb = tf32 ? (D + 31) / 32 : (D + BK - 1) / BK;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms > kTcMaxCtas) sms = kTcMaxCtas;
    const int num_m_tiles = fuse ? B * tiles_per_item : (int)(N_pad / BM);
    const int num_n_tiles = K_pad / BN;
    const TcPlan plan = tc_plan(fuse, with_tail, scores_dbg != nullptr, num_m_tiles, D, tf32, K_pad);
    const bool grouped = plan.grouped;
    if (!plan.ok) { set_error("tc_search: D=%d does not fit the shared-memory pipeline", D); return VQB_E_SHAPE; }
    const bool two = plan.two;
    const int cs = plan.cs, a_slots = plan.a_slots, b_stages = plan.b_stages, ev_sm = plan.ev_sm;
    const size_t smem = plan.smem;
    static bool attr_done_dev[64] = {};             // function attributes are per device (one process may drive several GPUs)
    bool& attr_done = attr_done_dev[dev & 63];
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(tc_search_kernel<false, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(tc_search_kernel<true, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(tc_search_kernel<false, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(tc_search_kernel<true, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(tc_search_kernel<false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(tc_search_kernel<true, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(tc_search_kernel<false, true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(tc_search_kernel<true, true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(tc_search_kernel<false, true, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(tc_search_kernel<true, true, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(tc_search_kernel<false, true, false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(tc_search_kernel<true, true, false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(tc_search_kernel)");
        attr_done = true;
    }
    CUtensorMap me_c, meh, mxt;
    if (with_tail) { if ((rc = make_map_z(&mxt, z_fused, (uint64_t)B, (uint64_t)D, (uint64_t)W, TAIL_CHUNK)) != 0) return rc; }
    else mxt = mx;
    if (tf32) { if ((rc = make_map_f32(&me_c, codebook_f32, (uint64_t)K, (uint64_t)D, two ? BN / 2 : BN / cs)) != 0) return rc; }
    else if ((rc = make_map(&me_c, eb, (uint64_t)K_pad, (uint64_t)D, two ? BN / 2 : BN / cs)) != 0) return rc;
    if ((rc = make_map_eh(&meh, eh, (uint64_t)K_pad)) != 0) return rc;
    int grid = num_m_tiles < sms ? num_m_tiles : sms;
    grid = grid / cs * cs;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(with_tail ? NUM_THREADS_TAIL : NUM_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)cs;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    TimingSlot* slot = timing_begin(s);
    cudaError_t le;
    const TailArgs targs = with_tail ? *tail_args : TailArgs{};
    const int l2_once = (fuse && latents_read_once((size_t)N * D * 4)) ? 1 : 0;   // stream the latents past the L2-resident working set
    // nap of the helper warps between two looks at a barrier they wait on (mbar_wait_sleep); VQB_TC_SLEEP (experiments): ns, 0 = spin
    const uint32_t spin_ns = (uint32_t)env_get(ENV_TC_SLEEP, kHelperNapNs);
#define VQB_TC_LAUNCH(TWO, FUSE, TAIL, ...)                                                                                                  \
    le = cudaLaunchKernelEx(&cfg, tc_search_kernel<TWO, FUSE, TAIL, ##__VA_ARGS__>, mx, me_c, meh, mxt, eh, W, tiles_per_item, D, (const WsMeta*)meta, band, N,   \
                            num_m_tiles, num_n_tiles, num_kb, a_slots, b_stages, cs, K, cand_cnt, cand_idx, fallback_rows, meta, best64,        \
                            scores_dbg, reinterpret_cast<uint32_t*>(ev_scratch), targs, ev_sm, l2_once, plan.eh_slots, plan.wq_cap, spin_ns)
    if (grouped && tf32 && two) VQB_TC_LAUNCH(true, true, false, true, true);
    else if (grouped && tf32) VQB_TC_LAUNCH(false, true, false, true, true);
    else if (grouped && two) VQB_TC_LAUNCH(true, true, false, false, true);
    else if (grouped) VQB_TC_LAUNCH(false, true, false, false, true);
    else if (tf32 && two) VQB_TC_LAUNCH(true, true, false, true);
    else if (tf32) VQB_TC_LAUNCH(false, true, false, true);
    else if (two && with_tail) VQB_TC_LAUNCH(true, true, true);
    else if (with_tail) VQB_TC_LAUNCH(false, true, true);
    else if (two && fuse) VQB_TC_LAUNCH(true, true, false);
    else if (two) VQB_TC_LAUNCH(true, false, false);
    else if (fuse) VQB_TC_LAUNCH(false, true, false);
    else VQB_TC_LAUNCH(false, false, false);
#undef VQB_TC_LAUNCH
    if (le != cudaSuccess) return cuda_fail(le, "tc_search_kernel launch");
    cudaError_t e = cudaGetLastError();
    timing_end(slot, s);
    note_launch();
    if (e != cudaSuccess) return cuda_fail(e, "tc_search_kernel launch");
    return 0;
}

}  // namespace vqb

extern "C" {
#ifdef VQB_TC_TRACE
// debug builds only: device buffer of 2 x cap records (CTA 0 and CTA 1) for the hand-off timeline; cap = 0 switches tracing off
__attribute__((visibility("default"))) int vqb_debug_set_trace(void* buf, unsigned int cap) {
    unsigned long long* b = static_cast<unsigned long long*>(buf);
    if (cudaMemcpyToSymbol(vqb::tc::g_trace_buf, &b, sizeof(b)) != cudaSuccess) return 1;
    if (cudaMemcpyToSymbol(vqb::tc::g_trace_cap, &cap, sizeof(cap)) != cudaSuccess) return 1;
    return 0;
}
#endif
// enable != 0: record a CUDA-event pair around every stage of vqb_forward from now on (and forget earlier ones)
int vqb_debug_kernel_timing(int enable) {
    vqb::g_timing = enable != 0;
    vqb::g_slots_used = 0;
    return 0;
}
// total milliseconds and number of timed launches of one stage (VQB_STAGE_*) since timing was enabled (synchronises the events)
int vqb_debug_stage_time_ms(int stage, double* total_ms, int* launches) {
    if (!total_ms || !launches) { vqb::set_error("vqb_debug_stage_time_ms: NULL output"); return VQB_E_NULL; }
    double t = 0.0;
    int n = 0;
    for (int i = 0; i < vqb::g_slots_used; ++i) {
        if (vqb::g_slots[i].stage != stage) continue;
        cudaError_t e = cudaEventSynchronize(vqb::g_slots[i].stop);
        if (e != cudaSuccess) return vqb::cuda_fail(e, "cudaEventSynchronize");
        float ms = 0.f;
        e = cudaEventElapsedTime(&ms, vqb::g_slots[i].start, vqb::g_slots[i].stop);
        if (e != cudaSuccess) return vqb::cuda_fail(e, "cudaEventElapsedTime");
        t += ms;
        ++n;
    }
    *total_ms = t;
    *launches = n;
    return 0;
}
// the dominant kernel (tc_search_kernel): VQB_STAGE_SEARCH
int vqb_debug_kernel_time_ms(double* total_ms, int* launches) { return vqb_debug_stage_time_ms(VQB_STAGE_SEARCH, total_ms, launches); }
}
