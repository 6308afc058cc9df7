// tcgen05 (5th-generation tensor core) building blocks for sm_100a: TMEM allocation, shared-memory matrix
// descriptors, instruction descriptors, single-thread MMA issue, commit -> mbarrier, TMEM -> register loads.
// Bit layouts follow the PTX ISA "tcgen05 matrix descriptor" / "instruction descriptor" tables (the same
// fields CUTLASS names in cute/arch/mma_sm100_desc.hpp).
#pragma once
#include <stdint.h>

#include "ptx.cuh"

namespace pyvb {
namespace umma {

// ---- shared-memory matrix descriptor (64 bit) ---------------------------------------------------------------
//   [ 0,14) start address >> 4      [16,30) leading-dimension byte offset >> 4   [32,46) stride byte offset >> 4
//   [46,48) version = 1 (Blackwell) [49,52) base offset (0: tile aligned to the swizzle repeat)
//   [61,64) layout: 0 none, 1 128B(base 32B), 2 SWIZZLE_128B, 4 SWIZZLE_64B, 6 SWIZZLE_32B
enum Layout : uint64_t { SW_NONE = 0, SW_128B = 2, SW_64B = 4, SW_32B = 6 };

__device__ __forceinline__ uint64_t smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint64_t layout) {
    return (uint64_t)((smem_addr & 0x3ffff) >> 4) | ((uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32) | (1ULL << 46) | (layout << 61);
}

// K-major operand tile [rows][BK elements], BK * elemsize = 64 bytes per row (TMA SWIZZLE_64B box): rows are
// 64 bytes apart, groups of 8 rows 512 bytes apart.  One MMA consumes 32 bytes of K: advance the start by 32.
__device__ __forceinline__ uint64_t desc_kmajor_sw64(uint32_t tile_addr, int kstep) {
    return smem_desc(tile_addr + kstep * 32, 16, 512, SW_64B);
}
// K-major, 128 bytes per row (SWIZZLE_128B): groups of 8 rows 1024 bytes apart.
__device__ __forceinline__ uint64_t desc_kmajor_sw128(uint32_t tile_addr, int kstep) {
    return smem_desc(tile_addr + kstep * 32, 16, 1024, SW_128B);
}
// MN-major operand: memory is [K rows][64 MN elements (128 bytes)] per atom column (a TMA SWIZZLE_128B box with the
// MN dimension innermost); 64-element MN blocks are `mn_block_bytes` apart (LBO), groups of 8 K rows 1024 bytes
// apart (SBO).  One bf16 MMA consumes 16 K rows = 2048 bytes.
__device__ __forceinline__ uint64_t desc_mnmajor_sw128(uint32_t tile_addr, int kstep, uint32_t mn_block_bytes) {
    return smem_desc(tile_addr + kstep * 2048, mn_block_bytes, 1024, SW_128B);
}

// ---- instruction descriptor (32 bit), kind::f16 / kind::tf32 -------------------------------------------------
//   [4,6) D format: 1 = f32   [7,10) A format, [10,13) B format: 0 f16, 1 bf16, 2 tf32
//   [15] A major, [16] B major: 0 = K, 1 = MN   [17,23) N >> 3   [24,29) M >> 4
__host__ __device__ constexpr uint32_t idesc_bf16_f32(int M, int N, int a_mn_major, int b_mn_major) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// kind::i8: D format 2 = s32, A / B format 1 = signed 8 bit; K = 32 per instruction
__host__ __device__ constexpr uint32_t idesc_s8_s32(int M, int N) {
    return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---- TMEM ------------------------------------------------------------------------------------------------------
// one warp, all 32 lanes; ncols a power of two in [32, 512]; the base address lands in *smem_dst
__device__ __forceinline__ void tmem_alloc(uint32_t *smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---- CTA pair (cta_group::2): the two CTAs of a cluster act as ONE 256-row MMA.  Each CTA holds its own 128 rows of A and
// HALF of the N rows of B in its shared memory (same offsets in both CTAs); the leader CTA (cluster rank 0) issues the MMAs
// for both, each CTA gets its 128 accumulator lanes in its own TMEM.  Allocation, MMA, commit are the cta_group::2 forms.
__device__ __forceinline__ void tmem_alloc2(uint32_t *smem_dst, uint32_t ncols) {        // one warp in EACH CTA of the pair
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void mma_i8_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// all MMAs issued so far arrive on the barrier at this offset in BOTH CTAs of the pair when they have completed
__device__ __forceinline__ void mma_commit_pair(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     smem_u32(bar)),
                 "h"((uint16_t)3)
                 : "memory");
}
// shared::cluster address of `p` (a shared-memory address of this CTA) in the CTA with cluster rank `rank`
__device__ __forceinline__ uint32_t map_to_cta(const void *p, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(p)), "r"(rank));
    return r;
}
// Arrive on a barrier of another CTA of the cluster.  Default semantics (.release.cta), as CUTLASS's ClusterBarrier::arrive: the
// barriers signalled this way guard TMEM accumulators, whose accesses are ordered by the tcgen05 fences around the barrier, not
// by a memory fence.  With .release.cluster every arrive drained the thread's memory operations at cluster scope: ncu showed 18 %
// of all stall samples of K1-i8 (CTA-pair mode) on this one instruction (`membar`), once per tile and epilogue warp.
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}

// ---- MMA issue (one thread) ------------------------------------------------------------------------------------
// D[tmem] (+)= A[smem] * B[smem];  accumulate = 0 overwrites D
__device__ __forceinline__ void mma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void mma_i8(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// all MMAs issued so far by this thread arrive on `bar` when they have completed (implies fence::before_thread_sync)
__device__ __forceinline__ void mma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- TMEM -> registers: lane l of the warp reads TMEM lane (warp % 4) * 32 + l, 16 consecutive 32-bit columns ----
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- bounded mbarrier wait: a protocol bug traps instead of hanging the GPU -------------------------------------
__device__ __forceinline__ void mbar_wait_bounded(uint64_t *bar, uint32_t parity) {
    for (uint32_t it = 0; it < (1u << 26); ++it)
        if (mbar_try_wait(bar, parity)) return;
    __trap();
}

// 2-D TMA load with an L2 cache hint left at default; coordinates (c0 = innermost element index, c1 = row)
__device__ __forceinline__ void tma_2d(void *dst_smem, const void *tmap, int c0, int c1, uint64_t *bar) {
    tma_load_2d(dst_smem, tmap, c0, c1, bar);
}

}  // namespace umma
}  // namespace pyvb
