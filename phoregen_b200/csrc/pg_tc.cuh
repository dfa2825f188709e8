// sm_100a tensor-core plumbing: tcgen05 / TMEM / mbarrier PTX wrappers and UMMA descriptors.
// (Blackwell guide: blackwell_cuda_programming.md §1-2; descriptor bit layouts as in CUTLASS cute/arch/mma_sm100_desc.hpp.)
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// Plain parity wait (row / producer / epilogue warps: the register-tight hot loops).
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}
// Parity wait for LONG waits (a producer that is a whole tile ahead): between polls the warp sleeps, so that it neither
// takes issue slots nor shared-memory cycles (every try_wait is a shared-memory access) from the warps on the critical path.
__device__ __forceinline__ void mbar_wait_sleep(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "WAIT_LOOP:\n\t"
        "nanosleep.u32 256;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}
// Parity wait with a watchdog, used by the auxiliary warps (MMA issuer, loaders, feature warps).  In every hand-off cycle
// of these kernels one of them is among the waiters, so a protocol error ("mbarrier parity waits")
// surfaces as a launch failure of this kernel instead of a GPU that never comes back.  The wall clock is read every 4096
// failed polls; 10 s are orders of magnitude above any legitimate wait here (tens of microseconds).
__device__ __forceinline__ void mbar_wait_wd(uint64_t* bar, uint32_t parity) {
#ifdef PG_NO_WATCHDOG
    mbar_wait(bar, parity);
#else
    asm volatile(
        "{\n\t"
        ".reg .pred p, q;\n\t"
        ".reg .u32 polls, m;\n\t"
        ".reg .u64 t0, t1;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "mov.u32 polls, 0;\n\t"
        "mov.u64 t0, 0;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "add.u32 polls, polls, 1;\n\t"
        "and.b32 m, polls, 0xfff;\n\t"
        "setp.ne.u32 q, m, 0;\n\t"
        "@q bra WAIT_LOOP;\n\t"
        "mov.u64 t1, %%globaltimer;\n\t"
        "setp.eq.u64 q, t0, 0;\n\t"
        "@q mov.u64 t0, t1;\n\t"
        "sub.u64 t1, t1, t0;\n\t"
        "setp.lt.u64 q, t1, 10000000000;\n\t"
        "@q bra WAIT_LOOP;\n\t"
        "trap;\n\t"
        "DONE:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
#endif
}
// The same for the LONG waits of warps that run ahead of the critical path (loader, feature warps): the warp sleeps
// between polls, so it takes neither issue slots nor shared-memory cycles (every try_wait is a shared-memory access)
// from the row warps that share its scheduler.
__device__ __forceinline__ void mbar_wait_wd_sleep(uint64_t* bar, uint32_t parity) {
#ifdef PG_NO_WATCHDOG
    mbar_wait_sleep(bar, parity);
#else
    asm volatile(
        "{\n\t"
        ".reg .pred p, q;\n\t"
        ".reg .u32 polls, m;\n\t"
        ".reg .u64 t0, t1;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "mov.u32 polls, 0;\n\t"
        "mov.u64 t0, 0;\n\t"
        "WAIT_LOOP:\n\t"
        "nanosleep.u32 64;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "add.u32 polls, polls, 1;\n\t"
        "and.b32 m, polls, 0xfff;\n\t"
        "setp.ne.u32 q, m, 0;\n\t"
        "@q bra WAIT_LOOP;\n\t"
        "mov.u64 t1, %%globaltimer;\n\t"
        "setp.eq.u64 q, t0, 0;\n\t"
        "@q mov.u64 t0, t1;\n\t"
        "sub.u64 t1, t1, t0;\n\t"
        "setp.lt.u64 q, t1, 10000000000;\n\t"
        "@q bra WAIT_LOOP;\n\t"
        "trap;\n\t"
        "DONE:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
#endif
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// generic-proxy smem writes -> visible to the async proxy (tensor core / TMA reads)
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---------------------------------------------------------------- TMEM
template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem) {   // one full warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "n"(COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {     // the same warp that allocated
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// 32 lanes x 32 consecutive fp32 columns: thread t of the warp receives row (lane base + t), columns [col, col+32)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; i++) v[i] = __uint_as_float(r[i]);
}

// same without the trailing wait: issue several loads, then tmem_ld_wait() once
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t* r) {   // r: 32 registers (compile-time indexed slice)
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
// 4 consecutive columns of this thread's lane (waits for completion)
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, float (&v)[4]) {
    uint32_t r0, r1, r2, r3;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];\ntcgen05.wait::ld.sync.aligned;"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(taddr) : "memory");
    v[0] = __uint_as_float(r0); v[1] = __uint_as_float(r1); v[2] = __uint_as_float(r2); v[3] = __uint_as_float(r3);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// 16 TMEM lanes x 32 columns in the MMA-fragment shape: register j of thread t holds
//   lane (taddr.lane + 8 ((j >> 1) & 1) + t / 4), column (taddr.col + 8 (j >> 2) + 2 (t % 4) + (j & 1))
// (checked on the device by tools/micro/tmem_transpose_probe.cu).  No trailing wait.
__device__ __forceinline__ void tmem_ld_16x256b_x4_nowait(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}

// thread t of the warp writes 32 consecutive 32-bit columns [col, col+32) of TMEM lane (lane base + t)
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
        "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]),
        "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]),
        "r"(r[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
        "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// Column sums of a 32 x 32 block held one row per lane (p[c] = this lane's row, column c): the block goes through 32 TMEM
// columns of the warp's own lanes (taddr: lane base + first column; the columns are scratch) and comes back in the
// 16x256b fragment shape, where a thread holds 4 rows x 8 columns: 24 additions in registers, then 7 shuffles instead of
// the 31 of transpose_reduce32.  Returns the sum of column tmem_reduce_col(lane) -- note the permutation.
__device__ __forceinline__ int tmem_reduce_col(int lane) { return (lane & 24) | ((lane & 3) << 1) | ((lane >> 2) & 1); }
__device__ __forceinline__ float tmem_transpose_reduce32(uint32_t taddr, const float (&p)[32], int lane) {
    uint32_t u[32];
#pragma unroll
    for (int i = 0; i < 32; i++) u[i] = __float_as_uint(p[i]);
    tmem_st32(taddr, u);
    tmem_st_wait();
    uint32_t v0[16], v1[16];
    tmem_ld_16x256b_x4_nowait(taddr, v0);
    tmem_ld_16x256b_x4_nowait(taddr + (16u << 16), v1);
    tmem_ld_wait();
    float s[8];                                    // s[2 b + e]: rows = lane / 4 (mod 8), column 8 b + 2 (lane % 4) + e
#pragma unroll
    for (int b = 0; b < 4; b++)
#pragma unroll
        for (int e = 0; e < 2; e++)
            s[2 * b + e] = (__uint_as_float(v0[4 * b + e]) + __uint_as_float(v0[4 * b + 2 + e])) +
                           (__uint_as_float(v1[4 * b + e]) + __uint_as_float(v1[4 * b + 2 + e]));
    {
        const bool up = lane & 16;                 // keeps blocks b = 2, 3
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const float send = up ? s[i] : s[i + 4], keep = up ? s[i + 4] : s[i];
            s[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
        }
    }
    {
        const bool up = lane & 8;                  // keeps the odd block of its pair
#pragma unroll
        for (int i = 0; i < 2; i++) {
            const float send = up ? s[i] : s[i + 2], keep = up ? s[i + 2] : s[i];
            s[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
        }
    }
    const bool up = lane & 4;                      // keeps e = 1
    const float send = up ? s[0] : s[1], keep = up ? s[1] : s[0];
    return keep + __shfl_xor_sync(0xffffffffu, send, 4);
}

// ---------------------------------------------------------------- UMMA descriptors
// K-major operand tile in shared memory, 128-byte swizzle: rows of 64 bf16 (128 B), 8-row groups 1024 B apart.
// start address (>>4) | LBO=1 (ignored for swizzled K-major) | SBO = 1024 B | version 1 | SWIZZLE_128B (2).
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)2 << 61);
}
// MN-major operand tile (element (n, k): n contiguous), 128-byte swizzle: rows of 64 bf16 along N (128 B) per k, 8 k-rows per
// 1024-byte swizzle atom (16-byte chunk j of row k stored at chunk j ^ (k % 8)); LBO = byte stride between 64-element
// blocks along N, SBO = byte stride between 8-row groups along K (canonical layout ((8,n),(8,k)):((1,LBO),(8,SBO)) in
// 16-byte units, CUTLASS cute/atom/mma_traits_sm100.hpp).  Needs the instruction descriptor's major bit of that operand.
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes = 1024) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// kind::f16 instruction descriptor: D = F32, A = B = BF16, both K-major, shape M x N (x 16).
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// D[tmem] (+)= A[smem] * B[smem]^T ; issued by ONE thread
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// K-major operand tile WITHOUT swizzle ("interleaved" core matrices of 8 rows x 16 bytes):
// element (r, k) of a [rows x 16] bf16 tile lives at (r/8)*256 + (k/8)*128 + (r%8)*16 + (k%8)*2 bytes
// -> LBO (next core matrix along K) = 128 B, SBO (next 8-row group) = 256 B.
__device__ __forceinline__ uint64_t umma_desc_k16_noswizzle(uint32_t smem_addr) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFF) | ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(256 >> 4) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ uint32_t k16_offset(int r, int k) { return (uint32_t)((r >> 3) * 256 + (k >> 3) * 128 + (r & 7) * 16 + (k & 7) * 2); }

// D[tmem] (+)= A[tmem] * B[smem]^T : A is M x K bf16 in TMEM (lane = row, two K elements per 32-bit column)
__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "}" ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Warp-collective forms of the three calls above: EVERY lane of the converged issuer warp calls them with identical
// operands and one elected lane issues the instruction.  Branching on `lane == 0` around the issue code instead puts the
// descriptor arithmetic into a divergent region, where ptxas cannot use the uniform datapath and feeds every UTCHMMA
// through an ELECT / R2UR.BROADCAST / BRA.U.ANY waterfall loop (~100 cycles per MMA: the tensor pipe then idles behind
// the issuer; measured on the triplet kernel).  elect.sync always picks the same lane while the mask is unchanged, so the
// MMAs and their commit are issued by one thread, as tcgen05.commit requires.
__device__ __forceinline__ void umma_bf16_w(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p, q;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_bf16_ts_w(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p, q;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "}" ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit_w(uint64_t* bar) {
    asm volatile(
        "{\n\t"
        ".reg .pred q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t"
        "}" ::"r"(smem_u32(bar))
        : "memory");
}
// all previously issued MMAs of this thread arrive on the mbarrier when complete (implies fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// byte offset of the 16-byte chunk (row r, chunk j in 0..7) inside one 128B-swizzled K-block tile
__device__ __forceinline__ uint32_t sw128_chunk(int r, int j) { return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((j ^ (r & 7)) << 4)); }

// fp32 -> (bf16 hi, bf16 lo) with hi + lo ~ x to 2^-17 relative (bf16x3 product scheme)
__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
    hi = __float2bfloat16_rn(x);
    lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}
// Cheap pairwise split for the hot loops: hi = fp32 truncated to its upper 16 bits (an exact bf16 value, no
// conversion instruction), lo = bf16_rn(x - hi) with one packed cvt for two values.  |x - hi - lo| <= 2^-17 |x|.
__device__ __forceinline__ void split_pair_trunc(float x0, float x1, uint32_t& hi, uint32_t& lo) {
    const uint32_t b0 = __float_as_uint(x0), b1 = __float_as_uint(x1);
    hi = __byte_perm(b0, b1, 0x7632);                               // {x0.hi16, x1.hi16}
    const float r0 = x0 - __uint_as_float(b0 & 0xFFFF0000u), r1 = x1 - __uint_as_float(b1 & 0xFFFF0000u);
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(r1), "f"(r0));   // upper half <- r1, lower half <- r0
}
// ReLU fused into the split: hi = bf16_rz(max(y, 0)) and lo = bf16_rn(max(y - hi, 0)), one packed cvt each.
// For y < 0 both are 0; for y >= 0 truncation gives hi <= y, so the second ReLU is a no-op.  Same error bound as above.
__device__ __forceinline__ void split_pair_relu(float2 y, uint32_t& hi, uint32_t& lo) {
    asm("cvt.rz.relu.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(y.y), "f"(y.x));
    float2 hf = make_float2(__uint_as_float(hi << 16), __uint_as_float(hi & 0xFFFF0000u));
    unsigned long long r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(*reinterpret_cast<unsigned long long*>(&y)), "l"(*reinterpret_cast<unsigned long long*>(&hf)));
    const float2 d = *reinterpret_cast<float2*>(&r);
    asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(d.y), "f"(d.x));
}
// The same with an fp16 pair: hi = f16_rz(max(y, 0)), lo = f16_rn(max(y - hi, 0)); |y - hi - lo| <= 2^-22 |y| in the normal range.
__device__ __forceinline__ void split_pair_relu_f16(float2 y, uint32_t& hi, uint32_t& lo) {
    asm("cvt.rz.relu.f16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(y.y), "f"(y.x));
    float2 hf;
    asm("{ .reg .b16 l, h; mov.b32 {l, h}, %2; cvt.f32.f16 %0, l; cvt.f32.f16 %1, h; }" : "=f"(hf.x), "=f"(hf.y) : "r"(hi));
    unsigned long long r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(*reinterpret_cast<unsigned long long*>(&y)), "l"(*reinterpret_cast<unsigned long long*>(&hf)));
    const float2 d = *reinterpret_cast<float2*>(&r);
    asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(d.y), "f"(d.x));
}
// Warp-wide reductions on the REDUX unit (one instruction instead of a 5-round shuffle butterfly).
__device__ __forceinline__ float warp_max_redux(float x) {        // sm_100a: redux.sync on f32
    float r;
    asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(x));
    return r;
}
// Sum of 32 values in [0, 1]: fixed point with 23 fractional bits (sum <= 2^28), every term rounded to nearest, so the
// result is within 32 * 2^-24 of the exact sum and independent of the lane order.
__device__ __forceinline__ float warp_sum01_redux(float x) {
    const uint32_t q = __float2uint_rn(x * 8388608.0f);
    uint32_t r;
    asm volatile("redux.sync.add.u32 %0, %1, 0xffffffff;" : "=r"(r) : "r"(q));
    return (float)r * (1.0f / 8388608.0f);
}
// 1 / x for the softmax denominators (x >= 1: the row with the maximum contributes 2^0): one MUFU instead of the
// ~8 instructions of the correctly rounded __frcp_rn; 1 ulp.
__device__ __forceinline__ float rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float ex2_approx(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ uint32_t pack_bf16(__nv_bfloat16 a, __nv_bfloat16 b) {
    return (uint32_t)__bfloat16_as_ushort(a) | ((uint32_t)__bfloat16_as_ushort(b) << 16);
}

// ---------------------------------------------------------------- packed fp32x2 arithmetic (sm_100: FADD2 / FMUL2 / FFMA2)
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
    unsigned long long r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)));
    return *reinterpret_cast<float2*>(&r);
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
    unsigned long long r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)));
    return *reinterpret_cast<float2*>(&r);
}
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
    unsigned long long r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)),
        "l"(*reinterpret_cast<unsigned long long*>(&c)));
    return *reinterpret_cast<float2*>(&r);
}

// ---------------------------------------------------------------- 1-D bulk copy global -> shared with mbarrier completion (TMA unit)
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void* dst_smem, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

}  // namespace tc
