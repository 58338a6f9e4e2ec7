// tcgen05 / TMEM / mbarrier helpers shared by the tensor-core kernels (gemm_tc.cu, cin_tc.cu).
// PTX forms follow cute/arch/{mma_sm100_umma,copy_sm100,tmem_allocator_sm100}.hpp and
// cutlass/arch/barrier.h of the CUTLASS headers vendored in this image (read for reference only).
#pragma once
#include <stdint.h>
#include <stdio.h>

namespace {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    // bounded spin: a protocol bug becomes a trap (CUDA error) instead of a hung GPU
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 24)) {
            if ((threadIdx.x & 31) == 0)
                printf("ctr_b200: mbarrier wait timed out (block %d,%d,%d thread %d parity %u)\n", blockIdx.x,
                       blockIdx.y, blockIdx.z, threadIdx.x, parity);
            __trap();
        }
    }
}

// K-major, SWIZZLE_NONE shared-memory matrix descriptor (cute::UMMA::SmemDescriptor):
//   start address >> 4 [0,14) | LBO >> 4 [16,30) | SBO >> 4 [32,46) | version = 1 [46,48) | layout 0
// canonical layout in 16-byte units: ((8,n),2):((1,SBO),LBO)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc,
                                          uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
}


// Split an fp32 value into two TF32-exact parts with ROUND-TO-NEAREST at both levels:
//   hi = RN_tf32(x)            (11 significant bits)
//   lo = RN_tf32(x - hi)       (x - hi is exact in fp32; after rounding |x - hi - lo| <= 2^-24 |x|)
// Truncation (just clearing the low bits) leaves a residual of up to 2^-21 |x| that always points
// toward zero; in sums with heavy cancellation (weight gradients over a batch of near-identical rows)
// that bias does not average out and showed up as 5e-4 relative errors.  With RN the residual is
// sign-symmetric and 8x smaller, and both parts are exactly representable, so the result does not
// depend on how the tensor core narrows its fp32 inputs.
__device__ __forceinline__ float rn_tf32(float x) {
    return __uint_as_float((__float_as_uint(x) + 0x00001000u) & 0xFFFFE000u);
}
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
    hi = rn_tf32(x);
    lo = rn_tf32(x - hi);
}

__device__ __forceinline__ void tmem_alloc_warp(uint32_t* slot, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_warp(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// Shared-memory operand tile (R rows x 32 k, fp32): K-major core matrices, 16-byte chunk c of row r
// at byte offset c*(R+1)*16 + r*16.  The +1 row of padding per chunk staggers the chunks over the
// banks so that transposing loaders store conflict-free; UMMA descriptor: LBO = (R+1)*16 (between
// the two chunks of a K atom), SBO = 128 (between 8-row core matrices).
__device__ __forceinline__ int tile_off(int R, int r, int c) { return (c * (R + 1) + r) * 4; }   // in floats
__device__ __forceinline__ uint32_t tile_bytes(int R) { return (uint32_t)(R + 1) * 128u; }

__device__ __forceinline__ void split_store(float* hi_ptr, float* lo_ptr, float4 v) {
    float4 hi, lo;
    split_tf32(v.x, hi.x, lo.x);
    split_tf32(v.y, hi.y, lo.y);
    split_tf32(v.z, hi.z, lo.z);
    split_tf32(v.w, hi.w, lo.w);
    *reinterpret_cast<float4*>(hi_ptr) = hi;
    *reinterpret_cast<float4*>(lo_ptr) = lo;
}

// issue the 3xTF32 MMAs of one 32-wide K stage (4 atoms x {lo*hi, hi*lo, hi*hi}); tiles as above
__device__ __forceinline__ void issue_stage_mmas(uint32_t tmem_d, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi,
                                                 uint32_t b_lo, int RA, int RB, uint32_t idesc, bool first) {
    const uint32_t a_lbo = (uint32_t)(RA + 1) * 16u, b_lbo = (uint32_t)(RB + 1) * 16u;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const uint32_t ao = (uint32_t)j * 2u * a_lbo, bo = (uint32_t)j * 2u * b_lbo;
        const uint64_t dah = make_smem_desc(a_hi + ao, a_lbo, 128);
        const uint64_t dal = make_smem_desc(a_lo + ao, a_lbo, 128);
        const uint64_t dbh = make_smem_desc(b_hi + bo, b_lbo, 128);
        const uint64_t dbl = make_smem_desc(b_lo + bo, b_lbo, 128);
        umma_tf32(tmem_d, dal, dbh, idesc, (first && j == 0) ? 0u : 1u);   // small terms first
        umma_tf32(tmem_d, dah, dbl, idesc, 1u);
        umma_tf32(tmem_d, dah, dbh, idesc, 1u);
    }
}

// tf32 x tf32 -> f32 instruction descriptor, K-major A and B, M = 128
__device__ __forceinline__ uint32_t tf32_idesc(int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

}  // namespace
