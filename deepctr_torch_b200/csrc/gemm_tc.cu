// tcgen05 GEMM with fp32-grade accuracy (3xTF32 split) — the tensor-core path of the dense tower.
//
//   C[m,n] (+)= sum_k A(m,k) * B(n,k)            A, B: fp32 with arbitrary element strides
//
// Why 3xTF32: the parity gate is 1e-5 relative on the logits; single-pass TF32/BF16 operands miss
// it by 10-100x (SURVEY.md §7 hard part 1).  Every fp32 operand x is split on the fly into
//   hi = x with the 13 low mantissa bits cleared (exactly representable in TF32)
//   lo = x - hi (exact in fp32)
// and the tile product is accumulated as  lo*hi + hi*lo + hi*hi  by three tcgen05.mma.kind::tf32
// instructions per K atom into ONE fp32 accumulator in tensor memory (error ~2^-21 per product).
//
// CTA = 128 x BN output tile (BN multiple of 32, <= 256), 9 warps:
//   warps 0-7  producers: LDG (any strides, bounds, optional act'(mask) prologue) -> hi/lo split
//              -> STS.128 into the canonical K-major no-swizzle UMMA layout (8x16B core matrices),
//              fence.proxy.async + mbarrier arrive;  after the main loop the same warps run the
//              epilogue: tcgen05.ld (32 lanes x 16 columns) -> bias/activation/mask -> global.
//   warp 8     TMEM alloc/dealloc; one elected lane issues the MMAs and tcgen05.commit's.
// Pipeline: `stages` shared-memory stages of 32 fp32 of K (full/empty mbarriers), split-K across
// gridDim.z with fp32 atomics for the weight-gradient shapes.
#include "gemm.cuh"
#include "tc_common.cuh"

namespace {

constexpr int TC_BM = 128;
constexpr int TC_BK = 32;                 // fp32 elements of K per stage = 8 chunks of 16 B
constexpr int TC_PROD_WARPS = 8;
constexpr int TC_THREADS = (TC_PROD_WARPS + 1) * 32;

struct TcParams {
    GemmArgs g;
    int64_t k_chunk;      // K range per split (multiple of TC_BK)
    int BN;               // N tile (multiple of 32)
    int stages;
    int tmem_cols;        // power of two >= max(32, BN)
};

__global__ void __launch_bounds__(TC_THREADS, 1) gemm_tc_kernel(TcParams p) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const GemmArgs& g = p.g;
    const int BN = p.BN, S = p.stages;
    const uint32_t a_tile = TC_BM * TC_BK * 4;          // bytes of one A tile (hi or lo)
    const uint32_t b_tile = (uint32_t)BN * TC_BK * 4;
    const uint32_t stage_bytes = 2 * a_tile + 2 * b_tile;
    unsigned char* tiles = smem_raw;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_raw + (size_t)S * stage_bytes);
    uint64_t* empty_bar = full_bar + S;
    uint64_t* accum_bar = empty_bar + S;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t m0 = (int64_t)blockIdx.y * TC_BM, n0 = (int64_t)blockIdx.x * BN;
    const int64_t kbeg = (int64_t)blockIdx.z * p.k_chunk;
    const int64_t kend = (kbeg + p.k_chunk < g.K) ? kbeg + p.k_chunk : g.K;
    const int nkb = (int)((kend - kbeg + TC_BK - 1) / TC_BK);

    if (tid == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(&full_bar[s], TC_PROD_WARPS);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(accum_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == TC_PROD_WARPS) {   // TMEM allocation by the MMA warp
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"((uint32_t)p.tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp < TC_PROD_WARPS) {
        // ------------------------------ producers ------------------------------------------
        const int r_lo = lane & 7, c_lo = lane >> 3;
        for (int kb = 0; kb < nkb; ++kb) {
            const int s = kb % S;
            const uint32_t ph = (uint32_t)(kb / S) & 1u;
            mbar_wait(&empty_bar[s], ph ^ 1u);
            unsigned char* st = tiles + (size_t)s * stage_bytes;
            float* a_hi = reinterpret_cast<float*>(st);
            float* a_lo = reinterpret_cast<float*>(st + a_tile);
            float* b_hi = reinterpret_cast<float*>(st + 2 * a_tile);
            float* b_lo = reinterpret_cast<float*>(st + 2 * a_tile + b_tile);
            const int64_t k0 = kbeg + (int64_t)kb * TC_BK;
            // A: 16 row groups x 2 chunk halves = 32 units of (8 rows x 4 chunks)
            for (int U = warp; U < (TC_BM / 8) * 2; U += TC_PROD_WARPS) {
                const int r = (U >> 1) * 8 + r_lo;
                const int c = (U & 1) * 4 + c_lo;
                const int64_t m = m0 + r;
                float v[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int64_t k = k0 + c * 4 + e;
                    float x = 0.f;
                    if (m < g.M && k < kend) {
                        x = __ldg(g.A + m * g.sam + k * g.sak);
                        if (g.amask) x *= act_grad_from_y(g.amask_act, __ldg(g.amask + m * g.smm + k * g.smk));
                    }
                    v[e] = x;
                }
                float4 hi, lo;
                hi.x = __uint_as_float(__float_as_uint(v[0]) & 0xFFFFE000u); lo.x = v[0] - hi.x;
                hi.y = __uint_as_float(__float_as_uint(v[1]) & 0xFFFFE000u); lo.y = v[1] - hi.y;
                hi.z = __uint_as_float(__float_as_uint(v[2]) & 0xFFFFE000u); lo.z = v[2] - hi.z;
                hi.w = __uint_as_float(__float_as_uint(v[3]) & 0xFFFFE000u); lo.w = v[3] - hi.w;
                const int off = (c * TC_BM + r) * 4;     // floats: chunk-major, 16 B per row
                *reinterpret_cast<float4*>(a_hi + off) = hi;
                *reinterpret_cast<float4*>(a_lo + off) = lo;
            }
            for (int U = warp; U < (BN / 8) * 2; U += TC_PROD_WARPS) {
                const int r = (U >> 1) * 8 + r_lo;
                const int c = (U & 1) * 4 + c_lo;
                const int64_t n = n0 + r;
                float v[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int64_t k = k0 + c * 4 + e;
                    float x = 0.f;
                    if (n < g.N && k < kend) {
                        x = __ldg(g.B + n * g.sbn + k * g.sbk);
                        if (g.bmask) x *= act_grad_from_y(g.bmask_act, __ldg(g.bmask + n * g.sbmn + k * g.sbmk));
                    }
                    v[e] = x;
                }
                float4 hi, lo;
                hi.x = __uint_as_float(__float_as_uint(v[0]) & 0xFFFFE000u); lo.x = v[0] - hi.x;
                hi.y = __uint_as_float(__float_as_uint(v[1]) & 0xFFFFE000u); lo.y = v[1] - hi.y;
                hi.z = __uint_as_float(__float_as_uint(v[2]) & 0xFFFFE000u); lo.z = v[2] - hi.z;
                hi.w = __uint_as_float(__float_as_uint(v[3]) & 0xFFFFE000u); lo.w = v[3] - hi.w;
                const int off = (c * BN + r) * 4;
                *reinterpret_cast<float4*>(b_hi + off) = hi;
                *reinterpret_cast<float4*>(b_lo + off) = lo;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic -> async proxy
            __syncwarp();
            if (lane == 0) mbar_arrive(&full_bar[s]);
        }
        // ------------------------------ epilogue -------------------------------------------
        mbar_wait(accum_bar, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int quad = warp & 3;                       // TMEM lane quadrant of this warp
        const int half = warp >> 2;                      // column half
        const int64_t m = m0 + quad * 32 + lane;
        const bool split = gridDim.z > 1;
        const int col_beg = half * (BN / 2), col_end = col_beg + BN / 2;
        for (int c0 = col_beg; c0 < col_end; c0 += 16) {
            uint32_t raw[16];
            tmem_ld16(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)c0, raw);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (m < g.M) {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int64_t n = n0 + c0 + j;
                    if (n >= g.N) break;
                    float v = __uint_as_float(raw[j]);
                    float* cp = g.C + m * g.ldc + n;
                    switch (g.epilogue) {
                        case EPI_BIAS_ACT:
                            if (g.bias) v += __ldg(g.bias + n);
                            v = act_apply(g.act, v);
                            break;
                        case EPI_MUL_ACTGRAD:
                            v *= act_grad_from_y(g.act, __ldg(g.aux + m * g.ldaux + n));
                            break;
                        case EPI_CROSS: {
                            const float u = v + __ldg(g.bias + n);
                            if (g.out2) g.out2[m * g.ldout2 + n] = u;
                            v = __ldg(g.aux + m * g.ldaux + n) * u + __ldg(g.aux2 + m * g.ldaux2 + n);
                            break;
                        }
                        default: break;
                    }
                    if (split) atomicAdd(cp, v);
                    else if (g.accumulate) *cp += v;
                    else *cp = v;
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    } else {
        // ------------------------------ MMA issuer -----------------------------------------
        // instruction descriptor (cute::UMMA::InstrDescriptor): c=F32 [4,6), a=b=TF32 [7,10)/[10,13),
        // K-major A and B, N>>3 at [17,23), M>>4 at [24,29)
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) |
                               ((uint32_t)(TC_BM >> 4) << 24);
        for (int kb = 0; kb < nkb; ++kb) {
            const int s = kb % S;
            const uint32_t ph = (uint32_t)(kb / S) & 1u;
            mbar_wait(&full_bar[s], ph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (lane == 0) {
                const uint32_t base = smem_u32(tiles + (size_t)s * stage_bytes);
                const uint32_t a_hi = base, a_lo = base + a_tile, b_hi = base + 2 * a_tile,
                               b_lo = base + 2 * a_tile + b_tile;
                const uint32_t a_lbo = TC_BM * 16, b_lbo = (uint32_t)BN * 16;   // stride between the two K chunks
#pragma unroll
                for (int j = 0; j < TC_BK / 8; ++j) {
                    const uint32_t ao = (uint32_t)(2 * j) * a_lbo, bo = (uint32_t)(2 * j) * b_lbo;
                    const uint64_t dah = make_smem_desc(a_hi + ao, a_lbo, 128);
                    const uint64_t dal = make_smem_desc(a_lo + ao, a_lbo, 128);
                    const uint64_t dbh = make_smem_desc(b_hi + bo, b_lbo, 128);
                    const uint64_t dbl = make_smem_desc(b_lo + bo, b_lbo, 128);
                    umma_tf32(tmem_base, dal, dbh, idesc, (kb | j) != 0 ? 1u : 0u);   // small terms first
                    umma_tf32(tmem_base, dah, dbl, idesc, 1u);
                    umma_tf32(tmem_base, dah, dbh, idesc, 1u);
                }
                umma_commit(&empty_bar[s]);              // frees the stage when these MMAs retire
                if (kb == nkb - 1) umma_commit(accum_bar);
            }
            __syncwarp();
        }
    }
    __syncthreads();
    if (warp == TC_PROD_WARPS) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols)
                     : "memory");
    }
}

}  // namespace

bool gemm_tc_supported(const GemmArgs& g) {
    return g.M > 0 && g.N > 0 && g.K > 0;
}

int launch_gemm_tc(const GemmArgs& g, cudaStream_t st) {
    TcParams p;
    p.g = g;
    int BN;
    if (g.N >= 256) {
        // balance the N tiles: e.g. N = 429 -> two tiles of 224
        const int64_t tiles = ceil_div64(g.N, 256);
        BN = (int)(ceil_div64(ceil_div64(g.N, tiles), 32) * 32);
    } else {
        BN = (int)(ceil_div64(g.N, 32) * 32);
    }
    p.BN = BN;
    p.tmem_cols = 32;
    while (p.tmem_cols < BN) p.tmem_cols <<= 1;
    const size_t stage_bytes = 2 * (size_t)TC_BM * TC_BK * 4 + 2 * (size_t)BN * TC_BK * 4;
    int stages = (int)((200 * 1024) / stage_bytes);
    if (stages > 4) stages = 4;
    if (stages < 2) {
        ctr_set_error("launch_gemm_tc: tile does not fit shared memory");
        return -1;
    }
    p.stages = stages;
    const int64_t gm = ceil_div64(g.M, TC_BM), gn = ceil_div64(g.N, BN);
    int64_t splits = 1;
    if (g.allow_split_k && g.epilogue == EPI_STORE) {
        const int64_t tiles = gm * gn, target = (int64_t)ctr_sm_count();
        if (tiles < target && g.K >= 8 * TC_BK) {
            splits = ceil_div64(target, tiles);
            const int64_t max_splits = g.K / (4 * TC_BK);
            if (splits > max_splits) splits = max_splits;
            if (splits < 1) splits = 1;
        }
    }
    p.k_chunk = ceil_div64(ceil_div64(g.K, splits), TC_BK) * TC_BK;
    splits = ceil_div64(g.K, p.k_chunk);
    if (gm > 65535 || splits > 65535) {
        ctr_set_error("launch_gemm_tc: grid too large");
        return -1;
    }
    if (splits > 1 && !g.accumulate)
        CTR_CUDA(cudaMemset2DAsync(g.C, g.ldc * sizeof(float), 0, g.N * sizeof(float), g.M, st));
    const size_t smem = (size_t)stages * stage_bytes + (2 * stages + 1) * sizeof(uint64_t) + 16;
    static size_t configured = 0;
    if (smem > configured) {
        CTR_CUDA(cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(220 * 1024)));
        configured = 220 * 1024;
    }
    dim3 grid((unsigned)gn, (unsigned)gm, (unsigned)splits);
    gemm_tc_kernel<<<grid, TC_THREADS, smem, st>>>(p);
    CTR_LAUNCH_OK("gemm_tc_kernel");
    return 0;
}
