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
//   warps 0-7  producers: global -> registers (128-bit loads) -> hi/lo split -> STS.128 into the
//              canonical no-swizzle UMMA core-matrix layout, fence.proxy.async + mbarrier arrive;
//              after the main loop the same warps run the epilogue: tcgen05.ld (32 lanes x 16
//              columns) -> bias/activation/mask -> global.
//   warp 8     TMEM alloc/dealloc; one elected lane issues the MMAs and tcgen05.commit's.
// Operand staging modes (chosen per operand on the host):
//   K-contiguous  (forward X, W; dgrad dY): 128-bit loads along k -> K-major core matrices
//   MN-contiguous (wgrad dY^T, X^T; dgrad W^T): 128-bit loads along m/n -> MN-major core matrices
//                 (instruction descriptor a_major/b_major = 1), no register transpose needed
//   scalar        any strides / unaligned: 32-bit loads -> K-major core matrices
// Pipeline: `stages` shared-memory stages of 32 fp32 of K (full/empty mbarriers), split-K across
// gridDim.z with fp32 atomics for the weight-gradient shapes.
#include <stdlib.h>

#include "gemm.cuh"
#include "tc_common.cuh"

namespace {

constexpr int TC_BM = 128;
constexpr int TC_BK = 32;                 // fp32 elements of K per stage
constexpr int TC_PROD_WARPS = 8;
constexpr int TC_THREADS = (TC_PROD_WARPS + 1) * 32;

enum { LD_SCALAR = 0, LD_KVEC = 1, LD_MNVEC = 2 };

struct TcParams {
    GemmArgs g;
    int64_t k_chunk;      // K range per split (multiple of TC_BK)
    int BN;               // N tile (multiple of 32)
    int stages;
    int tmem_cols;        // power of two >= max(32, BN)
    int a_mode, b_mode;
};

__device__ __forceinline__ float4 mask4(float4 v, float4 y, int act) {
    v.x *= act_grad_from_y(act, y.x);
    v.y *= act_grad_from_y(act, y.y);
    v.z *= act_grad_from_y(act, y.z);
    v.w *= act_grad_from_y(act, y.w);
    return v;
}

__device__ __forceinline__ void split_store(float* hi_ptr, float* lo_ptr, float4 v) {
    float4 hi, lo;
    split_tf32(v.x, hi.x, lo.x);
    split_tf32(v.y, hi.y, lo.y);
    split_tf32(v.z, hi.z, lo.z);
    split_tf32(v.w, hi.w, lo.w);
    *reinterpret_cast<float4*>(hi_ptr) = hi;
    *reinterpret_cast<float4*>(lo_ptr) = lo;
}

// Stage one operand tile (R rows x 32 k) of matrix P(row, k) = P[row*s_row + k*s_k] into shared
// memory as hi/lo, in the layout selected by `mode`.
template <int MODE>
__device__ __forceinline__ void stage_operand(const float* __restrict__ P, int64_t s_row, int64_t s_k,
                                              const float* __restrict__ mask, int64_t m_row, int64_t m_k,
                                              int mask_act, int64_t row0, int64_t n_rows, int64_t k0,
                                              int64_t kend, int R, float* hi, float* lo, int warp, int lane) {
    if (MODE == LD_MNVEC) {
        // unit = 4 consecutive rows; lane -> (k within group of 8, unit within 4); MN-major layout:
        // float offset = ((g * (R/4) + unit) * 8 + kk) * 4
        const int kk = lane & 7, ui = lane >> 3;
        for (int unit = warp * 4 + ui; unit < R / 4; unit += TC_PROD_WARPS * 4) {
            const int64_t row = row0 + unit * 4;
#pragma unroll
            for (int g = 0; g < TC_BK / 8; ++g) {
                const int64_t k = k0 + g * 8 + kk;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (k < kend) {
                    if (row + 3 < n_rows) {
                        v = __ldg(reinterpret_cast<const float4*>(P + row + k * s_k));
                        if (mask) v = mask4(v, __ldg(reinterpret_cast<const float4*>(mask + row + k * m_k)), mask_act);
                    } else {
                        float t[4] = {0.f, 0.f, 0.f, 0.f};
                        for (int e = 0; e < 4; ++e)
                            if (row + e < n_rows) {
                                t[e] = __ldg(P + row + e + k * s_k);
                                if (mask) t[e] *= act_grad_from_y(mask_act, __ldg(mask + row + e + k * m_k));
                            }
                        v = make_float4(t[0], t[1], t[2], t[3]);
                    }
                }
                const int off = ((g * (R / 4) + unit) * 8 + kk) * 4;
                split_store(hi + off, lo + off, v);
            }
        }
    } else {
        // K-major layout: float offset = (c * R + r) * 4, c = 16-byte chunk along k (0..7)
        const int r_lo = lane & 7, c_lo = lane >> 3;
        for (int U = warp; U < (R / 8) * 2; U += TC_PROD_WARPS) {
            const int r = (U >> 1) * 8 + r_lo;
            const int c = (U & 1) * 4 + c_lo;
            const int64_t row = row0 + r;
            const int64_t k = k0 + c * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (row < n_rows && k < kend) {
                if (MODE == LD_KVEC && k + 3 < kend) {
                    v = __ldg(reinterpret_cast<const float4*>(P + row * s_row + k));
                    if (mask) v = mask4(v, __ldg(reinterpret_cast<const float4*>(mask + row * m_row + k)), mask_act);
                } else {
                    float t[4] = {0.f, 0.f, 0.f, 0.f};
                    for (int e = 0; e < 4; ++e)
                        if (k + e < kend) {
                            t[e] = __ldg(P + row * s_row + (k + e) * s_k);
                            if (mask) t[e] *= act_grad_from_y(mask_act, __ldg(mask + row * m_row + (k + e) * m_k));
                        }
                    v = make_float4(t[0], t[1], t[2], t[3]);
                }
            }
            const int off = (c * R + r) * 4;
            split_store(hi + off, lo + off, v);
        }
    }
}

__device__ __forceinline__ void stage_dispatch(int mode, const float* P, int64_t s_row, int64_t s_k,
                                               const float* mask, int64_t m_row, int64_t m_k, int mask_act,
                                               int64_t row0, int64_t n_rows, int64_t k0, int64_t kend, int R,
                                               float* hi, float* lo, int warp, int lane) {
    if (mode == LD_KVEC)
        stage_operand<LD_KVEC>(P, s_row, s_k, mask, m_row, m_k, mask_act, row0, n_rows, k0, kend, R, hi, lo, warp, lane);
    else if (mode == LD_MNVEC)
        stage_operand<LD_MNVEC>(P, s_row, s_k, mask, m_row, m_k, mask_act, row0, n_rows, k0, kend, R, hi, lo, warp, lane);
    else
        stage_operand<LD_SCALAR>(P, s_row, s_k, mask, m_row, m_k, mask_act, row0, n_rows, k0, kend, R, hi, lo, warp, lane);
}

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
    if (warp == TC_PROD_WARPS) tmem_alloc_warp(tmem_slot, (uint32_t)p.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < TC_PROD_WARPS) {
        // ------------------------------ producers ------------------------------------------
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
            stage_dispatch(p.a_mode, g.A, g.sam, g.sak, g.amask, g.smm, g.smk, g.amask_act, m0, g.M, k0, kend,
                           TC_BM, a_hi, a_lo, warp, lane);
            stage_dispatch(p.b_mode, g.B, g.sbn, g.sbk, g.bmask, g.sbmn, g.sbmk, g.bmask_act, n0, g.N, k0, kend,
                           BN, b_hi, b_lo, warp, lane);
            fence_async_smem();                                           // generic -> async proxy
            __syncwarp();
            if (lane == 0) mbar_arrive(&full_bar[s]);
        }
        // ------------------------------ epilogue -------------------------------------------
        mbar_wait(accum_bar, 0);
        tc_fence_after();
        const int quad = warp & 3;                       // TMEM lane quadrant of this warp
        const int half = warp >> 2;                      // column half
        const int64_t m = m0 + quad * 32 + lane;
        const bool split = gridDim.z > 1;
        const int col_beg = half * (BN / 2), col_end = col_beg + BN / 2;
        const bool vec_store = !split && !g.accumulate && (g.ldc % 4 == 0) &&
                               ((reinterpret_cast<uintptr_t>(g.C) & 15) == 0) && g.epilogue != EPI_CROSS;
        for (int c0 = col_beg; c0 < col_end; c0 += 16) {
            uint32_t raw[16];
            tmem_ld16(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)c0, raw);
            tmem_ld_wait();
            if (m < g.M && n0 + c0 < g.N) {
                float out[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int64_t n = n0 + c0 + j;
                    float v = __uint_as_float(raw[j]);
                    if (n < g.N) {
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
                    }
                    out[j] = v;
                }
                float* crow = g.C + m * g.ldc + n0 + c0;
                if (vec_store && n0 + c0 + 15 < g.N) {
#pragma unroll
                    for (int j = 0; j < 16; j += 4)
                        *reinterpret_cast<float4*>(crow + j) = make_float4(out[j], out[j + 1], out[j + 2], out[j + 3]);
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        if (n0 + c0 + j < g.N) {
                            if (split) atomicAdd(crow + j, out[j]);
                            else if (g.accumulate) crow[j] += out[j];
                            else crow[j] = out[j];
                        }
                    }
                }
            }
        }
        tc_fence_before();
    } else {
        // ------------------------------ MMA issuer -----------------------------------------
        const uint32_t idesc = tf32_idesc(BN) | ((p.a_mode == LD_MNVEC) ? (1u << 15) : 0u) |
                               ((p.b_mode == LD_MNVEC) ? (1u << 16) : 0u);
        // per K atom (8 k): K-major tiles advance by two 16-byte chunks (2 * R * 16 B);
        // MN-major tiles by one 8-k group ((R/4) * 128 B)
        const uint32_t a_step = (p.a_mode == LD_MNVEC) ? (TC_BM / 4) * 128u : 2u * TC_BM * 16u;
        const uint32_t b_step = (p.b_mode == LD_MNVEC) ? ((uint32_t)BN / 4) * 128u : 2u * (uint32_t)BN * 16u;
        const uint32_t a_lbo = (p.a_mode == LD_MNVEC) ? (TC_BM / 4) * 128u : TC_BM * 16u;
        const uint32_t b_lbo = (p.b_mode == LD_MNVEC) ? ((uint32_t)BN / 4) * 128u : (uint32_t)BN * 16u;
        for (int kb = 0; kb < nkb; ++kb) {
            const int s = kb % S;
            const uint32_t ph = (uint32_t)(kb / S) & 1u;
            mbar_wait(&full_bar[s], ph);
            tc_fence_after();
            if (lane == 0) {
                const uint32_t base = smem_u32(tiles + (size_t)s * stage_bytes);
                const uint32_t a_hi = base, a_lo = base + a_tile, b_hi = base + 2 * a_tile,
                               b_lo = base + 2 * a_tile + b_tile;
#pragma unroll
                for (int j = 0; j < TC_BK / 8; ++j) {
                    const uint32_t ao = (uint32_t)j * a_step, bo = (uint32_t)j * b_step;
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
        tc_fence_after();
        tmem_dealloc_warp(tmem_base, (uint32_t)p.tmem_cols);
    }
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// pick the staging mode of one operand P(row,k) = P[row*s_row + k*s_k] (+ optional mask)
int pick_mode(const float* P, int64_t s_row, int64_t s_k, const float* mask, int64_t m_row, int64_t m_k) {
    const char* e = getenv("CTR_TC_LOAD");
    if (e && e[0] == 's') return LD_SCALAR;
    if (s_k == 1 && s_row % 4 == 0 && aligned16(P) &&
        (!mask || (m_k == 1 && m_row % 4 == 0 && aligned16(mask))))
        return LD_KVEC;
    if (s_row == 1 && s_k % 4 == 0 && aligned16(P) &&
        (!mask || (m_row == 1 && m_k % 4 == 0 && aligned16(mask))))
        return (e && e[0] == 'k') ? LD_SCALAR : LD_MNVEC;
    return LD_SCALAR;
}

}  // namespace

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
    p.a_mode = pick_mode(g.A, g.sam, g.sak, g.amask, g.smm, g.smk);
    p.b_mode = pick_mode(g.B, g.sbn, g.sbk, g.bmask, g.sbmn, g.sbmk);
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
    static bool configured = false;
    if (!configured) {
        CTR_CUDA(cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(220 * 1024)));
        configured = true;
    }
    dim3 grid((unsigned)gn, (unsigned)gm, (unsigned)splits);
    gemm_tc_kernel<<<grid, TC_THREADS, smem, st>>>(p);
    CTR_LAUNCH_OK("gemm_tc_kernel");
    return 0;
}
