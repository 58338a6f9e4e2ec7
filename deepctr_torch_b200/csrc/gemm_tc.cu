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
constexpr int TC_PROD_WARPS = 8;          // producer warps per group
constexpr int TC_GROUPS = 2;              // producer groups working on alternating K stages
constexpr int TC_MMA_WARP = TC_PROD_WARPS * TC_GROUPS;
constexpr int TC_THREADS = (TC_MMA_WARP + 1) * 32;

enum { LD_SCALAR = 0, LD_KVEC = 1, LD_TRANS = 2 };

struct TcParams {
    unsigned long long* dbg;   // optional timeline buffer (ctr_debug_set_buffer), CTA (0,0,0) only
    GemmArgs g;
    int64_t k_chunk;      // K range per split (multiple of TC_BK)
    int BN;               // N tile (multiple of 32)
    int stages;
    int tmem_cols;        // power of two >= max(32, BN)
    int a_mode, b_mode;
};

struct OperandView {
    const float* P; int64_t s_row, s_k;          // P(row,k) = P[row*s_row + k*s_k]
    const float* mask; int64_t m_row, m_k; int mask_act;
    int64_t row0, n_rows;
};

constexpr int kMaxItems = 8;   // float4 registers per operand per stage (R <= 256)
constexpr int kItemsA = 4;     // the A tile always has 128 rows

// ---- phase 1: issue every global load of a stage into registers — loads ONLY, nothing here reads
// the loaded values.  They are consumed one pipeline iteration later by tile_store(), so the memory
// latency (measured: ~3.4k cycles per stage when exposed) overlaps the MMA of the previous stage.
// The loads are UNCONDITIONAL: indices are clamped to a valid element and out-of-range lanes are
// zeroed with selects in tile_store() (a load inside a data-dependent branch is waited for at the
// join).  Preconditions checked by pick_mode(): KVEC/TRANS operands are readable up to the next
// multiple of 4 elements along their contiguous axis (leading dimension >= round_up(extent, 4)).
__device__ __forceinline__ float4 sel4(float4 v, bool k0, bool k1, bool k2, bool k3) {
    return make_float4(k0 ? v.x : 0.f, k1 ? v.y : 0.f, k2 ? v.z : 0.f, k3 ? v.w : 0.f);
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

template <int MODE, int NIT, bool MASK>
__device__ __forceinline__ void tile_load(const OperandView& o, int R, int64_t k0, int64_t kend, int warp,
                                          int lane, float4 (&v)[NIT], float4 (&vm)[NIT]) {
    if (MODE == LD_TRANS) {
        // item = (row quad, chunk): 4 loads of 4 consecutive rows at k, k+1, k+2, k+3
        const int rq_l = (lane >> 3) * 2 + (lane & 1), c_l = (lane >> 1) & 3;
#pragma unroll
        for (int it = 0; it < NIT / 4; ++it) {
            const int U = warp + TC_PROD_WARPS * it;
            const bool live = U < (R / 32) * 2;
            const int64_t row = o.row0 + ((U >> 1) * 8 + rq_l) * 4;
            const int64_t k = k0 + ((U & 1) * 4 + c_l) * 4;
            const int64_t rowc = (live && row < o.n_rows) ? row : 0;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int64_t kc = (k + e < kend) ? k + e : k0;
                v[it * 4 + e] = __ldg(reinterpret_cast<const float4*>(o.P + rowc + kc * o.s_k));
                if (MASK) vm[it * 4 + e] = __ldg(reinterpret_cast<const float4*>(o.mask + rowc + kc * o.m_k));
            }
        }
    } else {
        const int r_l = lane & 7, c_l = lane >> 3;
#pragma unroll
        for (int it = 0; it < NIT; ++it) {
            const int U = warp + TC_PROD_WARPS * it;
            const bool live = U < (R / 8) * 2;
            const int64_t row = o.row0 + (U >> 1) * 8 + r_l;
            const int64_t k = k0 + ((U & 1) * 4 + c_l) * 4;
            const bool ok = live && row < o.n_rows && k < kend;
            const int64_t rowc = ok ? row : o.row0, kc = ok ? k : k0;
            if (MODE == LD_KVEC) {
                v[it] = __ldg(reinterpret_cast<const float4*>(o.P + rowc * o.s_row + kc));
                if (MASK) vm[it] = __ldg(reinterpret_cast<const float4*>(o.mask + rowc * o.m_row + kc));
            } else {
                const bool e1 = ok && k + 1 < kend, e2 = ok && k + 2 < kend, e3 = ok && k + 3 < kend;
                const float* src = o.P + rowc * o.s_row + kc * o.s_k;
                v[it].x = __ldg(src);
                v[it].y = __ldg(src + (e1 ? o.s_k : 0));
                v[it].z = __ldg(src + (e2 ? 2 * o.s_k : 0));
                v[it].w = __ldg(src + (e3 ? 3 * o.s_k : 0));
                if (MASK) {
                    const float* ms = o.mask + rowc * o.m_row + kc * o.m_k;
                    vm[it].x = __ldg(ms);
                    vm[it].y = __ldg(ms + (e1 ? o.m_k : 0));
                    vm[it].z = __ldg(ms + (e2 ? 2 * o.m_k : 0));
                    vm[it].w = __ldg(ms + (e3 ? 3 * o.m_k : 0));
                }
            }
        }
    }
}

// L2 prefetch of a future stage's operand lines (fire and forget; one request per 128-byte line)
template <int MODE>
__device__ __forceinline__ void tile_prefetch(const OperandView& o, int R, int64_t k0, int64_t kend, int warp,
                                              int lane) {
    if (k0 >= kend) return;
    if (MODE == LD_TRANS) {
        // rows are contiguous: one line = 32 rows of one k; lanes 0..(R/32-1) x 32 k rows
        const int k_l = lane;                       // 32 k per stage
        for (int blk = warp; blk < R / 32; blk += TC_PROD_WARPS) {
            const int64_t row = o.row0 + blk * 32;
            if (row < o.n_rows && k0 + k_l < kend) prefetch_l2(o.P + row + (k0 + k_l) * o.s_k);
        }
    } else if (MODE == LD_KVEC) {
        // one line = the 32 k of one row
        for (int r = warp * 32 + lane; r < R; r += TC_PROD_WARPS * 32) {
            const int64_t row = o.row0 + r;
            if (row < o.n_rows) prefetch_l2(o.P + row * o.s_row + k0);
        }
    }
}

__device__ __forceinline__ float4 apply_mask(float4 x, float4 y, int act) {
    // zeroed (out-of-range) lanes stay exactly zero even if the clamped mask element is not finite
    x.x = (x.x == 0.f) ? 0.f : x.x * act_grad_from_y(act, y.x);
    x.y = (x.y == 0.f) ? 0.f : x.y * act_grad_from_y(act, y.y);
    x.z = (x.z == 0.f) ? 0.f : x.z * act_grad_from_y(act, y.z);
    x.w = (x.w == 0.f) ? 0.f : x.w * act_grad_from_y(act, y.w);
    return x;
}

// ---- phase 2 (one iteration later): zero the out-of-range lanes, mask, split into hi/lo, store
template <int MODE, int NIT, bool MASK>
__device__ __forceinline__ void tile_store(const OperandView& o, int R, int64_t k0, int64_t kend, int warp, int lane,
                                           float4 (&v)[NIT], float4 (&vm)[NIT], float* hi, float* lo) {
    if (MODE == LD_TRANS) {
        const int rq_l = (lane >> 3) * 2 + (lane & 1), c_l = (lane >> 1) & 3;
#pragma unroll
        for (int it = 0; it < NIT / 4; ++it) {
            const int U = warp + TC_PROD_WARPS * it;
            if (U < (R / 32) * 2) {
                const int r = ((U >> 1) * 8 + rq_l) * 4, c = (U & 1) * 4 + c_l;
                const int64_t row = o.row0 + r, k = k0 + c * 4;
                const bool r0 = row < o.n_rows, r1 = row + 1 < o.n_rows, r2 = row + 2 < o.n_rows, r3 = row + 3 < o.n_rows;
                float4 x[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const bool k_in = k + e < kend;
                    x[e] = sel4(v[it * 4 + e], r0 && k_in, r1 && k_in, r2 && k_in, r3 && k_in);
                    if (MASK) x[e] = apply_mask(x[e], vm[it * 4 + e], o.mask_act);
                }
                // 4x4 register transpose: row r+i gets (k, k+1, k+2, k+3)
                split_store(hi + tile_off(R, r + 0, c), lo + tile_off(R, r + 0, c), make_float4(x[0].x, x[1].x, x[2].x, x[3].x));
                split_store(hi + tile_off(R, r + 1, c), lo + tile_off(R, r + 1, c), make_float4(x[0].y, x[1].y, x[2].y, x[3].y));
                split_store(hi + tile_off(R, r + 2, c), lo + tile_off(R, r + 2, c), make_float4(x[0].z, x[1].z, x[2].z, x[3].z));
                split_store(hi + tile_off(R, r + 3, c), lo + tile_off(R, r + 3, c), make_float4(x[0].w, x[1].w, x[2].w, x[3].w));
            }
        }
    } else {
        const int r_l = lane & 7, c_l = lane >> 3;
#pragma unroll
        for (int it = 0; it < NIT; ++it) {
            const int U = warp + TC_PROD_WARPS * it;
            if (U < (R / 8) * 2) {
                const int r = (U >> 1) * 8 + r_l, c = (U & 1) * 4 + c_l;
                const int64_t row = o.row0 + r, k = k0 + c * 4;
                const bool ok = row < o.n_rows && k < kend;
                float4 x = sel4(v[it], ok, ok && k + 1 < kend, ok && k + 2 < kend, ok && k + 3 < kend);
                if (MASK) x = apply_mask(x, vm[it], o.mask_act);
                split_store(hi + tile_off(R, r, c), lo + tile_off(R, r, c), x);
            }
        }
    }
}

#define TC_MODE_SWITCH(mode, has_mask, CALL)                                  \
    do {                                                                      \
        if (has_mask) {                                                       \
            if ((mode) == LD_KVEC) { CALL(LD_KVEC, true); }                   \
            else if ((mode) == LD_TRANS) { CALL(LD_TRANS, true); }            \
            else { CALL(LD_SCALAR, true); }                                   \
        } else {                                                              \
            if ((mode) == LD_KVEC) { CALL(LD_KVEC, false); }                  \
            else if ((mode) == LD_TRANS) { CALL(LD_TRANS, false); }           \
            else { CALL(LD_SCALAR, false); }                                  \
        }                                                                     \
    } while (0)

__global__ void __launch_bounds__(TC_THREADS, 1) gemm_tc_kernel(TcParams p) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const GemmArgs& g = p.g;
    const int BN = p.BN, S = p.stages;
    const uint32_t a_tile = (TC_BM + 1) * 128;          // bytes of one A tile (hi or lo), see tile_off
    const uint32_t b_tile = ((uint32_t)BN + 1) * 128;
    const uint32_t stage_bytes = 2 * a_tile + 2 * b_tile;
    unsigned char* tiles = smem_raw;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_raw + (size_t)S * stage_bytes);
    uint64_t* empty_bar = full_bar + S;
    uint64_t* accum_bar = empty_bar + S;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

    const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
    const int group = wid / TC_PROD_WARPS;    // producer group (TC_GROUPS = the MMA warp)
    const int warp = wid % TC_PROD_WARPS;     // warp index inside the group (item mapping)
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
    if (wid == TC_MMA_WARP) tmem_alloc_warp(tmem_slot, (uint32_t)p.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (wid < TC_MMA_WARP) {
        // ------------------------------ producers ------------------------------------------
        // group g stages the K blocks kb = g, g + TC_GROUPS, ...: while one group waits for its loads
        // the other one is splitting/storing, so two stages of global loads are always in flight
        OperandView oa{g.A, g.sam, g.sak, g.amask, g.smm, g.smk, g.amask_act, m0, g.M};
        OperandView ob{g.B, g.sbn, g.sbk, g.bmask, g.sbmn, g.sbmk, g.bmask_act, n0, g.N};
        const bool dbg = p.dbg && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && tid == 0;
        constexpr int NG = TC_GROUPS;
        if (dbg) p.dbg[0] = clock64();
        float4 va[kItemsA], vam[kItemsA], vb[kMaxItems], vbm[kMaxItems];
#define TC_LOAD_A(MODE, MK) tile_load<MODE, kItemsA, MK>(oa, TC_BM, kl, kend, warp, lane, va, vam)
#define TC_LOAD_B(MODE, MK) tile_load<MODE, kMaxItems, MK>(ob, BN, kl, kend, warp, lane, vb, vbm)
#define TC_STORE_A(MODE, MK) tile_store<MODE, kItemsA, MK>(oa, TC_BM, k0, kend, warp, lane, va, vam, a_hi, a_lo)
#define TC_STORE_B(MODE, MK) tile_store<MODE, kMaxItems, MK>(ob, BN, k0, kend, warp, lane, vb, vbm, b_hi, b_lo)
#define TC_PF_A(MODE, MK) tile_prefetch<MODE>(oa, TC_BM, kp, kend, warp, lane)
#define TC_PF_B(MODE, MK) tile_prefetch<MODE>(ob, BN, kp, kend, warp, lane)
        if (group < nkb) {   // prologue: L2 prefetch of this group's next stages, register loads of its first
            for (int d = 1; d <= 2; ++d) {
                const int64_t kp = kbeg + (int64_t)(group + d * NG) * TC_BK;
                TC_MODE_SWITCH(p.a_mode, false, TC_PF_A);
                TC_MODE_SWITCH(p.b_mode, false, TC_PF_B);
            }
            const int64_t kl = kbeg + (int64_t)group * TC_BK;
            TC_MODE_SWITCH(p.a_mode, oa.mask != nullptr, TC_LOAD_A);
            TC_MODE_SWITCH(p.b_mode, ob.mask != nullptr, TC_LOAD_B);
        }
        for (int kb = group; kb < nkb; kb += NG) {
            const int s = kb % S;
            const uint32_t ph = (uint32_t)(kb / S) & 1u;
            const int64_t k0 = kbeg + (int64_t)kb * TC_BK;
            if (dbg && kb < 30) p.dbg[8 + kb * 8 + 0] = clock64();
            mbar_wait(&empty_bar[s], ph ^ 1u);
            if (dbg && kb < 30) p.dbg[8 + kb * 8 + 2] = clock64();
            unsigned char* st = tiles + (size_t)s * stage_bytes;
            float* a_hi = reinterpret_cast<float*>(st);
            float* a_lo = reinterpret_cast<float*>(st + a_tile);
            float* b_hi = reinterpret_cast<float*>(st + 2 * a_tile);
            float* b_lo = reinterpret_cast<float*>(st + 2 * a_tile + b_tile);
            // consume the registers loaded one iteration ago
            TC_MODE_SWITCH(p.a_mode, oa.mask != nullptr, TC_STORE_A);
            TC_MODE_SWITCH(p.b_mode, ob.mask != nullptr, TC_STORE_B);
            if (dbg && kb < 30) p.dbg[8 + kb * 8 + 3] = clock64();
            fence_async_smem();                                           // generic -> async proxy
            __syncwarp();
            if (lane == 0) mbar_arrive(&full_bar[s]);
            if (dbg && kb < 30) p.dbg[8 + kb * 8 + 4] = clock64();
            // issue the next stage's loads (consumed after the next empty-barrier wait) and the L2
            // prefetch of the stage four ahead
            if (kb + NG < nkb) {
                const int64_t kl = k0 + (int64_t)NG * TC_BK;
                TC_MODE_SWITCH(p.a_mode, oa.mask != nullptr, TC_LOAD_A);
                TC_MODE_SWITCH(p.b_mode, ob.mask != nullptr, TC_LOAD_B);
                const int64_t kp = k0 + (int64_t)3 * NG * TC_BK;
                TC_MODE_SWITCH(p.a_mode, false, TC_PF_A);
                TC_MODE_SWITCH(p.b_mode, false, TC_PF_B);
            }
            if (dbg && kb < 30) p.dbg[8 + kb * 8 + 1] = clock64();
        }
        // ------------------------------ epilogue (group 0) ---------------------------------
        if (dbg) p.dbg[1] = clock64();
        if (group == 0) {
        mbar_wait(accum_bar, 0);
        if (dbg) p.dbg[2] = clock64();
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
                            case EPI_MUL:
                                v *= __ldg(g.aux + m * g.ldaux + n);
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
        if (dbg) p.dbg[3] = clock64();
        }
    } else {
        // ------------------------------ MMA issuer -----------------------------------------
        const bool dbgm = p.dbg && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && lane == 0;
        const uint32_t idesc = tf32_idesc(BN);
        // per K atom (8 k = two 16-byte chunks) the tiles advance by 2 * (R+1) * 16 bytes
        const uint32_t a_lbo = (TC_BM + 1) * 16u, b_lbo = ((uint32_t)BN + 1) * 16u;
        const uint32_t a_step = 2u * a_lbo, b_step = 2u * b_lbo;
        for (int kb = 0; kb < nkb; ++kb) {
            const int s = kb % S;
            const uint32_t ph = (uint32_t)(kb / S) & 1u;
            mbar_wait(&full_bar[s], ph);
            if (dbgm && kb < 30) p.dbg[8 + kb * 8 + 5] = clock64();
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
                if (dbgm && kb < 30) p.dbg[8 + kb * 8 + 6] = clock64();
            }
            __syncwarp();
        }
    }
    __syncthreads();
    if (wid == TC_MMA_WARP) {
        tc_fence_after();
        tmem_dealloc_warp(tmem_base, (uint32_t)p.tmem_cols);
    }
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// pick the staging mode of one operand P(row,k) = P[row*s_row + k*s_k] (+ optional mask).
// Vector modes read whole 16-byte groups, so the operand must be readable up to the next multiple
// of 4 along its contiguous axis: guaranteed when the leading dimension covers round_up(extent, 4).
int pick_mode(const float* P, int64_t s_row, int64_t s_k, const float* mask, int64_t m_row, int64_t m_k,
              int64_t n_rows, int64_t K) {
    const char* e = getenv("CTR_TC_LOAD");
    if (e && e[0] == 's') return LD_SCALAR;
    const int64_t K4 = (K + 3) / 4 * 4, R4 = (n_rows + 3) / 4 * 4;
    if (s_k == 1 && s_row % 4 == 0 && s_row >= K4 && aligned16(P) &&
        (!mask || (m_k == 1 && m_row % 4 == 0 && m_row >= K4 && aligned16(mask))))
        return LD_KVEC;
    if (s_row == 1 && s_k % 4 == 0 && s_k >= R4 && aligned16(P) &&
        (!mask || (m_row == 1 && m_k % 4 == 0 && m_k >= R4 && aligned16(mask))))
        return (e && e[0] == 'k') ? LD_SCALAR : LD_TRANS;
    return LD_SCALAR;
}

}  // namespace

static int g_gemm_passes = 3;
int ctr_gemm_passes() { return g_gemm_passes; }
extern "C" int ctr_set_gemm_passes(int passes) {
    const int prev = g_gemm_passes;
    g_gemm_passes = passes == 1 ? 1 : 3;
    return prev;
}

static unsigned long long* g_dbg_buf = nullptr;
unsigned long long* ctr_debug_buffer() { return g_dbg_buf; }
extern "C" int ctr_debug_set_buffer(void* ptr) {
    g_dbg_buf = reinterpret_cast<unsigned long long*>(ptr);
    return 0;
}

int launch_gemm_tc(const GemmArgs& g, cudaStream_t st) {
    TcParams p;
    p.dbg = g_dbg_buf;
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
    p.a_mode = pick_mode(g.A, g.sam, g.sak, g.amask, g.smm, g.smk, g.M, g.K);
    p.b_mode = pick_mode(g.B, g.sbn, g.sbk, g.bmask, g.sbmn, g.sbmk, g.N, g.K);
    const size_t stage_bytes = 2 * (size_t)(TC_BM + 1) * 128 + 2 * (size_t)(BN + 1) * 128;
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
