// CIN forward and weight gradient, second generation.  Written at the end of round 1, validated on hardware in
// round 2 (tests/test_gpu_cin_v2.py; xDeepFM config #3: forward 9.9 -> 6.2 ms, backward 24.9 -> 22.0 ms per step)
// and the default since; CTR_CIN_V2=0 selects the round-1 kernels of cin_tc.cu.
//
//   Z[(b,d), n] = sum_{hm} P[(b,d), hm] * W[n, hm],   P[(b,d), h*M+m] = Xp[b,h,d] * X0[b,m,d]
//   (reference layers/interaction.py:207-248; bias + activation, Y store and the direct-connect sum over d
//   in the epilogue)
//
// What changes against cin_tc_fwd_kernel (profiles/SUMMARY.md: 9.9 ms per xDeepFM step, 13-15 % tensor
// activity, producer-bound), applying what the GEMM engine (gemm_pk.cu) taught this round:
//   * W arrives pre-split (hi, lo) in packed K-major tiles by 1-D TMA — no weight work in the CTA at all;
//   * 16 generator warps instead of 8, and a generator thread owns ONE (b,d) row for the whole kernel:
//     x0[b,:,d] and xp[b,:,d] sit transposed in shared memory ([m][row] / [h][row]: lanes = consecutive rows,
//     conflict-free scalar LDS), (h, m) of the running k are tracked incrementally — no division, no
//     k -> (h,m) lookup table, one multiply + split per generated element;
//   * CTA tile = 256 rows (two TMEM accumulators) x BN channels, so a W stage is used twice.
#include <stdlib.h>

#include "gemm.cuh"
#include "tc_common.cuh"

namespace {

constexpr int V2_GEN_WARPS = 16;
constexpr int V2_THREADS = 64 + V2_GEN_WARPS * 32;
constexpr int V2_ROWS = 256;             // (b,d) rows per CTA = two 128-row MMA tiles
constexpr int V2_KB = 16;                // k per stage

struct CinV2Fwd {
    const float* Xp; int64_t sxp; int H;
    const float* X0; int64_t sx0; int M;
    int D;
    const float* Wp; int wp_nkb;         // packed W: [N tiles][wp_nkb][hi | lo][4 chunks][BN rows][4]
    const float* bias; int N; int direct_start; int act;
    float* Y; float* out; int64_t ld_out; int64_t B;
    int BN, TB, SB, tmem_cols;
    int chains;                          // accumulation chains per row tile (2: even / odd k stages, summed in the epilogue)
    uint32_t off_b, off_xp, off_x0, off_bar;
};

__device__ __forceinline__ void v2_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void v2_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void v2_tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
}
__device__ __forceinline__ void v2_split_store(float* tile, int off, int lo_off, float4 v) {
    float4 hi, lo;
    split_tf32(v.x, hi.x, lo.x);
    split_tf32(v.y, hi.y, lo.y);
    split_tf32(v.z, hi.z, lo.z);
    split_tf32(v.w, hi.w, lo.w);
    *reinterpret_cast<float4*>(tile + off) = hi;
    *reinterpret_cast<float4*>(tile + off + lo_off) = lo;
}

__global__ void __launch_bounds__(V2_THREADS, 1) cin_v2_fwd_kernel(CinV2Fwd a) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const int D = a.D, H = a.H, M = a.M, N = a.N, BN = a.BN, SB = a.SB;
    constexpr int SA = 2;
    const uint32_t a_stage = V2_ROWS * 128u;                 // two 128-row sub-tiles, [hi | lo] each
    const uint32_t b_stage = (uint32_t)BN * 128u;
    unsigned char* ringA = smem_raw;
    unsigned char* ringB = smem_raw + a.off_b;
    float* sXpT = reinterpret_cast<float*>(smem_raw + a.off_xp);     // [H][256]
    float* sX0T = reinterpret_cast<float*>(smem_raw + a.off_x0);     // [M][256]
    uint64_t* a_full = reinterpret_cast<uint64_t*>(smem_raw + a.off_bar);
    uint64_t* a_empty = a_full + SA;
    uint64_t* b_full = a_empty + SA;
    uint64_t* b_empty = b_full + SB;
    uint64_t* accum_bar = b_empty + SB;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

    const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
    const int64_t b0 = (int64_t)blockIdx.x * a.TB;
    const int n0 = blockIdx.y * BN;
    const int tb = (int)((a.B - b0 < a.TB) ? a.B - b0 : a.TB);
    const int rows = tb * D;
    const int HM = H * M;
    const int nkb = (HM + V2_KB - 1) / V2_KB;

    if (tid == 0) {
        for (int s = 0; s < SA; ++s) {
            mbar_init(&a_full[s], V2_GEN_WARPS);
            mbar_init(&a_empty[s], 1);
        }
        for (int s = 0; s < SB; ++s) {
            mbar_init(&b_full[s], 1);
            mbar_init(&b_empty[s], 1);
        }
        mbar_init(accum_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (wid == 1) tmem_alloc_warp(tmem_slot, (uint32_t)a.tmem_cols);
    // x_p and x_0 of the CTA's samples, transposed: [h][row], row = bl*D + d (zeros beyond the batch)
    for (int i = tid; i < H * V2_ROWS; i += V2_THREADS) {
        const int h = i / V2_ROWS, r = i - h * V2_ROWS;
        const int bl = r / D, d = r - bl * D;
        sXpT[i] = (r < rows) ? __ldg(a.Xp + (b0 + bl) * a.sxp + (int64_t)h * D + d) : 0.f;
    }
    for (int i = tid; i < M * V2_ROWS; i += V2_THREADS) {
        const int m = i / V2_ROWS, r = i - m * V2_ROWS;
        const int bl = r / D, d = r - bl * D;
        sX0T[i] = (r < rows) ? __ldg(a.X0 + (b0 + bl) * a.sx0 + (int64_t)m * D + d) : 0.f;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (wid == 0) {
        // ------------------------------ TMA producer: packed W tiles ------------------------
        if (lane == 0) {
            const unsigned char* src = reinterpret_cast<const unsigned char*>(a.Wp) + (size_t)blockIdx.y * a.wp_nkb * b_stage;
            int sb = 0;
            uint32_t ph = 1;
            for (int i = 0; i < nkb; ++i) {
                mbar_wait(&b_empty[sb], ph);
                v2_expect_tx(&b_full[sb], b_stage);
                v2_bulk_g2s(ringB + (size_t)sb * b_stage, src + (size_t)i * b_stage, b_stage, &b_full[sb]);
                if (++sb == SB) { sb = 0; ph ^= 1u; }
            }
        }
        __syncwarp();
    } else if (wid == 1) {
        // ------------------------------ MMA issuer ------------------------------------------
        const uint32_t idesc = tf32_idesc(BN);
        const uint32_t a_lbo = 128u * 16u, b_lbo = (uint32_t)BN * 16u, a_tile = 128u * 128u;
        int sa = 0, sb = 0;
        uint32_t pha = 0, phb = 0;
        for (int i = 0; i < nkb; ++i) {
            mbar_wait(&a_full[sa], pha);
            mbar_wait(&b_full[sb], phb);
            tc_fence_after();
            if (lane == 0) {
                const uint32_t a_base = smem_u32(ringA + (size_t)sa * a_stage);
                const uint32_t b_hi = smem_u32(ringB + (size_t)sb * b_stage), b_lo = b_hi + b_stage / 2;
#pragma unroll
                for (int j = 0; j < V2_KB / 8; ++j) {
                    const uint64_t dbh = make_smem_desc(b_hi + (uint32_t)j * 2u * b_lbo, b_lbo, 128);
                    const uint64_t dbl = make_smem_desc(b_lo + (uint32_t)j * 2u * b_lbo, b_lbo, 128);
#pragma unroll
                    for (int mt = 0; mt < 2; ++mt) {
                        const uint32_t a_hi = a_base + (uint32_t)mt * a_tile + (uint32_t)j * 2u * a_lbo;
                        const uint64_t dah = make_smem_desc(a_hi, a_lbo, 128);
                        const uint64_t dal = make_smem_desc(a_hi + a_tile / 2, a_lbo, 128);
                        // two accumulation chains per row tile (even / odd stages) halve the number of truncating
                        // fp32 adds a sum goes through inside the tensor core (K = 1664 on one chain: 1.05e-5)
                        const int ch = (a.chains == 2) ? (i & 1) : 0;
                        const uint32_t dacc = tmem_base + (uint32_t)((mt * a.chains + ch) * BN);
                        umma_tf32(dacc, dal, dbh, idesc, (i >= a.chains || j != 0) ? 1u : 0u);      // small terms first
                        umma_tf32(dacc, dah, dbl, idesc, 1u);
                        umma_tf32(dacc, dah, dbh, idesc, 1u);
                    }
                }
                umma_commit(&a_empty[sa]);
                umma_commit(&b_empty[sb]);
                if (i == nkb - 1) umma_commit(accum_bar);
            }
            __syncwarp();
            if (++sa == SA) { sa = 0; pha ^= 1u; }
            if (++sb == SB) { sb = 0; phb ^= 1u; }
        }
    } else {
        // ------------------------------ generators (16 warps) -------------------------------
        const int ct = tid - 64;
        const int r = ct & (V2_ROWS - 1), half = ct >> 8;            // row of this thread, which 8 of the 16 k
        const float* xpr = sXpT + r;
        const float* x0r = sX0T + r;
        // destination of this thread's two chunks inside a stage: sub-tile r / 128, chunks 2*half, 2*half+1
        const int st = r >> 7, rr = r & 127;
        const int off0 = st * (128 * 32) + ((2 * half) * 128 + rr) * 4, off1 = off0 + 128 * 4, lo_off = 128 * 16;
        int h = (8 * half) / M, m = (8 * half) - h * M;            // (h, m) of the first k of the running stage
        int sa = 0;
        uint32_t pha = 1;
        for (int i = 0; i < nkb; ++i) {
            float v[8];
            {
                int hh = h, mm = m;
                float xp = (hh < H) ? xpr[hh * V2_ROWS] : 0.f;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    v[j] = (hh < H) ? xp * x0r[mm * V2_ROWS] : 0.f;
                    if (++mm == M) {
                        mm = 0;
                        ++hh;
                        xp = (hh < H) ? xpr[hh * V2_ROWS] : 0.f;
                    }
                }
            }
            mbar_wait(&a_empty[sa], pha);
            float* tile = reinterpret_cast<float*>(ringA + (size_t)sa * a_stage);
            v2_split_store(tile, off0, lo_off, make_float4(v[0], v[1], v[2], v[3]));
            v2_split_store(tile, off1, lo_off, make_float4(v[4], v[5], v[6], v[7]));
            fence_async_smem();                                   // generic-proxy stores -> async proxy (tcgen05.mma)
            __syncwarp();
            if (lane == 0) mbar_arrive(&a_full[sa]);
            if (++sa == SA) { sa = 0; pha ^= 1u; }
            m += V2_KB;                                           // next stage: k advances by 16
            while (m >= M) {
                m -= M;
                ++h;
            }
        }
        // ------------------------------ epilogue (same 16 warps) ----------------------------
        mbar_wait(accum_bar, 0);
        tc_fence_after();
        const int quad = wid & 3, cg = (wid - 2) >> 2;
        const int nchunks = (BN + 31) / 32;
#pragma unroll 1
        for (int mt = 0; mt < 2; ++mt) {
            const int er = mt * 128 + quad * 32 + lane;
            const int bl = er / D, d = er - bl * D;
            const bool row_ok = er < rows;
            const int64_t b = b0 + bl;
            if (mt * 128 + quad * 32 >= rows) continue;           // warp-uniform: nothing in these 32 rows
            for (int ci = cg; ci < nchunks; ci += V2_GEN_WARPS / 4) {
                const int c0 = ci * 32;
                if (n0 + c0 >= N) break;                          // warp-uniform
                uint32_t raw[32];
                v2_tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(mt * a.chains * BN + c0), raw);
                tmem_ld_wait();
                if (a.chains == 2 && nkb > 1) {
                    uint32_t raw2[32];
                    v2_tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)((mt * 2 + 1) * BN + c0), raw2);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j) raw[j] = __float_as_uint(__uint_as_float(raw[j]) + __uint_as_float(raw2[j]));
                }
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const int n = n0 + c0 + j;
                    if (n >= N || c0 + j >= BN) break;            // warp-uniform
                    const float y = act_apply(a.act, __uint_as_float(raw[j]) + (a.bias ? __ldg(a.bias + n) : 0.f));
                    if (row_ok) a.Y[(b * N + n) * D + d] = y;
                    if (n >= a.direct_start && a.out) {
                        float sacc = row_ok ? y : 0.f;
                        for (int o = 1; o < D; o <<= 1) sacc += __shfl_xor_sync(0xffffffffu, sacc, o);
                        if (row_ok && d == 0) a.out[b * a.ld_out + (n - a.direct_start)] = sacc;
                    }
                }
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (wid == 1) {
        tc_fence_after();
        tmem_dealloc_warp(tmem_base, (uint32_t)a.tmem_cols);
    }
}


// =============================================================================================
// dW[n, hm] = sum_{(b,d)} dZ[b,n,d] * Xp[b,h,d] * X0[b,m,d]     (D == 16)
//
// GEMM with M = channels (128 per CTA), N' = 256 (h,m) columns, K = (b,d): one 16-k stage = one sample.
//   A(n, (b,d)) = dZ[b, n, d]            : the sample's [N][16] block IS a K-contiguous tile
//   B(hm, (b,d)) = Xp[b,h,d] * X0[b,m,d] : generated; (h, m) of a thread's two rows are kernel constants
// 16 generator warps, loads of sample b+1 in registers while sample b is split and stored, fp32
// reductions (red.global.add.v4 through a transposing staging tile) because K is split across CTAs.
// =============================================================================================
struct CinV2Dw {
    const float* Xp; int64_t sxp; int H;
    const float* X0; int64_t sx0; int M;
    int N;
    const float* dZ; float* dW;
    int64_t B; int64_t b_per_cta;
    int S; uint32_t off_bar;
};

constexpr int V2_STG_PITCH = 36;

__global__ void __launch_bounds__(V2_THREADS, 1) cin_v2_dw_kernel(CinV2Dw a) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const int M = a.M, N = a.N, S = a.S;
    const int HM = a.H * M;
    const uint32_t a_tile = 128u * 128u, b_tile = 256u * 128u, stage_bytes = a_tile + b_tile;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_raw + a.off_bar);
    uint64_t* empty_bar = full_bar + S;
    uint64_t* accum_bar = empty_bar + S;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

    const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
    const int hm0 = blockIdx.x * 256, n0 = blockIdx.y * 128;
    const int64_t b_beg = (int64_t)blockIdx.z * a.b_per_cta;
    const int64_t b_end = (b_beg + a.b_per_cta < a.B) ? b_beg + a.b_per_cta : a.B;
    const int nst = (int)(b_end - b_beg);                       // stages = samples of this CTA

    if (tid == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(&full_bar[s], V2_GEN_WARPS);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(accum_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (wid == 1) tmem_alloc_warp(tmem_slot, 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (wid == 1) {
        const uint32_t idesc = tf32_idesc(256);
        const uint32_t a_lbo = 128u * 16u, b_lbo = 256u * 16u;
        int s = 0;
        uint32_t ph = 0;
        for (int i = 0; i < nst; ++i) {
            mbar_wait(&full_bar[s], ph);
            tc_fence_after();
            if (lane == 0) {
                const uint32_t base = smem_u32(smem_raw + (size_t)s * stage_bytes);
                const uint32_t a_hi = base, a_lo = base + a_tile / 2, b_hi = base + a_tile, b_lo = b_hi + b_tile / 2;
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const uint64_t dah = make_smem_desc(a_hi + (uint32_t)j * 2u * a_lbo, a_lbo, 128);
                    const uint64_t dal = make_smem_desc(a_lo + (uint32_t)j * 2u * a_lbo, a_lbo, 128);
                    const uint64_t dbh = make_smem_desc(b_hi + (uint32_t)j * 2u * b_lbo, b_lbo, 128);
                    const uint64_t dbl = make_smem_desc(b_lo + (uint32_t)j * 2u * b_lbo, b_lbo, 128);
                    umma_tf32(tmem_base, dal, dbh, idesc, (i | j) != 0 ? 1u : 0u);
                    umma_tf32(tmem_base, dah, dbl, idesc, 1u);
                    umma_tf32(tmem_base, dah, dbh, idesc, 1u);
                }
                umma_commit(&empty_bar[s]);
                if (i == nst - 1) umma_commit(accum_bar);
            }
            __syncwarp();
            if (++s == S) { s = 0; ph ^= 1u; }
        }
    } else if (wid >= 2) {
        const int ct = tid - 64;
        // A piece of this thread: channel row ra, chunk ca;  B pieces: rows rb (q = 0, 1), chunk cb
        const int ra = ct & 127, ca = ct >> 7;
        const bool a_ok = n0 + ra < N;
        const int64_t a_off = (int64_t)(a_ok ? n0 + ra : 0) * 16 + 4 * ca;
        const int a_toff = (ca * 128 + ra) * 4;
        int64_t xp_off[2], x0_off[2];
        int b_toff[2];
        bool b_ok[2];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int idx = ct + V2_GEN_WARPS * 32 * q;
            const int rb = idx & 255, cb = idx >> 8;
            const int hm = hm0 + rb;
            b_ok[q] = hm < HM;
            const int h = b_ok[q] ? hm / M : 0, m = b_ok[q] ? hm - h * M : 0;
            xp_off[q] = (int64_t)h * 16 + 4 * cb;
            x0_off[q] = (int64_t)m * 16 + 4 * cb;
            b_toff[q] = (cb * 256 + rb) * 4;
        }
        const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
        float4 av, pv[2], xv[2];
        auto load = [&](int64_t b) {
            const bool in = b < b_end;
            const int64_t bc = in ? b : b_beg;
            av = (in && a_ok) ? __ldg(reinterpret_cast<const float4*>(a.dZ + bc * N * 16 + a_off)) : z4;
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                pv[q] = (in && b_ok[q]) ? __ldg(reinterpret_cast<const float4*>(a.Xp + bc * a.sxp + xp_off[q])) : z4;
                xv[q] = (in && b_ok[q]) ? __ldg(reinterpret_cast<const float4*>(a.X0 + bc * a.sx0 + x0_off[q])) : z4;
            }
        };
        load(b_beg);
        int s = 0;
        uint32_t ph = 1;
        for (int i = 0; i < nst; ++i) {
            const float4 a_cur = av;
            const float4 p0 = make_float4(pv[0].x * xv[0].x, pv[0].y * xv[0].y, pv[0].z * xv[0].z, pv[0].w * xv[0].w);
            const float4 p1 = make_float4(pv[1].x * xv[1].x, pv[1].y * xv[1].y, pv[1].z * xv[1].z, pv[1].w * xv[1].w);
            load(b_beg + i + 1);                                  // next sample's loads fly during the stores below
            mbar_wait(&empty_bar[s], ph);
            float* st = reinterpret_cast<float*>(smem_raw + (size_t)s * stage_bytes);
            v2_split_store(st, a_toff, 128 * 16, a_cur);
            v2_split_store(st + a_tile / 4, b_toff[0], 256 * 16, p0);
            v2_split_store(st + a_tile / 4, b_toff[1], 256 * 16, p1);
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&full_bar[s]);
            if (++s == S) { s = 0; ph ^= 1u; }
        }
        // epilogue: accumulate the 128 x 256 tile into dW (K is split across CTAs)
        mbar_wait(accum_bar, 0);
        tc_fence_after();
        const int ew = wid - 2, quad = wid & 3, cg = ew >> 2;
        float* stg = reinterpret_cast<float*>(smem_raw) + (size_t)ew * (32 * V2_STG_PITCH);
        const bool vec = (HM % 4 == 0) && ((reinterpret_cast<uintptr_t>(a.dW) & 15) == 0);
        for (int ci = cg; ci < 8; ci += V2_GEN_WARPS / 4) {
            const int c0 = ci * 32;
            if (hm0 + c0 >= HM) break;                            // warp-uniform
            uint32_t raw[32];
            v2_tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)c0, raw);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 8; ++j)
                *reinterpret_cast<uint4*>(stg + lane * V2_STG_PITCH + 4 * j) =
                    make_uint4(raw[4 * j], raw[4 * j + 1], raw[4 * j + 2], raw[4 * j + 3]);
            __syncwarp();
            const int colq = lane & 7;
            const int hm = hm0 + c0 + 4 * colq;
#pragma unroll 1
            for (int i2 = 0; i2 < 8; ++i2) {
                const int r = 4 * i2 + (lane >> 3);
                const int n = n0 + quad * 32 + r;
                const float4 v = *reinterpret_cast<const float4*>(stg + r * V2_STG_PITCH + 4 * colq);
                if (n < N && hm < HM) {
                    float* dst = a.dW + (size_t)n * HM + hm;
                    if (vec && hm + 3 < HM) {
                        red_add4(dst, v);
                    } else {
                        atomicAdd(dst, v.x);
                        if (hm + 1 < HM) atomicAdd(dst + 1, v.y);
                        if (hm + 2 < HM) atomicAdd(dst + 2, v.z);
                        if (hm + 3 < HM) atomicAdd(dst + 3, v.w);
                    }
                }
            }
            __syncwarp();
        }
        tc_fence_before();
    }
    __syncthreads();
    if (wid == 1) {
        tc_fence_after();
        tmem_dealloc_warp(tmem_base, 256);
    }
}

}  // namespace

// returns 1 if the kernel was launched, 0 if the shape / environment is not supported (caller falls back
// to cin_tc_fwd_kernel), any other value on error
int launch_cin_v2_fwd(const float* Xp, int64_t sxp, int H, const float* X0, int64_t sx0, int M, int D, const float* W,
                      const float* bias, int N, int direct_start, int act, float* Y, float* out, int64_t ld_out, int64_t B,
                      cudaStream_t st) {
    const char* e = getenv("CTR_CIN_V2");      // default since round 2 (validated on hardware); CTR_CIN_V2=0: round-1 kernels
    if (e && e[0] == '0') return 0;
    if (D < 4 || D > 32 || (32 % D) != 0 || H < 1 || M < 1 || N < 1) return 0;
    CinV2Fwd a{};
    a.Xp = Xp; a.sxp = sxp; a.H = H; a.X0 = X0; a.sx0 = sx0; a.M = M; a.D = D;
    a.bias = bias; a.N = N; a.direct_start = direct_start; a.act = act;
    a.Y = Y; a.out = out; a.ld_out = ld_out; a.B = B;
    const int64_t ntiles = ceil_div64(N, 256);
    a.BN = (int)(ceil_div64(ceil_div64(N, ntiles), 16) * 16);
    a.TB = V2_ROWS / D;
    // the epilogue reads whole 32-column chunks: the last chunk of the second accumulator must stay inside
    // the allocation
    a.chains = (4 * a.BN <= 512 && a.BN % 32 == 0) ? 2 : 1;
    const int n_acc = 2 * a.chains;
    const int tm_need = (n_acc * a.BN > (n_acc - 1) * a.BN + 32 * ((a.BN + 31) / 32)) ? n_acc * a.BN
                                                                                      : (n_acc - 1) * a.BN + 32 * ((a.BN + 31) / 32);
    a.tmem_cols = 32;
    while (a.tmem_cols < tm_need) a.tmem_cols <<= 1;
    if (a.tmem_cols > 512) return 0;
    const int HM = H * M;
    const int64_t nkb = ceil_div64(HM, V2_KB), n_rb = ceil_div64(N, a.BN);
    const int64_t need = n_rb * nkb * (int64_t)a.BN * 128;
    float* wp = reinterpret_cast<float*>(gemm_scratch_ptr(need));
    if (!wp) return 0;
    const int64_t a_ring = 2 * (int64_t)V2_ROWS * 128, b_stage = (int64_t)a.BN * 128;
    const int64_t xbytes = (int64_t)(H + M) * V2_ROWS * 4;
    for (a.SB = 4; a.SB >= 2; --a.SB)
        if (a_ring + a.SB * b_stage + xbytes + 256 <= 232448 - 512) break;
    if (a.SB < 2) return 0;
    a.off_b = (uint32_t)a_ring;
    a.off_xp = (uint32_t)(a_ring + a.SB * b_stage);
    a.off_x0 = a.off_xp + (uint32_t)H * V2_ROWS * 4;
    a.off_bar = a.off_x0 + (uint32_t)M * V2_ROWS * 4;
    const size_t smem = a.off_bar + (2 * 2 + 2 * a.SB + 1) * sizeof(uint64_t) + 16;
    int rc = gemm_pack_operand(W, HM, 1, N, HM, a.BN, nkb, wp, st);
    if (rc) return rc;
    a.Wp = wp;
    a.wp_nkb = (int)nkb;
    static bool configured = false;
    if (!configured) {
        cudaError_t ce = cudaFuncSetAttribute(cin_v2_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
        if (ce != cudaSuccess) {
            ctr_set_error("cin_v2_fwd: cudaFuncSetAttribute failed: %s", cudaGetErrorString(ce));
            return (int)ce + 1000;
        }
        configured = true;
    }
    dim3 grid((unsigned)ceil_div64(B, a.TB), (unsigned)n_rb);
    cin_v2_fwd_kernel<<<grid, V2_THREADS, smem, st>>>(a);
    ctr_count_launch();
    cudaError_t le = cudaGetLastError();
    if (le != cudaSuccess) {
        ctr_set_error("launch of cin_v2_fwd_kernel failed: %s", cudaGetErrorString(le));
        return (int)le + 1000;
    }
    return 1;
}

// dW of one CIN layer (dW must have been zeroed).  Returns 1 if launched, 0 if not supported / not enabled.
int launch_cin_v2_dw(const float* Xp, int64_t sxp, int H, const float* X0, int64_t sx0, int M, int D, int N, const float* dZ,
                     float* dW, int64_t B, cudaStream_t st) {
    const char* e = getenv("CTR_CIN_V2");      // default since round 2 (validated on hardware); CTR_CIN_V2=0: round-1 kernels
    if (e && e[0] == '0') return 0;
    if (D != 16 || B < 1) return 0;
    if ((sxp % 4) != 0 || (sx0 % 4) != 0 || (reinterpret_cast<uintptr_t>(Xp) & 15) || (reinterpret_cast<uintptr_t>(X0) & 15) ||
        (reinterpret_cast<uintptr_t>(dZ) & 15))
        return 0;
    const int HM = H * M;
    CinV2Dw a{Xp, sxp, H, X0, sx0, M, N, dZ, dW, B, 0, 4, 0};
    const int64_t tiles = ceil_div64(HM, 256) * ceil_div64(N, 128);
    int64_t splits = ceil_div64((int64_t)ctr_sm_count(), tiles);
    if (splits > B) splits = B;
    if (splits < 1) splits = 1;
    a.b_per_cta = ceil_div64(B, splits);
    splits = ceil_div64(B, a.b_per_cta);
    if (splits > 65535) return 0;
    const size_t stage_bytes = 128 * 128 + 256 * 128;
    a.off_bar = (uint32_t)(a.S * stage_bytes);
    const size_t smem = a.off_bar + (2 * a.S + 1) * sizeof(uint64_t) + 16;
    static bool configured = false;
    if (!configured) {
        cudaError_t ce = cudaFuncSetAttribute(cin_v2_dw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
        if (ce != cudaSuccess) {
            ctr_set_error("cin_v2_dw: cudaFuncSetAttribute failed: %s", cudaGetErrorString(ce));
            return (int)ce + 1000;
        }
        configured = true;
    }
    dim3 grid((unsigned)ceil_div64(HM, 256), (unsigned)ceil_div64(N, 128), (unsigned)splits);
    cin_v2_dw_kernel<<<grid, V2_THREADS, smem, st>>>(a);
    ctr_count_launch();
    cudaError_t le = cudaGetLastError();
    if (le != cudaSuccess) {
        ctr_set_error("launch of cin_v2_dw_kernel failed: %s", cudaGetErrorString(le));
        return (int)le + 1000;
    }
    return 1;
}
