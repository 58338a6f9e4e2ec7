// tcgen05 GEMM over PRE-SPLIT, PRE-TILED operands (3xTF32, fp32-grade accuracy) — the tensor-core
// engine of the dense tower (reference layers/core.py:120-134 and every other GEMM-shaped op).
//
//   C[m,n] (+)= sum_k A(m,k) * B(n,k)
//
// Two phases per GEMM, both in this file:
//
//  1. pack_*_kernel  (HBM-bound, elementwise): reads an fp32 operand P(row,k) with ANY strides
//     (row-major, transposed, strided views; optional act'(mask) prologue of the backward), splits
//     every value into two TF32-exact parts hi = RN_tf32(x), lo = RN_tf32(x - hi) and writes them as
//     the exact shared-memory image the tensor core wants: tiles of R rows x 16 k, each tile =
//     [hi | lo] x [4 chunks of 16 B] x [R rows] (K-major SWIZZLE_NONE core matrices: 8 rows x 16 B
//     = 128 contiguous bytes).  Out-of-range rows / k are written as zeros, so the GEMM has no
//     bounds logic on its operand path.
//
//  2. gemm_pk_kernel (tensor-pipe-bound): warp-specialised, no operand ALU work at all:
//       warp 0   one lane: cp.async.bulk (TMA, 1-D) of whole tiles global -> shared, completion
//                on an mbarrier (expect_tx), S-stage ring
//       warp 1   TMEM alloc; one lane issues tcgen05.mma.kind::tf32: per 8-wide K atom and per
//                accumulator lo*hi + hi*lo + hi*hi into fp32 TMEM; tcgen05.commit frees the stage
//       warps 2-5 epilogue: tcgen05.ld -> bias / activation / act' mask / CrossNet -> global
//                (plain 128-bit stores, or fp32 reductions when K is split across CTAs)
//     CTA tile = (MT x 128) x BN: MT = 2 keeps two accumulators in TMEM (2 x BN <= 512 columns)
//     against ONE B stage, halving the L2->SM bytes per flop of the weight operand.
//
// Why pre-split: 3xTF32 needs hi AND lo of both operands in shared memory (16 B per fp32 pair).
// Splitting inside the GEMM (gemm_tc.cu, the first engine) made 16 producer warps the bottleneck
// (LDG -> 5 ALU ops per value -> 2 STS, 96 live registers, spills) and every CTA re-split the same
// weights; measured 34 TFLOP/s effective on the DeepFM tower.  Here the split costs one streaming
// pass per operand and the GEMM runs at the tensor pipe's pace.
#include <stdlib.h>

#include "gemm.cuh"
#include "tc_common.cuh"

namespace {

constexpr int PK_KB = 16;            // k per packed tile (4 chunks of 16 bytes)
constexpr int PK_AR = 128;           // rows of an A tile (one UMMA M)
constexpr int PK_THREADS = 192;      // producer warp, MMA warp, 4 epilogue warps

// ---------------------------------------------------------------------------------------------
// phase 1: pack
// ---------------------------------------------------------------------------------------------
struct PackArgs {
    const float* P; int64_t s_row, s_k;            // P(row,k) = P[row*s_row + k*s_k]
    const float* mask; int64_t m_row, m_k; int mask_act;
    int64_t n_rows, K;                             // logical extents
    int R;                                         // rows per tile (multiple of 16)
    int64_t n_rb;                                  // row blocks emitted (extra blocks are zeros)
    int64_t nkb;                                   // ceil(K / 16)
    float* out;                                    // n_rb * nkb tiles of 32*R floats
};

__device__ __forceinline__ int64_t pk_tile_base(const PackArgs& a, int64_t rb, int64_t kb) {
    return (rb * a.nkb + kb) * (int64_t)(32 * a.R);
}

__device__ __forceinline__ void pk_store(float* out, int64_t base, int R, int c, int r, float4 v) {
    float4 hi, lo;
    split_tf32(v.x, hi.x, lo.x);
    split_tf32(v.y, hi.y, lo.y);
    split_tf32(v.z, hi.z, lo.z);
    split_tf32(v.w, hi.w, lo.w);
    float* ph = out + base + ((int64_t)c * R + r) * 4;
    *reinterpret_cast<float4*>(ph) = hi;
    *reinterpret_cast<float4*>(ph + 16 * R) = lo;
}

__device__ __forceinline__ float4 pk_mask4(float4 x, float4 y, int act) {
    x.x *= act_grad_from_y(act, y.x);
    x.y *= act_grad_from_y(act, y.y);
    x.z *= act_grad_from_y(act, y.z);
    x.w *= act_grad_from_y(act, y.w);
    return x;
}

// K-contiguous operand (s_k == 1): one thread = (row, 16-k block): 64 contiguous bytes in, 8 x 16
// bytes out; lanes map to consecutive rows so every store instruction writes 512 contiguous bytes.
__global__ void __launch_bounds__(256) pack_kvec_kernel(PackArgs a) {
    const int64_t rows_pad = a.n_rb * a.R;
    const int64_t total = rows_pad * a.nkb;
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t kb = i / rows_pad, row = i - kb * rows_pad;
        const int64_t rb = row / a.R;
        const int r = (int)(row - rb * a.R);
        const int64_t k0 = kb * PK_KB;
        float4 v[4];
        if (row < a.n_rows) {
            const float* src = a.P + row * a.s_row + k0;
            const float* msk = a.mask ? a.mask + row * a.m_row + k0 : nullptr;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int64_t k = k0 + c * 4;
                if (k < a.K) {      // whole 16-byte groups are readable (host checks ld >= round4(K))
                    v[c] = __ldg(reinterpret_cast<const float4*>(src + c * 4));
                    if (msk) v[c] = pk_mask4(v[c], __ldg(reinterpret_cast<const float4*>(msk + c * 4)), a.mask_act);
                    if (k + 1 >= a.K) v[c].y = 0.f;
                    if (k + 2 >= a.K) v[c].z = 0.f;
                    if (k + 3 >= a.K) v[c].w = 0.f;
                } else {
                    v[c] = z;
                }
            }
        } else {
            v[0] = v[1] = v[2] = v[3] = z;
        }
        const int64_t base = pk_tile_base(a, rb, kb);
#pragma unroll
        for (int c = 0; c < 4; ++c) pk_store(a.out, base, a.R, c, r, v[c]);
    }
}

// Row-contiguous operand (s_row == 1; the transposed views of the weight-gradient GEMMs): one
// thread = (4 rows, 4 k): four 128-bit loads along the rows, 4x4 register transpose, 4 x (hi, lo)
// stores of 16 bytes to consecutive rows (64 contiguous bytes per thread and part).
__global__ void __launch_bounds__(256) pack_trans_kernel(PackArgs a) {
    const int64_t rq_pad = a.n_rb * a.R / 4;
    const int64_t total = rq_pad * a.nkb * 4;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t cg = i / rq_pad, rq = i - cg * rq_pad;      // global chunk index, row quad
        const int64_t row = rq * 4;
        const int64_t kb = cg >> 2;
        const int c = (int)(cg & 3);
        const int64_t k0 = cg * 4;
        const int64_t rb = row / a.R;
        const int r = (int)(row - rb * a.R);
        float4 x[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int64_t k = k0 + e;
            if (row < a.n_rows && k < a.K) {
                x[e] = __ldg(reinterpret_cast<const float4*>(a.P + row + k * a.s_k));
                if (a.mask)
                    x[e] = pk_mask4(x[e], __ldg(reinterpret_cast<const float4*>(a.mask + row + k * a.m_k)), a.mask_act);
                if (row + 1 >= a.n_rows) x[e].y = 0.f;
                if (row + 2 >= a.n_rows) x[e].z = 0.f;
                if (row + 3 >= a.n_rows) x[e].w = 0.f;
            } else {
                x[e] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        const int64_t base = pk_tile_base(a, rb, kb);
        pk_store(a.out, base, a.R, c, r + 0, make_float4(x[0].x, x[1].x, x[2].x, x[3].x));
        pk_store(a.out, base, a.R, c, r + 1, make_float4(x[0].y, x[1].y, x[2].y, x[3].y));
        pk_store(a.out, base, a.R, c, r + 2, make_float4(x[0].z, x[1].z, x[2].z, x[3].z));
        pk_store(a.out, base, a.R, c, r + 3, make_float4(x[0].w, x[1].w, x[2].w, x[3].w));
    }
}

// any strides / alignment: one thread = (row, chunk), scalar loads
__global__ void __launch_bounds__(256) pack_generic_kernel(PackArgs a) {
    const int64_t rows_pad = a.n_rb * a.R;
    const int64_t total = rows_pad * a.nkb * 4;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t cg = i / rows_pad, row = i - cg * rows_pad;
        const int64_t kb = cg >> 2;
        const int c = (int)(cg & 3);
        const int64_t rb = row / a.R;
        const int r = (int)(row - rb * a.R);
        float v[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int64_t k = cg * 4 + e;
            float x = 0.f;
            if (row < a.n_rows && k < a.K) {
                x = __ldg(a.P + row * a.s_row + k * a.s_k);
                if (a.mask) x *= act_grad_from_y(a.mask_act, __ldg(a.mask + row * a.m_row + k * a.m_k));
            }
            v[e] = x;
        }
        pk_store(a.out, pk_tile_base(a, rb, kb), a.R, c, r, make_float4(v[0], v[1], v[2], v[3]));
    }
}

bool pk_al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

int launch_pack(const PackArgs& a, cudaStream_t st) {
    const int64_t K4 = (a.K + 3) / 4 * 4, R4 = (a.n_rows + 3) / 4 * 4;
    const bool kvec = a.s_k == 1 && a.s_row % 4 == 0 && a.s_row >= K4 && pk_al16(a.P) &&
                      (!a.mask || (a.m_k == 1 && a.m_row % 4 == 0 && a.m_row >= K4 && pk_al16(a.mask)));
    const bool trans = a.s_row == 1 && a.s_k % 4 == 0 && a.s_k >= R4 && pk_al16(a.P) &&
                       (!a.mask || (a.m_row == 1 && a.m_k % 4 == 0 && a.m_k >= R4 && pk_al16(a.mask)));
    const int64_t cap = (int64_t)ctr_sm_count() * 8;
    if (kvec) {
        int64_t blocks = ceil_div64(a.n_rb * a.R * a.nkb, 256);
        if (blocks > cap) blocks = cap;
        pack_kvec_kernel<<<(unsigned)blocks, 256, 0, st>>>(a);
        CTR_LAUNCH_OK("pack_kvec_kernel");
    } else if (trans) {
        int64_t blocks = ceil_div64(a.n_rb * a.R * a.nkb, 256);
        if (blocks > cap) blocks = cap;
        pack_trans_kernel<<<(unsigned)blocks, 256, 0, st>>>(a);
        CTR_LAUNCH_OK("pack_trans_kernel");
    } else {
        int64_t blocks = ceil_div64(a.n_rb * a.R * a.nkb * 4, 256);
        if (blocks > cap) blocks = cap;
        pack_generic_kernel<<<(unsigned)blocks, 256, 0, st>>>(a);
        CTR_LAUNCH_OK("pack_generic_kernel");
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------
// phase 2: GEMM over packed tiles
// ---------------------------------------------------------------------------------------------
struct PkParams {
    GemmArgs g;                  // M, N, epilogue fields (A/B pointers unused here)
    const float* Ap;             // packed A: row blocks of 128
    const float* Bp;             // packed B: row blocks of BN
    int64_t nkb;                 // K blocks of 16 in the packed operands
    int64_t kb_per_split;
    int BN, MT, stages, tmem_cols;
};

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

__global__ void __launch_bounds__(PK_THREADS, 1) gemm_pk_kernel(PkParams p) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const GemmArgs& g = p.g;
    const int BN = p.BN, MT = p.MT, S = p.stages;
    const uint32_t a_tile = PK_AR * 128u;                 // hi+lo of a 128 x 16 tile
    const uint32_t b_tile = (uint32_t)BN * 128u;
    const uint32_t stage_bytes = (uint32_t)MT * a_tile + b_tile;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_raw + (size_t)S * stage_bytes);
    uint64_t* empty_bar = full_bar + S;
    uint64_t* accum_bar = empty_bar + S;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

    const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
    const int64_t mblk = blockIdx.y, nblk = blockIdx.x;
    const int64_t kb_beg = (int64_t)blockIdx.z * p.kb_per_split;
    const int64_t kb_end = (kb_beg + p.kb_per_split < p.nkb) ? kb_beg + p.kb_per_split : p.nkb;
    const int nkb = (int)(kb_end - kb_beg);

    if (tid == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(accum_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (wid == 1) tmem_alloc_warp(tmem_slot, (uint32_t)p.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (wid == 0) {
        // ------------------------------ TMA producer (one lane) ----------------------------
        if (lane == 0) {
            const unsigned char* a_src = reinterpret_cast<const unsigned char*>(p.Ap);
            const unsigned char* b_src = reinterpret_cast<const unsigned char*>(p.Bp);
            for (int i = 0; i < nkb; ++i) {
                const int s = i % S;
                mbar_wait(&empty_bar[s], ((uint32_t)(i / S) & 1u) ^ 1u);
                unsigned char* st = smem_raw + (size_t)s * stage_bytes;
                mbar_expect_tx(&full_bar[s], stage_bytes);
                const int64_t kb = kb_beg + i;
                for (int mt = 0; mt < MT; ++mt)
                    bulk_g2s(st + (size_t)mt * a_tile, a_src + ((mblk * MT + mt) * p.nkb + kb) * (int64_t)a_tile, a_tile,
                             &full_bar[s]);
                bulk_g2s(st + (size_t)MT * a_tile, b_src + (nblk * p.nkb + kb) * (int64_t)b_tile, b_tile, &full_bar[s]);
            }
        }
        __syncwarp();
    } else if (wid == 1) {
        // ------------------------------ MMA issuer ------------------------------------------
        const uint32_t idesc = tf32_idesc(BN);
        const uint32_t a_lbo = PK_AR * 16u, b_lbo = (uint32_t)BN * 16u;     // between the two chunks of a K atom
        for (int i = 0; i < nkb; ++i) {
            const int s = i % S;
            mbar_wait(&full_bar[s], (uint32_t)(i / S) & 1u);
            tc_fence_after();
            if (lane == 0) {
                const uint32_t base = smem_u32(smem_raw + (size_t)s * stage_bytes);
                const uint32_t b_hi = base + (uint32_t)MT * a_tile, b_lo = b_hi + b_tile / 2;
#pragma unroll
                for (int j = 0; j < PK_KB / 8; ++j) {
                    const uint64_t dbh = make_smem_desc(b_hi + (uint32_t)j * 2u * b_lbo, b_lbo, 128);
                    const uint64_t dbl = make_smem_desc(b_lo + (uint32_t)j * 2u * b_lbo, b_lbo, 128);
                    for (int mt = 0; mt < MT; ++mt) {
                        const uint32_t a_hi = base + (uint32_t)mt * a_tile + (uint32_t)j * 2u * a_lbo;
                        const uint64_t dah = make_smem_desc(a_hi, a_lbo, 128);
                        const uint64_t dal = make_smem_desc(a_hi + a_tile / 2, a_lbo, 128);
                        const uint32_t d = tmem_base + (uint32_t)(mt * BN);
                        umma_tf32(d, dal, dbh, idesc, (i | j) != 0 ? 1u : 0u);      // small terms first
                        umma_tf32(d, dah, dbl, idesc, 1u);
                        umma_tf32(d, dah, dbh, idesc, 1u);
                    }
                }
                umma_commit(&empty_bar[s]);               // frees the stage when these MMAs retire
                if (i == nkb - 1) umma_commit(accum_bar);
            }
            __syncwarp();
        }
    } else {
        // ------------------------------ epilogue (4 warps) ----------------------------------
        mbar_wait(accum_bar, 0);
        tc_fence_after();
        const int quad = wid & 3;                         // TMEM lane quadrant this warp may read
        const bool split = gridDim.z > 1;
        const bool vec_store = !split && !g.accumulate && (g.ldc % 4 == 0) &&
                               ((reinterpret_cast<uintptr_t>(g.C) & 15) == 0) && g.epilogue != EPI_CROSS;
        const int64_t n0 = nblk * BN;
        for (int mt = 0; mt < MT; ++mt) {
            const int64_t m = (mblk * MT + mt) * PK_AR + quad * 32 + lane;
            if ((mblk * MT + mt) * PK_AR >= g.M) break;   // whole tile out of range (warp-uniform)
            for (int c0 = 0; c0 < BN; c0 += 16) {
                if (n0 + c0 >= g.N) break;                // warp-uniform
                uint32_t raw[16];
                tmem_ld16(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(mt * BN + c0), raw);
                tmem_ld_wait();
                if (m < g.M) {
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
        }
        tc_fence_before();
    }
    __syncthreads();
    if (wid == 1) {
        tc_fence_after();
        tmem_dealloc_warp(tmem_base, (uint32_t)p.tmem_cols);
    }
}

// ---------------------------------------------------------------------------------------------
// configuration shared by the launcher and the scratch-size query
// ---------------------------------------------------------------------------------------------
struct PkConfig {
    int BN, MT, stages, tmem_cols;
    int64_t gm, gn, splits, nkb, kb_per_split;
    int64_t a_bytes, b_bytes;    // packed operand sizes
};

PkConfig pk_config(int64_t M, int64_t N, int64_t K, bool allow_split) {
    PkConfig c;
    const int64_t ntiles = ceil_div64(N, 256);
    c.BN = (int)(ceil_div64(ceil_div64(N, ntiles), 16) * 16);
    if (c.BN < 16) c.BN = 16;
    c.gn = ceil_div64(N, c.BN);
    c.nkb = ceil_div64(K, PK_KB);
    const int64_t sms = ctr_sm_count();
    // two accumulators per CTA when there is enough work to fill the machine anyway
    c.MT = (ceil_div64(M, 2 * PK_AR) * c.gn >= sms || (allow_split && M > PK_AR)) ? 2 : 1;
    if (M <= PK_AR) c.MT = 1;
    c.gm = ceil_div64(M, (int64_t)PK_AR * c.MT);
    c.splits = 1;
    if (allow_split) {
        const int64_t tiles = c.gm * c.gn;
        if (tiles < sms && c.nkb >= 16) {
            c.splits = ceil_div64(sms, tiles);
            const int64_t max_splits = c.nkb / 8;
            if (c.splits > max_splits) c.splits = max_splits;
            if (c.splits < 1) c.splits = 1;
        }
    }
    c.kb_per_split = ceil_div64(c.nkb, c.splits);
    c.splits = ceil_div64(c.nkb, c.kb_per_split);
    const int64_t stage_bytes = (int64_t)c.MT * PK_AR * 128 + (int64_t)c.BN * 128;
    c.stages = (int)((200 * 1024) / stage_bytes);
    if (c.stages > 6) c.stages = 6;
    c.tmem_cols = 32;
    while (c.tmem_cols < c.MT * c.BN) c.tmem_cols <<= 1;
    c.a_bytes = c.gm * c.MT * c.nkb * (int64_t)PK_AR * 128;
    c.b_bytes = c.gn * c.nkb * (int64_t)c.BN * 128;
    return c;
}

// scratch registered per device (autograd runs the backward on its own host thread, so this cannot
// be thread-local); written by ctr_set_scratch before any launch that uses it
constexpr int kMaxDevices = 64;
void* g_scratch[kMaxDevices] = {};
int64_t g_scratch_bytes[kMaxDevices] = {};

int current_device() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) dev = 0;
    return dev;
}

}  // namespace

int64_t gemm_pk_scratch_bytes(int64_t M, int64_t N, int64_t K) {
    if (M <= 0 || N <= 0 || K <= 0) return 0;
    const PkConfig a = pk_config(M, N, K, false), b = pk_config(M, N, K, true);
    const int64_t x = a.a_bytes + a.b_bytes, y = b.a_bytes + b.b_bytes;
    return (x > y ? x : y) + 256;
}

bool gemm_pk_has_scratch(int64_t M, int64_t N, int64_t K) {
    const int dev = current_device();
    return g_scratch[dev] && g_scratch_bytes[dev] >= gemm_pk_scratch_bytes(M, N, K);
}

extern "C" int64_t ctr_gemm_scratch_bytes(int64_t M, int64_t N, int64_t K) { return gemm_pk_scratch_bytes(M, N, K); }

extern "C" int ctr_set_scratch(void* ptr, int64_t bytes) {
    CTR_ARG(bytes >= 0 && (ptr || bytes == 0), "ctr_set_scratch: bad arguments");
    CTR_ARG((reinterpret_cast<uintptr_t>(ptr) & 127) == 0, "ctr_set_scratch: pointer must be 128-byte aligned");
    const int dev = current_device();
    g_scratch[dev] = ptr;
    g_scratch_bytes[dev] = bytes;
    return 0;
}

int launch_gemm_pk(const GemmArgs& g, cudaStream_t st) {
    const bool allow_split = g.allow_split_k && g.epilogue == EPI_STORE;
    const PkConfig c = pk_config(g.M, g.N, g.K, allow_split);
    if (c.stages < 2) {
        ctr_set_error("launch_gemm_pk: tile does not fit shared memory");
        return -1;
    }
    const int64_t need = c.a_bytes + c.b_bytes;
    const int dev = current_device();
    void* scratch = g_scratch[dev];
    if (!scratch || g_scratch_bytes[dev] < need) {
        ctr_set_error("launch_gemm_pk: scratch of %lld bytes required for M=%lld N=%lld K=%lld, %lld registered "
                      "(ctr_set_scratch / ctr_gemm_scratch_bytes)",
                      (long long)need, (long long)g.M, (long long)g.N, (long long)g.K, (long long)g_scratch_bytes[dev]);
        return -2;
    }
    if (c.gm > 65535 || c.splits > 65535) {
        ctr_set_error("launch_gemm_pk: grid too large");
        return -1;
    }
    float* Ap = reinterpret_cast<float*>(scratch);
    float* Bp = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(scratch) + c.a_bytes);
    int rc;
    PackArgs pa{g.A, g.sam, g.sak, g.amask, g.smm, g.smk, g.amask_act, g.M, g.K, PK_AR, c.gm * c.MT, c.nkb, Ap};
    if ((rc = launch_pack(pa, st)) != 0) return rc;
    PackArgs pb{g.B, g.sbn, g.sbk, g.bmask, g.sbmn, g.sbmk, g.bmask_act, g.N, g.K, c.BN, c.gn, c.nkb, Bp};
    if ((rc = launch_pack(pb, st)) != 0) return rc;
    if (c.splits > 1 && !g.accumulate)
        CTR_CUDA(cudaMemset2DAsync(g.C, g.ldc * sizeof(float), 0, g.N * sizeof(float), g.M, st));
    PkParams p;
    p.g = g;
    p.Ap = Ap;
    p.Bp = Bp;
    p.nkb = c.nkb;
    p.kb_per_split = c.kb_per_split;
    p.BN = c.BN;
    p.MT = c.MT;
    p.stages = c.stages;
    p.tmem_cols = c.tmem_cols;
    const size_t stage_bytes = (size_t)c.MT * PK_AR * 128 + (size_t)c.BN * 128;
    const size_t smem = (size_t)c.stages * stage_bytes + (2 * c.stages + 1) * sizeof(uint64_t) + 16;
    static bool configured = false;
    if (!configured) {
        CTR_CUDA(cudaFuncSetAttribute(gemm_pk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(220 * 1024)));
        configured = true;
    }
    dim3 grid((unsigned)c.gn, (unsigned)c.gm, (unsigned)c.splits);
    gemm_pk_kernel<<<grid, PK_THREADS, smem, st>>>(p);
    CTR_LAUNCH_OK("gemm_pk_kernel");
    return 0;
}
