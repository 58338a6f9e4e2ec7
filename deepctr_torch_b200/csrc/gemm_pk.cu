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
#include <cuda.h>          // CUtensorMap (the encoder is fetched through cudaGetDriverEntryPoint: no libcuda link)

#include "gemm.cuh"
#include "tc_common.cuh"

namespace {

constexpr int PK_KB = 16;            // k per packed tile (4 chunks of 16 bytes)
constexpr int PK_AR = 128;           // rows of an A tile (one UMMA M)
constexpr int PK_CONV_WARPS = 16;    // converter warps (they also run the epilogue)
constexpr int PK_THREADS = 64 + PK_CONV_WARPS * 32;   // + TMA producer warp + MMA warp
constexpr int PK_STG_PITCH = 36;     // floats per row of an epilogue staging tile (32 + 4: conflict-free)

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
// bytes out.  Lanes map to consecutive rows, so every store instruction writes 512 contiguous bytes;
// a block owns 128 rows x 4 consecutive k blocks and its two halves read neighbouring 64-byte
// pieces of the same rows at the same time, so every 128-byte line of the source is consumed while
// it is hot (sweeping one k block over all rows first re-fetched each line 27 / 2 times apart and
// ran at 3.8 TB/s).
constexpr int PKV_KG = 4;     // k blocks per CTA item
template <bool MASK>
__global__ void __launch_bounds__(256, 2) pack_kvec_kernel(PackArgs a) {
    const int64_t rows_pad = a.n_rb * a.R;
    const int64_t n_rg = (rows_pad + 127) / 128;
    const int64_t n_kg = (a.nkb + PKV_KG - 1) / PKV_KG;
    const int r_l = threadIdx.x & 127, kpar = threadIdx.x >> 7;
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int64_t item = blockIdx.x; item < n_rg * n_kg; item += gridDim.x) {
        const int64_t rg = item / n_kg, kg = item - rg * n_kg;
        const int64_t row = rg * 128 + r_l;
        if (row >= rows_pad) continue;
        const int64_t rb = row / a.R;
        const int r = (int)(row - rb * a.R);
        const bool row_ok = row < a.n_rows;
        const float* src = a.P + (row_ok ? row : 0) * a.s_row;
        const float* msk = MASK ? a.mask + (row_ok ? row : 0) * a.m_row : nullptr;
        float4 v[PKV_KG / 2][4], mv[MASK ? PKV_KG / 2 : 1][4];
#pragma unroll
        for (int j = 0; j < PKV_KG / 2; ++j) {
            const int64_t kb = kg * PKV_KG + kpar + 2 * j;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int64_t k = kb * PK_KB + c * 4;
                const bool ok = row_ok && k < a.K;       // whole 16-byte groups are readable (host: ld >= round4(K))
                v[j][c] = ok ? __ldg(reinterpret_cast<const float4*>(src + k)) : z;
                if (MASK) mv[j][c] = ok ? __ldg(reinterpret_cast<const float4*>(msk + k)) : z;
            }
        }
#pragma unroll
        for (int j = 0; j < PKV_KG / 2; ++j) {
            const int64_t kb = kg * PKV_KG + kpar + 2 * j;
            if (kb >= a.nkb) continue;
            const int64_t base = pk_tile_base(a, rb, kb);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int64_t k = kb * PK_KB + c * 4;
                float4 x = v[j][c];
                if (MASK) x = pk_mask4(x, mv[j][c], a.mask_act);
                if (k + 1 >= a.K) x.y = 0.f;
                if (k + 2 >= a.K) x.z = 0.f;
                if (k + 3 >= a.K) x.w = 0.f;
                pk_store(a.out, base, a.R, c, r, x);
            }
        }
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

// fp32 tiles WITHOUT the split (TS engine, b_raw): same chunk layout, one part only — tile = R rows x 16 k =
// [4 chunks] x [R rows] x 16 bytes.  kind::tf32 ignores the low 13 mantissa bits of what it reads (probe mode 4:
// 16384 / 16384 products match truncation), so the raw tile IS the hi operand; the CTA derives lo itself.
__global__ void __launch_bounds__(256) pack_raw_kernel(PackArgs a) {
    const int64_t rows_pad = a.n_rb * a.R;
    const int64_t total = rows_pad * a.nkb * 4;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t cg = i / rows_pad, row = i - cg * rows_pad;
        const int64_t kb = cg >> 2;
        const int c = (int)(cg & 3);
        const int64_t rb = row / a.R;
        const int r = (int)(row - rb * a.R);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        float* pv = &v.x;
        for (int e = 0; e < 4; ++e) {
            const int64_t k = cg * 4 + e;
            if (row < a.n_rows && k < a.K) pv[e] = __ldg(a.P + row * a.s_row + k * a.s_k);
        }
        *reinterpret_cast<float4*>(a.out + (rb * a.nkb + kb) * (int64_t)(16 * a.R) + ((int64_t)c * a.R + r) * 4) = v;
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
        int64_t blocks = ceil_div64(a.n_rb * a.R, 128) * ceil_div64(a.nkb, PKV_KG);
        if (blocks > cap) blocks = cap;
        if (a.mask) pack_kvec_kernel<true><<<(unsigned)blocks, 256, 0, st>>>(a);
        else pack_kvec_kernel<false><<<(unsigned)blocks, 256, 0, st>>>(a);
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
enum { OP_PACKED = 0, OP_KVEC = 1, OP_TRANS = 2, OP_SCALAR = 3 };

// An operand that is converted INSIDE the GEMM (mode != OP_PACKED): the converter warps fetch fp32
// pieces with cp.async straight into shared memory, split them into hi/lo and write the MMA-ready
// tile — no packed copy of the operand ever exists in HBM.
struct StreamOp {
    const float* P; int64_t s_row, s_k;            // P(row,k) = P[row*s_row + k*s_k]
    const float* mask; int64_t m_row, m_k; int mask_act;
    int64_t n_rows;
    int mode;
};

struct PkParams {
    GemmArgs g;                  // M, N, K, epilogue fields
    const float* Ap;             // packed A (row blocks of 128) when sa.mode == OP_PACKED
    const float* Bp;             // packed B (row blocks of BN)  when sb.mode == OP_PACKED
    StreamOp sa, sb;
    int64_t nkb;                 // K blocks of 16
    int64_t kb_per_split;
    int BN, MT, SA, SB, depth, tmem_cols;
    uint32_t off_b, off_raw, off_bar;     // shared-memory layout (bytes): A ring at 0
    uint32_t raw_a_bytes, raw_b_bytes;    // raw cp.async region of each streamed operand, per depth
    int t0_b;                             // first converter thread of B's OP_TRANS groups
    unsigned long long* dbg;              // optional clock64 timeline of CTA (0,0,0) (ctr_debug_set_buffer): [8 events][64 stages]
    int fast;                             // single-pass TF32 (non-parity fast mode): only the hi*hi MMA is issued
    int a_tma;                            // K-contiguous streamed A: raw [MT*128 x 16] fp32 tiles arrive by 2-D tiled TMA loads
                                          // (64-byte swizzle) instead of the converters' own cp.async pieces
};
#define PK_DBG(ev, idx)                                                                                   \
    do {                                                                                                  \
        if (p.dbg && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && (idx) < 64) p.dbg[(ev) * 64 + (idx)] = clock64(); \
    } while (0)

constexpr int PK_CONV_THREADS = PK_CONV_WARPS * 32;

__device__ __forceinline__ void cp_async16(void* dst, const void* src, int bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async4(void* dst, const void* src, int bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// address (floats) of chunk c of row r inside an MMA tile made of sub-tiles of `sub` rows
// ([hi | lo] x [4 chunks] x [sub rows] x 16 bytes per sub-tile)
__device__ __forceinline__ int tile_float_off(int sub, int r, int c) {
    const int st = r / sub, rr = r - st * sub;
    return st * (sub * 32) + (c * sub + rr) * 4;
}
__device__ __forceinline__ void pin_op(StreamOp& o) {
    // opaque to the optimiser: the values now live in registers, not in the parameter bank
    asm volatile("" : "+l"(o.P), "+l"(o.mask), "+l"(o.s_k), "+l"(o.m_k), "+r"(o.mode), "+r"(o.mask_act));
}

// ---- converter state.  512 converter threads share one 16-k stage of an operand:
//   OP_KVEC / OP_SCALAR : 4*R pieces of 16 bytes (row r, chunk c) -> up to 2 per thread
//                         (idx = ct + 512 q: c = idx / R, r = idx % R: lanes = consecutive rows)
//   OP_TRANS            : R groups (4 consecutive rows x 4 consecutive k) -> one group for each of
//                         the 256 threads starting at thread t0 (A: 0, B: 256 when both operands are
//                         transposed streams, so the two operands use disjoint halves of the warps)
// Everything that does not change from one stage to the next — source pointers, destination
// offsets inside the MMA tile, row validity — is computed ONCE; per stage a thread only bumps its
// pointers (the first version recomputed idx / R, 64-bit addresses and bounds for every piece of
// every stage: instruction-bound at 32 % issue utilisation, tensor pipe 28 %).
struct PieceSet {
    const float* src[2];     // piece source at the CTA's first k block (TRANS: src[0] = first k of the group)
    const float* msk[2];
    int toff[2];             // float offset of the destination chunk inside the MMA tile, < 0: nothing to store
    int koff[2];             // k offset of the piece inside a stage
    int maxb[2];             // bytes the row geometry allows (0: outside the matrix -> zero fill)
    int64_t step, mstep;     // element step of src / msk per stage
    int tloc, nthr;          // thread index inside the operand's raw-slot region, threads in that region
};

__device__ __forceinline__ void piece_setup(const StreamOp& o, int R, int sub, int64_t row0, int64_t k_first, int ct, int t0,
                                            PieceSet& ps) {
    ps.step = PK_KB * o.s_k;
    ps.mstep = PK_KB * o.m_k;
    if (o.mode == OP_TRANS) {
        const int t = ct - t0;
        const int RQ = R >> 2;
        const bool live = t >= 0 && t < R;
        const int tt = live ? t : 0;
        const int rq = tt % RQ, c = tt / RQ;
        const int64_t row = row0 + 4 * rq;
        int rb = 0;
        if (live && row < o.n_rows) rb = (o.n_rows - row >= 4) ? 16 : (int)(o.n_rows - row) * 4;
        ps.koff[0] = 4 * c;
        ps.maxb[0] = rb;
        ps.src[0] = o.P + (rb ? row + (k_first + 4 * c) * o.s_k : 0);
        ps.msk[0] = o.mask ? o.mask + (rb ? row + (k_first + 4 * c) * o.m_k : 0) : nullptr;
        ps.toff[0] = live ? tile_float_off(sub, 4 * rq, c) : -1;     // rows 4rq+e follow at +4 floats each
        ps.tloc = (t >= 0 && t < 256) ? t : 0;
        ps.nthr = 256;
        ps.koff[1] = 0; ps.maxb[1] = 0; ps.src[1] = o.P; ps.msk[1] = o.mask; ps.toff[1] = -1;
    } else {
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int idx = ct + PK_CONV_THREADS * q;
            const int c = idx / R, r = idx - c * R;
            const int64_t row = row0 + r;
            const bool in_tile = c < 4, ok = in_tile && row < o.n_rows;
            ps.koff[q] = 4 * c;
            ps.maxb[q] = ok ? 16 : 0;
            ps.src[q] = o.P + (ok ? row * o.s_row + (k_first + 4 * c) * o.s_k : 0);
            ps.msk[q] = o.mask ? o.mask + (ok ? row * o.m_row + (k_first + 4 * c) * o.m_k : 0) : nullptr;
            ps.toff[q] = in_tile ? tile_float_off(sub, r, c) : -1;
        }
        ps.tloc = ct;
        ps.nthr = PK_CONV_THREADS;
    }
}

// phase 1: cp.async of one stage (kr = min(K - k0, 64) of the stage).  `slots` = the operand's raw
// region of this depth + tloc; slot q is slots[q * nthr] (data first, then the mask slots); dead
// pieces / k tails are zero-filled by the copy itself (src-size < cp-size).
__device__ __forceinline__ void stream_issue(const StreamOp& o, PieceSet& ps, int kr, float4* slots) {
    const int nthr = ps.nthr;
    if (o.mode == OP_TRANS) {
        if (ps.toff[0] < 0) return;          // thread outside the operand's group range: owns no slot
        const int rb = ps.maxb[0];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int bytes = (ps.koff[0] + e < kr) ? rb : 0;
            cp_async16(&slots[e * nthr], bytes ? ps.src[0] + e * o.s_k : o.P, bytes);
            if (o.mask) cp_async16(&slots[(4 + e) * nthr], bytes ? ps.msk[0] + e * o.m_k : o.mask, bytes);
        }
        ps.src[0] += ps.step;
        if (o.mask) ps.msk[0] += ps.mstep;
    } else {
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            if (ps.toff[q] < 0) continue;     // piece outside the tile (R < 256)
            if (o.mode == OP_KVEC) {
                int bytes = (kr - ps.koff[q]) * 4;
                bytes = bytes > 16 ? 16 : (bytes < 0 ? 0 : bytes);
                if (!ps.maxb[q]) bytes = 0;
                cp_async16(&slots[q * nthr], bytes ? ps.src[q] : o.P, bytes);
                if (o.mask) cp_async16(&slots[(2 + q) * nthr], bytes ? ps.msk[q] : o.mask, bytes);
            } else {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const bool ok = ps.maxb[q] && ps.koff[q] + e < kr;
                    cp_async4(reinterpret_cast<float*>(&slots[q * nthr]) + e, ok ? ps.src[q] + e * o.s_k : o.P, ok ? 4 : 0);
                    if (o.mask)
                        cp_async4(reinterpret_cast<float*>(&slots[(2 + q) * nthr]) + e, ok ? ps.msk[q] + e * o.m_k : o.mask,
                                  ok ? 4 : 0);
                }
            }
            ps.src[q] += ps.step;
            if (o.mask) ps.msk[q] += ps.mstep;
        }
    }
}

__device__ __forceinline__ void split_store(float* tile, int off, int lo_off, float4 v) {
    float4 hi, lo;
    split_tf32(v.x, hi.x, lo.x);
    split_tf32(v.y, hi.y, lo.y);
    split_tf32(v.z, hi.z, lo.z);
    split_tf32(v.w, hi.w, lo.w);
    *reinterpret_cast<float4*>(tile + off) = hi;
    *reinterpret_cast<float4*>(tile + off + lo_off) = lo;
}

// phase 2: own raw slots -> (mask) -> hi/lo split -> MMA tile (sub = rows of a sub-tile: lo part at +sub*16 floats)
__device__ __forceinline__ void stream_convert(const StreamOp& o, const PieceSet& ps, int sub, float* tile, const float4* slots) {
    const int lo_off = sub * 16;
    const int nthr = ps.nthr;
    if (o.mode == OP_TRANS) {
        if (ps.toff[0] < 0) return;
        float4 x[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            x[e] = slots[e * nthr];
            if (o.mask) x[e] = pk_mask4(x[e], slots[(4 + e) * nthr], o.mask_act);
        }
        // 4x4 register transpose: output row i gets (k, k+1, k+2, k+3)
        split_store(tile, ps.toff[0], lo_off, make_float4(x[0].x, x[1].x, x[2].x, x[3].x));
        split_store(tile, ps.toff[0] + 4, lo_off, make_float4(x[0].y, x[1].y, x[2].y, x[3].y));
        split_store(tile, ps.toff[0] + 8, lo_off, make_float4(x[0].z, x[1].z, x[2].z, x[3].z));
        split_store(tile, ps.toff[0] + 12, lo_off, make_float4(x[0].w, x[1].w, x[2].w, x[3].w));
    } else {
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            if (ps.toff[q] < 0) continue;
            float4 v = slots[q * nthr];
            if (o.mask) v = pk_mask4(v, slots[(2 + q) * nthr], o.mask_act);
            split_store(tile, ps.toff[q], lo_off, v);
        }
    }
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// one 2-D tiled TMA load: box (x = first column, y = first row) of a row-major fp32 matrix; elements outside the
// matrix arrive as zeros (batch tail, feature tails) and still count towards the mbarrier's transaction bytes
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int x, int y, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(x), "r"(y), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}


__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
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

// 4 consecutive floats of one row with the same (warp-uniform) addressing rules: 128-bit access when
// the 4 elements exist and the address is 16-byte aligned, element-wise otherwise
__device__ __forceinline__ float4 ld4_or_scalar(const float* p, int nvalid, bool vec) {
    if (vec && nvalid == 4) return __ldg(reinterpret_cast<const float4*>(p));
    float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
    r.x = __ldg(p);
    if (nvalid > 1) r.y = __ldg(p + 1);
    if (nvalid > 2) r.z = __ldg(p + 2);
    if (nvalid > 3) r.w = __ldg(p + 3);
    return r;
}
__device__ __forceinline__ void st4_or_scalar(float* p, float4 v, int nvalid, bool vec) {
    if (vec && nvalid == 4) {
        *reinterpret_cast<float4*>(p) = v;
        return;
    }
    p[0] = v.x;
    if (nvalid > 1) p[1] = v.y;
    if (nvalid > 2) p[2] = v.z;
    if (nvalid > 3) p[3] = v.w;
}
__device__ __forceinline__ bool row_vec_ok(const float* base, int64_t ld, int64_t n) {
    return ((reinterpret_cast<uintptr_t>(base) & 15) == 0) && (ld % 4 == 0) && (n % 4 == 0);
}

__device__ __forceinline__ float4 act4(int act, float4 v) {
    if (act == CTR_ACT_RELU) return make_float4(fmaxf(v.x, 0.f), fmaxf(v.y, 0.f), fmaxf(v.z, 0.f), fmaxf(v.w, 0.f));
    if (act == CTR_ACT_LINEAR) return v;
    return make_float4(act_apply(act, v.x), act_apply(act, v.y), act_apply(act, v.z), act_apply(act, v.w));
}
__device__ __forceinline__ float4 actgrad4(int act, float4 v, float4 y) {
    if (act == CTR_ACT_RELU)
        return make_float4(y.x > 0.f ? v.x : 0.f, y.y > 0.f ? v.y : 0.f, y.z > 0.f ? v.z : 0.f, y.w > 0.f ? v.w : 0.f);
    if (act == CTR_ACT_LINEAR) return v;
    return make_float4(v.x * act_grad_from_y(act, y.x), v.y * act_grad_from_y(act, y.y), v.z * act_grad_from_y(act, y.z),
                       v.w * act_grad_from_y(act, y.w));
}

// epilogue of 4 consecutive outputs C[m, n..n+3] (nvalid of them exist); b4 = bias[n..n+3] (or zeros)
template <int EPI>
__device__ __forceinline__ void epi_store(const GemmArgs& g, float4 v, float4 b4, int64_t m, int64_t n, int nvalid,
                                          bool split) {
    if (EPI == EPI_BIAS_ACT) {
        v = act4(g.act, make_float4(v.x + b4.x, v.y + b4.y, v.z + b4.z, v.w + b4.w));
    } else if (EPI == EPI_MUL_ACTGRAD) {
        v = actgrad4(g.act, v, ld4_or_scalar(g.aux + m * g.ldaux + n, nvalid, row_vec_ok(g.aux, g.ldaux, n)));
    } else if (EPI == EPI_MUL) {
        const float4 y = ld4_or_scalar(g.aux + m * g.ldaux + n, nvalid, row_vec_ok(g.aux, g.ldaux, n));
        v = make_float4(v.x * y.x, v.y * y.y, v.z * y.z, v.w * y.w);
    } else if (EPI == EPI_CROSS) {
        const float4 u = make_float4(v.x + b4.x, v.y + b4.y, v.z + b4.z, v.w + b4.w);
        if (g.out2) st4_or_scalar(g.out2 + m * g.ldout2 + n, u, nvalid, row_vec_ok(g.out2, g.ldout2, n));
        const float4 x0 = ld4_or_scalar(g.aux + m * g.ldaux + n, nvalid, row_vec_ok(g.aux, g.ldaux, n));
        const float4 xl = ld4_or_scalar(g.aux2 + m * g.ldaux2 + n, nvalid, row_vec_ok(g.aux2, g.ldaux2, n));
        v = make_float4(x0.x * u.x + xl.x, x0.y * u.y + xl.y, x0.z * u.z + xl.z, x0.w * u.w + xl.w);
    }
    float* crow = g.C + m * g.ldc + n;
    const bool cvec = row_vec_ok(g.C, g.ldc, n);
    if (split) {
        if (cvec && nvalid == 4) {
            red_add4(crow, v);
        } else {
            atomicAdd(crow, v.x);
            if (nvalid > 1) atomicAdd(crow + 1, v.y);
            if (nvalid > 2) atomicAdd(crow + 2, v.z);
            if (nvalid > 3) atomicAdd(crow + 3, v.w);
        }
    } else {
        if (g.accumulate) {
            const float4 c = ld4_or_scalar(crow, nvalid, cvec);
            v.x += c.x; v.y += c.y; v.z += c.z; v.w += c.w;
        }
        st4_or_scalar(crow, v, nvalid, cvec);
    }
}

// Epilogue of one CTA tile (MT accumulators of 128 x BN in tensor memory) by 16 warps (CTA warps 2..17).
// Four warps per TMEM lane quadrant, interleaved over the 32-column chunks.  tcgen05.ld hands every
// thread ONE row (32 consecutive columns); written like that to C a warp would touch 32 different rows
// per store.  So each chunk is transposed through a padded per-warp staging tile in shared memory (the
// pipeline stages are free once accum_bar fired) and the epilogue math + stores run in the coalesced
// domain: 8 lanes cover 128 contiguous bytes of a row.
template <int EPI>
__device__ __forceinline__ void tile_epilogue(const GemmArgs& g, uint32_t tmem_base, unsigned char* smem_raw, int wid, int lane,
                                              int64_t m_tile0, int MT, int BN, int64_t n0, bool split) {
    const int ew = wid - 2;
    const int quad = wid & 3;                         // TMEM lane quadrant this warp may read
    const int cgrp = ew >> 2;
    float* stg = reinterpret_cast<float*>(smem_raw) + (size_t)ew * (32 * PK_STG_PITCH);
    const int nchunks = (BN + 31) / 32;
    for (int mt = 0; mt < MT; ++mt) {
        const int64_t m_base = m_tile0 + (int64_t)mt * PK_AR + quad * 32;
        if (m_base >= g.M) break;                     // warp-uniform
        for (int ci = cgrp; ci < nchunks; ci += PK_CONV_WARPS / 4) {
            const int c0 = ci * 32;
            if (n0 + c0 >= g.N) break;                // warp-uniform
            uint32_t raw[32];
            tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(mt * BN + c0), raw);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 8; ++j)
                *reinterpret_cast<uint4*>(stg + lane * PK_STG_PITCH + 4 * j) =
                    make_uint4(raw[4 * j], raw[4 * j + 1], raw[4 * j + 2], raw[4 * j + 3]);
            __syncwarp();
            const int colq = lane & 7;
            const int col = c0 + 4 * colq;            // column inside the CTA's N tile
            const int64_t n = n0 + col;
            // columns of this chunk that belong to the tile AND to the matrix
            int nvalid = 0;
            if (col < BN && n < g.N) {
                nvalid = BN - col < 4 ? BN - col : 4;
                if (g.N - n < nvalid) nvalid = (int)(g.N - n);
            }
            float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if ((EPI == EPI_BIAS_ACT || EPI == EPI_CROSS) && g.bias && nvalid > 0)
                b4 = ld4_or_scalar(g.bias + n, nvalid, row_vec_ok(g.bias, 4, n));
            // not unrolled on purpose: the body (with its activation variants) stays a few dozen
            // instructions, so the whole epilogue fits the instruction cache
#pragma unroll 1
            for (int i = 0; i < 8; ++i) {
                const int r = 4 * i + (lane >> 3);
                const int64_t m = m_base + r;
                const float4 v = *reinterpret_cast<const float4*>(stg + r * PK_STG_PITCH + 4 * colq);
                if (m < g.M && nvalid > 0) epi_store<EPI>(g, v, b4, m, n, nvalid, split);
            }
            __syncwarp();
        }
    }
}

template <int EPI>
__global__ void __launch_bounds__(PK_THREADS, 1) gemm_pk_kernel(PkParams p, const __grid_constant__ CUtensorMap map_a,
                                                                const __grid_constant__ CUtensorMap map_am) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const GemmArgs& g = p.g;
    const int BN = p.BN, MT = p.MT, SA = p.SA, SB = p.SB;
    const uint32_t a_tile = PK_AR * 128u;                 // hi+lo of a 128 x 16 tile
    const uint32_t a_stage = (uint32_t)MT * a_tile;
    const uint32_t b_stage = (uint32_t)BN * 128u;
    unsigned char* ringA = smem_raw;
    unsigned char* ringB = smem_raw + p.off_b;
    uint64_t* a_full = reinterpret_cast<uint64_t*>(smem_raw + p.off_bar);
    uint64_t* a_empty = a_full + SA;
    uint64_t* b_full = a_empty + SA;
    uint64_t* b_empty = b_full + SB;
    uint64_t* accum_bar = b_empty + SB;
    uint64_t* raw_full = accum_bar + 1;                   // [depth] raw A tiles delivered by TMA (a_tma)
    uint64_t* raw_empty = raw_full + p.depth;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(raw_empty + p.depth);

    const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
    const int64_t mblk = blockIdx.y, nblk = blockIdx.x;
    const int64_t kb_beg = (int64_t)blockIdx.z * p.kb_per_split;
    const int64_t kb_end = (kb_beg + p.kb_per_split < p.nkb) ? kb_beg + p.kb_per_split : p.nkb;
    const int nkb = (int)(kb_end - kb_beg);
    const bool split = gridDim.z > 1;
    const bool a_stream = p.sa.mode != OP_PACKED, b_stream = p.sb.mode != OP_PACKED;
    const bool a_tma = p.a_tma != 0;

    if (tid == 0) {
        // a stage is full after all converter warps arrived (streamed operand) or after the TMA
        // transaction armed by the producer completed (packed operand)
        for (int s = 0; s < SA; ++s) {
            mbar_init(&a_full[s], a_stream ? PK_CONV_WARPS : 1);
            mbar_init(&a_empty[s], 1);
        }
        for (int s = 0; s < SB; ++s) {
            mbar_init(&b_full[s], b_stream ? PK_CONV_WARPS : 1);
            mbar_init(&b_empty[s], 1);
        }
        mbar_init(accum_bar, 1);
        for (int s = 0; s < p.depth; ++s) {
            mbar_init(&raw_full[s], 1);
            mbar_init(&raw_empty[s], PK_CONV_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (wid == 1) tmem_alloc_warp(tmem_slot, (uint32_t)p.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (tid == 0) PK_DBG(7, 2);

    if (wid == 0) {
        // ------------------------------ TMA producer (one lane): packed operands, raw A tiles ------------
        if (lane == 0 && (!a_stream || !b_stream || a_tma)) {
            const unsigned char* a_src = reinterpret_cast<const unsigned char*>(p.Ap);
            const unsigned char* b_src = reinterpret_cast<const unsigned char*>(p.Bp);
            int sa = 0, sb = 0, dr = 0;
            uint32_t pha = 1, phb = 1, phr = 1;           // parity of the "empty" phase to wait for
            const uint32_t raw_tile = (uint32_t)MT * PK_AR * 64u;          // [MT*128 rows x 16 k] fp32
            const uint32_t raw_depth_bytes = p.raw_a_bytes + p.raw_b_bytes;
            for (int i = 0; i < nkb; ++i) {
                const int64_t kb = kb_beg + i;
                if (a_tma) {
                    mbar_wait(&raw_empty[dr], phr);
                    mbar_expect_tx(&raw_full[dr], raw_tile * (p.sa.mask ? 2u : 1u));
                    unsigned char* dst = smem_raw + p.off_raw + (size_t)dr * raw_depth_bytes;
                    tma_load_2d(dst, &map_a, (int)(kb * PK_KB), (int)(mblk * MT * PK_AR), &raw_full[dr]);
                    if (p.sa.mask) tma_load_2d(dst + 16384, &map_am, (int)(kb * PK_KB), (int)(mblk * MT * PK_AR), &raw_full[dr]);
                    if (++dr == p.depth) { dr = 0; phr ^= 1u; }
                }
                if (!a_stream) {
                    mbar_wait(&a_empty[sa], pha);
                    mbar_expect_tx(&a_full[sa], a_stage);
                    for (int mt = 0; mt < MT; ++mt)
                        bulk_g2s(ringA + (size_t)sa * a_stage + (size_t)mt * a_tile,
                                 a_src + ((mblk * MT + mt) * p.nkb + kb) * (int64_t)a_tile, a_tile, &a_full[sa]);
                    if (++sa == SA) { sa = 0; pha ^= 1u; }
                }
                if (!b_stream) {
                    mbar_wait(&b_empty[sb], phb);
                    mbar_expect_tx(&b_full[sb], b_stage);
                    bulk_g2s(ringB + (size_t)sb * b_stage, b_src + (nblk * p.nkb + kb) * (int64_t)b_stage, b_stage, &b_full[sb]);
                    if (++sb == SB) { sb = 0; phb ^= 1u; }
                }
            }
        }
        __syncwarp();
    } else if (wid == 1) {
        // ------------------------------ MMA issuer ------------------------------------------
        const uint32_t idesc = tf32_idesc(BN);
        const uint32_t a_lbo = PK_AR * 16u, b_lbo = (uint32_t)BN * 16u;     // between the two chunks of a K atom
        int sa = 0, sb = 0;
        uint32_t pha = 0, phb = 0;
        for (int i = 0; i < nkb; ++i) {
            mbar_wait(&a_full[sa], pha);
            if (lane == 0) PK_DBG(0, i);
            mbar_wait(&b_full[sb], phb);
            if (lane == 0) PK_DBG(1, i);
            tc_fence_after();
            if (lane == 0) {
                const uint32_t a_base = smem_u32(ringA + (size_t)sa * a_stage);
                const uint32_t b_hi = smem_u32(ringB + (size_t)sb * b_stage), b_lo = b_hi + b_stage / 2;
                if (!p.fast) {
#pragma unroll
                    for (int j = 0; j < PK_KB / 8; ++j) {
                        const uint64_t dbh = make_smem_desc(b_hi + (uint32_t)j * 2u * b_lbo, b_lbo, 128);
                        const uint64_t dbl = make_smem_desc(b_lo + (uint32_t)j * 2u * b_lbo, b_lbo, 128);
                        for (int mt = 0; mt < MT; ++mt) {
                            const uint32_t a_hi = a_base + (uint32_t)mt * a_tile + (uint32_t)j * 2u * a_lbo;
                            const uint64_t dah = make_smem_desc(a_hi, a_lbo, 128);
                            const uint64_t dal = make_smem_desc(a_hi + a_tile / 2, a_lbo, 128);
                            const uint32_t d = tmem_base + (uint32_t)(mt * BN);
                            umma_tf32(d, dal, dbh, idesc, (i | j) != 0 ? 1u : 0u);      // small terms first
                            umma_tf32(d, dah, dbl, idesc, 1u);
                            umma_tf32(d, dah, dbh, idesc, 1u);
                        }
                    }
                } else {                                  // labelled non-parity mode: single-pass TF32
#pragma unroll
                    for (int j = 0; j < PK_KB / 8; ++j) {
                        const uint64_t dbh = make_smem_desc(b_hi + (uint32_t)j * 2u * b_lbo, b_lbo, 128);
                        for (int mt = 0; mt < MT; ++mt) {
                            const uint64_t dah = make_smem_desc(a_base + (uint32_t)mt * a_tile + (uint32_t)j * 2u * a_lbo, a_lbo, 128);
                            umma_tf32(tmem_base + (uint32_t)(mt * BN), dah, dbh, idesc, (i | j) != 0 ? 1u : 0u);
                        }
                    }
                }
                umma_commit(&a_empty[sa]);                // frees the stages when these MMAs retire
                umma_commit(&b_empty[sb]);
                if (i == nkb - 1) umma_commit(accum_bar);
                PK_DBG(2, i);
            }
            __syncwarp();
            if (++sa == SA) { sa = 0; pha ^= 1u; }
            if (++sb == SB) { sb = 0; phb ^= 1u; }
        }
    } else {
        // ------------------------------ converters (16 warps), then epilogue ----------------
        if (a_stream || b_stream) {
            const int ct = tid - 64;
            unsigned char* raw = smem_raw + p.off_raw;
            const uint32_t raw_depth = p.raw_a_bytes + p.raw_b_bytes;
            const int64_t a_row0 = mblk * (int64_t)MT * PK_AR, b_row0 = nblk * (int64_t)BN;
            // operand descriptors in REGISTERS for the whole stage loop: read from the kernel-parameter
            // bank they were re-fetched (LDCU) before every dependent branch, the top stall of run 18
            StreamOp oa = p.sa, ob = p.sb;
            pin_op(oa);
            pin_op(ob);
            PieceSet pa, pb;
            if (a_stream) piece_setup(p.sa, MT * PK_AR, PK_AR, a_row0, kb_beg * PK_KB, ct, 0, pa);
            if (b_stream) piece_setup(p.sb, BN, BN, b_row0, kb_beg * PK_KB, ct, p.t0_b, pb);
            // a_tma: piece q of this thread = (row r, 16-byte chunk c) of the raw tile the TMA unit wrote with the
            // 64-byte swizzle: chunk c of row r sits at r*64 + ((c ^ ((r >> 1) & 3)) * 16) — a quarter warp (8 consecutive
            // rows, same c) then reads 8 different 16-byte bank groups
            int raw_off[2] = {0, 0};
            if (a_tma) {
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const int idx = ct + PK_CONV_THREADS * q, R = MT * PK_AR;
                    const int c = idx / R, r = idx - c * R;
                    raw_off[q] = r * 64 + ((c ^ ((r >> 1) & 3)) << 4);
                }
            }
            int64_t krem_issue = g.K - kb_beg * PK_KB;    // K left at the stage being issued
            int issued = 0, d_issue = 0;
            auto issue = [&]() {
                if (issued < nkb) {
                    const int kr = krem_issue > 64 ? 64 : (int)krem_issue;
                    unsigned char* base = raw + (size_t)d_issue * raw_depth;
                    if (a_stream && !a_tma) stream_issue(oa, pa, kr, reinterpret_cast<float4*>(base) + pa.tloc);
                    if (b_stream) stream_issue(ob, pb, kr, reinterpret_cast<float4*>(base + p.raw_a_bytes) + pb.tloc);
                }
                cp_async_commit();
                ++issued;
                krem_issue -= PK_KB;
                if (++d_issue == p.depth) d_issue = 0;
            };
            for (int i = 0; i < p.depth; ++i) issue();
            int sa = 0, sb = 0, d = 0;
            uint32_t pha = 1, phb = 1, phr = 0;
            for (int i = 0; i < nkb; ++i) {
                if (b_stream || !a_tma) {
                    if (p.depth == 3) cp_async_wait<2>();
                    else cp_async_wait<1>();
                }
                const unsigned char* base = raw + (size_t)d * raw_depth;
                if (a_tma) {
                    mbar_wait(&raw_full[d], phr);
                    if (ct == 0) PK_DBG(3, i);
                    float4 v[2];
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        if (pa.toff[q] < 0) continue;
                        v[q] = *reinterpret_cast<const float4*>(base + raw_off[q]);
                        if (oa.mask) v[q] = pk_mask4(v[q], *reinterpret_cast<const float4*>(base + 16384 + raw_off[q]), oa.mask_act);
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&raw_empty[d]);        // the raw tile may be refilled
                    mbar_wait(&a_empty[sa], pha);
                    if (ct == 0) PK_DBG(4, i);
                    float* tile = reinterpret_cast<float*>(ringA + (size_t)sa * a_stage);
#pragma unroll
                    for (int q = 0; q < 2; ++q)
                        if (pa.toff[q] >= 0) split_store(tile, pa.toff[q], PK_AR * 16, v[q]);
                } else if (a_stream) {
                    if (ct == 0) PK_DBG(3, i);
                    mbar_wait(&a_empty[sa], pha);
                    if (ct == 0) PK_DBG(4, i);
                    stream_convert(oa, pa, PK_AR, reinterpret_cast<float*>(ringA + (size_t)sa * a_stage),
                                   reinterpret_cast<const float4*>(base) + pa.tloc);
                }
                if (b_stream) {
                    mbar_wait(&b_empty[sb], phb);
                    stream_convert(ob, pb, BN, reinterpret_cast<float*>(ringB + (size_t)sb * b_stage),
                                   reinterpret_cast<const float4*>(base + p.raw_a_bytes) + pb.tloc);
                }
                fence_async_smem();                       // generic-proxy stores -> async proxy (tcgen05.mma)
                __syncwarp();
                if (lane == 0) {
                    if (a_stream) mbar_arrive(&a_full[sa]);
                    if (b_stream) mbar_arrive(&b_full[sb]);
                }
                if (ct == 0) PK_DBG(5, i);
                if (++sa == SA) { sa = 0; pha ^= 1u; }
                if (++sb == SB) { sb = 0; phb ^= 1u; }
                if (++d == p.depth) { d = 0; phr ^= 1u; }
                issue();
            }
            cp_async_wait<0>();
        }
        // ------------------------------ epilogue (16 warps) ---------------------------------
        mbar_wait(accum_bar, 0);
        if (wid == 2 && lane == 0) PK_DBG(7, 0);
        tc_fence_after();
        tile_epilogue<EPI>(g, tmem_base, smem_raw, wid, lane, mblk * MT * PK_AR, MT, BN, nblk * BN, split);
        tc_fence_before();
        if (wid == 2 && lane == 0) PK_DBG(7, 1);
    }
    __syncthreads();
    if (wid == 1) {
        tc_fence_after();
        tmem_dealloc_warp(tmem_base, (uint32_t)p.tmem_cols);
    }
}

// ---------------------------------------------------------------------------------------------
// TS engine (round 2): the streamed operand A goes through TENSOR MEMORY instead of shared memory.
//
//   C[m,n] = epilogue( sum_k A(m,k) * W(n,k) )     A: activations [M, K] row-major (K contiguous),
//                                                  W: weights, pre-split + pre-tiled once per call (pack_*)
//
// Why: in the SS engine above every fp32 of A crosses shared memory four times (cp.async in, read, hi and
// lo written back) and is then read three more times by the MMAs; with the weight tiles that is ~156 B/clk
// against the 128 B/clk an SM's shared memory delivers — the tensor pipe idles (run 18: 15-37 %).
// tcgen05.mma takes its A operand from tensor memory as well (probe: scripts/probe_umma.cu, bit-identical
// to the shared-memory form), one TMEM lane per row and one 32-bit column per k.  So:
//   * a converter warp copies ITS 32 rows x 16 k of A with coalesced cp.async (4 lanes = 64 contiguous
//     bytes of a row) into a raw staging tile, every thread then reads its own row back (conflict-free,
//     80-byte pitch), splits it into hi/lo and writes both halves straight into its TMEM lane with
//     tcgen05.st: no hi/lo tile in shared memory, no generic->async proxy fence, no A reads by the MMA;
//   * shared memory only carries the weight stages (TMA in, MMA out) and the 10 KB raw tiles: ~125 B/clk
//     at the MMA-bound rate;
//   * 16 converter warps = 4 groups of 4 warps (one warp per TMEM lane quadrant); a group owns every 4th
//     (k stage, M tile) item, so four items are in conversion while the MMAs of earlier ones run;
//   * TMEM: MT accumulators of BN columns (MT * BN <= 256) + a ring of A slots (32 columns = 16 k x {hi, lo}).
// ---------------------------------------------------------------------------------------------
constexpr int TS_NG = 4;                          // converter groups
constexpr int TS_RAW_PITCH = 20;                  // floats per row of a raw item (16 k + 4 padding: conflict-free LDS.128)
constexpr uint32_t TS_RAW_TILE = PK_AR * TS_RAW_PITCH * 4;    // 10 240 bytes: 128 rows
constexpr int TS_THREADS = PK_THREADS + 128;                  // + 4 warps that derive the lo half of raw weight stages

struct TsParams {
    GemmArgs g;
    const float* Bp;                 // packed W tiles (row blocks of BN)
    int64_t nkb;
    int BN, MT, SB, SLOTS, RAWD, tmem_cols, a_col0, has_mask;
    uint32_t off_raw, off_bar, raw_item_bytes;
    unsigned long long* dbg;         // optional clock64 timeline of CTA (0,0) (ctr_debug_set_buffer): [event][64 stages]
    int b_raw;                       // weight stages arrive as raw fp32 tiles (half the bytes): hi = the raw tile, lo derived in the CTA
    int CL;                          // thread-block cluster size along the M tiles: every weight stage is read from L2 once
                                     // per cluster (each CTA fetches 1/CL of it and multicasts it to all of them)
};
#define TS_DBG(ev, idx)                                                             \
    do {                                                                            \
        if (p.dbg && blockIdx.x == 0 && blockIdx.y == 0 && (idx) < 64) p.dbg[(ev) * 64 + (idx)] = clock64(); \
    } while (0)

__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
        "r"(tmem_a), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(
            taddr),
        "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
        "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
        "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
        "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void bulk_g2s_mcast(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint16_t mask) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(
            smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar)), "h"(mask)
        : "memory");
}
__device__ __forceinline__ void umma_commit_mcast(uint64_t* bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     smem_u32(bar)),
                 "h"(mask)
                 : "memory");
}

template <int EPI>
__global__ void __launch_bounds__(TS_THREADS, 1) gemm_ts_kernel(TsParams p) {
    if (p.dbg && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) p.dbg[7 * 64 + 2] = clock64();
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const GemmArgs& g = p.g;
    const int BN = p.BN, MT = p.MT, SB = p.SB, SLOTS = p.SLOTS;
    const uint32_t b_stage = (uint32_t)BN * 128u;
    unsigned char* ringB = smem_raw;
    uint64_t* b_full = reinterpret_cast<uint64_t*>(smem_raw + p.off_bar);
    uint64_t* b_empty = b_full + SB;
    uint64_t* a_full = b_empty + SB;                  // [MT * SLOTS]
    uint64_t* a_empty = a_full + MT * SLOTS;
    uint64_t* accum_bar = a_empty + MT * SLOTS;
    uint64_t* b_rawfull = accum_bar + 1;              // [SB] raw fp32 weight tile landed (b_raw)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(b_rawfull + SB);

    const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
    const int64_t mblk = blockIdx.y, nblk = blockIdx.x;
    const int nkb = (int)p.nkb;

    if (tid == 0) {
        for (int s = 0; s < SB; ++s) {
            mbar_init(&b_rawfull[s], 1);
            // TMA transaction — or, b_raw, the four warps that derived the lo half
            mbar_init(&b_full[s], p.b_raw ? 4 : 1);
            mbar_init(&b_empty[s], p.CL);             // every CTA of the cluster has finished reading the stage
        }
        for (int s = 0; s < MT * SLOTS; ++s) {
            mbar_init(&a_full[s], 4);                 // the four warps of the group that filled the slot
            mbar_init(&a_empty[s], 1);
        }
        mbar_init(accum_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (wid == 1) tmem_alloc_warp(tmem_slot, (uint32_t)p.tmem_cols);
    tc_fence_before();
    __syncthreads();
    if (p.CL > 1) cluster_sync_all();                 // peers' barriers exist before anything is signalled remotely
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint16_t cl_mask = (uint16_t)((1u << p.CL) - 1u);

    if (wid == 0) {
        // ------------------------------ TMA producer: weight stages -------------------------------
        if (lane == 0) {
            const unsigned char* b_src = reinterpret_cast<const unsigned char*>(p.Bp);
            const uint32_t tma_bytes = p.b_raw ? b_stage / 2 : b_stage;
            uint64_t* const tma_bar = p.b_raw ? b_rawfull : b_full;
            const uint32_t slice = tma_bytes / (uint32_t)p.CL, my = (p.CL > 1) ? cluster_ctarank() * slice : 0u;
            int sb = 0;
            uint32_t phb = 1;
            for (int i = 0; i < nkb; ++i) {
                mbar_wait(&b_empty[sb], phb);
                TS_DBG(6, i);
                mbar_expect_tx(&tma_bar[sb], tma_bytes);    // own slice + the peers' multicast slices
                const unsigned char* src = b_src + (nblk * p.nkb + i) * (int64_t)tma_bytes + my;
                unsigned char* dst = ringB + (size_t)sb * b_stage + my;
                if (p.CL > 1) bulk_g2s_mcast(dst, src, slice, &tma_bar[sb], cl_mask);
                else bulk_g2s(dst, src, slice, &tma_bar[sb]);
                if (++sb == SB) { sb = 0; phb ^= 1u; }
            }
        }
        __syncwarp();
    } else if (wid == 1) {
        // ------------------------------ MMA issuer ------------------------------------------------
        const uint32_t idesc = tf32_idesc(BN);
        const uint32_t b_lbo = (uint32_t)BN * 16u;
        int sb = 0;
        uint32_t phb = 0;
        int slot = 0;                                      // i % SLOTS and (i / SLOTS) & 1 kept as counters: a runtime
        uint32_t pa = 0;                                   // integer division costs this single thread ~200 cycles per stage
        for (int i = 0; i < nkb; ++i) {
            mbar_wait(&b_full[sb], phb);
            if (lane == 0) TS_DBG(0, i);
            for (int mt = 0; mt < MT; ++mt) mbar_wait(&a_full[mt * SLOTS + slot], pa);
            if (lane == 0) TS_DBG(1, i);
            tc_fence_after();
            if (lane == 0) {
                const uint32_t b_hi = smem_u32(ringB + (size_t)sb * b_stage), b_lo = b_hi + b_stage / 2;
#pragma unroll
                for (int j = 0; j < PK_KB / 8; ++j) {
                    const uint64_t dbh = make_smem_desc(b_hi + (uint32_t)j * 2u * b_lbo, b_lbo, 128);
                    const uint64_t dbl = make_smem_desc(b_lo + (uint32_t)j * 2u * b_lbo, b_lbo, 128);
                    for (int mt = 0; mt < MT; ++mt) {
                        const uint32_t a_hi = tmem_base + (uint32_t)(p.a_col0 + (mt * SLOTS + slot) * 32 + 8 * j);
                        const uint32_t a_lo = a_hi + 16u;
                        const uint32_t d = tmem_base + (uint32_t)(mt * BN);
                        umma_tf32_ts(d, a_lo, dbh, idesc, (i | j) != 0 ? 1u : 0u);      // small terms first
                        umma_tf32_ts(d, a_hi, dbl, idesc, 1u);
                        umma_tf32_ts(d, a_hi, dbh, idesc, 1u);
                    }
                }
                // frees the stage (in every CTA of the cluster) / the slots when these MMAs retire
                if (p.CL > 1) umma_commit_mcast(&b_empty[sb], cl_mask);
                else umma_commit(&b_empty[sb]);
                for (int mt = 0; mt < MT; ++mt) umma_commit(&a_empty[mt * SLOTS + slot]);
                if (i == nkb - 1) umma_commit(accum_bar);
                TS_DBG(2, i);
            }
            __syncwarp();
            if (++sb == SB) { sb = 0; phb ^= 1u; }
            if (++slot == SLOTS) { slot = 0; pa ^= 1u; }
        }
    } else if (wid >= 2 + PK_CONV_WARPS) {
        // ------------------------------ weight-lo warps (b_raw) -----------------------------------
        // The weight stage arrived as raw fp32 (= hi after the tensor core's truncation): lo = RN_tf32(w - trunc(w)),
        // elementwise over the tile (the chunk layout does not matter).  Half the bytes cross L2 -> SM.
        if (p.b_raw) {
            const int t = tid - (2 + PK_CONV_WARPS) * 32;
            const int n4 = (int)(b_stage / 32);               // float4 per half
            int sb = 0;
            uint32_t ph = 0;
            for (int i = 0; i < nkb; ++i) {
                mbar_wait(&b_rawfull[sb], ph);
                float4* hi4 = reinterpret_cast<float4*>(ringB + (size_t)sb * b_stage);
                float4* lo4 = reinterpret_cast<float4*>(ringB + (size_t)sb * b_stage + b_stage / 2);
                for (int q = t; q < n4; q += 128) {
                    const float4 w = hi4[q];
                    float4 l;
                    l.x = rn_tf32(w.x - __uint_as_float(__float_as_uint(w.x) & 0xFFFFE000u));
                    l.y = rn_tf32(w.y - __uint_as_float(__float_as_uint(w.y) & 0xFFFFE000u));
                    l.z = rn_tf32(w.z - __uint_as_float(__float_as_uint(w.z) & 0xFFFFE000u));
                    l.w = rn_tf32(w.w - __uint_as_float(__float_as_uint(w.w) & 0xFFFFE000u));
                    lo4[q] = l;
                }
                fence_async_smem();                           // generic-proxy stores -> async proxy (tcgen05.mma)
                __syncwarp();
                if (lane == 0) mbar_arrive(&b_full[sb]);
                if (++sb == SB) { sb = 0; ph ^= 1u; }
            }
        }
    } else {
        // ------------------------------ converters (4 groups x 4 warps), then epilogue ------------
        const int cw = wid - 2;
        const int grp = cw >> 2;
        const int quad = wid & 3;                          // TMEM lane quadrant of this warp = its 32 rows of the M tile
        const int n_items = nkb * MT;
        unsigned char* raw_grp = smem_raw + p.off_raw + (size_t)grp * p.RAWD * p.raw_item_bytes;
        const float* Aptr = g.A;
        const float* Mptr = g.amask;
        const int64_t sam = g.sam, smm = g.smm;
        const int mask_act = g.amask_act;
        // cp.async mapping: instruction j moves rows 8j + lane/4 of the warp's 32, chunk lane%4 (64 contiguous bytes per row)
        const int cp_row = lane >> 2, cp_chunk = lane & 3;
        int issued = 0, raw_issue = 0;                     // items of this group issued so far / their raw buffer
        auto issue = [&]() {
            const int it = grp + issued * TS_NG;
            if (it < n_items) {
                const int i = MT == 2 ? it >> 1 : it, mt = MT == 2 ? it & 1 : 0;        // MT is 1 or 2
                unsigned char* buf = raw_grp + (size_t)raw_issue * p.raw_item_bytes;
                const int64_t k0 = (int64_t)i * PK_KB + cp_chunk * 4;
                int kbytes = (int)(g.K - k0) * 4;
                kbytes = kbytes > 16 ? 16 : (kbytes < 0 ? 0 : kbytes);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int rt = quad * 32 + 8 * j + cp_row;
                    const int64_t row = (mblk * MT + mt) * (int64_t)PK_AR + rt;
                    const int bytes = row < g.M ? kbytes : 0;
                    float* dst = reinterpret_cast<float*>(buf) + rt * TS_RAW_PITCH + cp_chunk * 4;
                    cp_async16(dst, bytes ? Aptr + row * sam + k0 : Aptr, bytes);
                    if (p.has_mask)
                        cp_async16(reinterpret_cast<float*>(buf + TS_RAW_TILE) + rt * TS_RAW_PITCH + cp_chunk * 4,
                                   bytes ? Mptr + row * smm + k0 : Mptr, bytes);
                }
            }
            cp_async_commit();
            ++issued;
            if (++raw_issue == p.RAWD) raw_issue = 0;
        };
        for (int d = 0; d < p.RAWD; ++d) issue();
        // stage i = it / MT advances by TS_NG / MT per item: slot = i % SLOTS and the slot's use parity are counters
        const int i_step = MT == 2 ? TS_NG / 2 : TS_NG;
        int raw_conv = 0, slot = (MT == 2 ? grp >> 1 : grp);
        uint32_t slot_par = 0;
        while (slot >= SLOTS) { slot -= SLOTS; slot_par ^= 1u; }
        for (int it = grp; it < n_items; it += TS_NG) {
            const int i = MT == 2 ? it >> 1 : it, mt = MT == 2 ? it & 1 : 0;
            if (p.RAWD == 3) cp_async_wait<2>();
            else if (p.RAWD == 2) cp_async_wait<1>();
            else cp_async_wait<0>();
            __syncwarp();                                  // the warp's lanes copied each other's rows
            if (quad == 0 && lane == 0) TS_DBG(3, i);
            const unsigned char* buf = raw_grp + (size_t)raw_conv * p.raw_item_bytes;
            const float* myrow = reinterpret_cast<const float*>(buf) + (quad * 32 + lane) * TS_RAW_PITCH;
            float hi[16], lo[16];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                float4 v = *reinterpret_cast<const float4*>(myrow + 4 * c);
                if (p.has_mask) v = pk_mask4(v, *reinterpret_cast<const float4*>(myrow + TS_RAW_TILE / 4 + 4 * c), mask_act);
                split_tf32(v.x, hi[4 * c + 0], lo[4 * c + 0]);
                split_tf32(v.y, hi[4 * c + 1], lo[4 * c + 1]);
                split_tf32(v.z, hi[4 * c + 2], lo[4 * c + 2]);
                split_tf32(v.w, hi[4 * c + 3], lo[4 * c + 3]);
            }
            __syncwarp();                                  // every lane has read its row: the raw tile may be refilled
            issue();
            mbar_wait(&a_empty[mt * SLOTS + slot], slot_par ^ 1u);
            if (quad == 0 && lane == 0) TS_DBG(4, i);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(p.a_col0 + (mt * SLOTS + slot) * 32);
            tmem_st16(taddr, hi);
            tmem_st16(taddr + 16u, lo);
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&a_full[mt * SLOTS + slot]);
            if (quad == 0 && lane == 0) TS_DBG(5, i);
            if (++raw_conv == p.RAWD) raw_conv = 0;
            slot += i_step;
            while (slot >= SLOTS) { slot -= SLOTS; slot_par ^= 1u; }
        }
        cp_async_wait<0>();
        // ------------------------------ epilogue (16 warps) ---------------------------------------
        mbar_wait(accum_bar, 0);
        if (wid == 2 && lane == 0) TS_DBG(7, 0);
        tc_fence_after();
        tile_epilogue<EPI>(g, tmem_base, smem_raw, wid, lane, mblk * MT * PK_AR, MT, BN, nblk * BN, false);
        if (wid == 2 && lane == 0) TS_DBG(7, 1);
        tc_fence_before();
    }
    __syncthreads();
    if (p.CL > 1) cluster_sync_all();                 // nobody leaves while a peer may still multicast into / signal this CTA
    if (wid == 1) {
        tc_fence_after();
        tmem_dealloc_warp(tmem_base, (uint32_t)p.tmem_cols);
    }
}

// ---------------------------------------------------------------------------------------------
// PP engine (round 2): the SS engine as a PERSISTENT, tile-pipelined kernel for the tall GEMMs of the tower
// (M = batch, K-contiguous streamed A by TMA, packed weights, no split-K).  One CTA per SM loops over 128 x BN
// tiles; the accumulator is double-buffered in tensor memory (2 x BN <= 512 columns) and the roles are split:
//   warp 0      TMA producer (raw A tiles + packed weight stages, rings run across tile boundaries)
//   warp 1      MMA issuer (one lane)
//   warps 2-9   converters (raw fp32 -> hi/lo MMA tile): two groups of 4 warps on alternating stages
//   warps 10-17 epilogue of tile t while the main loop of tile t+1 runs
// Why: in the one-tile-per-CTA kernel the prologue (4.5 k cycles) and the epilogue (10 k) of every tile are exposed
// and 256-512 tiles leave a partial last wave on 148 SMs; with K = 256..432 that is a third of a tile's time
// (profiles/r02_gemm_engines.md section 5).
// ---------------------------------------------------------------------------------------------
struct PpParams {
    GemmArgs g;
    const float* Bp;                 // packed B (row blocks of BN)
    int64_t nkb, gm, gn;
    int BN, SA, SB, RD, tmem_cols, has_mask, fast;
    uint32_t off_b, off_raw, off_stg, off_bar;
};
constexpr int PP_CONV_WARPS = 8, PP_EPI_WARPS = 8;

template <int EPI>
__device__ __forceinline__ void pp_tile_epilogue(const GemmArgs& g, uint32_t tmem_acc, float* stg_base, int ew, int lane,
                                                 int64_t m_tile0, int BN, int64_t n0) {
    const int quad = ew & 3;                          // TMEM lane quadrant of this warp (= CTA warp index & 3, hardware rule)
    const int cgrp = ew >> 2;                         // which half of the 32-column chunks
    float* stg = stg_base + (size_t)(ew & 7) * (32 * PK_STG_PITCH);
    const int nchunks = (BN + 31) / 32;
    const int64_t m_base = m_tile0 + quad * 32;
    if (m_base >= g.M) return;
    for (int ci = cgrp; ci < nchunks; ci += PP_EPI_WARPS / 4) {
        const int c0 = ci * 32;
        if (n0 + c0 >= g.N) break;
        uint32_t raw[32];
        tmem_ld32(tmem_acc + ((uint32_t)(quad * 32) << 16) + (uint32_t)c0, raw);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 8; ++j)
            *reinterpret_cast<uint4*>(stg + lane * PK_STG_PITCH + 4 * j) =
                make_uint4(raw[4 * j], raw[4 * j + 1], raw[4 * j + 2], raw[4 * j + 3]);
        __syncwarp();
        const int colq = lane & 7;
        const int col = c0 + 4 * colq;
        const int64_t n = n0 + col;
        int nvalid = 0;
        if (col < BN && n < g.N) {
            nvalid = BN - col < 4 ? BN - col : 4;
            if (g.N - n < nvalid) nvalid = (int)(g.N - n);
        }
        float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if ((EPI == EPI_BIAS_ACT || EPI == EPI_CROSS) && g.bias && nvalid > 0)
            b4 = ld4_or_scalar(g.bias + n, nvalid, row_vec_ok(g.bias, 4, n));
#pragma unroll 1
        for (int i = 0; i < 8; ++i) {
            const int r = 4 * i + (lane >> 3);
            const int64_t m = m_base + r;
            const float4 v = *reinterpret_cast<const float4*>(stg + r * PK_STG_PITCH + 4 * colq);
            if (m < g.M && nvalid > 0) epi_store<EPI>(g, v, b4, m, n, nvalid, false);
        }
        __syncwarp();
    }
}

template <int EPI>
__global__ void __launch_bounds__(PK_THREADS, 1) gemm_pp_kernel(PpParams p, const __grid_constant__ CUtensorMap map_a,
                                                                const __grid_constant__ CUtensorMap map_am) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const GemmArgs& g = p.g;
    const int BN = p.BN, SA = p.SA, SB = p.SB, RD = p.RD;
    const uint32_t a_stage = PK_AR * 128u;                // hi+lo of a 128 x 16 tile
    const uint32_t b_stage = (uint32_t)BN * 128u;
    const uint32_t raw_item = PK_AR * 64u * (p.has_mask ? 2u : 1u);
    unsigned char* ringA = smem_raw;
    unsigned char* ringB = smem_raw + p.off_b;
    unsigned char* ringR = smem_raw + p.off_raw;
    uint64_t* a_full = reinterpret_cast<uint64_t*>(smem_raw + p.off_bar);
    uint64_t* a_empty = a_full + SA;
    uint64_t* b_full = a_empty + SA;
    uint64_t* b_empty = b_full + SB;
    uint64_t* raw_full = b_empty + SB;
    uint64_t* raw_empty = raw_full + RD;
    uint64_t* acc_full = raw_empty + RD;                  // [2]
    uint64_t* acc_empty = acc_full + 2;                   // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

    const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
    const int nkb = (int)p.nkb;
    const int64_t n_tiles = p.gm * p.gn;

    if (tid == 0) {
        for (int s = 0; s < SA; ++s) { mbar_init(&a_full[s], PP_CONV_WARPS / 2); mbar_init(&a_empty[s], 1); }
        for (int s = 0; s < SB; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
        for (int s = 0; s < RD; ++s) { mbar_init(&raw_full[s], 1); mbar_init(&raw_empty[s], PP_CONV_WARPS / 2); }
        for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], PP_EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (wid == 1) tmem_alloc_warp(tmem_slot, (uint32_t)p.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (wid == 0) {
        // ------------------------------ TMA producer ---------------------------------------------
        if (lane == 0) {
            const unsigned char* b_src = reinterpret_cast<const unsigned char*>(p.Bp);
            int sb = 0, dr = 0;
            uint32_t phb = 1, phr = 1;
            for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
                const int64_t mblk = t / p.gn, nblk = t - mblk * p.gn;
                for (int i = 0; i < nkb; ++i) {
                    mbar_wait(&raw_empty[dr], phr);
                    mbar_expect_tx(&raw_full[dr], raw_item);
                    unsigned char* dst = ringR + (size_t)dr * raw_item;
                    tma_load_2d(dst, &map_a, i * PK_KB, (int)(mblk * PK_AR), &raw_full[dr]);
                    if (p.has_mask) tma_load_2d(dst + PK_AR * 64u, &map_am, i * PK_KB, (int)(mblk * PK_AR), &raw_full[dr]);
                    if (++dr == RD) { dr = 0; phr ^= 1u; }
                    mbar_wait(&b_empty[sb], phb);
                    mbar_expect_tx(&b_full[sb], b_stage);
                    bulk_g2s(ringB + (size_t)sb * b_stage, b_src + (nblk * p.nkb + i) * (int64_t)b_stage, b_stage, &b_full[sb]);
                    if (++sb == SB) { sb = 0; phb ^= 1u; }
                }
            }
        }
        __syncwarp();
    } else if (wid == 1) {
        // ------------------------------ MMA issuer -----------------------------------------------
        if (lane == 0) {
            const uint32_t idesc = tf32_idesc(BN);
            const uint32_t a_lbo = PK_AR * 16u, b_lbo = (uint32_t)BN * 16u;
            int sa = 0, sb = 0, buf = 0;
            uint32_t pha = 0, phb = 0, phe[2] = {1, 1};
            for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
                mbar_wait(&acc_empty[buf], phe[buf]);         // the epilogue drained this accumulator
                phe[buf] ^= 1u;
                tc_fence_after();
                const uint32_t d = tmem_base + (uint32_t)(buf * BN);
                for (int i = 0; i < nkb; ++i) {
                    mbar_wait(&a_full[sa], pha);
                    mbar_wait(&b_full[sb], phb);
                    tc_fence_after();
                    const uint32_t a_base = smem_u32(ringA + (size_t)sa * a_stage);
                    const uint32_t b_hi = smem_u32(ringB + (size_t)sb * b_stage), b_lo = b_hi + b_stage / 2;
                    if (!p.fast) {
#pragma unroll
                        for (int j = 0; j < PK_KB / 8; ++j) {
                            const uint64_t dbh = make_smem_desc(b_hi + (uint32_t)j * 2u * b_lbo, b_lbo, 128);
                            const uint64_t dbl = make_smem_desc(b_lo + (uint32_t)j * 2u * b_lbo, b_lbo, 128);
                            const uint32_t a_hi = a_base + (uint32_t)j * 2u * a_lbo;
                            const uint64_t dah = make_smem_desc(a_hi, a_lbo, 128);
                            const uint64_t dal = make_smem_desc(a_hi + a_stage / 2, a_lbo, 128);
                            umma_tf32(d, dal, dbh, idesc, (i | j) != 0 ? 1u : 0u);      // small terms first
                            umma_tf32(d, dah, dbl, idesc, 1u);
                            umma_tf32(d, dah, dbh, idesc, 1u);
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < PK_KB / 8; ++j) {
                            const uint64_t dbh = make_smem_desc(b_hi + (uint32_t)j * 2u * b_lbo, b_lbo, 128);
                            const uint64_t dah = make_smem_desc(a_base + (uint32_t)j * 2u * a_lbo, a_lbo, 128);
                            umma_tf32(d, dah, dbh, idesc, (i | j) != 0 ? 1u : 0u);
                        }
                    }
                    umma_commit(&a_empty[sa]);
                    umma_commit(&b_empty[sb]);
                    if (i == nkb - 1) umma_commit(&acc_full[buf]);
                    if (++sa == SA) { sa = 0; pha ^= 1u; }
                    if (++sb == SB) { sb = 0; phb ^= 1u; }
                }
                buf ^= 1;
            }
        }
        __syncwarp();
    } else if (wid < 2 + PP_CONV_WARPS) {
        // ------------------------------ converters: 2 groups x 4 warps on alternating stages -----
        // One group converts a whole 128 x 16 stage (4 pieces per thread); the per-stage chain of a group (wait raw ->
        // LDS -> wait free stage -> split -> STS -> proxy fence -> arrive) is latency-bound, so two groups in flight
        // double the conversion rate (one group of 8 warps: 2 000 cycles per stage against 1 026 of MMA work).
        const int grp = (wid - 2) >> 2;                       // 0 / 1
        const int ct = tid - 64 - grp * 128;                  // 0..127 inside the group
        int raw_off[4], toff[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int idx = ct + 128 * q;
            const int c = idx >> 7, r = idx & 127;           // 512 pieces = 4 chunks x 128 rows
            raw_off[q] = r * 64 + ((c ^ ((r >> 1) & 3)) << 4);
            toff[q] = (c * PK_AR + r) * 4;
        }
        const int mask_act = g.amask_act;
        // ring positions of this group's FIRST stage (global stage index = grp), then +2 per iteration
        int sa = grp % SA, dr = grp % RD;
        uint32_t pha = 1u ^ (uint32_t)((grp / SA) & 1), phr = (uint32_t)((grp / RD) & 1);
        int64_t my_tiles = 0;
        for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) ++my_tiles;
        const int64_t total_stages = my_tiles * nkb;
        for (int64_t gs = grp; gs < total_stages; gs += 2) {
            mbar_wait(&raw_full[dr], phr);
            const unsigned char* base = ringR + (size_t)dr * raw_item;
            float4 v[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                v[q] = *reinterpret_cast<const float4*>(base + raw_off[q]);
                if (p.has_mask) v[q] = pk_mask4(v[q], *reinterpret_cast<const float4*>(base + PK_AR * 64u + raw_off[q]), mask_act);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&raw_empty[dr]);
            dr += 2;
            if (dr >= RD) { dr -= RD; phr ^= 1u; }
            mbar_wait(&a_empty[sa], pha);
            float* tile = reinterpret_cast<float*>(ringA + (size_t)sa * a_stage);
#pragma unroll
            for (int q = 0; q < 4; ++q) split_store(tile, toff[q], PK_AR * 16, v[q]);
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&a_full[sa]);
            sa += 2;
            if (sa >= SA) { sa -= SA; pha ^= 1u; }
        }
    } else {
        // ------------------------------ epilogue warps -------------------------------------------
        const int ew = wid - 10;                               // 0..7; TMEM quadrant = wid & 3 = (ew + 2) & 3
        const int ewq = ((wid & 3)) | ((ew >> 2) << 2);        // quadrant in the low bits, column group above
        float* stg_base = reinterpret_cast<float*>(smem_raw + p.off_stg);
        int buf = 0;
        uint32_t phf[2] = {0, 0};
        for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
            const int64_t mblk = t / p.gn, nblk = t - mblk * p.gn;
            mbar_wait(&acc_full[buf], phf[buf]);
            phf[buf] ^= 1u;
            tc_fence_after();
            pp_tile_epilogue<EPI>(g, tmem_base + (uint32_t)(buf * BN), stg_base, ewq, lane, mblk * PK_AR, BN, nblk * BN);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[buf]);
            buf ^= 1;
        }
    }
    __syncthreads();
    if (wid == 1) {
        tc_fence_after();
        tmem_dealloc_warp(tmem_base, (uint32_t)p.tmem_cols);
    }
}

// ---------------------------------------------------------------------------------------------
// TSW engine (round 2): the weight-gradient GEMM   dW[n, k] (+)= sum_b dZ[b, n] * X[b, k]
// (reference: autograd of nn.Linear inside DNN, layers/core.py:120-134) with BOTH operands streamed from
// the batch-major activations and the contraction over the BATCH:
//   * "A" = dZ^T goes through tensor memory: a converter thread owns ONE output feature n (= TMEM lane) and reads
//     its COLUMN of the raw [16 samples x 128 features] tile (conflict-free LDS.32) — the transposition the SS
//     engine does with 4x4 register shuffles and transposed shared-memory stores is free here; the same thread
//     also accumulates db[n] = sum_b dZ[b, n] (the bias gradient: no separate column-sum kernel);
//   * "B" = X^T is converted by two other warp groups into the K-major [hi | lo] tile the MMA reads from shared
//     memory (4 LDS.32 down a column -> split -> two 128-bit stores, conflict-free);
//   * K (the batch) is split across CTAs; partial tiles are reduced into dW with 128-bit fp32 reductions.
// ---------------------------------------------------------------------------------------------
constexpr int TW_A_PITCH = 128;                  // floats per row of a raw A tile (dense TMA box: 16 samples x 128 features)
constexpr uint32_t TW_A_TILE = 16 * TW_A_PITCH * 4;           // 8 192 bytes

struct TwParams {
    GemmArgs g;                      // M = out features, N = in features, K = batch
    int64_t nkb, kb_per_split;
    int BN, SLOTS, SBW, RA, RB, tmem_cols, a_col0, has_mask;
    uint32_t off_braw, off_araw, off_bar, b_raw_bytes, a_raw_bytes;
    int b_pitch;                     // floats per row of a raw B tile (BN + 4)
    int fast;                        // single-pass TF32 (non-parity fast mode)
    float* db;                       // bias gradient [M] (NULL: not wanted), accumulated with atomics
    unsigned long long* dbg;         // optional clock64 timeline of CTA (0,0,0) (ctr_debug_set_buffer): [10 events][64 stages]
};
#define TW_DBG(ev, idx)                                                                                   \
    do {                                                                                                  \
        if (p.dbg && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && (idx) < 64) p.dbg[(ev) * 64 + (idx)] = clock64(); \
    } while (0)

template <int DUMMY>
__global__ void __launch_bounds__(PK_THREADS, 1) gemm_tsw_kernel(TwParams p, const __grid_constant__ CUtensorMap map_a,
                                                                 const __grid_constant__ CUtensorMap map_m,
                                                                 const __grid_constant__ CUtensorMap map_b) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const GemmArgs& g = p.g;
    const int BN = p.BN, SLOTS = p.SLOTS, SBW = p.SBW, RA = p.RA, RB = p.RB;
    const uint32_t b_stage = (uint32_t)BN * 128u;
    unsigned char* ringB = smem_raw;                                  // converted B stages
    uint64_t* b_full = reinterpret_cast<uint64_t*>(smem_raw + p.off_bar);
    uint64_t* b_empty = b_full + SBW;
    uint64_t* a_full = b_empty + SBW;
    uint64_t* a_empty = a_full + SLOTS;
    uint64_t* braw_full = a_empty + SLOTS;                            // raw fp32 tiles delivered by the TMA producer
    uint64_t* braw_empty = braw_full + RB;
    uint64_t* araw_full = braw_empty + RB;                            // [2 groups][RA]
    uint64_t* araw_empty = araw_full + 2 * RA;
    uint64_t* accum_bar = araw_empty + 2 * RA;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

    const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
    const int64_t mblk = blockIdx.y, nblk = blockIdx.x;
    const int64_t kb_beg = (int64_t)blockIdx.z * p.kb_per_split;
    const int64_t kb_end = (kb_beg + p.kb_per_split < p.nkb) ? kb_beg + p.kb_per_split : p.nkb;
    const int nkb = (int)(kb_end - kb_beg);
    const int64_t m0 = mblk * PK_AR, n0 = nblk * BN;

    if (tid == 0) {
        for (int s = 0; s < SBW; ++s) {
            mbar_init(&b_full[s], 8);                 // the eight B-converter warps
            mbar_init(&b_empty[s], 1);
        }
        for (int s = 0; s < SLOTS; ++s) {
            mbar_init(&a_full[s], 4);
            mbar_init(&a_empty[s], 1);
        }
        for (int s = 0; s < RB; ++s) {
            mbar_init(&braw_full[s], 1);              // the producer's arrive.expect_tx
            mbar_init(&braw_empty[s], 8);
        }
        for (int s = 0; s < 2 * RA; ++s) {
            mbar_init(&araw_full[s], 1);
            mbar_init(&araw_empty[s], 4);
        }
        mbar_init(accum_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (wid == 1) tmem_alloc_warp(tmem_slot, (uint32_t)p.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (tid == 0) TW_DBG(7, 2);

    if (wid == 0) {
        // ------------------------------ TMA producer: raw fp32 tiles by 2-D tiled loads ----------------------------
        // One instruction per operand and stage: a [16 samples x BN] box of X, a [16 x 128] box of dZ (and of the
        // activation mask).  The converters never touch global memory; tails are zero-filled by the TMA unit.
        // (One bulk copy per sample ROW — 32 to 48 copies of 0.5-0.9 KB per stage — was tried first: ~55 cycles per
        // copy in the TMA unit, 2 700 cycles per stage.)
        if (lane == 0) {
            int rb = 0, ra[2] = {0, 0};
            uint32_t phb = 1, pha[2] = {1, 1};
            for (int i = 0; i < nkb; ++i) {
                const int grp = i & 1;
                const int y = (int)((kb_beg + i) * PK_KB);
                uint64_t* bfull = &braw_full[rb];
                uint64_t* afull = &araw_full[grp * RA + ra[grp]];
                mbar_wait(&braw_empty[rb], phb);
                mbar_expect_tx(bfull, (uint32_t)(PK_KB * BN * 4));
                tma_load_2d(smem_raw + p.off_braw + (size_t)rb * p.b_raw_bytes, &map_b, (int)n0, y, bfull);
                mbar_wait(&araw_empty[grp * RA + ra[grp]], pha[grp]);
                mbar_expect_tx(afull, TW_A_TILE * (p.has_mask ? 2u : 1u));
                unsigned char* dst = smem_raw + p.off_araw + (size_t)(grp * RA + ra[grp]) * p.a_raw_bytes;
                tma_load_2d(dst, &map_a, (int)m0, y, afull);
                if (p.has_mask) tma_load_2d(dst + TW_A_TILE, &map_m, (int)m0, y, afull);
                if (++rb == RB) { rb = 0; phb ^= 1u; }
                if (++ra[grp] == RA) { ra[grp] = 0; pha[grp] ^= 1u; }
            }
        }
        __syncwarp();
    } else if (wid == 1) {
        // ------------------------------ MMA issuer ------------------------------------------------
        const uint32_t idesc = tf32_idesc(BN);
        const uint32_t b_lbo = (uint32_t)BN * 16u;
        int sb = 0;
        uint32_t phb = 0;
        int slot = 0;                                      // counters instead of i % SLOTS, (i / SLOTS) & 1: the divisions
        uint32_t pa = 0;                                   // cost this single thread ~400 cycles of a 1 750-cycle stage
        // one lane runs the whole loop (waits included): no warp-level reconvergence between stages
        if (lane == 0)
        for (int i = 0; i < nkb; ++i) {
            mbar_wait(&b_full[sb], phb);
            TW_DBG(0, i);
            mbar_wait(&a_full[slot], pa);
            TW_DBG(1, i);
            tc_fence_after();
            {
                const uint32_t b_hi = smem_u32(ringB + (size_t)sb * b_stage), b_lo = b_hi + b_stage / 2;
                if (!p.fast) {
#pragma unroll
                    for (int j = 0; j < PK_KB / 8; ++j) {
                        const uint64_t dbh = make_smem_desc(b_hi + (uint32_t)j * 2u * b_lbo, b_lbo, 128);
                        const uint64_t dbl = make_smem_desc(b_lo + (uint32_t)j * 2u * b_lbo, b_lbo, 128);
                        const uint32_t a_hi = tmem_base + (uint32_t)(p.a_col0 + slot * 32 + 8 * j), a_lo = a_hi + 16u;
                        umma_tf32_ts(tmem_base, a_lo, dbh, idesc, (i | j) != 0 ? 1u : 0u);
                        umma_tf32_ts(tmem_base, a_hi, dbl, idesc, 1u);
                        umma_tf32_ts(tmem_base, a_hi, dbh, idesc, 1u);
                    }
                } else {                                  // labelled non-parity mode: single-pass TF32
#pragma unroll
                    for (int j = 0; j < PK_KB / 8; ++j) {
                        const uint64_t dbh = make_smem_desc(b_hi + (uint32_t)j * 2u * b_lbo, b_lbo, 128);
                        umma_tf32_ts(tmem_base, tmem_base + (uint32_t)(p.a_col0 + slot * 32 + 8 * j), dbh, idesc, (i | j) != 0 ? 1u : 0u);
                    }
                }
                umma_commit(&b_empty[sb]);
                umma_commit(&a_empty[slot]);
                if (i == nkb - 1) umma_commit(accum_bar);
                TW_DBG(2, i);
            }
            if (++sb == SBW) { sb = 0; phb ^= 1u; }
            if (++slot == SLOTS) { slot = 0; pa ^= 1u; }
        }
        __syncwarp();
        if (nkb == 0 && lane == 0) mbar_arrive(accum_bar);
    } else if (wid < 10) {
        // ------------------------------ A converters: 2 groups x 4 warps (dZ^T through TMEM) ------
        const int cw = wid - 2, grp = cw >> 2, quad = wid & 3;
        const int t128 = (cw & 3) * 32 + lane;
        const int feat = quad * 32 + lane;                      // the output feature this thread owns = its TMEM lane
        const unsigned char* araw = smem_raw + p.off_araw + (size_t)grp * RA * p.a_raw_bytes;
        const int mask_act = g.amask_act;
        float dbacc = 0.f;
        int ra = 0, slot = grp % SLOTS;
        uint32_t ph_raw = 0, slot_par = 0;                      // parity of the slot's current use (stage / SLOTS)
        for (int i = grp; i < nkb; i += 2) {
            mbar_wait(&araw_full[grp * RA + ra], ph_raw);
            if (t128 == 0) TW_DBG(3, i);
            const float* col = reinterpret_cast<const float*>(araw + (size_t)ra * p.a_raw_bytes) + feat;
            float hi[16], lo[16];
            // samples beyond the batch were zero-filled by the TMA unit; the activation kind is hoisted out of the
            // element loop (three straight-line variants)
            if (!p.has_mask) {
#pragma unroll
                for (int r = 0; r < 16; ++r) {
                    const float v = col[r * TW_A_PITCH];
                    dbacc += v;
                    split_tf32(v, hi[r], lo[r]);
                }
            } else if (mask_act == CTR_ACT_RELU) {
#pragma unroll
                for (int r = 0; r < 16; ++r) {
                    const float v = col[TW_A_TILE / 4 + r * TW_A_PITCH] > 0.f ? col[r * TW_A_PITCH] : 0.f;
                    dbacc += v;
                    split_tf32(v, hi[r], lo[r]);
                }
            } else {
#pragma unroll
                for (int r = 0; r < 16; ++r) {
                    const float v = col[r * TW_A_PITCH] * act_grad_from_y(mask_act, col[TW_A_TILE / 4 + r * TW_A_PITCH]);
                    dbacc += v;
                    split_tf32(v, hi[r], lo[r]);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&araw_empty[grp * RA + ra]);      // the raw tile may be refilled
            mbar_wait(&a_empty[slot], slot_par ^ 1u);
            if (t128 == 0) TW_DBG(4, i);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(p.a_col0 + slot * 32);
            tmem_st16(taddr, hi);
            tmem_st16(taddr + 16u, lo);
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&a_full[slot]);
            if (t128 == 0) TW_DBG(5, i);
            if (++ra == RA) { ra = 0; ph_raw ^= 1u; }
            slot += 2;                                          // next stage of this group: i + 2
            if (slot >= SLOTS) { slot -= SLOTS; slot_par ^= 1u; }
        }
        if (p.db && nblk == 0 && m0 + feat < g.M) atomicAdd(p.db + m0 + feat, dbacc);
    } else {
        // ------------------------------ B converters: 8 warps (X^T into the K-major [hi | lo] tile) -
        // Thread t (< BN) owns COLUMN t of the tile: it reads the 16 samples of its column from the raw tile
        // (conflict-free LDS.32) and writes the four 16-byte chunks (k chunk c, column t) of the MMA tile.  History of
        // this warp group (cycles per 16-sample stage, dW1 of the DeepFM tower): rolled loops with integer divisions
        // 3 400; precomputed offsets, cp.async by the converters themselves 2 060 -> 1 750 (column ownership): ncu showed
        // 4 700 warp-instructions per stage and every phase between two named barriers taking ~550 cycles; with the
        // raw tiles delivered by bulk copies the loop has no global-memory instruction and no CTA-level barrier left.
        const int t256 = (wid - 10) * 32 + lane;
        const bool act = t256 < BN;
        const float* braw_f = reinterpret_cast<const float*>(smem_raw + p.off_braw);
        const int raw_floats = (int)(p.b_raw_bytes >> 2);
        const int pitch = p.b_pitch;
        int sb = 0, rb = 0;
        uint32_t phb = 1, ph_raw = 0;
        for (int i = 0; i < nkb; ++i) {
            mbar_wait(&braw_full[rb], ph_raw);
            if (t256 == 0) TW_DBG(6, i);
            const float* raw = braw_f + rb * raw_floats + t256;
            float4 v[4];
            if (act) {
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const float* cell = raw + 4 * c * pitch;           // 4 samples of column t -> one 16-byte chunk of row t
                    v[c] = make_float4(cell[0], cell[pitch], cell[2 * pitch], cell[3 * pitch]);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&braw_empty[rb]);             // the raw tile may be refilled
            mbar_wait(&b_empty[sb], phb);
            if (t256 == 0) TW_DBG(8, i);
            if (act) {
                float* tile = reinterpret_cast<float*>(ringB + (size_t)sb * b_stage) + t256 * 4;
#pragma unroll
                for (int c = 0; c < 4; ++c) split_store(tile, c * BN * 4, BN * 16, v[c]);
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&b_full[sb]);
            if (t256 == 0) TW_DBG(9, i);
            if (++rb == RB) { rb = 0; ph_raw ^= 1u; }
            if (++sb == SBW) { sb = 0; phb ^= 1u; }
        }
    }
    if (wid >= 2) {
        // ------------------------------ epilogue (16 warps): reduce the partial tile into dW ------
        mbar_wait(accum_bar, 0);
        if (wid == 2 && lane == 0) TW_DBG(7, 0);
        tc_fence_after();
        if (nkb > 0) tile_epilogue<EPI_STORE>(g, tmem_base, smem_raw, wid, lane, m0, 1, BN, n0, true);
        tc_fence_before();
        if (wid == 2 && lane == 0) TW_DBG(7, 1);
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
    int BN, MT, SA, SB, depth, tmem_cols;
    int a_mode, b_mode, t0_b;
    uint32_t raw_a_bytes, raw_b_bytes;
    int64_t gm, gn, splits, nkb, kb_per_split;
    int64_t a_bytes, b_bytes;    // scratch for the packed operands (0 when streamed)
    uint32_t off_b, off_raw, off_bar, smem;
    bool ok;
};

bool pk_al16(const void* p);

// how the converters fetch a streamed operand (see stream_issue)
int stream_mode(const float* P, int64_t s_row, int64_t s_k, const float* mask, int64_t m_row, int64_t m_k) {
    if (s_k == 1 && s_row % 4 == 0 && pk_al16(P) && (!mask || (m_k == 1 && m_row % 4 == 0 && pk_al16(mask)))) return OP_KVEC;
    if (s_row == 1 && s_k % 4 == 0 && pk_al16(P) && (!mask || (m_row == 1 && m_k % 4 == 0 && pk_al16(mask)))) return OP_TRANS;
    return OP_SCALAR;
}

// Operands whose packed image would be larger than this are converted inside the GEMM instead of
// being packed into the scratch first (activations); small ones (weights) are packed once per call
// and fetched by TMA.  CTR_PK_STREAM=0 forces the packed path for everything (A/B profiling).
constexpr int64_t kStreamThresholdBytes = 4 << 20;

PkConfig pk_config(const GemmArgs& g, bool allow_split, bool force_mt1 = false) {
    PkConfig c{};
    const int64_t M = g.M, N = g.N, K = g.K;
    const int64_t ntiles = ceil_div64(N, 256);
    c.BN = (int)(ceil_div64(ceil_div64(N, ntiles), 16) * 16);
    if (c.BN < 16) c.BN = 16;
    c.gn = ceil_div64(N, c.BN);
    c.nkb = ceil_div64(K, PK_KB);
    const int64_t sms = ctr_sm_count();
    // two accumulators per CTA when there is enough work to fill the machine anyway
    c.MT = (ceil_div64(M, 2 * PK_AR) * c.gn >= sms || (allow_split && M > PK_AR)) ? 2 : 1;
    if (M <= PK_AR || force_mt1) c.MT = 1;
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
    // the epilogue reads whole 32-column chunks: the last chunk of the last accumulator must stay inside the
    // TMEM allocation (matters for BN < 32 with two accumulators)
    const int tm_need = (c.MT - 1) * c.BN + 32 * ((c.BN + 31) / 32);
    c.tmem_cols = 32;
    while (c.tmem_cols < c.MT * c.BN || c.tmem_cols < tm_need) c.tmem_cols <<= 1;

    const char* e = getenv("CTR_PK_STREAM");
    const bool stream_ok = !(e && e[0] == '0');
    c.a_bytes = c.gm * c.MT * c.nkb * (int64_t)PK_AR * 128;
    c.b_bytes = c.gn * c.nkb * (int64_t)c.BN * 128;
    c.a_mode = (stream_ok && c.a_bytes > kStreamThresholdBytes) ? stream_mode(g.A, g.sam, g.sak, g.amask, g.smm, g.smk) : OP_PACKED;
    c.b_mode = (stream_ok && c.b_bytes > kStreamThresholdBytes) ? stream_mode(g.B, g.sbn, g.sbk, g.bmask, g.sbmn, g.sbmk) : OP_PACKED;
    if (c.a_mode != OP_PACKED) c.a_bytes = 0;
    if (c.b_mode != OP_PACKED) c.b_bytes = 0;
    // raw cp.async region of a streamed operand per depth: 16 KB of data (+16 KB for its mask):
    // KVEC/SCALAR = 2 slots x 512 threads, TRANS = 4 slots x 256 threads, 16 bytes each
    c.raw_a_bytes = c.a_mode == OP_PACKED ? 0u : (g.amask ? 32768u : 16384u);
    c.raw_b_bytes = c.b_mode == OP_PACKED ? 0u : (g.bmask ? 32768u : 16384u);
    c.t0_b = (c.a_mode == OP_TRANS && c.b_mode == OP_TRANS) ? 256 : 0;

    // shared-memory plan: [A ring | B ring | raw cp.async slots | barriers]
    const int64_t a_stage = (int64_t)c.MT * PK_AR * 128, b_stage = (int64_t)c.BN * 128;
    const int64_t raw_per_depth = (int64_t)c.raw_a_bytes + c.raw_b_bytes;
    const int64_t budget = 232448 - 1024;                  // the 227 KB opt-in maximum minus barriers / slack
    const int64_t stg = (int64_t)PK_CONV_WARPS * 32 * PK_STG_PITCH * 4;       // epilogue staging overlays the rings
    c.ok = false;
    if (raw_per_depth == 0) {
        int S = (int)(budget / (a_stage + b_stage));
        if (S > 6) S = 6;
        while (S >= 2 && (int64_t)S * (a_stage + b_stage) < stg && S < 6) ++S;
        if (S >= 2) {
            c.SA = c.SB = S;
            c.depth = 0;
            c.ok = true;
        }
    } else {
        const int opts[5][3] = {{3, 4, 3}, {3, 3, 3}, {2, 3, 3}, {2, 2, 3}, {2, 2, 2}};
        for (int i = 0; i < 5 && !c.ok; ++i) {
            const int64_t need = opts[i][0] * a_stage + opts[i][1] * b_stage + opts[i][2] * raw_per_depth;
            if (need <= budget) {
                c.SA = opts[i][0];
                c.SB = opts[i][1];
                c.depth = opts[i][2];
                c.ok = true;
            }
        }
    }
    if (c.ok) {
        int64_t rings = c.SA * a_stage + c.SB * b_stage;
        c.off_b = (uint32_t)(c.SA * a_stage);
        if (rings < stg) rings = stg;
        c.off_raw = (uint32_t)rings;
        c.off_bar = (uint32_t)(rings + c.depth * raw_per_depth);
        c.smem = c.off_bar + (uint32_t)((2 * c.SA + 2 * c.SB + 1 + 2 * c.depth) * sizeof(uint64_t) + 16);
        if (c.smem > 232448) c.ok = false;
    }
    // nothing fits with two accumulators (e.g. both operands streamed, one of them masked): one accumulator
    if (!c.ok && c.MT == 2 && !force_mt1) return pk_config(g, allow_split, true);
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
    // layout-independent upper bound: an operand is either streamed (no scratch) or at most
    // kStreamThresholdBytes when packed; with CTR_PK_STREAM=0 both operands are packed in full
    GemmArgs g = gemm_args_default();
    g.M = M; g.N = N; g.K = K;
    int64_t worst = 0;
    for (int sp = 0; sp < 2; ++sp) {
        const PkConfig c = pk_config(g, sp != 0);
        const int64_t full_a = c.gm * c.MT * c.nkb * (int64_t)PK_AR * 128, full_b = c.gn * c.nkb * (int64_t)c.BN * 128;
        const char* e = getenv("CTR_PK_STREAM");
        const bool stream_ok = !(e && e[0] == '0');
        const int64_t a = (stream_ok && full_a > kStreamThresholdBytes) ? 0 : full_a;
        const int64_t b = (stream_ok && full_b > kStreamThresholdBytes) ? 0 : full_b;
        if (a + b > worst) worst = a + b;
    }
    return worst + 256;
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

// helpers for the other tensor-core kernels (cin_tc.cu): the registered scratch and the weight packer
void* gemm_scratch_ptr(int64_t need_bytes) {
    const int dev = current_device();
    return (g_scratch[dev] && g_scratch_bytes[dev] >= need_bytes) ? g_scratch[dev] : nullptr;
}

// pack P(row,k) (n_rows x K, any strides) into [hi | lo] K-major tiles of R rows x 16 k; out must hold
// ceil(n_rows/R) * nkb * R * 128 bytes (nkb >= ceil(K/16) k blocks, the extra ones zero)
int gemm_pack_operand(const float* P, int64_t s_row, int64_t s_k, int64_t n_rows, int64_t K, int R, int64_t nkb, float* out,
                      cudaStream_t st) {
    if (nkb < ceil_div64(K, PK_KB)) nkb = ceil_div64(K, PK_KB);
    PackArgs pa{P, s_row, s_k, nullptr, 0, 0, 0, n_rows, K, R, ceil_div64(n_rows, R), nkb, out};
    return launch_pack(pa, st);
}

// TS engine: eligible when A is a big K-contiguous fp32 matrix (activations) and B is small enough to be packed
// (weights), no split-K.  Returns -3 when the shape does not qualify (the caller continues with the SS engine).
static int launch_gemm_ts(const GemmArgs& g, cudaStream_t st) {
    // opt-in (CTR_GEMM_TS=1): measured within 10 % of the SS engine on the tower shapes (profiles/r02_gemm_engines.md:
    // both are paced by the weight-stage ring, not by shared memory or the MMA rate), so the proven engine stays default
    const char* e = getenv("CTR_GEMM_TS");
    if (!(e && e[0] == '1')) return -3;
    if (g.M < 4 * PK_AR || g.K < 16 || g.N < 16 || g.accumulate) return -3;
    if (stream_mode(g.A, g.sam, g.sak, g.amask, g.smm, g.smk) != OP_KVEC) return -3;
    if (g.bmask) return -3;
    const int64_t gn = ceil_div64(g.N, 256);
    const int BN = (int)(ceil_div64(ceil_div64(g.N, gn), 16) * 16);
    const int MT = (BN <= 128 && g.M >= 2 * PK_AR * ctr_sm_count()) ? 2 : 1;
    const int64_t nkb = ceil_div64(g.K, PK_KB);
    // fewer tiles than SMs: the SS engine splits K across CTAs (and its shorter accumulation chains lose less to the
    // tensor core's truncating fp32 adds: 1.2e-6 vs 1.2e-5 at K = 1664)
    if (ceil_div64(g.M, (int64_t)PK_AR * MT) * gn < ctr_sm_count() && nkb >= 16) return -3;
    const int64_t b_bytes = gn * nkb * (int64_t)BN * 128;
    if (b_bytes > kStreamThresholdBytes) return -3;           // B is not a "small weight" operand
    const int64_t a_packed = ceil_div64(g.M, PK_AR) * nkb * (int64_t)PK_AR * 128;
    if (a_packed <= kStreamThresholdBytes) return -3;         // small A: the all-packed SS path is fine
    const int dev = current_device();
    void* scratch = g_scratch[dev];
    if (!scratch || g_scratch_bytes[dev] < b_bytes) return -3;
    TsParams p{};
    p.g = g;
    p.Bp = reinterpret_cast<const float*>(scratch);
    p.nkb = nkb;
    p.BN = BN;
    p.MT = MT;
    p.has_mask = g.amask ? 1 : 0;
    p.dbg = ctr_debug_buffer();
    p.a_col0 = MT * BN;
    int cols_left = 512 - p.a_col0;
    p.SLOTS = cols_left / (32 * MT);
    if (p.SLOTS > 8) p.SLOTS = 8;
    if (p.SLOTS < 2) return -3;
    p.tmem_cols = 32;
    while (p.tmem_cols < p.a_col0 + MT * p.SLOTS * 32) p.tmem_cols <<= 1;
    p.raw_item_bytes = TS_RAW_TILE * (p.has_mask ? 2u : 1u);
    const int64_t budget = 232448 - 1024;
    const int64_t b_stage = (int64_t)BN * 128;
    const int64_t stg = (int64_t)PK_CONV_WARPS * 32 * PK_STG_PITCH * 4;
    // cluster of CL CTAs (consecutive M tiles, same weight tiles): every CTA fetches 1/CL of a weight stage and
    // multicasts it, so the L2 -> SM traffic of the weights drops CL-fold (run r2-5 timeline: 32 KB per 128 x 16
    // stage from L2 was the pacing item, 3 800 cycles per copy with four in flight).  CTR_TS_CLUSTER=1|2|4.
    {
        static int cl_env = -1;
        if (cl_env < 0) {
            const char* ce = getenv("CTR_TS_CLUSTER");
            cl_env = ce ? atoi(ce) : 1;
            if (cl_env != 1 && cl_env != 2 && cl_env != 4) cl_env = 1;
        }
        p.CL = cl_env;
        while (p.CL > 1 && (b_stage % (16 * p.CL)) != 0) p.CL >>= 1;
    }
    // the weight ring wants depth (the copies' latency is ~4 stages of MMA time), the raw tiles need little: the
    // converters run far ahead of the MMAs anyway
    static int sb_env = -1, rawd_env = -1, raw_env = -1;
    if (sb_env < 0) {
        const char* e1 = getenv("CTR_TS_SB");
        const char* e2 = getenv("CTR_TS_RAWD");
        const char* e6 = getenv("CTR_TS_BRAW");
        sb_env = e1 ? atoi(e1) : 0;
        rawd_env = e2 ? atoi(e2) : 0;
        raw_env = (e6 && e6[0] == '1') ? 1 : 0;
    }
    p.b_raw = raw_env;       // weight stages as raw fp32 (half the L2 -> SM bytes), lo derived by four extra warps
    p.SB = 0;
    for (int rawd = (rawd_env ? rawd_env : 2); rawd >= 1 && !p.SB; --rawd)
        for (int sbn = (sb_env ? sb_env : 6); sbn >= 3 && !p.SB; --sbn) {
            const int64_t need = sbn * b_stage + (int64_t)TS_NG * rawd * p.raw_item_bytes;
            if (need <= budget) {
                p.SB = sbn;
                p.RAWD = rawd;
            }
        }
    if (!p.SB) return -3;
    int64_t rings = p.SB * b_stage;
    p.off_raw = (uint32_t)rings;
    int64_t end = rings + (int64_t)TS_NG * p.RAWD * p.raw_item_bytes;
    if (end < stg) end = stg;
    p.off_bar = (uint32_t)((end + 15) / 16 * 16);
    const size_t smem = p.off_bar + (size_t)(3 * p.SB + 2 * MT * p.SLOTS + 1) * sizeof(uint64_t) + 16;
    if (smem > 232448) return -3;
    int64_t gm = ceil_div64(g.M, (int64_t)PK_AR * MT);
    gm = ceil_div64(gm, p.CL) * p.CL;               // whole clusters: a CTA past the matrix only feeds its peers' weight stages
    if (gm > 65535) return -3;
    PackArgs pb{g.B, g.sbn, g.sbk, nullptr, 0, 0, 0, g.N, g.K, BN, gn, nkb, reinterpret_cast<float*>(scratch)};
    int rc;
    if (p.b_raw) {
        int64_t blocks = ceil_div64(gn * BN * nkb * 4, 256);
        if (blocks > (int64_t)ctr_sm_count() * 8) blocks = (int64_t)ctr_sm_count() * 8;
        pack_raw_kernel<<<(unsigned)blocks, 256, 0, st>>>(pb);
        CTR_LAUNCH_OK("pack_raw_kernel");
    } else if ((rc = launch_pack(pb, st)) != 0) {
        return rc;
    }
    static bool configured = false;
    if (!configured) {
        const int max_smem = 232448;
        CTR_CUDA(cudaFuncSetAttribute(gemm_ts_kernel<EPI_STORE>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
        CTR_CUDA(cudaFuncSetAttribute(gemm_ts_kernel<EPI_BIAS_ACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
        CTR_CUDA(cudaFuncSetAttribute(gemm_ts_kernel<EPI_MUL_ACTGRAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
        CTR_CUDA(cudaFuncSetAttribute(gemm_ts_kernel<EPI_CROSS>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
        CTR_CUDA(cudaFuncSetAttribute(gemm_ts_kernel<EPI_MUL>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
        configured = true;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)gn, (unsigned)gm, 1);
    cfg.blockDim = dim3(TS_THREADS, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 1;
    attr[0].val.clusterDim.y = (unsigned)p.CL;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    switch (g.epilogue) {
        case EPI_BIAS_ACT: CTR_CUDA(cudaLaunchKernelEx(&cfg, gemm_ts_kernel<EPI_BIAS_ACT>, p)); break;
        case EPI_MUL_ACTGRAD: CTR_CUDA(cudaLaunchKernelEx(&cfg, gemm_ts_kernel<EPI_MUL_ACTGRAD>, p)); break;
        case EPI_CROSS: CTR_CUDA(cudaLaunchKernelEx(&cfg, gemm_ts_kernel<EPI_CROSS>, p)); break;
        case EPI_MUL: CTR_CUDA(cudaLaunchKernelEx(&cfg, gemm_ts_kernel<EPI_MUL>, p)); break;
        default: CTR_CUDA(cudaLaunchKernelEx(&cfg, gemm_ts_kernel<EPI_STORE>, p)); break;
    }
    CTR_LAUNCH_OK("gemm_ts_kernel");
    return 0;
}

static bool row_vec_ok_host(const float* base, int64_t ld) { return ((reinterpret_cast<uintptr_t>(base) & 15) == 0) && (ld % 4 == 0); }

// TSW: weight-gradient form  C[M,N] = A^T B  with A(m,k) = A[k*sak + m] (sam == 1), B(n,k) = B[k*sbk + n] (sbn == 1),
// K = batch large, C row-major.  db (may be NULL): column sums of the (masked) A operand, i.e. the bias gradient.
// Returns -3 when the shape does not qualify.
// cuTensorMapEncodeTiled through the runtime's driver entry point query (no link against libcuda)
typedef CUresult (*TensorMapEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                      const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                      CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static TensorMapEncodeFn tensor_map_encoder() {
    static TensorMapEncodeFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<TensorMapEncodeFn>(ptr);
        else
            (void)cudaGetLastError();
    }
    return fn;
}

// row-major fp32 matrix [rows, cols] with a row stride of ld floats, loaded in boxes of [16 rows x box_cols]
static bool make_tile_map(CUtensorMap* map, const float* base, int64_t rows, int64_t cols, int64_t ld, int box_cols,
                          int box_rows = PK_KB, bool swizzle64 = false) {
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
    const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    return tensor_map_encoder()(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

bool gemm_tsw_eligible(const GemmArgs& g) {
    const char* e = getenv("CTR_GEMM_TSW");
    if (e && e[0] == '0') return false;            // CTR_GEMM_TSW=0: SS engine + column-sum kernel (A/B baseline)
    if (g.sam != 1 || g.sbn != 1 || g.K < 4096 || g.M < 16 || g.N < 16 || g.epilogue != EPI_STORE || g.bmask) return false;
    if (g.sak % 4 != 0 || g.sbk % 4 != 0 || !pk_al16(g.A) || !pk_al16(g.B)) return false;
    if (g.amask && (g.smm != 1 || g.smk % 4 != 0 || !pk_al16(g.amask))) return false;
    if (!row_vec_ok_host(g.C, g.ldc)) return false;
    if (g.K > 0x7fffffff || g.M > 0x7fffffff || g.N > 0x7fffffff) return false;      // tensor-map coordinates are int32
    return ceil_div64(g.M, PK_AR) <= 65535 && tensor_map_encoder() != nullptr;
}

int launch_gemm_tsw(const GemmArgs& g, float* db, cudaStream_t st) {
    if (!gemm_tsw_eligible(g)) return -3;
    TwParams p{};
    p.g = g;
    p.db = db;
    p.fast = ctr_gemm_passes() == 1 ? 1 : 0;
    p.dbg = ctr_debug_buffer();
    const int64_t gn = ceil_div64(g.N, 256);
    p.BN = (int)(ceil_div64(ceil_div64(g.N, gn), 16) * 16);
    const int64_t gm = ceil_div64(g.M, PK_AR);
    p.nkb = ceil_div64(g.K, PK_KB);
    const int64_t tiles = gm * gn, sms = ctr_sm_count();
    int64_t splits = tiles >= sms ? 1 : sms / tiles;
    if (splits > p.nkb / 8) splits = p.nkb / 8;
    if (splits < 1) splits = 1;
    p.kb_per_split = ceil_div64(p.nkb, splits);
    splits = ceil_div64(p.nkb, p.kb_per_split);
    if (splits > 65535 || gm > 65535) return -3;
    p.has_mask = g.amask ? 1 : 0;
    p.a_col0 = p.BN;
    p.SLOTS = (512 - p.a_col0) / 32;
    if (p.SLOTS > 8) p.SLOTS = 8;
    p.tmem_cols = 32;
    while (p.tmem_cols < p.a_col0 + p.SLOTS * 32) p.tmem_cols <<= 1;
    p.b_pitch = p.BN;                                   // dense TMA box
    p.b_raw_bytes = (uint32_t)(16 * p.b_pitch * 4);
    p.a_raw_bytes = TW_A_TILE * (p.has_mask ? 2u : 1u);
    const int64_t budget = 232448 - 1024, b_stage = (int64_t)p.BN * 128;
    const int64_t stg = (int64_t)PK_CONV_WARPS * 32 * PK_STG_PITCH * 4;
    p.SBW = 0;
    for (int ra = 3; ra >= 2 && !p.SBW; --ra)
        for (int sbw = 3; sbw >= 2 && !p.SBW; --sbw)
            for (int rb = 4; rb >= 3 && !p.SBW; --rb) {
                const int64_t need = sbw * b_stage + rb * (int64_t)p.b_raw_bytes + 2 * ra * (int64_t)p.a_raw_bytes;
                if (need <= budget) {
                    p.SBW = sbw;
                    p.RA = ra;
                    p.RB = rb;
                }
            }
    if (!p.SBW) return -3;
    int64_t off = p.SBW * b_stage;
    p.off_braw = (uint32_t)off;
    off += (int64_t)p.RB * p.b_raw_bytes;
    p.off_araw = (uint32_t)((off + 127) / 128 * 128);
    off = p.off_araw + 2 * (int64_t)p.RA * p.a_raw_bytes;
    if (off < stg) off = stg;
    p.off_bar = (uint32_t)((off + 15) / 16 * 16);
    const size_t smem = p.off_bar + (size_t)(2 * p.SBW + 2 * p.SLOTS + 2 * p.RB + 4 * p.RA + 1) * sizeof(uint64_t) + 16;
    if (smem > 232448) return -3;
    CUtensorMap map_a, map_m, map_b;
    bool ok = make_tile_map(&map_a, g.A, g.K, g.M, g.sak, PK_AR) && make_tile_map(&map_b, g.B, g.K, g.N, g.sbk, p.BN);
    ok = ok && make_tile_map(&map_m, g.amask ? g.amask : g.A, g.K, g.M, g.amask ? g.smk : g.sak, PK_AR);
    if (!ok) return -3;
    if (!g.accumulate) CTR_CUDA(cudaMemset2DAsync(g.C, g.ldc * sizeof(float), 0, g.N * sizeof(float), g.M, st));
    if (db) CTR_CUDA(cudaMemsetAsync(db, 0, sizeof(float) * g.M, st));
    static bool configured = false;
    if (!configured) {
        CTR_CUDA(cudaFuncSetAttribute(gemm_tsw_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
        configured = true;
    }
    dim3 grid((unsigned)gn, (unsigned)gm, (unsigned)splits);
    gemm_tsw_kernel<0><<<grid, PK_THREADS, smem, st>>>(p, map_a, map_m, map_b);
    CTR_LAUNCH_OK("gemm_tsw_kernel");
    return 0;
}

// PP engine launcher: returns -3 when the shape / layout does not qualify (caller continues with the SS engine)
static int launch_gemm_pp(const GemmArgs& g, cudaStream_t st) {
    // Default: short contractions only (K <= 256), where the exposed prologue / epilogue of the one-tile-per-CTA kernel
    // is a third of a tile: input gradient L1 117.7 -> 98.8 us, L2 48.3 -> 43.6 us.  For longer K the 128-row tiles
    // fetch every weight stage twice as often as the 256-row tiles of the SS engine and the weight ring paces the loop
    // (forward L1 96.5 vs 94.7 us): CTR_GEMM_PP=1 forces the engine for any K, =0 disables it.
    const char* e = getenv("CTR_GEMM_PP");
    if (e && e[0] == '0') return -3;
    // masked A (act'(Y) prologue: a second raw tile per stage) leaves room for only two weight stages: 104 us against
    // 82 us of the SS engine on the top layer's input gradient (profiles/r02_final_deepfm_kernels.md) — SS keeps it
    if (!(e && e[0] == '1') && (g.K > 256 || g.amask)) return -3;
    const int64_t sms = ctr_sm_count();
    if (g.bmask || g.K < 32 || g.N < 16 || g.M < 2 * sms * PK_AR) return -3;       // >= 2 tiles per SM: no split-K needed
    if (g.M > 0x7fffffff || g.K > 0x7fffffff || !tensor_map_encoder()) return -3;
    if (stream_mode(g.A, g.sam, g.sak, g.amask, g.smm, g.smk) != OP_KVEC) return -3;
    PpParams p{};
    p.g = g;
    const int64_t ntn = ceil_div64(g.N, 256);
    p.BN = (int)(ceil_div64(ceil_div64(g.N, ntn), 16) * 16);
    p.gn = ceil_div64(g.N, p.BN);
    p.gm = ceil_div64(g.M, PK_AR);
    p.nkb = ceil_div64(g.K, PK_KB);
    p.has_mask = g.amask ? 1 : 0;
    p.fast = ctr_gemm_passes() == 1 ? 1 : 0;
    p.tmem_cols = 32;
    while (p.tmem_cols < 2 * p.BN) p.tmem_cols <<= 1;
    if (p.tmem_cols > 512) return -3;
    const int64_t b_bytes = p.gn * p.nkb * (int64_t)p.BN * 128;
    if (b_bytes > kStreamThresholdBytes) return -3;    // weights only
    void* scratch = gemm_scratch_ptr(b_bytes);
    if (!scratch) return -3;
    // shared-memory plan: [A ring | B ring | raw ring (1 KB aligned) | epilogue staging | barriers]
    const int64_t a_stage = (int64_t)PK_AR * 128, b_stage = (int64_t)p.BN * 128, raw_item = (int64_t)PK_AR * 64 * (p.has_mask ? 2 : 1);
    const int64_t stg = (int64_t)PP_EPI_WARPS * 32 * PK_STG_PITCH * 4;
    const int64_t budget = 232448 - 1024;
    p.SA = 0;
    const int opts[4][3] = {{4, 4, 4}, {3, 3, 4}, {3, 3, 3}, {3, 2, 3}};
    for (int i = 0; i < 4 && !p.SA; ++i) {
        const int64_t rings = (opts[i][0] * a_stage + opts[i][1] * b_stage + 1023) / 1024 * 1024;
        if (rings + opts[i][2] * raw_item + stg <= budget) {
            p.SA = opts[i][0];
            p.SB = opts[i][1];
            p.RD = opts[i][2];
        }
    }
    if (!p.SA) return -3;
    p.off_b = (uint32_t)(p.SA * a_stage);
    p.off_raw = (uint32_t)((p.SA * a_stage + p.SB * b_stage + 1023) / 1024 * 1024);
    p.off_stg = (uint32_t)(p.off_raw + p.RD * raw_item);
    p.off_bar = (uint32_t)(p.off_stg + stg);
    const size_t smem = p.off_bar + (size_t)(2 * p.SA + 2 * p.SB + 2 * p.RD + 4) * sizeof(uint64_t) + 16;
    if (smem > 232448) return -3;
    CUtensorMap map_a{}, map_am{};
    bool ok = make_tile_map(&map_a, g.A, g.M, g.K, g.sam, PK_KB, PK_AR, true);
    if (ok && g.amask) ok = make_tile_map(&map_am, g.amask, g.M, g.K, g.smm, PK_KB, PK_AR, true);
    if (!ok) return -3;
    float* Bp = reinterpret_cast<float*>(scratch);
    PackArgs pb{g.B, g.sbn, g.sbk, nullptr, 0, 0, 0, g.N, g.K, p.BN, p.gn, p.nkb, Bp};
    int rc;
    if ((rc = launch_pack(pb, st)) != 0) return rc;
    p.Bp = Bp;
    static bool configured = false;
    if (!configured) {
        const int max_smem = 232448;
        CTR_CUDA(cudaFuncSetAttribute(gemm_pp_kernel<EPI_STORE>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
        CTR_CUDA(cudaFuncSetAttribute(gemm_pp_kernel<EPI_BIAS_ACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
        CTR_CUDA(cudaFuncSetAttribute(gemm_pp_kernel<EPI_MUL_ACTGRAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
        CTR_CUDA(cudaFuncSetAttribute(gemm_pp_kernel<EPI_CROSS>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
        CTR_CUDA(cudaFuncSetAttribute(gemm_pp_kernel<EPI_MUL>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
        configured = true;
    }
    const int64_t tiles = p.gm * p.gn;
    const unsigned grid = (unsigned)(tiles < sms ? tiles : sms);
    switch (g.epilogue) {
        case EPI_BIAS_ACT: gemm_pp_kernel<EPI_BIAS_ACT><<<grid, PK_THREADS, smem, st>>>(p, map_a, map_am); break;
        case EPI_MUL_ACTGRAD: gemm_pp_kernel<EPI_MUL_ACTGRAD><<<grid, PK_THREADS, smem, st>>>(p, map_a, map_am); break;
        case EPI_CROSS: gemm_pp_kernel<EPI_CROSS><<<grid, PK_THREADS, smem, st>>>(p, map_a, map_am); break;
        case EPI_MUL: gemm_pp_kernel<EPI_MUL><<<grid, PK_THREADS, smem, st>>>(p, map_a, map_am); break;
        default: gemm_pp_kernel<EPI_STORE><<<grid, PK_THREADS, smem, st>>>(p, map_a, map_am); break;
    }
    CTR_LAUNCH_OK("gemm_pp_kernel");
    return 0;
}

int launch_gemm_pk(const GemmArgs& g, cudaStream_t st) {
    {
        const int rc_pp = launch_gemm_pp(g, st);
        if (rc_pp != -3) return rc_pp;
    }
    {
        const int rc_ts = launch_gemm_ts(g, st);
        if (rc_ts != -3) return rc_ts;
    }
    if (g.allow_split_k) {          // long-K products of two batch-major operands (bilinear dW, ...): the TSW engine
        const int rc_tw = launch_gemm_tsw(g, nullptr, st);
        if (rc_tw != -3) return rc_tw;
    }

    const bool allow_split = g.allow_split_k && g.epilogue == EPI_STORE;
    const PkConfig c = pk_config(g, allow_split);
    if (!c.ok) return -3;     // no shared-memory plan for this shape: the caller uses the first-generation engine
    const int64_t need = c.a_bytes + c.b_bytes;
    const int dev = current_device();
    void* scratch = g_scratch[dev];
    if (need > 0 && (!scratch || g_scratch_bytes[dev] < need)) {
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
    if (c.a_mode == OP_PACKED) {
        PackArgs pa{g.A, g.sam, g.sak, g.amask, g.smm, g.smk, g.amask_act, g.M, g.K, PK_AR, c.gm * c.MT, c.nkb, Ap};
        if ((rc = launch_pack(pa, st)) != 0) return rc;
    }
    if (c.b_mode == OP_PACKED) {
        PackArgs pb{g.B, g.sbn, g.sbk, g.bmask, g.sbmn, g.sbmk, g.bmask_act, g.N, g.K, c.BN, c.gn, c.nkb, Bp};
        if ((rc = launch_pack(pb, st)) != 0) return rc;
    }
    if (c.splits > 1 && !g.accumulate)
        CTR_CUDA(cudaMemset2DAsync(g.C, g.ldc * sizeof(float), 0, g.N * sizeof(float), g.M, st));
    PkParams p{};
    p.g = g;
    p.fast = ctr_gemm_passes() == 1 ? 1 : 0;
    p.dbg = ctr_debug_buffer();
    p.Ap = Ap;
    p.Bp = Bp;
    p.sa = StreamOp{g.A, g.sam, g.sak, g.amask, g.smm, g.smk, g.amask_act, g.M, c.a_mode};
    p.sb = StreamOp{g.B, g.sbn, g.sbk, g.bmask, g.sbmn, g.sbmk, g.bmask_act, g.N, c.b_mode};
    p.nkb = c.nkb;
    p.kb_per_split = c.kb_per_split;
    p.BN = c.BN;
    p.MT = c.MT;
    p.SA = c.SA;
    p.SB = c.SB;
    p.depth = c.depth;
    p.tmem_cols = c.tmem_cols;
    p.off_b = c.off_b;
    p.off_raw = c.off_raw;
    p.off_bar = c.off_bar;
    p.raw_a_bytes = c.raw_a_bytes;
    p.raw_b_bytes = c.raw_b_bytes;
    p.t0_b = c.t0_b;
    // K-contiguous streamed A: raw tiles by 2-D tiled TMA loads (CTR_PK_ATMA=0: the converters' own cp.async pieces)
    CUtensorMap map_a{}, map_am{};
    {
        const char* e = getenv("CTR_PK_ATMA");
        if (c.a_mode == OP_KVEC && !(e && e[0] == '0') && tensor_map_encoder() && g.M <= 0x7fffffff && g.K <= 0x7fffffff &&
            c.off_raw % 1024 == 0 && (c.raw_a_bytes + c.raw_b_bytes) % 1024 == 0) {
            bool ok = make_tile_map(&map_a, g.A, g.M, g.K, g.sam, PK_KB, c.MT * PK_AR, true);
            if (ok && g.amask) ok = make_tile_map(&map_am, g.amask, g.M, g.K, g.smm, PK_KB, c.MT * PK_AR, true);
            p.a_tma = ok ? 1 : 0;
        }
    }
    const size_t smem = c.smem;
    static bool configured = false;
    if (!configured) {
        const int max_smem = 232448;
        CTR_CUDA(cudaFuncSetAttribute(gemm_pk_kernel<EPI_STORE>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
        CTR_CUDA(cudaFuncSetAttribute(gemm_pk_kernel<EPI_BIAS_ACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
        CTR_CUDA(cudaFuncSetAttribute(gemm_pk_kernel<EPI_MUL_ACTGRAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
        CTR_CUDA(cudaFuncSetAttribute(gemm_pk_kernel<EPI_CROSS>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
        CTR_CUDA(cudaFuncSetAttribute(gemm_pk_kernel<EPI_MUL>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
        configured = true;
    }
    dim3 grid((unsigned)c.gn, (unsigned)c.gm, (unsigned)c.splits);
    switch (g.epilogue) {
        case EPI_BIAS_ACT: gemm_pk_kernel<EPI_BIAS_ACT><<<grid, PK_THREADS, smem, st>>>(p, map_a, map_am); break;
        case EPI_MUL_ACTGRAD: gemm_pk_kernel<EPI_MUL_ACTGRAD><<<grid, PK_THREADS, smem, st>>>(p, map_a, map_am); break;
        case EPI_CROSS: gemm_pk_kernel<EPI_CROSS><<<grid, PK_THREADS, smem, st>>>(p, map_a, map_am); break;
        case EPI_MUL: gemm_pk_kernel<EPI_MUL><<<grid, PK_THREADS, smem, st>>>(p, map_a, map_am); break;
        default: gemm_pk_kernel<EPI_STORE><<<grid, PK_THREADS, smem, st>>>(p, map_a, map_am); break;
    }
    CTR_LAUNCH_OK("gemm_pk_kernel");
    return 0;
}
