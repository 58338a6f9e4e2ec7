// Fused multi-slot embedding gather (+ linear term + FM + dnn_input assembly) and its backward.
//
// Replaces, in ONE pass over the batch, what the reference does with 52 nn.Embedding calls,
// four torch.cat's and the FM element-wise ops (reference models/basemodel.py:354-380,63-92;
// inputs.py:126-138; layers/interaction.py:26-34).  HBM-bound: per sample it reads the X row
// (C floats), F rows of D floats, F linear weights and writes the [F*D + n_dense] block once.
//
// Work decomposition: one warp per sample (grid-stride).  A row of D floats is moved by
// LPR = D/4 lanes, each with one 128-bit load/store, so a warp moves 32/LPR rows per step.
#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr int kMaxSmemSlots = 256;  // slot metadata staged in shared memory up to this many fields

__device__ __forceinline__ int decode_id(float xv, int vocab, int32_t* err_flag, int id_mode) {
    // id_mode 0 (reference): fp32 -> integer by truncation, exactly like `.long()` (reference
    // basemodel.py:369), exact only below 2^24.  id_mode 1 (CTR_IDS_I32BITS): the 4-byte cell holds the
    // int32 id itself (bit pattern), no 2^24 limit.  ids live in [0, vocab) with vocab < 2^31:
    // out-of-range / negative / huge values trip the unsigned compare
    int id = id_mode ? __float_as_int(xv) : __float2int_rz(xv);
    if ((unsigned)id >= (unsigned)vocab) {
        if (err_flag) atomicOr(err_flag, 1);
        id = 0;
    }
    return id;
}

struct GatherArgs {
    const float* X;
    int64_t ldx;
    int64_t B;
    int n_emb, D;
    const float* const* emb_tables;
    const int32_t* emb_cols;
    const int32_t* emb_vocab;
    int n_lin;
    const float* const* lin_tables;
    const int32_t* lin_cols;
    const int32_t* lin_vocab;
    int n_dense;
    const int32_t* dense_cols;
    int n_lin_dense;
    const int32_t* lin_dense_cols;
    const float* lin_dense_w;
    float* blk;
    int64_t ld_blk;
    float* lin;
    float* fm;
    int32_t* err_flag;
    int n_shards;   // tables row-sharded: pointer [f*n_shards + id % n_shards], row id / n_shards
    // exchanged mode (p2p.cu): rows of other shards were delivered into local response buffers;
    // where[b, plan col] = -1 (row is local) or (owner << 26) | slot
    const int32_t* where; int n_plan; int me;
    const int32_t* emb_plan_col; const int32_t* lin_plan_col;
    const float* const* resp_emb; const float* const* resp_lin;
    int id_mode;
};

// ---------------------------------------------------------------------------------------------
// forward, vector path: D % 4 == 0 and D/4 a power of two <= 32
// ---------------------------------------------------------------------------------------------
// Memory-level parallelism is what bounds this kernel: per sample the chain is X row -> ids ->
// rows.  A warp works on SPW samples at once: all their ids are fetched first, then all their row
// loads (SPW x 4 x 128 bit per lane) are issued before the first store, so ~2 * F rows per warp are
// in flight at the same time.
// STEPS: row loads in flight per lane and sample; XCH: exchanged mode compiled in (kept out of the
// single-GPU instantiation: its extra registers cost the local gather 25 % — run 23)
template <int LPR, int SPW, int STEPS = 4, bool XCH = false>
__global__ void __launch_bounds__(256) gather_fwd_vec_kernel(GatherArgs a) {
    constexpr int RPW = 32 / LPR;  // rows per warp step
    constexpr int D = LPR * 4;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // stage slot metadata (table pointers, column, vocab) in shared memory
    const float** s_tab = reinterpret_cast<const float**>(smem_raw);
    const int G = a.n_shards;
    int32_t* s_col = reinterpret_cast<int32_t*>(s_tab + a.n_emb * G);
    int32_t* s_voc = s_col + a.n_emb;
    int32_t* s_pc = s_voc + a.n_emb;            // plan column of each embedding slot (exchanged mode)
    for (int i = threadIdx.x; i < a.n_emb * G; i += blockDim.x) s_tab[i] = a.emb_tables[i];
    for (int i = threadIdx.x; i < a.n_emb; i += blockDim.x) {
        s_col[i] = a.emb_cols[i];
        s_voc[i] = a.emb_vocab[i];
        s_pc[i] = (XCH && a.where) ? a.emb_plan_col[i] : 0;
    }
    __syncthreads();

    const int lane = threadIdx.x & 31;
    const int sub = lane % LPR;
    const int rslot = lane / LPR;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;

    for (int64_t bb = warp0 * SPW; bb < a.B; bb += nwarps * SPW) {
        float4 S[SPW];
        float q[SPW];
#pragma unroll
        for (int t = 0; t < SPW; ++t) {
            S[t] = make_float4(0.f, 0.f, 0.f, 0.f);
            q[t] = 0.f;
        }
        for (int f0 = 0; f0 < a.n_emb; f0 += STEPS * RPW) {
            // every load below is UNCONDITIONAL (sample / field indices are clamped to a valid element and
            // the result is simply not used when out of range): a load inside a data-dependent branch is
            // waited for at the join, which would serialise the SPW x STEPS row loads of a lane
            float xr[SPW][STEPS];
#pragma unroll
            for (int t = 0; t < SPW; ++t) {
                const int64_t bc = (bb + t < a.B) ? bb + t : a.B - 1;
#pragma unroll
                for (int s = 0; s < STEPS; ++s) {
                    const int f = f0 + s * RPW + rslot;
                    xr[t][s] = __ldg(a.X + bc * a.ldx + s_col[f < a.n_emb ? f : a.n_emb - 1]);
                }
            }
            float4 v[SPW][STEPS];
#pragma unroll
            for (int t = 0; t < SPW; ++t) {
#pragma unroll
                for (int s = 0; s < STEPS; ++s) {
                    const int f = f0 + s * RPW + rslot;
                    const int fc = f < a.n_emb ? f : a.n_emb - 1;
                    const int id = decode_id(xr[t][s], s_voc[fc], a.err_flag, a.id_mode);
                    const float* tab;
                    int row;
                    if (G == 1) {
                        tab = s_tab[fc];
                        row = id;
                    } else if (XCH && a.where) {
                        const int64_t bc2 = (bb + t < a.B) ? bb + t : a.B - 1;
                        const int w = __ldg(a.where + bc2 * a.n_plan + s_pc[fc]);
                        if (w < 0) {
                            tab = s_tab[fc * G + a.me];
                            row = id / G;
                        } else {                               // delivered by the owner into a LOCAL buffer
                            tab = a.resp_emb[w >> 26];
                            row = w & ((1 << 26) - 1);
                        }
                    } else {
                        tab = s_tab[fc * G + id % G];
                        row = id / G;
                    }
                    // direct peer (NVLink-mapped) rows: plain coherent loads; everything local streams past L1
                    v[t][s] = (G == 1 || (XCH && a.where)) ? ld_stream4(tab + (size_t)row * D + sub * 4)
                                                  : *reinterpret_cast<const float4*>(tab + (size_t)row * D + sub * 4);
                }
            }
#pragma unroll
            for (int t = 0; t < SPW; ++t) {
                const int64_t b = bb + t;
#pragma unroll
                for (int s = 0; s < STEPS; ++s) {
                    const int f = f0 + s * RPW + rslot;
                    if (b < a.B && f < a.n_emb) {
                        if (a.blk) st_stream4(a.blk + b * a.ld_blk + (int64_t)f * D + sub * 4, v[t][s]);
                        S[t].x += v[t][s].x; S[t].y += v[t][s].y; S[t].z += v[t][s].z; S[t].w += v[t][s].w;
                        q[t] += v[t][s].x * v[t][s].x + v[t][s].y * v[t][s].y + v[t][s].z * v[t][s].z +
                                v[t][s].w * v[t][s].w;
                    }
                }
            }
        }
#pragma unroll
        for (int t = 0; t < SPW; ++t) {
            const int64_t b = bb + t;
            if (b >= a.B) break;                      // warp-uniform
            const float* xrow = a.X + b * a.ldx;
            float fmv = 0.f;
            if (a.fm) {
                float4 Ss = S[t];
                // sum S over the lanes that hold the same quad of d (stride LPR), q over all lanes
#pragma unroll
                for (int o = LPR; o < 32; o <<= 1) {
                    Ss.x += __shfl_xor_sync(0xffffffffu, Ss.x, o);
                    Ss.y += __shfl_xor_sync(0xffffffffu, Ss.y, o);
                    Ss.z += __shfl_xor_sync(0xffffffffu, Ss.z, o);
                    Ss.w += __shfl_xor_sync(0xffffffffu, Ss.w, o);
                }
                float tt = Ss.x * Ss.x + Ss.y * Ss.y + Ss.z * Ss.z + Ss.w * Ss.w;
#pragma unroll
                for (int o = 1; o < LPR; o <<= 1) tt += __shfl_xor_sync(0xffffffffu, tt, o);
                fmv = 0.5f * (tt - warp_sum(q[t]));
            }
            // linear term: sparse weights + dense dot; dense copy (and zero padding) into the block
            float lp = 0.f;
            for (int f = lane; f < a.n_lin; f += 32) {
                const int id = decode_id(__ldg(xrow + a.lin_cols[f]), a.lin_vocab[f], a.err_flag, a.id_mode);
                if (G == 1) {
                    lp += __ldg(a.lin_tables[f] + id);
                } else if (XCH && a.where) {
                    const int w = __ldg(a.where + b * a.n_plan + a.lin_plan_col[f]);
                    lp += (w < 0) ? __ldg(a.lin_tables[f * G + a.me] + id / G) : __ldg(a.resp_lin[w >> 26] + (w & ((1 << 26) - 1)));
                } else {
                    lp += a.lin_tables[f * G + id % G][id / G];
                }
            }
            for (int k = lane; k < a.n_lin_dense; k += 32)
                lp += __ldg(xrow + a.lin_dense_cols[k]) * __ldg(a.lin_dense_w + k);
            if (a.blk) {
                float* drow = a.blk + b * a.ld_blk + (int64_t)a.n_emb * D;
                const int n_pad = (int)(a.ld_blk - (int64_t)a.n_emb * D);  // dense columns, then zeros up to ld
                for (int k = lane; k < n_pad; k += 32) drow[k] = (k < a.n_dense) ? __ldg(xrow + a.dense_cols[k]) : 0.f;
            }
            if (a.lin) {
                lp = warp_sum(lp);
                if (lane == 0) a.lin[b] = lp;
            }
            if (a.fm && lane == 0) a.fm[b] = fmv;
        }
    }
}

// forward, generic path: any D (scalar loads); FM is computed by fm_fwd_kernel afterwards
__global__ void __launch_bounds__(256) gather_fwd_generic_kernel(GatherArgs a) {
    const int lane = threadIdx.x & 31;
    const int D = a.D;
    const int G = a.n_shards;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t b = warp0; b < a.B; b += nwarps) {
        const float* xrow = a.X + b * a.ldx;
        if (a.blk) {
            const int total = a.n_emb * D;
            for (int i = lane; i < total; i += 32) {
                const int f = i / D, d = i - f * D;
                const int64_t id = decode_id(__ldg(xrow + a.emb_cols[f]), a.emb_vocab[f], a.err_flag, a.id_mode);
                const float* src = a.emb_tables[f * G + (int)(id % G)] + (id / G) * D + d;
                a.blk[b * a.ld_blk + i] = (G == 1) ? __ldg(src) : *src;      // peer rows: plain loads
            }
            float* drow = a.blk + b * a.ld_blk + (int64_t)total;
            const int n_pad = (int)(a.ld_blk - total);
            for (int k = lane; k < n_pad; k += 32) drow[k] = (k < a.n_dense) ? __ldg(xrow + a.dense_cols[k]) : 0.f;
        }
        float lp = 0.f;
        for (int f = lane; f < a.n_lin; f += 32) {
            const int64_t id = decode_id(__ldg(xrow + a.lin_cols[f]), a.lin_vocab[f], a.err_flag, a.id_mode);
            const float* lsrc = a.lin_tables[f * G + (int)(id % G)] + id / G;
            lp += (G == 1) ? __ldg(lsrc) : *lsrc;
        }
        for (int k = lane; k < a.n_lin_dense; k += 32)
            lp += __ldg(xrow + a.lin_dense_cols[k]) * __ldg(a.lin_dense_w + k);
        if (a.lin) {
            lp = warp_sum(lp);
            if (lane == 0) a.lin[b] = lp;
        }
    }
}

// FM over an assembled block: one warp per sample, lanes over d
__global__ void __launch_bounds__(256) fm_fwd_kernel(const float* __restrict__ blk, int64_t ld,
                                                     int64_t B, int F, int D, float* fm) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t b = warp0; b < B; b += nwarps) {
        const float* row = blk + b * ld;
        float acc = 0.f;
        for (int d = lane; d < D; d += 32) {
            float s = 0.f, q = 0.f;
            for (int f = 0; f < F; ++f) {
                const float v = row[f * D + d];
                s += v;
                q += v * v;
            }
            acc += s * s - q;
        }
        acc = warp_sum(acc);
        if (lane == 0) fm[b] = 0.5f * acc;
    }
}

__global__ void __launch_bounds__(256) fm_bwd_kernel(const float* __restrict__ blk, int64_t ld,
                                                     int64_t B, int F, int D,
                                                     const float* __restrict__ g, float* d_blk,
                                                     int64_t ld_d) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t b = warp0; b < B; b += nwarps) {
        const float* row = blk + b * ld;
        float* drow = d_blk + b * ld_d;
        const float gb = g[b];
        for (int d = lane; d < D; d += 32) {
            float s = 0.f;
            for (int f = 0; f < F; ++f) s += row[f * D + d];
            for (int f = 0; f < F; ++f) drow[f * D + d] += gb * (s - row[f * D + d]);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------
struct ScatterArgs {
    const float* X;
    int64_t ldx;
    int64_t B;
    int n_emb, D;
    float* const* emb_out;         // dense: grad tables [V,D] (device pointer array)
    float* emb_rg;                 // rowwise: row-grad buffers [n_emb][B,D] at emb_rg + f*emb_rg_stride
    int64_t emb_rg_stride;
    float* lin_rg;                 // rowwise: [n_lin][B] at lin_rg + f*lin_rg_stride
    int64_t lin_rg_stride;
    const int32_t* emb_cols;       // dense: X column; rowwise: plan column
    const int32_t* emb_vocab;
    int n_lin;
    float* const* lin_out;
    const int32_t* lin_cols;
    const int32_t* lin_vocab;
    const float* blk;
    int64_t ld_blk;
    const float* d_blk;
    int64_t ld_dblk;
    const float* g_fm;
    const float* g_lin;
    // rowwise only
    int n_plan;
    const int32_t* inv;
    const int32_t* cnt;
    int id_mode;
};

// A warp works on SPW samples at once.  r[b,f,:] = d_blk + g_fm (S - E) is added to its destination
// row: dense mode -> red.global.add.v4.f32 into [V,D]; rowwise mode -> plain 128-bit store when the
// id is unique in the batch (cnt == 1), vector reduction otherwise.  All loads are unconditional
// (clamped indices) and issued before the first dependent instruction; only stores are predicated.
template <int LPR, bool ROWWISE, int SPW>
__global__ void __launch_bounds__(128) scatter_bwd_vec_kernel(ScatterArgs a) {
    constexpr int RPW = 32 / LPR;
    constexpr int STEPS = 4;
    constexpr int D = LPR * 4;
    const int lane = threadIdx.x & 31;
    const int sub = lane % LPR;
    const int rslot = lane / LPR;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);

    for (int64_t bb = warp0 * SPW; bb < a.B; bb += nwarps * SPW) {
        int64_t bc[SPW];
        float gfm[SPW];
        float4 S[SPW];
#pragma unroll
        for (int t = 0; t < SPW; ++t) {
            bc[t] = (bb + t < a.B) ? bb + t : a.B - 1;
            gfm[t] = a.g_fm ? __ldg(a.g_fm + bc[t]) : 0.f;
            S[t] = zero4;
        }
        float4 e0[SPW][STEPS];         // blk rows of the first chunk (reused in pass 2)
        if (a.g_fm) {
            for (int f0 = 0; f0 < a.n_emb; f0 += STEPS * RPW) {
                float4 v[SPW][STEPS];
#pragma unroll
                for (int t = 0; t < SPW; ++t)
#pragma unroll
                    for (int s = 0; s < STEPS; ++s) {
                        const int f = f0 + s * RPW + rslot;
                        const int fc = f < a.n_emb ? f : a.n_emb - 1;
                        v[t][s] = ld_stream4(a.blk + bc[t] * a.ld_blk + (int64_t)fc * D + sub * 4);
                    }
#pragma unroll
                for (int t = 0; t < SPW; ++t)
#pragma unroll
                    for (int s = 0; s < STEPS; ++s) {
                        const bool live = f0 + s * RPW + rslot < a.n_emb;
                        S[t].x += live ? v[t][s].x : 0.f; S[t].y += live ? v[t][s].y : 0.f;
                        S[t].z += live ? v[t][s].z : 0.f; S[t].w += live ? v[t][s].w : 0.f;
                        if (f0 == 0) e0[t][s] = v[t][s];
                    }
            }
#pragma unroll
            for (int t = 0; t < SPW; ++t)
#pragma unroll
                for (int o = LPR; o < 32; o <<= 1) {
                    S[t].x += __shfl_xor_sync(0xffffffffu, S[t].x, o);
                    S[t].y += __shfl_xor_sync(0xffffffffu, S[t].y, o);
                    S[t].z += __shfl_xor_sync(0xffffffffu, S[t].z, o);
                    S[t].w += __shfl_xor_sync(0xffffffffu, S[t].w, o);
                }
        }
        for (int f0 = 0; f0 < a.n_emb; f0 += STEPS * RPW) {
            float4 r[SPW][STEPS];
            int u[SPW][STEPS];
            int c[SPW][STEPS];
#pragma unroll
            for (int t = 0; t < SPW; ++t)
#pragma unroll
                for (int s = 0; s < STEPS; ++s) {
                    const int f = f0 + s * RPW + rslot;
                    const int fc = f < a.n_emb ? f : a.n_emb - 1;
                    r[t][s] = a.d_blk ? ld_stream4(a.d_blk + bc[t] * a.ld_dblk + (int64_t)fc * D + sub * 4) : zero4;
                    if (ROWWISE) u[t][s] = __ldg(a.inv + bc[t] * a.n_plan + a.emb_cols[fc]);
                    else u[t][s] = decode_id(__ldg(a.X + bc[t] * a.ldx + a.emb_cols[fc]), a.emb_vocab[fc], nullptr, a.id_mode);
                }
            if (ROWWISE) {
#pragma unroll
                for (int t = 0; t < SPW; ++t)
#pragma unroll
                    for (int s = 0; s < STEPS; ++s) {
                        const int f = f0 + s * RPW + rslot;
                        const int fc = f < a.n_emb ? f : a.n_emb - 1;
                        c[t][s] = __ldg(a.cnt + (int64_t)a.emb_cols[fc] * a.B + u[t][s]);
                    }
            }
#pragma unroll
            for (int t = 0; t < SPW; ++t)
#pragma unroll
                for (int s = 0; s < STEPS; ++s) {
                    const int f = f0 + s * RPW + rslot;
                    if (f < a.n_emb && bb + t < a.B) {
                        float4 rr = r[t][s];
                        if (a.g_fm) {
                            const float4 v = (f0 == 0) ? e0[t][s]
                                                       : ld_stream4(a.blk + bc[t] * a.ld_blk + (int64_t)f * D + sub * 4);
                            rr.x += gfm[t] * (S[t].x - v.x);
                            rr.y += gfm[t] * (S[t].y - v.y);
                            rr.z += gfm[t] * (S[t].z - v.z);
                            rr.w += gfm[t] * (S[t].w - v.w);
                        }
                        if (ROWWISE) {
                            float* dst = a.emb_rg + f * a.emb_rg_stride + (int64_t)u[t][s] * D + sub * 4;
                            if (c[t][s] == 1) st_stream4(dst, rr);
                            else red_add4(dst, rr);
                        } else {
                            red_add4(a.emb_out[f] + (size_t)u[t][s] * D + sub * 4, rr);
                        }
                    }
                }
        }
        if (a.g_lin) {
#pragma unroll
            for (int t = 0; t < SPW; ++t) {
                if (bb + t >= a.B) break;
                const float gl = __ldg(a.g_lin + bc[t]);
                for (int f = lane; f < a.n_lin; f += 32) {
                    if (ROWWISE) {
                        const int pc = a.lin_cols[f];
                        const int uu = __ldg(a.inv + bc[t] * a.n_plan + pc);
                        const int cc = __ldg(a.cnt + (int64_t)pc * a.B + uu);
                        float* dst = a.lin_rg + f * a.lin_rg_stride + uu;
                        if (cc == 1) *dst = gl;
                        else atomicAdd(dst, gl);
                    } else {
                        const int id = decode_id(__ldg(a.X + bc[t] * a.ldx + a.lin_cols[f]), a.lin_vocab[f], nullptr, a.id_mode);
                        atomicAdd(a.lin_out[f] + id, gl);
                    }
                }
            }
        }
    }
}

// generic (any D) backward: scalar atomics
template <bool ROWWISE>
__global__ void __launch_bounds__(256) scatter_bwd_generic_kernel(ScatterArgs a) {
    const int lane = threadIdx.x & 31;
    const int D = a.D;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t b = warp0; b < a.B; b += nwarps) {
        const float* xrow = a.X ? a.X + b * a.ldx : nullptr;
        const float gfm = a.g_fm ? a.g_fm[b] : 0.f;
        const int total = a.n_emb * D;
        for (int i = lane; i < total; i += 32) {
            const int f = i / D, d = i - f * D;
            float r = a.d_blk ? a.d_blk[b * a.ld_dblk + i] : 0.f;
            if (a.g_fm) {
                float s = 0.f;
                for (int f2 = 0; f2 < a.n_emb; ++f2) s += a.blk[b * a.ld_blk + f2 * D + d];
                r += gfm * (s - a.blk[b * a.ld_blk + i]);
            }
            if (ROWWISE) {
                const int pc = a.emb_cols[f];
                const int u = a.inv[b * a.n_plan + pc];
                atomicAdd(a.emb_rg + f * a.emb_rg_stride + (int64_t)u * D + d, r);
            } else {
                const int64_t id = decode_id(xrow[a.emb_cols[f]], a.emb_vocab[f], nullptr, a.id_mode);
                atomicAdd(a.emb_out[f] + id * D + d, r);
            }
        }
        if (a.g_lin) {
            const float gl = a.g_lin[b];
            for (int f = lane; f < a.n_lin; f += 32) {
                if (ROWWISE) {
                    const int u = a.inv[b * a.n_plan + a.lin_cols[f]];
                    atomicAdd(a.lin_rg + f * a.lin_rg_stride + u, gl);
                } else {
                    const int64_t id = decode_id(xrow[a.lin_cols[f]], a.lin_vocab[f], nullptr, a.id_mode);
                    atomicAdd(a.lin_out[f] + id, gl);
                }
            }
        }
    }
}

// rowwise prep: zero every destination row that will not be written by exactly one plain store
// (cnt == 0: padding beyond n_uniq; cnt > 1: accumulated with reductions).  force_all zeroes
// everything (generic path uses atomics only).
__global__ void __launch_bounds__(256) rowgrad_prep_kernel(int64_t B, const int32_t* __restrict__ cnt,
                                                           int n_emb, int D, float* emb_rg,
                                                           int64_t emb_rg_stride,
                                                           const int32_t* emb_plan, int n_lin,
                                                           float* lin_rg, int64_t lin_rg_stride,
                                                           const int32_t* lin_plan, int force_all) {
    // one thread per (field, unique slot): a single cnt lookup decides the whole row
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
    const int64_t total = (int64_t)(n_emb + n_lin) * B;
    const bool vec = (D % 4 == 0) && (emb_rg_stride % 4 == 0);
    for (int64_t i = tid; i < total; i += nthreads) {
        const int f = (int)(i / B);
        const int64_t u = i - (int64_t)f * B;
        if (f < n_emb) {
            if (force_all || __ldg(cnt + (int64_t)emb_plan[f] * B + u) != 1) {
                float* row = emb_rg + f * emb_rg_stride + u * D;
                if (vec) {
                    for (int d = 0; d < D; d += 4) *reinterpret_cast<float4*>(row + d) = make_float4(0.f, 0.f, 0.f, 0.f);
                } else {
                    for (int d = 0; d < D; ++d) row[d] = 0.f;
                }
            }
        } else {
            const int fl = f - n_emb;
            if (force_all || __ldg(cnt + (int64_t)lin_plan[fl] * B + u) != 1) lin_rg[fl * lin_rg_stride + u] = 0.f;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// unique plan (hash based, per id column)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t mix32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}

// Thread mapping is column-major (i = c*B + b): the lanes of a warp work on the same id column, so
// the "next unique index" counter of that column is bumped once per warp (ballot-aggregated)
// instead of once per inserted key.  The kernel is a chain of dependent L2 round trips per key
// (id -> CAS -> next probe ...), so each thread keeps PLAN_ILP independent keys in flight and the
// first probe is the CAS itself (no read-before-CAS): measured 140 us -> see profiles/.
constexpr int PLAN_ILP = 4;

// Ragged plans (receive lists of the sharded backward) size each column's hash table from the number of entries the
// column really holds: the storage stride stays H, but only the first plan_hash_slots(count) slots are cleared and
// probed — the clears no longer scale with the worst-case list capacity (world x batch).
__device__ __forceinline__ uint32_t plan_hash_slots(int64_t count, int64_t H) {
    uint32_t h = 64;
    while ((int64_t)h < 2 * count && (int64_t)h < H) h <<= 1;
    return h;
}

__global__ void __launch_bounds__(256) plan_clear_kernel(int64_t B, int64_t H, const int32_t* __restrict__ col_count,
                                                         int32_t* keys, int32_t* n_uniq, int32_t* uniq, int32_t* cnt) {
    const int c = blockIdx.y;
    int64_t n = __ldg(col_count + c);
    if (n > B) n = B;
    if (n < 0) n = 0;
    const int64_t hs = plan_hash_slots(n, H);
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nt = (int64_t)gridDim.x * blockDim.x;
    if (t == 0) n_uniq[c] = 0;
    int4* k4 = reinterpret_cast<int4*>(keys + (int64_t)c * H);          // H is a power of two >= 64
    for (int64_t i = t; i < hs / 4; i += nt) k4[i] = make_int4(-1, -1, -1, -1);
    for (int64_t i = t; i < n; i += nt) {
        uniq[(int64_t)c * B + i] = 0;
        cnt[(int64_t)c * B + i] = 0;
    }
}
__global__ void __launch_bounds__(256) plan_insert_kernel(const float* __restrict__ X, int64_t ldx,
                                                          int64_t B, int n_cols,
                                                          const int32_t* __restrict__ cols,
                                                          const int32_t* __restrict__ vocab,
                                                          int32_t* keys, int32_t* vals, int64_t H,
                                                          int32_t* n_uniq, int32_t* uniq,
                                                          int32_t* inv, int32_t* err_flag, int id_mode,
                                                          const int32_t* __restrict__ col_count) {
    // ragged plan: blockIdx.y is the column and only its col_count[c] entries are visited
    const int col_fixed = col_count ? (int)blockIdx.y : -1;
    int64_t total = B * n_cols;
    uint32_t col_mask = (uint32_t)(H - 1);
    if (col_count) {
        total = __ldg(col_count + col_fixed);
        if (total > B) total = B;
        col_mask = plan_hash_slots(total, H) - 1u;
    }
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t base = warp * (32 * PLAN_ILP); base < total; base += nwarps * (32 * PLAN_ILP)) {
        int c[PLAN_ILP];
        int64_t b[PLAN_ILP];
        bool valid[PLAN_ILP], won[PLAN_ILP];
        uint32_t slot[PLAN_ILP], mask[PLAN_ILP];
        int32_t key[PLAN_ILP], old[PLAN_ILP];
        float xv[PLAN_ILP];
#pragma unroll
        for (int t = 0; t < PLAN_ILP; ++t) {                 // all id loads in flight
            const int64_t i = base + t * 32 + lane;
            valid[t] = i < total;
            const int64_t ic = valid[t] ? i : 0;
            c[t] = col_fixed >= 0 ? col_fixed : (int)(ic / B);
            b[t] = col_fixed >= 0 ? ic : ic - (int64_t)c[t] * B;
            mask[t] = col_mask;
            xv[t] = __ldg(X + (valid[t] ? b[t] : 0) * ldx + cols[c[t]]);
        }
#pragma unroll
        for (int t = 0; t < PLAN_ILP; ++t) {                 // all first probes (CAS) in flight
            key[t] = (int32_t)decode_id(xv[t], vocab[c[t]], valid[t] ? err_flag : nullptr, id_mode);
            slot[t] = mix32((uint32_t)key[t]) & mask[t];
            old[t] = valid[t] ? atomicCAS(keys + (int64_t)c[t] * H + slot[t], -1, key[t]) : key[t];
        }
#pragma unroll
        for (int t = 0; t < PLAN_ILP; ++t) {
            won[t] = false;
            if (valid[t]) {
                int32_t* kc = keys + (int64_t)c[t] * H;
                int32_t o = old[t];
                while (true) {
                    if (o == -1) {
                        won[t] = true;
                        break;
                    }
                    if (o == key[t]) break;
                    slot[t] = (slot[t] + 1) & mask[t];         // linear probing (rare at load factor <= 0.5)
                    o = atomicCAS(kc + slot[t], -1, key[t]);
                }
                inv[b[t] * n_cols + c[t]] = (int32_t)slot[t];  // replaced by the unique index in plan_finalize_kernel
            }
        }
        __syncwarp();
#pragma unroll
        for (int t = 0; t < PLAN_ILP; ++t) {
            // lanes that inserted a key of the same column share one atomicAdd
            const unsigned winners = __ballot_sync(0xffffffffu, won[t]);
            if (won[t]) {
                const unsigned peers = __match_any_sync(winners, c[t]);
                const int leader = __ffs(peers) - 1;
                int32_t base_u = 0;
                if (lane == leader) base_u = atomicAdd(n_uniq + c[t], __popc(peers));
                base_u = __shfl_sync(peers, base_u, leader);
                const int32_t u = base_u + __popc(peers & ((1u << lane) - 1u));
                vals[(int64_t)c[t] * H + slot[t]] = u;
                uniq[(int64_t)c[t] * B + u] = key[t];
            }
        }
    }
}

__global__ void __launch_bounds__(256) plan_finalize_kernel(int64_t B, int n_cols,
                                                            const int32_t* __restrict__ vals,
                                                            int64_t H, int32_t* inv, int32_t* cnt,
                                                            const int32_t* __restrict__ col_count) {
    if (col_count) {                                   // ragged plan: blockIdx.y is the column
        const int c = blockIdx.y;
        int64_t n = __ldg(col_count + c);
        if (n > B) n = B;
        for (int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; b < n; b += (int64_t)gridDim.x * blockDim.x) {
            const int64_t i = b * n_cols + c;
            const int32_t u = __ldcg(vals + (int64_t)c * H + inv[i]);
            inv[i] = u;
            atomicAdd(cnt + (int64_t)c * B + u, 1);
        }
        return;
    }
    const int64_t total = B * n_cols;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % n_cols);
        const int32_t u = __ldcg(vals + (int64_t)c * H + inv[i]);
        inv[i] = u;
        atomicAdd(cnt + (int64_t)c * B + u, 1);
    }
}

// dw[k] = sum_b g[b] * X[b, cols[k]]: one thread per sample (grid-stride), up to 16 columns per pass in
// registers, warp-shuffle + shared-memory reduction, one atomicAdd per column and block
__global__ void __launch_bounds__(256) lin_dense_wgrad_kernel(const float* __restrict__ X,
                                                              int64_t ldx, int64_t B, int n,
                                                              const int32_t* __restrict__ cols,
                                                              const float* __restrict__ g, float* dw) {
    __shared__ float s_part[8][16];
    __shared__ int s_col[16];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int k0 = 0; k0 < n; k0 += 16) {
        const int nk = (n - k0 < 16) ? n - k0 : 16;
        __syncthreads();
        if (threadIdx.x < 16) s_col[threadIdx.x] = cols[k0 + (threadIdx.x < nk ? threadIdx.x : 0)];
        __syncthreads();
        float acc[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) acc[k] = 0.f;
        for (int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; b < B; b += (int64_t)gridDim.x * blockDim.x) {
            const float gb = __ldg(g + b);
            const float* xrow = X + b * ldx;
#pragma unroll
            for (int k = 0; k < 16; ++k)
                if (k < nk) acc[k] += gb * __ldg(xrow + s_col[k]);
        }
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const float v = warp_sum(acc[k]);
            if (lane == 0) s_part[wid][k] = v;
        }
        __syncthreads();
        if (threadIdx.x < nk) {
            float v = 0.f;
            for (int w = 0; w < 8; ++w) v += s_part[w][threadIdx.x];
            atomicAdd(dw + k0 + threadIdx.x, v);
        }
    }
}

int lpr_for_dim(int D) {
    if (D % 4 != 0) return 0;
    const int l = D / 4;
    if (l < 1 || l > 32 || (l & (l - 1)) != 0) return 0;
    return l;
}

unsigned sample_grid(int64_t B, int warps_per_block, int blocks_per_sm) {
    int64_t want = ceil_div64(B, warps_per_block);
    int64_t cap = (int64_t)ctr_sm_count() * blocks_per_sm;
    if (want > cap) want = cap;
    if (want < 1) want = 1;
    return (unsigned)want;
}

}  // namespace

// =============================================================================================
// C ABI
// =============================================================================================
static int gather_fwd_impl(const float* X, int64_t ldx, int64_t B, int n_emb, int D,
                              const float* const* emb_tables, const int32_t* emb_cols,
                              const int32_t* emb_vocab, int n_lin, const float* const* lin_tables,
                              const int32_t* lin_cols, const int32_t* lin_vocab, int n_dense,
                              const int32_t* dense_cols, int n_lin_dense,
                              const int32_t* lin_dense_cols, const float* lin_dense_w, float* blk,
                              int64_t ld_blk, float* lin, float* fm, int32_t* err_flag,
                              int n_shards, int id_mode, void* stream, const int32_t* where, int n_plan, int me,
                              const int32_t* emb_plan_col, const int32_t* lin_plan_col,
                              const float* const* resp_emb, const float* const* resp_lin) {
    CTR_ARG(X && B >= 0 && ldx >= 0, "ctr_gather_fwd: X/B/ldx invalid");
    CTR_ARG(n_shards >= 1, "ctr_gather_fwd: n_shards must be >= 1");
    CTR_ARG(n_emb >= 0 && n_lin >= 0 && n_dense >= 0 && n_lin_dense >= 0, "ctr_gather_fwd: negative count");
    CTR_ARG(n_emb == 0 || (D > 0 && emb_tables && emb_cols && emb_vocab), "ctr_gather_fwd: embedding slot arrays missing");
    CTR_ARG(n_lin == 0 || (lin_tables && lin_cols && lin_vocab), "ctr_gather_fwd: linear slot arrays missing");
    CTR_ARG(n_dense == 0 || dense_cols, "ctr_gather_fwd: dense_cols missing");
    CTR_ARG(n_lin_dense == 0 || (lin_dense_cols && lin_dense_w), "ctr_gather_fwd: linear dense arrays missing");
    CTR_ARG(!blk || ld_blk >= (int64_t)n_emb * D + n_dense, "ctr_gather_fwd: ld_blk too small");
    if (B == 0) return 0;
    GatherArgs a{X, ldx, B, n_emb, D, emb_tables, emb_cols, emb_vocab, n_lin, lin_tables, lin_cols,
                 lin_vocab, n_dense, dense_cols, n_lin_dense, lin_dense_cols, lin_dense_w, blk, ld_blk,
                 lin, fm, err_flag, n_shards, where, n_plan, me, emb_plan_col, lin_plan_col, resp_emb, resp_lin, id_mode};
    cudaStream_t st = as_stream(stream);
    const int lpr = (n_emb > 0) ? lpr_for_dim(D) : 1;
    const bool vec_ok = lpr > 0 && n_emb * n_shards <= kMaxSmemSlots &&
                        (!blk || ((ld_blk % 4 == 0) && ((reinterpret_cast<uintptr_t>(blk) & 15) == 0)));
    const unsigned grid = sample_grid((B + 1) / 2, 8, 8);
    if (where && !vec_ok) {
        ctr_set_error("ctr_gather_fwd_exchanged: needs the vector path (D %% 4 == 0, aligned block)");
        return -2;
    }
    if (vec_ok) {
        const size_t smem = (size_t)n_emb * n_shards * sizeof(void*) + (size_t)n_emb * 12;
        switch (lpr) {
            case 1:
                if (where) gather_fwd_vec_kernel<1, 2, 4, true><<<grid, 256, smem, st>>>(a);
                else gather_fwd_vec_kernel<1, 2><<<grid, 256, smem, st>>>(a);
                break;
            case 2:
                if (where) gather_fwd_vec_kernel<2, 2, 4, true><<<grid, 256, smem, st>>>(a);
                else gather_fwd_vec_kernel<2, 2><<<grid, 256, smem, st>>>(a);
                break;
            case 4: {
                // CTR_GATHER_VARIANT (diagnostic): fewer loads in flight per lane — peer (NVLink) rows were
                // 3.7x slower than torch's one-load-per-thread gather with the default 2 samples x 4 steps
                const char* e = getenv("CTR_GATHER_VARIANT");
                const int var = e ? atoi(e) : 0;
                if (var == 1) gather_fwd_vec_kernel<4, 1, 4><<<sample_grid(B, 8, 8), 256, smem, st>>>(a);
                else if (var == 2) gather_fwd_vec_kernel<4, 1, 2><<<sample_grid(B, 8, 8), 256, smem, st>>>(a);
                else if (var == 3) gather_fwd_vec_kernel<4, 1, 1><<<sample_grid(B, 8, 8), 256, smem, st>>>(a);
                else if (var == 4) gather_fwd_vec_kernel<4, 1, 1><<<sample_grid(B, 8, 2), 256, smem, st>>>(a);
                else if (where) gather_fwd_vec_kernel<4, 2, 4, true><<<grid, 256, smem, st>>>(a);
                else gather_fwd_vec_kernel<4, 2><<<grid, 256, smem, st>>>(a);
                break;
            }
            case 8:
                if (where) gather_fwd_vec_kernel<8, 2, 4, true><<<grid, 256, smem, st>>>(a);
                else gather_fwd_vec_kernel<8, 2><<<grid, 256, smem, st>>>(a);
                break;
            case 16:
                if (where) gather_fwd_vec_kernel<16, 2, 4, true><<<grid, 256, smem, st>>>(a);
                else gather_fwd_vec_kernel<16, 2><<<grid, 256, smem, st>>>(a);
                break;
            default:
                if (where) gather_fwd_vec_kernel<32, 2, 4, true><<<grid, 256, smem, st>>>(a);
                else gather_fwd_vec_kernel<32, 2><<<grid, 256, smem, st>>>(a);
                break;
        }
        CTR_LAUNCH_OK("gather_fwd_vec_kernel");
    } else {
        CTR_ARG(blk || !fm || n_emb == 0, "ctr_gather_fwd: FM on the generic path needs blk");
        gather_fwd_generic_kernel<<<grid, 256, 0, st>>>(a);
        CTR_LAUNCH_OK("gather_fwd_generic_kernel");
        if (fm) {
            if (n_emb > 0) {
                fm_fwd_kernel<<<grid, 256, 0, st>>>(blk, ld_blk, B, n_emb, D, fm);
                CTR_LAUNCH_OK("fm_fwd_kernel");
            } else {
                CTR_CUDA(cudaMemsetAsync(fm, 0, sizeof(float) * B, st));
            }
        }
    }
    return 0;
}

extern "C" int ctr_gather_fwd(const float* X, int64_t ldx, int64_t B, int n_emb, int D,
                              const float* const* emb_tables, const int32_t* emb_cols,
                              const int32_t* emb_vocab, int n_lin, const float* const* lin_tables,
                              const int32_t* lin_cols, const int32_t* lin_vocab, int n_dense,
                              const int32_t* dense_cols, int n_lin_dense,
                              const int32_t* lin_dense_cols, const float* lin_dense_w, float* blk,
                              int64_t ld_blk, float* lin, float* fm, int32_t* err_flag,
                              int n_shards, int id_mode, void* stream) {
    return gather_fwd_impl(X, ldx, B, n_emb, D, emb_tables, emb_cols, emb_vocab, n_lin, lin_tables, lin_cols, lin_vocab,
                           n_dense, dense_cols, n_lin_dense, lin_dense_cols, lin_dense_w, blk, ld_blk, lin, fm, err_flag,
                           n_shards, id_mode, stream, nullptr, 0, 0, nullptr, nullptr, nullptr, nullptr);
}

extern "C" int ctr_gather_fwd_exchanged(const float* X, int64_t ldx, int64_t B, int n_emb, int D,
                                        const float* const* emb_tables, const int32_t* emb_cols,
                                        const int32_t* emb_vocab, int n_lin, const float* const* lin_tables,
                                        const int32_t* lin_cols, const int32_t* lin_vocab, int n_dense,
                                        const int32_t* dense_cols, int n_lin_dense,
                                        const int32_t* lin_dense_cols, const float* lin_dense_w, float* blk,
                                        int64_t ld_blk, float* lin, float* fm, int32_t* err_flag,
                                        int n_shards, int rank, const int32_t* where, int n_plan,
                                        const int32_t* emb_plan_col, const int32_t* lin_plan_col,
                                        const float* const* resp_emb, const float* const* resp_lin, int id_mode,
                                        void* stream) {
    CTR_ARG(where && n_plan > 0 && rank >= 0 && rank < n_shards, "ctr_gather_fwd_exchanged: bad exchange arguments");
    CTR_ARG((n_emb == 0 || (emb_plan_col && resp_emb)) && (n_lin == 0 || (lin_plan_col && resp_lin)),
            "ctr_gather_fwd_exchanged: plan columns / response buffers missing");
    return gather_fwd_impl(X, ldx, B, n_emb, D, emb_tables, emb_cols, emb_vocab, n_lin, lin_tables, lin_cols, lin_vocab,
                           n_dense, dense_cols, n_lin_dense, lin_dense_cols, lin_dense_w, blk, ld_blk, lin, fm, err_flag,
                           n_shards, id_mode, stream, where, n_plan, rank, emb_plan_col, lin_plan_col, resp_emb, resp_lin);
}

extern "C" int ctr_fm_fwd(const float* blk, int64_t ld, int64_t B, int F, int D, float* fm,
                          void* stream) {
    CTR_ARG(blk && fm && F > 0 && D > 0 && ld >= (int64_t)F * D, "ctr_fm_fwd: bad arguments");
    if (B == 0) return 0;
    fm_fwd_kernel<<<sample_grid(B, 8, 8), 256, 0, as_stream(stream)>>>(blk, ld, B, F, D, fm);
    CTR_LAUNCH_OK("fm_fwd_kernel");
    return 0;
}

extern "C" int ctr_fm_bwd(const float* blk, int64_t ld, int64_t B, int F, int D, const float* g,
                          float* d_blk, int64_t ld_d, void* stream) {
    CTR_ARG(blk && g && d_blk && F > 0 && D > 0, "ctr_fm_bwd: bad arguments");
    if (B == 0) return 0;
    fm_bwd_kernel<<<sample_grid(B, 8, 8), 256, 0, as_stream(stream)>>>(blk, ld, B, F, D, g, d_blk, ld_d);
    CTR_LAUNCH_OK("fm_bwd_kernel");
    return 0;
}

static int launch_scatter(ScatterArgs& a, bool rowwise, cudaStream_t st, bool force_generic = false) {
    const int lpr = (a.n_emb > 0) ? lpr_for_dim(a.D) : 1;
    const bool aligned =
        (!a.blk || (a.ld_blk % 4 == 0 && (reinterpret_cast<uintptr_t>(a.blk) & 15) == 0)) &&
        (!a.d_blk || (a.ld_dblk % 4 == 0 && (reinterpret_cast<uintptr_t>(a.d_blk) & 15) == 0));
    const unsigned grid = sample_grid((a.B + 1) / 2, 8, 8);
    // samples per warp: 2 keeps more loads in flight per warp but needs 144 registers (18 % occupancy);
    // 1 halves the register arrays.  CTR_SCATTER_SPW=1|2 selects (A/B profiling), default 2.
    static int spw = 0;
    if (spw == 0) {
        const char* e = getenv("CTR_SCATTER_SPW");
        spw = (e && e[0] == '1') ? 1 : 2;
    }
    const unsigned grid_vec = (spw == 1) ? sample_grid(a.B, 4, 16) : sample_grid((a.B + 1) / 2, 4, 12);
    if (lpr > 0 && aligned && !force_generic) {
#define LAUNCH_SC(L)                                                                     \
    if (spw == 1) {                                                                      \
        if (rowwise) scatter_bwd_vec_kernel<L, true, 1><<<grid_vec, 128, 0, st>>>(a);    \
        else scatter_bwd_vec_kernel<L, false, 1><<<grid_vec, 128, 0, st>>>(a);           \
    } else {                                                                             \
        if (rowwise) scatter_bwd_vec_kernel<L, true, 2><<<grid_vec, 128, 0, st>>>(a);    \
        else scatter_bwd_vec_kernel<L, false, 2><<<grid_vec, 128, 0, st>>>(a);           \
    }
        switch (lpr) {
            case 1: LAUNCH_SC(1) break;
            case 2: LAUNCH_SC(2) break;
            case 4: LAUNCH_SC(4) break;
            case 8: LAUNCH_SC(8) break;
            case 16: LAUNCH_SC(16) break;
            default: LAUNCH_SC(32) break;
        }
#undef LAUNCH_SC
        CTR_LAUNCH_OK("scatter_bwd_vec_kernel");
        return 1;
    }
    if (rowwise) scatter_bwd_generic_kernel<true><<<grid, 256, 0, st>>>(a);
    else scatter_bwd_generic_kernel<false><<<grid, 256, 0, st>>>(a);
    CTR_LAUNCH_OK("scatter_bwd_generic_kernel");
    return 2;
}

extern "C" int ctr_scatter_bwd_dense(const float* X, int64_t ldx, int64_t B, int n_emb, int D,
                                     float* const* emb_grads, const int32_t* emb_cols,
                                     const int32_t* emb_vocab, int n_lin, float* const* lin_grads,
                                     const int32_t* lin_cols, const int32_t* lin_vocab,
                                     const float* blk, int64_t ld_blk, const float* d_blk,
                                     int64_t ld_dblk, const float* g_fm, const float* g_lin,
                                     int id_mode, void* stream) {
    CTR_ARG(X && B >= 0, "ctr_scatter_bwd_dense: X/B invalid");
    CTR_ARG(n_emb == 0 || (D > 0 && emb_grads && emb_cols && emb_vocab), "ctr_scatter_bwd_dense: embedding arrays missing");
    CTR_ARG(n_lin == 0 || !g_lin || (lin_grads && lin_cols && lin_vocab), "ctr_scatter_bwd_dense: linear arrays missing");
    CTR_ARG(!g_fm || blk, "ctr_scatter_bwd_dense: FM gradient needs blk");
    if (B == 0) return 0;
    ScatterArgs a{};
    a.X = X; a.ldx = ldx; a.B = B; a.n_emb = n_emb; a.D = D;
    a.emb_out = emb_grads; a.emb_cols = emb_cols; a.emb_vocab = emb_vocab;
    a.n_lin = g_lin ? n_lin : 0; a.lin_out = lin_grads; a.lin_cols = lin_cols; a.lin_vocab = lin_vocab;
    a.blk = blk; a.ld_blk = ld_blk; a.d_blk = d_blk; a.ld_dblk = ld_dblk; a.g_fm = g_fm; a.g_lin = g_lin;
    a.id_mode = id_mode;
    if (!d_blk && !g_fm) a.n_emb = 0;
    const int r = launch_scatter(a, false, as_stream(stream));
    return r < 0 ? r : (r > 2 ? r : 0);
}

extern "C" int64_t ctr_unique_plan_hash_slots(int64_t B) {
    int64_t h = 64;
    while (h < 2 * B) h <<= 1;
    return h;
}

extern "C" int ctr_unique_plan(const float* X, int64_t ldx, int64_t B, int n_cols,
                               const int32_t* cols, const int32_t* vocab, int32_t* hash_keys,
                               int32_t* hash_vals, int64_t H, int32_t* n_uniq, int32_t* uniq,
                               int32_t* inv, int32_t* cnt, int32_t* err_flag, int id_mode,
                               const int32_t* col_count, void* stream) {
    CTR_ARG(X && cols && vocab && hash_keys && hash_vals && n_uniq && uniq && inv && cnt,
            "ctr_unique_plan: null argument");
    CTR_ARG(n_cols > 0 && B >= 0, "ctr_unique_plan: bad sizes");
    CTR_ARG(H >= 2 * B && (H & (H - 1)) == 0, "ctr_unique_plan: H must be a power of two >= 2B");
    if (B == 0) return 0;
    cudaStream_t st = as_stream(stream);
    int64_t blocks = ceil_div64(B * n_cols, 256);
    const int64_t cap = (int64_t)ctr_sm_count() * 8;
    if (blocks > cap) blocks = cap;
    if (col_count) {
        int64_t bx = ceil_div64(blocks, n_cols);
        if (bx < 1) bx = 1;
        plan_clear_kernel<<<dim3((unsigned)bx, (unsigned)n_cols), 256, 0, st>>>(B, H, col_count, hash_keys, n_uniq, uniq, cnt);
        CTR_LAUNCH_OK("plan_clear_kernel");
    } else {
        CTR_CUDA(cudaMemsetAsync(hash_keys, 0xFF, sizeof(int32_t) * H * n_cols, st));
        CTR_CUDA(cudaMemsetAsync(n_uniq, 0, sizeof(int32_t) * n_cols, st));
        CTR_CUDA(cudaMemsetAsync(uniq, 0, sizeof(int32_t) * B * n_cols, st));
        CTR_CUDA(cudaMemsetAsync(cnt, 0, sizeof(int32_t) * B * n_cols, st));
    }
    dim3 grid((unsigned)blocks, 1);
    if (col_count) {
        // expected list length is about B / n_shards: a few blocks per column keep every SM busy without
        // launching the worst-case grid
        int64_t bx = ceil_div64(cap, n_cols);
        if (bx < 1) bx = 1;
        grid = dim3((unsigned)bx, (unsigned)n_cols);
    }
    plan_insert_kernel<<<grid, 256, 0, st>>>(X, ldx, B, n_cols, cols, vocab, hash_keys, hash_vals, H, n_uniq, uniq, inv,
                                            err_flag, id_mode, col_count);
    CTR_LAUNCH_OK("plan_insert_kernel");
    plan_finalize_kernel<<<grid, 256, 0, st>>>(B, n_cols, hash_vals, H, inv, cnt, col_count);
    CTR_LAUNCH_OK("plan_finalize_kernel");
    return 0;
}

extern "C" int ctr_scatter_bwd_rowwise(int64_t B, int n_plan_cols, const int32_t* inv,
                                       const int32_t* cnt, const int32_t* n_uniq, int n_emb, int D,
                                       float* emb_rowgrad, int64_t emb_rowgrad_stride,
                                       const int32_t* emb_plan_col, int n_lin, float* lin_rowgrad,
                                       int64_t lin_rowgrad_stride, const int32_t* lin_plan_col,
                                       const float* blk, int64_t ld_blk, const float* d_blk,
                                       int64_t ld_dblk, const float* g_fm, const float* g_lin,
                                       void* stream) {
    (void)n_uniq;
    CTR_ARG(inv && cnt && n_plan_cols > 0 && B >= 0, "ctr_scatter_bwd_rowwise: plan missing");
    CTR_ARG(n_emb == 0 || (D > 0 && emb_rowgrad && emb_plan_col && emb_rowgrad_stride >= B * D),
            "ctr_scatter_bwd_rowwise: embedding row-gradient buffer missing or too small");
    CTR_ARG(n_lin == 0 || (lin_rowgrad && lin_plan_col && g_lin && lin_rowgrad_stride >= B),
            "ctr_scatter_bwd_rowwise: linear row-gradient buffer / g_lin missing");
    CTR_ARG(!g_fm || blk, "ctr_scatter_bwd_rowwise: FM gradient needs blk");
    if (B == 0) return 0;
    cudaStream_t st = as_stream(stream);
    ScatterArgs a{};
    a.B = B; a.n_emb = n_emb; a.D = D;
    a.emb_rg = emb_rowgrad; a.emb_rg_stride = emb_rowgrad_stride; a.emb_cols = emb_plan_col;
    a.n_lin = n_lin; a.lin_rg = lin_rowgrad; a.lin_rg_stride = lin_rowgrad_stride; a.lin_cols = lin_plan_col;
    a.blk = blk; a.ld_blk = ld_blk; a.d_blk = d_blk; a.ld_dblk = ld_dblk; a.g_fm = g_fm; a.g_lin = g_lin;
    a.n_plan = n_plan_cols; a.inv = inv; a.cnt = cnt;
    const int lpr = (n_emb > 0) ? lpr_for_dim(D) : 1;
    const bool aligned =
        (!blk || (ld_blk % 4 == 0 && (reinterpret_cast<uintptr_t>(blk) & 15) == 0)) &&
        (!d_blk || (ld_dblk % 4 == 0 && (reinterpret_cast<uintptr_t>(d_blk) & 15) == 0)) &&
        (n_emb == 0 || (emb_rowgrad_stride % 4 == 0 && (reinterpret_cast<uintptr_t>(emb_rowgrad) & 15) == 0));
    const int force_all = (lpr > 0 && aligned) ? 0 : 1;
    {
        const int64_t work = (int64_t)(n_emb + n_lin) * B;
        int64_t blocks = ceil_div64(work, 256);
        const int64_t cap = (int64_t)ctr_sm_count() * 8;
        if (blocks > cap) blocks = cap;
        if (blocks < 1) blocks = 1;
        rowgrad_prep_kernel<<<(unsigned)blocks, 256, 0, st>>>(B, cnt, n_emb, D, emb_rowgrad, emb_rowgrad_stride,
                                                              emb_plan_col, n_lin, lin_rowgrad,
                                                              lin_rowgrad_stride, lin_plan_col, force_all);
        CTR_LAUNCH_OK("rowgrad_prep_kernel");
    }
    const int r = launch_scatter(a, true, st, force_all != 0);
    return r < 0 ? r : (r > 2 ? r : 0);
}

extern "C" int ctr_lin_dense_wgrad(const float* X, int64_t ldx, int64_t B, int n,
                                   const int32_t* cols, const float* g, float* dw, void* stream) {
    CTR_ARG(X && cols && g && dw && n > 0 && B >= 0, "ctr_lin_dense_wgrad: bad arguments");
    cudaStream_t st = as_stream(stream);
    CTR_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * n, st));
    if (B == 0) return 0;
    lin_dense_wgrad_kernel<<<sample_grid(B, 256, 2), 256, 0, st>>>(X, ldx, B, n, cols, g, dw);
    CTR_LAUNCH_OK("lin_dense_wgrad_kernel");
    return 0;
}
