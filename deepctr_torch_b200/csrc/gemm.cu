// fp32 GEMM engine + the dense-tower entry points (DNN layer fwd/bwd, row-dot head, prediction).
//
// Parity mode of the tower: every product is an FP32 FFMA with FP32 accumulation, so logits
// agree with the reference's fp32 MKL/oneDNN path to round-off (SURVEY.md §7 hard part 1:
// plain TF32/BF16 tensor-core inputs miss the 1e-5 gate).  Tiling: 128x128x16 CTA tile, 256
// threads, 8x8 register micro-tile (two 4-wide halves so that every shared-memory read is a
// conflict-free 128-bit load), double-buffered through registers; operands with any element
// strides, optional activation-gradient prologue on A and fused bias/activation/mask epilogues;
// split-K with fp32 atomics for the weight-gradient shapes (M,N small, K = batch).
#include <stdlib.h>

#include "gemm.cuh"

namespace {

template <int BM, int BN, int BK, int TM, int TN, bool A_KMAJ, bool B_KMAJ>
__global__ void __launch_bounds__((BM / TM) * (BN / TN))
sgemm_kernel(GemmArgs g, int64_t k_chunk) {
    constexpr int NT = (BM / TM) * (BN / TN);
    constexpr int A_ELEMS = BM * BK / NT;
    constexpr int B_ELEMS = BN * BK / NT;
    static_assert(TM % 4 == 0 && TN % 4 == 0, "micro-tile must be a multiple of 4");
    __shared__ __align__(16) float As[2][BK][BM + 4];
    __shared__ __align__(16) float Bs[2][BK][BN + 4];

    const int tid = threadIdx.x;
    const int tx = tid % (BN / TN), ty = tid / (BN / TN);
    const int64_t m0 = (int64_t)blockIdx.y * BM, n0 = (int64_t)blockIdx.x * BN;
    const int64_t kbeg = (int64_t)blockIdx.z * k_chunk;
    const int64_t kend = (kbeg + k_chunk < g.K) ? kbeg + k_chunk : g.K;

    float ra[A_ELEMS], rb[B_ELEMS];
    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    auto load_tile = [&](int64_t k0) {
#pragma unroll
        for (int i = 0; i < A_ELEMS; ++i) {
            const int e = tid + i * NT;
            const int kk = A_KMAJ ? e % BK : e / BM;
            const int mm = A_KMAJ ? e / BK : e % BM;
            const int64_t m = m0 + mm, k = k0 + kk;
            float v = 0.f;
            if (m < g.M && k < kend) {
                v = __ldg(g.A + m * g.sam + k * g.sak);
                if (g.amask) v *= act_grad_from_y(g.amask_act, __ldg(g.amask + m * g.smm + k * g.smk));
            }
            ra[i] = v;
        }
#pragma unroll
        for (int i = 0; i < B_ELEMS; ++i) {
            const int e = tid + i * NT;
            const int kk = B_KMAJ ? e % BK : e / BN;
            const int nn = B_KMAJ ? e / BK : e % BN;
            const int64_t n = n0 + nn, k = k0 + kk;
            float v = 0.f;
            if (n < g.N && k < kend) {
                v = __ldg(g.B + n * g.sbn + k * g.sbk);
                if (g.bmask) v *= act_grad_from_y(g.bmask_act, __ldg(g.bmask + n * g.sbmn + k * g.sbmk));
            }
            rb[i] = v;
        }
    };
    auto store_tile = [&](int buf) {
#pragma unroll
        for (int i = 0; i < A_ELEMS; ++i) {
            const int e = tid + i * NT;
            const int kk = A_KMAJ ? e % BK : e / BM;
            const int mm = A_KMAJ ? e / BK : e % BM;
            As[buf][kk][mm] = ra[i];
        }
#pragma unroll
        for (int i = 0; i < B_ELEMS; ++i) {
            const int e = tid + i * NT;
            const int kk = B_KMAJ ? e % BK : e / BN;
            const int nn = B_KMAJ ? e / BK : e % BN;
            Bs[buf][kk][nn] = rb[i];
        }
    };

    if (kbeg < kend) {
        load_tile(kbeg);
        store_tile(0);
    }
    __syncthreads();
    int buf = 0;
    for (int64_t k0 = kbeg; k0 < kend; k0 += BK) {
        const bool has_next = k0 + BK < kend;
        if (has_next) load_tile(k0 + BK);
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float a[TM], b[TN];
#pragma unroll
            for (int h = 0; h < TM / 4; ++h) {
                const float4 v = *reinterpret_cast<const float4*>(&As[buf][kk][h * (BM / (TM / 4)) + ty * 4]);
                a[h * 4 + 0] = v.x; a[h * 4 + 1] = v.y; a[h * 4 + 2] = v.z; a[h * 4 + 3] = v.w;
            }
#pragma unroll
            for (int h = 0; h < TN / 4; ++h) {
                const float4 v = *reinterpret_cast<const float4*>(&Bs[buf][kk][h * (BN / (TN / 4)) + tx * 4]);
                b[h * 4 + 0] = v.x; b[h * 4 + 1] = v.y; b[h * 4 + 2] = v.z; b[h * 4 + 3] = v.w;
            }
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (has_next) store_tile(buf ^ 1);
        __syncthreads();
        buf ^= 1;
    }

    const bool split = gridDim.z > 1;
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int64_t m = m0 + (i / 4) * (BM / (TM / 4)) + ty * 4 + (i % 4);
        if (m >= g.M) continue;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int64_t n = n0 + (j / 4) * (BN / (TN / 4)) + tx * 4 + (j % 4);
            if (n >= g.N) continue;
            float v = acc[i][j];
            float* cp = g.C + m * g.ldc + n;
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
            if (split) atomicAdd(cp, v);
            else if (g.accumulate) *cp += v;
            else *cp = v;
        }
    }
}

template <int BM, int BN, int BK, int TM, int TN>
int dispatch_layout(const GemmArgs& g, dim3 grid, int64_t k_chunk, cudaStream_t st) {
    const bool a_k = (g.sak == 1) || (g.sam != 1);
    const bool b_k = (g.sbk == 1) || (g.sbn != 1);
    constexpr int NT = (BM / TM) * (BN / TN);
    if (a_k && b_k) sgemm_kernel<BM, BN, BK, TM, TN, true, true><<<grid, NT, 0, st>>>(g, k_chunk);
    else if (a_k && !b_k) sgemm_kernel<BM, BN, BK, TM, TN, true, false><<<grid, NT, 0, st>>>(g, k_chunk);
    else if (!a_k && b_k) sgemm_kernel<BM, BN, BK, TM, TN, false, true><<<grid, NT, 0, st>>>(g, k_chunk);
    else sgemm_kernel<BM, BN, BK, TM, TN, false, false><<<grid, NT, 0, st>>>(g, k_chunk);
    CTR_LAUNCH_OK("sgemm_kernel");
    return 0;
}

// column sums with optional activation-gradient mask and per-row weight
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ A, int64_t sam,
                                                     int64_t san, const float* __restrict__ mask,
                                                     int64_t smm, int64_t smn, int mask_act,
                                                     const float* __restrict__ w, int64_t M,
                                                     int64_t N, int64_t rows_per_block, float* out) {
    __shared__ float part[8][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int64_t n = (int64_t)blockIdx.x * 32 + tx;
    const int64_t mbeg = (int64_t)blockIdx.y * rows_per_block;
    const int64_t mend = (mbeg + rows_per_block < M) ? mbeg + rows_per_block : M;
    float acc = 0.f;
    if (n < N) {
        for (int64_t m = mbeg + ty; m < mend; m += 8) {
            float v = __ldg(A + m * sam + n * san);
            if (mask) v *= act_grad_from_y(mask_act, __ldg(mask + m * smm + n * smn));
            if (w) v *= __ldg(w + m);
            acc += v;
        }
    }
    part[ty][tx] = acc;
    __syncthreads();
    if (ty == 0 && n < N) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) s += part[i][tx];
        atomicAdd(out + n, s);
    }
}

// contiguous rows (san == 1), 128-bit loads: thread = (column quad, row lane)
__global__ void __launch_bounds__(256) colsum_vec_kernel(const float* __restrict__ A, int64_t sam,
                                                         const float* __restrict__ mask, int64_t smm,
                                                         int mask_act, const float* __restrict__ w,
                                                         int64_t M, int nq, int64_t rows_per_block,
                                                         float* out, int N) {
    extern __shared__ float4 s_part[];           // [ylanes][nq_tile]
    const int xq = blockDim.x;                   // column quads per block
    const int q = blockIdx.x * xq + threadIdx.x;
    const int64_t mbeg = (int64_t)blockIdx.y * rows_per_block;
    const int64_t mend = (mbeg + rows_per_block < M) ? mbeg + rows_per_block : M;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (q < nq) {
        for (int64_t m = mbeg + threadIdx.y; m < mend; m += blockDim.y) {
            float4 v = __ldg(reinterpret_cast<const float4*>(A + m * sam) + q);
            if (mask) {
                const float4 y = __ldg(reinterpret_cast<const float4*>(mask + m * smm) + q);
                v.x *= act_grad_from_y(mask_act, y.x); v.y *= act_grad_from_y(mask_act, y.y);
                v.z *= act_grad_from_y(mask_act, y.z); v.w *= act_grad_from_y(mask_act, y.w);
            }
            if (w) {
                const float ww = __ldg(w + m);
                v.x *= ww; v.y *= ww; v.z *= ww; v.w *= ww;
            }
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
    }
    s_part[threadIdx.y * xq + threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.y == 0 && q < nq) {
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int i = 0; i < (int)blockDim.y; ++i) {
            const float4 v = s_part[i * xq + threadIdx.x];
            t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
        }
        atomicAdd(out + q * 4 + 0, t.x);              // columns past N are row padding: read, never published
        if (q * 4 + 1 < N) atomicAdd(out + q * 4 + 1, t.y);
        if (q * 4 + 2 < N) atomicAdd(out + q * 4 + 2, t.z);
        if (q * 4 + 3 < N) atomicAdd(out + q * 4 + 3, t.w);
    }
}

__global__ void __launch_bounds__(256) rowdot_fwd_kernel(const float* __restrict__ H, int64_t ldh,
                                                         const float* __restrict__ w, int64_t B,
                                                         int N, float* out, int accumulate) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t b = warp0; b < B; b += nwarps) {
        const float* row = H + b * ldh;
        float acc = 0.f;
        for (int n = lane; n < N; n += 32) acc = fmaf(__ldg(row + n), __ldg(w + n), acc);
        acc = warp_sum(acc);
        if (lane == 0) out[b] = accumulate ? out[b] + acc : acc;
    }
}

// dH[b, n] (+)= g[b] * w[n]: a block is (256 / nq) rows x nq column quads, every thread keeps its four w values
// in registers and walks down the rows: 128-bit coalesced stores, no per-element index arithmetic
// (round 1: one scalar store and a 64-bit division per element, 0.2 ms for [65536, 557]).
template <bool VEC>
__global__ void __launch_bounds__(256) rowdot_bwd_dh_kernel(const float* __restrict__ w,
                                                            const float* __restrict__ g, int64_t B,
                                                            int N, int nq_pad, float* dH, int64_t lddh,
                                                            int accumulate) {
    const int q = threadIdx.x % nq_pad, rsub = threadIdx.x / nq_pad, rows_per_block = blockDim.x / nq_pad;
    const int step = VEC ? 4 * nq_pad : nq_pad;
    for (int n0 = VEC ? 4 * q : q; n0 < N; n0 += step) {
        float4 w4 = make_float4(0.f, 0.f, 0.f, 0.f);
        w4.x = __ldg(w + n0);
        if (VEC) {                                  // columns past N are row padding: they receive zeros
            if (n0 + 1 < N) w4.y = __ldg(w + n0 + 1);
            if (n0 + 2 < N) w4.z = __ldg(w + n0 + 2);
            if (n0 + 3 < N) w4.w = __ldg(w + n0 + 3);
        }
        for (int64_t b = (int64_t)blockIdx.x * rows_per_block + rsub; b < B; b += (int64_t)gridDim.x * rows_per_block) {
            const float gv = __ldg(g + b);
            float* p = dH + b * lddh + n0;
            if (VEC) {
                float4 v = make_float4(gv * w4.x, gv * w4.y, gv * w4.z, gv * w4.w);
                if (accumulate) {
                    const float4 o = *reinterpret_cast<const float4*>(p);
                    v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
                }
                *reinterpret_cast<float4*>(p) = v;
            } else {
                const float v = gv * w4.x;
                *p = accumulate ? *p + v : v;
            }
        }
    }
}

struct TermList {
    const float* p[8];
    int n;
};

__global__ void __launch_bounds__(256) predict_fwd_kernel(TermList t, const float* bias, int64_t B,
                                                          int binary, float* logit, float* y) {
    for (int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; b < B;
         b += (int64_t)gridDim.x * blockDim.x) {
        float z = 0.f;
        for (int i = 0; i < t.n; ++i) z += t.p[i][b];
        if (bias) z += bias[0];
        if (logit) logit[b] = z;
        y[b] = binary ? 1.f / (1.f + expf(-z)) : z;
    }
}

__global__ void __launch_bounds__(256) predict_bwd_kernel(const float* __restrict__ y,
                                                          const float* __restrict__ dy, int64_t B,
                                                          int binary, float* dlogit, float* dbias) {
    float acc = 0.f;
    for (int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; b < B;
         b += (int64_t)gridDim.x * blockDim.x) {
        const float yy = y[b];
        const float v = binary ? dy[b] * yy * (1.f - yy) : dy[b];
        dlogit[b] = v;
        acc += v;
    }
    if (dbias) {
        acc = warp_sum(acc);
        __shared__ float part[8];
        const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
        if (lane == 0) part[wid] = acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            float s = 0.f;
            for (int i = 0; i < 8; ++i) s += part[i];
            atomicAdd(dbias, s);
        }
    }
}

}  // namespace

// Engines: 0 = FP32 FFMA tiles (tiny problems), 1 = tcgen05 3xTF32 with in-kernel operand split
// (gemm_tc.cu; needs no scratch), 2 = tcgen05 3xTF32 over pre-split packed operands (gemm_pk.cu;
// the default whenever the caller registered enough scratch).  CTR_GEMM=simt | tc1 | pk forces one
// (tests, A/B profiling).
static int gemm_engine(const GemmArgs& g) {
    const char* e = getenv("CTR_GEMM");
    if (e && e[0] == 's') return 0;
    if (e && e[0] == 't') return 1;
    if (e && e[0] == 'p') return 2;
    const double macs = (double)g.M * (double)g.N * (double)g.K;
    if (macs < 2097152.0) return 0;
    return gemm_pk_has_scratch(g.M, g.N, g.K) ? 2 : 1;
}

int launch_sgemm(const GemmArgs& g, cudaStream_t st) {
    if (g.M <= 0 || g.N <= 0) return 0;
    if (g.K <= 0) {
        // empty contraction: result is the epilogue of zero — only the plain store case is used
        if (!g.accumulate && g.epilogue == EPI_STORE)
            CTR_CUDA(cudaMemset2DAsync(g.C, g.ldc * sizeof(float), 0, g.N * sizeof(float), g.M, st));
        return 0;
    }
    const int engine = gemm_engine(g);
    if (engine == 2) {
        const int rc = launch_gemm_pk(g, st);
        if (rc != -3) return rc;          // -3: no plan for this shape in the streaming engine (gemm_pk.cu)
        return launch_gemm_tc(g, st);
    }
    if (engine == 1) return launch_gemm_tc(g, st);
    const bool big = (g.M >= 128 && g.N >= 96) || (g.N >= 128 && g.M >= 96);
    const int BM = big ? 128 : 64, BN = big ? 128 : 64, BK = 16;
    const int64_t gm = ceil_div64(g.M, BM), gn = ceil_div64(g.N, BN);
    int64_t splits = 1;
    if (g.allow_split_k && g.epilogue == EPI_STORE) {
        const int64_t tiles = gm * gn, target = 2LL * ctr_sm_count();
        if (tiles < target && g.K >= 8 * BK) {
            splits = ceil_div64(target, tiles);
            const int64_t max_splits = g.K / (4 * BK);
            if (splits > max_splits) splits = max_splits;
            if (splits < 1) splits = 1;
        }
    }
    int64_t k_chunk = ceil_div64(ceil_div64(g.K, splits), BK) * BK;
    splits = ceil_div64(g.K, k_chunk);
    if (splits > 1 && !g.accumulate)
        CTR_CUDA(cudaMemset2DAsync(g.C, g.ldc * sizeof(float), 0, g.N * sizeof(float), g.M, st));
    if (gm > 65535 || splits > 65535) {
        ctr_set_error("launch_sgemm: grid too large (M tiles %lld, splits %lld)", (long long)gm, (long long)splits);
        return -1;
    }
    dim3 grid((unsigned)gn, (unsigned)gm, (unsigned)splits);
    if (big) return dispatch_layout<128, 128, 16, 8, 8>(g, grid, k_chunk, st);
    return dispatch_layout<64, 64, 16, 4, 4>(g, grid, k_chunk, st);
}

int launch_colsum(const float* A, int64_t sam, int64_t san, const float* mask, int64_t smm,
                  int64_t smn, int mask_act, const float* w, int64_t M, int64_t N, float* out,
                  cudaStream_t st) {
    CTR_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * N, st));
    if (M <= 0 || N <= 0) return 0;
    const int64_t n4 = (N + 3) / 4 * 4;          // rows padded to a multiple of 4 floats may be read up to the padding
    if (san == 1 && sam % 4 == 0 && sam >= n4 && (reinterpret_cast<uintptr_t>(A) & 15) == 0 &&
        (!mask || (smn == 1 && smm % 4 == 0 && smm >= n4 && (reinterpret_cast<uintptr_t>(mask) & 15) == 0))) {
        const int nq = (int)(n4 / 4);
        int xq = 1;
        while (xq < nq && xq < 64) xq <<= 1;
        const int ylanes = 256 / xq;
        const int64_t gx = ceil_div64(nq, xq);
        int64_t gy = ceil_div64(4LL * ctr_sm_count(), gx);
        const int64_t max_gy = ceil_div64(M, (int64_t)ylanes * 8);
        if (gy > max_gy) gy = max_gy;
        if (gy < 1) gy = 1;
        if (gy > 65535) gy = 65535;
        const int64_t rows_per_block = ceil_div64(M, gy);
        gy = ceil_div64(M, rows_per_block);
        dim3 grid((unsigned)gx, (unsigned)gy), block(xq, ylanes);
        colsum_vec_kernel<<<grid, block, sizeof(float4) * 256, st>>>(A, sam, mask, smm, mask_act, w, M, nq,
                                                                     rows_per_block, out, (int)N);
        CTR_LAUNCH_OK("colsum_vec_kernel");
        return 0;
    }
    const int64_t gx = ceil_div64(N, 32);
    int64_t gy = ceil_div64(4LL * ctr_sm_count(), gx);
    const int64_t max_gy = ceil_div64(M, 64);
    if (gy > max_gy) gy = max_gy;
    if (gy < 1) gy = 1;
    if (gy > 65535) gy = 65535;
    const int64_t rows_per_block = ceil_div64(M, gy);
    gy = ceil_div64(M, rows_per_block);
    dim3 grid((unsigned)gx, (unsigned)gy);
    colsum_kernel<<<grid, 256, 0, st>>>(A, sam, san, mask, smm, smn, mask_act, w, M, N, rows_per_block, out);
    CTR_LAUNCH_OK("colsum_kernel");
    return 0;
}

// =============================================================================================
// C ABI
// =============================================================================================
extern "C" int ctr_sgemm(int64_t M, int64_t N, int64_t K, const float* A, int64_t sam, int64_t sak,
                         const float* Bm, int64_t sbn, int64_t sbk, float* C, int64_t ldc,
                         int accumulate, void* stream) {
    CTR_ARG(M >= 0 && N >= 0 && K >= 0 && C && (K == 0 || (A && Bm)), "ctr_sgemm: bad arguments");
    GemmArgs g = gemm_args_default();
    g.M = M; g.N = N; g.K = K;
    g.A = A; g.sam = sam; g.sak = sak;
    g.B = Bm; g.sbn = sbn; g.sbk = sbk;
    g.C = C; g.ldc = ldc; g.accumulate = accumulate;
    g.allow_split_k = 1;
    return launch_sgemm(g, as_stream(stream));
}

extern "C" int ctr_dnn_layer_fwd(const float* X, int64_t ldx, const float* W, int64_t swn,
                                 int64_t swk, const float* bias, float* Y, int64_t ldy, int64_t B,
                                 int K, int N, int act, void* stream) {
    CTR_ARG(X && W && Y && B >= 0 && K > 0 && N > 0, "ctr_dnn_layer_fwd: bad arguments");
    CTR_ARG(ldx >= K && ldy >= N, "ctr_dnn_layer_fwd: leading dimension too small");
    CTR_ARG(act >= CTR_ACT_LINEAR && act <= CTR_ACT_TANH, "ctr_dnn_layer_fwd: unknown activation %d", act);
    GemmArgs g = gemm_args_default();
    g.M = B; g.N = N; g.K = K;
    g.A = X; g.sam = ldx; g.sak = 1;
    g.B = W; g.sbn = swn; g.sbk = swk;
    g.C = Y; g.ldc = ldy;
    g.epilogue = EPI_BIAS_ACT; g.act = act; g.bias = bias;
    return launch_sgemm(g, as_stream(stream));
}

// dW[n, k] = sum_b dZ[b, n] X[b, k], dW row-major [N, K] (nn.Linear layout)
static GemmArgs wgrad_args(const float* X, int64_t ldx, const float* mask, int64_t ldy, const float* dY, int64_t lddy,
                           float* dW, int64_t sdwn, int64_t B, int K, int N, int act) {
    GemmArgs g = gemm_args_default();
    g.K = B;
    g.allow_split_k = 1;
    g.M = N; g.N = K;
    g.A = dY; g.sam = 1; g.sak = lddy;
    g.amask = mask; g.smm = 1; g.smk = ldy; g.amask_act = act;
    g.B = X; g.sbn = 1; g.sbk = ldx;
    g.C = dW; g.ldc = sdwn;
    return g;
}

static int dnn_layer_bwd_impl(const float* X, int64_t ldx, const float* W, int64_t swn, int64_t swk,
                              const float* Y, int64_t ldy, const float* dY, int64_t lddy, float* dX,
                              int64_t lddx, int accumulate_dx, float* dW, int64_t sdwn, int64_t sdwk,
                              float* db, int64_t B, int K, int N, int act, int dy_is_dz, int dx_act,
                              cudaStream_t st) {
    // dZ = dY (.) act'(Y) is applied on the fly (operand prologue) unless the caller already holds dZ
    const float* mask = (act == CTR_ACT_LINEAR || dy_is_dz) ? nullptr : Y;
    int rc;
    if (dX) {  // dX[b,k] (+)= sum_n dZ[b,n] W(n,k)   [ (.) dx_act'(X[b,k]) : the previous layer's dZ ]
        GemmArgs g = gemm_args_default();
        g.M = B; g.N = K; g.K = N;
        g.A = dY; g.sam = lddy; g.sak = 1;
        g.amask = mask; g.smm = ldy; g.smk = 1; g.amask_act = act;
        g.B = W; g.sbn = swk; g.sbk = swn;
        g.C = dX; g.ldc = lddx; g.accumulate = accumulate_dx;
        if (dx_act != CTR_ACT_LINEAR) {
            g.epilogue = EPI_MUL_ACTGRAD;
            g.act = dx_act;
            g.aux = X; g.ldaux = ldx;
        }
        if ((rc = launch_sgemm(g, st)) != 0) return rc;
    }
    if (dW) {  // dW(n,k) = sum_b dZ[b,n] X[b,k]
        GemmArgs g = gemm_args_default();
        g.K = B;
        g.allow_split_k = 1;
        if (sdwk == 1) {  // rows = n, cols = k
            g = wgrad_args(X, ldx, mask, ldy, dY, lddy, dW, sdwn, B, K, N, act);
            // TSW engine: dZ^T through tensor memory, bias gradient fused (no colsum launch)
            rc = launch_gemm_tsw(g, db, st);
            if (rc == 0) return 0;
            if (rc != -3) return rc;
        } else {          // transposed storage: rows = k, cols = n
            g.M = K; g.N = N;
            g.A = X; g.sam = 1; g.sak = ldx;
            g.B = dY; g.sbn = 1; g.sbk = lddy;
            g.bmask = mask; g.sbmn = 1; g.sbmk = ldy; g.bmask_act = act;
            g.C = dW; g.ldc = sdwk;
        }
        if ((rc = launch_sgemm(g, st)) != 0) return rc;
    }
    if (db) {
        if ((rc = launch_colsum(dY, lddy, 1, mask, ldy, 1, act, nullptr, B, N, db, st)) != 0) return rc;
    }
    return 0;
}

extern "C" int ctr_dnn_layer_bwd(const float* X, int64_t ldx, const float* W, int64_t swn,
                                 int64_t swk, const float* Y, int64_t ldy, const float* dY,
                                 int64_t lddy, float* dX, int64_t lddx, int accumulate_dx, float* dW,
                                 int64_t sdwn, int64_t sdwk, float* db, int64_t B, int K, int N,
                                 int act, void* stream) {
    CTR_ARG(W && dY && B >= 0 && K > 0 && N > 0, "ctr_dnn_layer_bwd: bad arguments");
    CTR_ARG(act == CTR_ACT_LINEAR || Y, "ctr_dnn_layer_bwd: Y required for a non-linear activation");
    CTR_ARG(!dW || X, "ctr_dnn_layer_bwd: X required for dW");
    CTR_ARG(!dW || sdwn == 1 || sdwk == 1, "ctr_dnn_layer_bwd: dW must be contiguous along n or k");
    return dnn_layer_bwd_impl(X, ldx, W, swn, swk, Y, ldy, dY, lddy, dX, lddx, accumulate_dx, dW, sdwn, sdwk, db, B,
                              K, N, act, 0, CTR_ACT_LINEAR, as_stream(stream));
}

extern "C" int ctr_dnn_wgrad_is_scratch_free(const float* X, int64_t ldx, const float* Y, int64_t ldy, const float* dY,
                                             int64_t lddy, const float* dW, int64_t sdwn, int64_t sdwk, int64_t B, int K,
                                             int N, int act, int dy_is_dz) {
    if (!X || !dY || !dW || sdwk != 1 || B <= 0 || K <= 0 || N <= 0) return 0;
    const float* mask = (act == CTR_ACT_LINEAR || dy_is_dz) ? nullptr : Y;
    return gemm_tsw_eligible(wgrad_args(X, ldx, mask, ldy, dY, lddy, const_cast<float*>(dW), sdwn, B, K, N, act)) ? 1 : 0;
}

extern "C" int ctr_dnn_layer_bwd_chain(const float* X, int64_t ldx, const float* W, int64_t swn,
                                       int64_t swk, const float* Y, int64_t ldy, const float* dY,
                                       int64_t lddy, float* dX, int64_t lddx, float* dW, int64_t sdwn,
                                       int64_t sdwk, float* db, int64_t B, int K, int N, int act,
                                       int dy_is_dz, int dx_act, void* stream) {
    CTR_ARG(W && dY && B >= 0 && K > 0 && N > 0, "ctr_dnn_layer_bwd_chain: bad arguments");
    CTR_ARG(act == CTR_ACT_LINEAR || dy_is_dz || Y, "ctr_dnn_layer_bwd_chain: Y required for a non-linear activation");
    CTR_ARG(!dW || X, "ctr_dnn_layer_bwd_chain: X required for dW");
    CTR_ARG(dx_act == CTR_ACT_LINEAR || (X && dX), "ctr_dnn_layer_bwd_chain: X and dX required when dx_act is set");
    CTR_ARG(dx_act >= CTR_ACT_LINEAR && dx_act <= CTR_ACT_TANH, "ctr_dnn_layer_bwd_chain: unknown dx_act %d", dx_act);
    CTR_ARG(!dW || sdwn == 1 || sdwk == 1, "ctr_dnn_layer_bwd_chain: dW must be contiguous along n or k");
    return dnn_layer_bwd_impl(X, ldx, W, swn, swk, Y, ldy, dY, lddy, dX, lddx, 0, dW, sdwn, sdwk, db, B, K, N, act,
                              dy_is_dz, dx_act, as_stream(stream));
}

extern "C" int ctr_rowdot_fwd(const float* H, int64_t ldh, const float* w, int64_t B, int N,
                              float* out, int accumulate, void* stream) {
    CTR_ARG(H && w && out && B >= 0 && N > 0 && ldh >= N, "ctr_rowdot_fwd: bad arguments");
    if (B == 0) return 0;
    int64_t blocks = ceil_div64(B, 8);
    const int64_t cap = (int64_t)ctr_sm_count() * 8;
    if (blocks > cap) blocks = cap;
    rowdot_fwd_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(H, ldh, w, B, N, out, accumulate);
    CTR_LAUNCH_OK("rowdot_fwd_kernel");
    return 0;
}

extern "C" int ctr_rowdot_bwd(const float* H, int64_t ldh, const float* w, const float* g, int64_t B,
                              int N, float* dH, int64_t lddh, int accumulate_dh, float* dw,
                              void* stream) {
    CTR_ARG(w && g && B >= 0 && N > 0, "ctr_rowdot_bwd: bad arguments");
    CTR_ARG(!dw || H, "ctr_rowdot_bwd: H required for dw");
    cudaStream_t st = as_stream(stream);
    if (dH && B > 0) {
        // vector form: 16-byte aligned rows whose padding up to round_up(N, 4) exists (it is written with zeros)
        const int64_t n4 = ((int64_t)N + 3) / 4 * 4;
        const bool vec = (lddh % 4 == 0) && lddh >= n4 && ((reinterpret_cast<uintptr_t>(dH) & 15) == 0);
        const int nq = vec ? (int)(n4 / 4) : N;
        int nq_pad = 1;
        while (nq_pad < nq && nq_pad < 256) nq_pad <<= 1;
        const int rows_per_block = 256 / nq_pad;
        int64_t blocks = ceil_div64(B, (int64_t)rows_per_block * 8);
        const int64_t cap = (int64_t)ctr_sm_count() * 8;
        if (blocks > cap) blocks = cap;
        if (blocks < 1) blocks = 1;
        if (vec) rowdot_bwd_dh_kernel<true><<<(unsigned)blocks, 256, 0, st>>>(w, g, B, N, nq_pad, dH, lddh, accumulate_dh);
        else rowdot_bwd_dh_kernel<false><<<(unsigned)blocks, 256, 0, st>>>(w, g, B, N, nq_pad, dH, lddh, accumulate_dh);
        CTR_LAUNCH_OK("rowdot_bwd_dh_kernel");
    }
    if (dw) return launch_colsum(H, ldh, 1, nullptr, 0, 0, 0, g, B, N, dw, st);
    return 0;
}

extern "C" int ctr_predict_fwd(const float* const* terms, int n_terms, const float* bias, int64_t B,
                               int task_binary, float* logit, float* y, void* stream) {
    CTR_ARG(n_terms >= 0 && n_terms <= 8 && y && B >= 0, "ctr_predict_fwd: bad arguments (<= 8 terms)");
    CTR_ARG(n_terms == 0 || terms, "ctr_predict_fwd: terms missing");
    if (B == 0) return 0;
    TermList t{};
    t.n = n_terms;
    for (int i = 0; i < n_terms; ++i) {
        CTR_ARG(terms[i], "ctr_predict_fwd: null term %d", i);
        t.p[i] = terms[i];
    }
    int64_t blocks = ceil_div64(B, 256);
    const int64_t cap = (int64_t)ctr_sm_count() * 8;
    if (blocks > cap) blocks = cap;
    predict_fwd_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(t, bias, B, task_binary, logit, y);
    CTR_LAUNCH_OK("predict_fwd_kernel");
    return 0;
}

extern "C" int ctr_predict_bwd(const float* y, const float* dy, int64_t B, int task_binary,
                               float* dlogit, float* dbias, void* stream) {
    CTR_ARG(y && dy && dlogit && B >= 0, "ctr_predict_bwd: bad arguments");
    cudaStream_t st = as_stream(stream);
    if (dbias) CTR_CUDA(cudaMemsetAsync(dbias, 0, sizeof(float), st));
    if (B == 0) return 0;
    int64_t blocks = ceil_div64(B, 256 * 4);
    const int64_t cap = (int64_t)ctr_sm_count() * 4;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    predict_bwd_kernel<<<(unsigned)blocks, 256, 0, st>>>(y, dy, B, task_binary, dlogit, dbias);
    CTR_LAUNCH_OK("predict_bwd_kernel");
    return 0;
}
