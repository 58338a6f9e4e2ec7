// CrossNet (vector / matrix) and the fused pieces of CrossNetMix.
// Reference: layers/interaction.py:438-453 (CrossNet), :499-534 (CrossNetMix).
//
// vector form is HBM-bound: all L layers run in one kernel with x0 / x_l resident in shared
// memory (one warp per sample), 2 reads + 1 write of the [B,n] activations in total.  The
// backward keeps lane-private weight-gradient accumulators in shared memory (no atomics in the
// sample loop) and issues one global fp32 atomic per (block, parameter) at the end.
#include "gemm.cuh"

namespace {

constexpr int kMaxSmemBytes = 200 * 1024;

__global__ void cross_vector_fwd_kernel(const float* __restrict__ x0, int64_t ldx,
                                        const float* __restrict__ kernels,
                                        const float* __restrict__ bias, int L, int n, float* out,
                                        int64_t ldo, float* s, int64_t B) {
    extern __shared__ float smem[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    float* s_x0 = smem + (size_t)wid * 2 * n;
    float* s_xl = s_x0 + n;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t b = warp0; b < B; b += nwarps) {
        for (int i = lane; i < n; i += 32) {
            const float v = __ldg(x0 + b * ldx + i);
            s_x0[i] = v;
            s_xl[i] = v;
        }
        __syncwarp();
        for (int l = 0; l < L; ++l) {
            float dot = 0.f;
            for (int i = lane; i < n; i += 32) dot = fmaf(s_xl[i], __ldg(kernels + (size_t)l * n + i), dot);
            dot = warp_sum(dot);
            if (lane == 0 && s) s[b * L + l] = dot;
            for (int i = lane; i < n; i += 32)
                s_xl[i] = s_x0[i] * dot + __ldg(bias + (size_t)l * n + i) + s_xl[i];
            __syncwarp();
        }
        for (int i = lane; i < n; i += 32) out[b * ldo + i] = s_xl[i];
        __syncwarp();
    }
}

// per warp shared layout: x[L][n] (x_0..x_{L-1}), g[n], dx0[n], accw[L][n], accb[L][n]
__global__ void cross_vector_bwd_kernel(const float* __restrict__ x0, int64_t ldx,
                                        const float* __restrict__ kernels,
                                        const float* __restrict__ bias, int L, int n,
                                        const float* __restrict__ s,
                                        const float* __restrict__ dout, int64_t lddo, float* dx0,
                                        int64_t lddx, int accumulate_dx, float* dkernels,
                                        float* dbias, int64_t B) {
    extern __shared__ float smem[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const size_t per_warp = (size_t)(3 * L + 2) * n;
    float* s_x = smem + wid * per_warp;
    float* s_g = s_x + (size_t)L * n;
    float* s_dx0 = s_g + n;
    float* s_accw = s_dx0 + n;
    float* s_accb = s_accw + (size_t)L * n;
    for (int i = lane; i < L * n; i += 32) {
        s_accw[i] = 0.f;
        s_accb[i] = 0.f;
    }
    __syncwarp();
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t b = warp0; b < B; b += nwarps) {
        // recompute x_0 .. x_{L-1} from the saved dots
        for (int i = lane; i < n; i += 32) {
            float v = __ldg(x0 + b * ldx + i);
            const float v0 = v;
            s_x[i] = v;
            for (int l = 1; l < L; ++l) {
                v = v0 * __ldg(s + b * L + (l - 1)) + __ldg(bias + (size_t)(l - 1) * n + i) + v;
                s_x[(size_t)l * n + i] = v;
            }
            s_g[i] = __ldg(dout + b * lddo + i);
            s_dx0[i] = 0.f;
        }
        __syncwarp();
        for (int l = L - 1; l >= 0; --l) {
            float ds = 0.f;
            for (int i = lane; i < n; i += 32) ds = fmaf(s_g[i], s_x[i], ds);  // g . x_0
            ds = warp_sum(ds);
            const float sl = __ldg(s + b * L + l);
            for (int i = lane; i < n; i += 32) {
                const float g = s_g[i];
                s_accw[(size_t)l * n + i] += s_x[(size_t)l * n + i] * ds;
                s_accb[(size_t)l * n + i] += g;
                s_dx0[i] += g * sl;
                s_g[i] = g + __ldg(kernels + (size_t)l * n + i) * ds;
            }
            __syncwarp();
        }
        for (int i = lane; i < n; i += 32) {
            const float v = s_dx0[i] + s_g[i];
            float* p = dx0 + b * lddx + i;
            *p = accumulate_dx ? *p + v : v;
        }
        __syncwarp();
    }
    __syncthreads();
    // reduce the per-warp accumulators of this block and publish with one atomic per entry
    const int nw = blockDim.x >> 5;
    for (int i = threadIdx.x; i < L * n; i += blockDim.x) {
        float aw = 0.f, ab = 0.f;
        for (int w = 0; w < nw; ++w) {
            const float* base = smem + w * per_warp + (size_t)(L + 2) * n;
            aw += base[i];
            ab += base[(size_t)L * n + i];
        }
        atomicAdd(dkernels + i, aw);
        atomicAdd(dbias + i, ab);
    }
}

// ---------------------------------------------------------------------------------------------
// Register-resident CrossNet-vector kernels (round 2).  One warp per sample; lane `lane` owns the
// float4 quads q = j*32 + lane (j < NQ) of the padded row, i.e. up to 16 coordinates, in REGISTERS for all
// L layers: x0, x_l, g, dx0 and — backward — the lane-private partial sums of dkernels / dbias over every
// sample the warp processes.  128-bit coalesced loads/stores, no shared-memory round trips in the sample
// loop (the first version kept everything in per-warp shared memory with scalar accesses: 0.58 ms for the
// backward at [65536, 429] against a 52 us HBM floor).  Weights live in shared memory ([2L][4*32*NQ] floats).
// ---------------------------------------------------------------------------------------------
template <int NQ>
__device__ __forceinline__ void cv_load_row(const float* row, int n, int lane, float (&v)[NQ * 4]) {
#pragma unroll
    for (int j = 0; j < NQ; ++j) {
        const int c0 = (j * 32 + lane) * 4;
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c0 < n) t = ld_stream4(row + c0);             // padding up to round_up(n,4) is readable (host check)
        v[4 * j + 0] = t.x;
        v[4 * j + 1] = (c0 + 1 < n) ? t.y : 0.f;
        v[4 * j + 2] = (c0 + 2 < n) ? t.z : 0.f;
        v[4 * j + 3] = (c0 + 3 < n) ? t.w : 0.f;
    }
}
template <int NQ>
__device__ __forceinline__ void cv_store_row(float* row, int n, int lane, const float (&v)[NQ * 4], bool accumulate) {
#pragma unroll
    for (int j = 0; j < NQ; ++j) {
        const int c0 = (j * 32 + lane) * 4;
        if (c0 < n) {
            float4 t = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            if (accumulate) {
                const float4 o = *reinterpret_cast<const float4*>(row + c0);
                t.x += o.x; t.y += o.y; t.z += o.z; t.w += o.w;
            }
            st_stream4(row + c0, t);
        }
    }
}

template <int NQ, int L>
__global__ void __launch_bounds__(256) cross_vector_fwd_reg_kernel(const float* __restrict__ x0, int64_t ldx,
                                                                   const float* __restrict__ kernels,
                                                                   const float* __restrict__ bias, int n, float* out,
                                                                   int64_t ldo, float* s, int64_t B) {
    constexpr int W = NQ * 128;                        // padded row width held by a warp
    __shared__ __align__(16) float s_w[L * W], s_b[L * W];
    for (int i = threadIdx.x; i < L * W; i += blockDim.x) {
        const int l = i / W, c = i - l * W;
        s_w[i] = c < n ? kernels[(size_t)l * n + c] : 0.f;
        s_b[i] = c < n ? bias[(size_t)l * n + c] : 0.f;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t b = warp0; b < B; b += nwarps) {
        float a0[NQ * 4], xl[NQ * 4];
        cv_load_row<NQ>(x0 + b * ldx, n, lane, a0);
#pragma unroll
        for (int i = 0; i < NQ * 4; ++i) xl[i] = a0[i];
#pragma unroll
        for (int l = 0; l < L; ++l) {
            float dot = 0.f;
#pragma unroll
            for (int j = 0; j < NQ; ++j) {
                const float4 w4 = *reinterpret_cast<const float4*>(s_w + l * W + (j * 32 + lane) * 4);
                dot = fmaf(xl[4 * j], w4.x, dot);
                dot = fmaf(xl[4 * j + 1], w4.y, dot);
                dot = fmaf(xl[4 * j + 2], w4.z, dot);
                dot = fmaf(xl[4 * j + 3], w4.w, dot);
            }
            dot = warp_sum(dot);
            if (lane == 0 && s) s[b * L + l] = dot;
#pragma unroll
            for (int j = 0; j < NQ; ++j) {
                const float4 b4 = *reinterpret_cast<const float4*>(s_b + l * W + (j * 32 + lane) * 4);
                xl[4 * j] = a0[4 * j] * dot + b4.x + xl[4 * j];
                xl[4 * j + 1] = a0[4 * j + 1] * dot + b4.y + xl[4 * j + 1];
                xl[4 * j + 2] = a0[4 * j + 2] * dot + b4.z + xl[4 * j + 2];
                xl[4 * j + 3] = a0[4 * j + 3] * dot + b4.w + xl[4 * j + 3];
            }
        }
        cv_store_row<NQ>(out + b * ldo, n, lane, xl, false);
    }
}

template <int NQ, int L>
__global__ void __launch_bounds__(256) cross_vector_bwd_reg_kernel(const float* __restrict__ x0, int64_t ldx,
                                                                   const float* __restrict__ kernels,
                                                                   const float* __restrict__ bias, int n,
                                                                   const float* __restrict__ s,
                                                                   const float* __restrict__ dout, int64_t lddo,
                                                                   float* dx0, int64_t lddx, int accumulate_dx,
                                                                   float* dkernels, float* dbias, int64_t B) {
    constexpr int W = NQ * 128;
    __shared__ __align__(16) float s_w[L * W], s_b[L * W], s_aw[L * W], s_ab[L * W];
    for (int i = threadIdx.x; i < L * W; i += blockDim.x) {
        const int l = i / W, c = i - l * W;
        s_w[i] = c < n ? kernels[(size_t)l * n + c] : 0.f;
        s_b[i] = c < n ? bias[(size_t)l * n + c] : 0.f;
        s_aw[i] = 0.f;
        s_ab[i] = 0.f;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    float accw[L][NQ * 4], accb[L][NQ * 4];
#pragma unroll
    for (int l = 0; l < L; ++l)
#pragma unroll
        for (int i = 0; i < NQ * 4; ++i) accw[l][i] = accb[l][i] = 0.f;
    for (int64_t b = warp0; b < B; b += nwarps) {
        float a0[NQ * 4], xl[NQ * 4], g[NQ * 4], d0[NQ * 4], sl[L];
        cv_load_row<NQ>(x0 + b * ldx, n, lane, a0);
        cv_load_row<NQ>(dout + b * lddo, n, lane, g);
#pragma unroll
        for (int l = 0; l < L; ++l) sl[l] = __ldg(s + b * L + l);
        // x_{L-1} by the forward recurrence from the saved dots (x_l = x_0 s_{l-1} + b_{l-1} + x_{l-1})
#pragma unroll
        for (int i = 0; i < NQ * 4; ++i) {
            xl[i] = a0[i];
            d0[i] = 0.f;
        }
#pragma unroll
        for (int l = 1; l < L; ++l)
#pragma unroll
            for (int j = 0; j < NQ; ++j) {
                const float4 b4 = *reinterpret_cast<const float4*>(s_b + (l - 1) * W + (j * 32 + lane) * 4);
                xl[4 * j] += a0[4 * j] * sl[l - 1] + b4.x;
                xl[4 * j + 1] += a0[4 * j + 1] * sl[l - 1] + b4.y;
                xl[4 * j + 2] += a0[4 * j + 2] * sl[l - 1] + b4.z;
                xl[4 * j + 3] += a0[4 * j + 3] * sl[l - 1] + b4.w;
            }
#pragma unroll
        for (int l = L - 1; l >= 0; --l) {
            float ds = 0.f;
#pragma unroll
            for (int i = 0; i < NQ * 4; ++i) ds = fmaf(g[i], a0[i], ds);       // g . x_0
            ds = warp_sum(ds);
#pragma unroll
            for (int j = 0; j < NQ; ++j) {
                const float4 w4 = *reinterpret_cast<const float4*>(s_w + l * W + (j * 32 + lane) * 4);
                const float wv[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int i = 4 * j + c;
                    accw[l][i] += xl[i] * ds;
                    accb[l][i] += g[i];
                    d0[i] += g[i] * sl[l];
                    g[i] += wv[c] * ds;
                }
            }
            if (l > 0) {                                   // step back to x_{l-1}
#pragma unroll
                for (int j = 0; j < NQ; ++j) {
                    const float4 b4 = *reinterpret_cast<const float4*>(s_b + (l - 1) * W + (j * 32 + lane) * 4);
                    xl[4 * j] -= a0[4 * j] * sl[l - 1] + b4.x;
                    xl[4 * j + 1] -= a0[4 * j + 1] * sl[l - 1] + b4.y;
                    xl[4 * j + 2] -= a0[4 * j + 2] * sl[l - 1] + b4.z;
                    xl[4 * j + 3] -= a0[4 * j + 3] * sl[l - 1] + b4.w;
                }
            }
        }
#pragma unroll
        for (int i = 0; i < NQ * 4; ++i) d0[i] += g[i];
        cv_store_row<NQ>(dx0 + b * lddx, n, lane, d0, accumulate_dx != 0);
    }
    // lane-private partial sums -> block accumulators (lanes own distinct columns) -> one global atomic per entry
#pragma unroll
    for (int l = 0; l < L; ++l)
#pragma unroll
        for (int j = 0; j < NQ; ++j)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int col = (j * 32 + lane) * 4 + c;
                atomicAdd(&s_aw[l * W + col], accw[l][4 * j + c]);
                atomicAdd(&s_ab[l * W + col], accb[l][4 * j + c]);
            }
    __syncthreads();
    for (int i = threadIdx.x; i < L * W; i += blockDim.x) {
        const int l = i / W, c = i - l * W;
        if (c < n) {
            atomicAdd(dkernels + (size_t)l * n + c, s_aw[i]);
            atomicAdd(dbias + (size_t)l * n + c, s_ab[i]);
        }
    }
}

bool cv_vec_ok(const void* p, int64_t ld, int n) {
    return p && ld % 4 == 0 && ld >= ((int64_t)n + 3) / 4 * 4 && (reinterpret_cast<uintptr_t>(p) & 15) == 0;
}

// dU = g (.) x0 ; dx0 += g (.) U ; gprev = g
__global__ void __launch_bounds__(256) cross_matrix_bwd_elem_kernel(
    const float* __restrict__ x0, int64_t ldx0, const float* __restrict__ U, int64_t ldu,
    const float* __restrict__ g, int64_t ldg, int n, float* dU, int64_t lddu, float* dx0,
    int64_t lddx0, float* gprev, int64_t ldgp, int64_t B) {
    const int64_t total = B * n;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = i / n;
        const int j = (int)(i - b * n);
        const float gv = g[b * ldg + j];
        dU[b * lddu + j] = gv * x0[b * ldx0 + j];
        dx0[b * lddx0 + j] += gv * U[b * ldu + j];
        gprev[b * ldgp + j] = gv;
    }
}

constexpr int kMaxExperts = 16;

__global__ void __launch_bounds__(256) cross_mix_fwd_kernel(
    const float* __restrict__ x0, int64_t ldx0, const float* __restrict__ xl, int64_t ldxl,
    const float* __restrict__ uv, int64_t ldu, const float* __restrict__ gate,
    const float* __restrict__ bias, int E, int n, float* xnext, int64_t ldn, int64_t B) {
    const int64_t total = B * n;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = i / n;
        const int j = (int)(i - b * n);
        float p[kMaxExperts];
        float mx = -INFINITY;
        for (int e = 0; e < E; ++e) {
            p[e] = __ldg(gate + b * E + e);
            mx = fmaxf(mx, p[e]);
        }
        float den = 0.f;
        for (int e = 0; e < E; ++e) {
            p[e] = expf(p[e] - mx);
            den += p[e];
        }
        const float x0v = x0[b * ldx0 + j], bj = __ldg(bias + j);
        float acc = 0.f;
        for (int e = 0; e < E; ++e)
            acc += (p[e] / den) * (x0v * (uv[((int64_t)e * B + b) * ldu + j] + bj));
        xnext[b * ldn + j] = acc + xl[b * ldxl + j];
    }
}

// one warp per sample: d uv_e = p_e g x0 ; dx0 += sum_e p_e g (uv_e + bias) ; dgate via softmax
__global__ void __launch_bounds__(256) cross_mix_bwd_kernel(
    const float* __restrict__ x0, int64_t ldx0, const float* __restrict__ uv, int64_t ldu,
    const float* __restrict__ gate, const float* __restrict__ bias, const float* __restrict__ g,
    int64_t ldg, int E, int n, float* duv, float* dgate, float* dx0, int64_t lddx0, int64_t B) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t b = warp0; b < B; b += nwarps) {
        float p[kMaxExperts], dp[kMaxExperts];
        float mx = -INFINITY, den = 0.f;
        for (int e = 0; e < E; ++e) {
            p[e] = __ldg(gate + b * E + e);
            mx = fmaxf(mx, p[e]);
        }
        for (int e = 0; e < E; ++e) {
            p[e] = expf(p[e] - mx);
            den += p[e];
        }
        for (int e = 0; e < E; ++e) {
            p[e] /= den;
            dp[e] = 0.f;
        }
        for (int j = lane; j < n; j += 32) {
            const float gv = g[b * ldg + j], x0v = x0[b * ldx0 + j], bj = __ldg(bias + j);
            float dx = 0.f;
            for (int e = 0; e < E; ++e) {
                const int64_t idx = ((int64_t)e * B + b) * ldu + j;
                const float t = uv[idx] + bj;
                duv[idx] = p[e] * gv * x0v;
                dx += p[e] * gv * t;
                dp[e] += gv * x0v * t;
            }
            dx0[b * lddx0 + j] += dx;
        }
        float dot = 0.f;
        for (int e = 0; e < E; ++e) {
            dp[e] = warp_sum(dp[e]);
            dot += p[e] * dp[e];
        }
        if (lane == 0)
            for (int e = 0; e < E; ++e) dgate[b * E + e] = p[e] * (dp[e] - dot);
    }
}

unsigned elem_grid(int64_t total) {
    int64_t blocks = ceil_div64(total, 256 * 4);
    const int64_t cap = (int64_t)ctr_sm_count() * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (unsigned)blocks;
}

}  // namespace

extern "C" int ctr_cross_vector_fwd(const float* x0, int64_t ldx, const float* kernels,
                                    const float* bias, int L, int n, float* out, int64_t ldo,
                                    float* s, int64_t B, void* stream) {
    CTR_ARG(x0 && kernels && bias && out && L >= 0 && n > 0 && B >= 0, "ctr_cross_vector_fwd: bad arguments");
    if (B == 0) return 0;
    if (L >= 1 && L <= 4 && n <= 512 && cv_vec_ok(x0, ldx, n) && cv_vec_ok(out, ldo, n)) {
        cudaStream_t st = as_stream(stream);
        int64_t blocks = ceil_div64(B, 8 * 4);
        const int64_t cap = (int64_t)ctr_sm_count() * 4;
        if (blocks > cap) blocks = cap;
#define CV_FWD(NQ, LL) cross_vector_fwd_reg_kernel<NQ, LL><<<(unsigned)blocks, 256, 0, st>>>(x0, ldx, kernels, bias, n, out, ldo, s, B)
#define CV_FWD_L(NQ) switch (L) { case 1: CV_FWD(NQ, 1); break; case 2: CV_FWD(NQ, 2); break; case 3: CV_FWD(NQ, 3); break; default: CV_FWD(NQ, 4); break; }
        if (n <= 128) { CV_FWD_L(1) } else if (n <= 256) { CV_FWD_L(2) } else { CV_FWD_L(4) }
#undef CV_FWD_L
#undef CV_FWD
        CTR_LAUNCH_OK("cross_vector_fwd_reg_kernel");
        return 0;
    }
    int warps = 4;
    while (warps > 1 && (size_t)warps * 2 * n * sizeof(float) > (size_t)kMaxSmemBytes) warps >>= 1;
    const size_t smem = (size_t)warps * 2 * n * sizeof(float);
    CTR_ARG(smem <= (size_t)kMaxSmemBytes, "ctr_cross_vector_fwd: n=%d too large for shared memory", n);
    if (smem > 48 * 1024)
        CTR_CUDA(cudaFuncSetAttribute(cross_vector_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int64_t blocks = ceil_div64(B, warps);
    const int64_t cap = (int64_t)ctr_sm_count() * 8;
    if (blocks > cap) blocks = cap;
    cross_vector_fwd_kernel<<<(unsigned)blocks, warps * 32, smem, as_stream(stream)>>>(
        x0, ldx, kernels, bias, L, n, out, ldo, s, B);
    CTR_LAUNCH_OK("cross_vector_fwd_kernel");
    return 0;
}

extern "C" int ctr_cross_vector_bwd(const float* x0, int64_t ldx, const float* kernels,
                                    const float* bias, int L, int n, const float* s,
                                    const float* dout, int64_t lddo, float* dx0, int64_t lddx,
                                    int accumulate_dx, float* dkernels, float* dbias, int64_t B,
                                    void* stream) {
    CTR_ARG(x0 && kernels && bias && dout && dx0 && dkernels && dbias && L >= 0 && n > 0 && B >= 0,
            "ctr_cross_vector_bwd: bad arguments");
    CTR_ARG(L == 0 || s, "ctr_cross_vector_bwd: saved dots missing");
    cudaStream_t st = as_stream(stream);
    if (L > 0) {
        CTR_CUDA(cudaMemsetAsync(dkernels, 0, sizeof(float) * L * n, st));
        CTR_CUDA(cudaMemsetAsync(dbias, 0, sizeof(float) * L * n, st));
    }
    if (B == 0) return 0;
    if (L >= 1 && L <= 4 && n <= 512 && cv_vec_ok(x0, ldx, n) && cv_vec_ok(dout, lddo, n) && cv_vec_ok(dx0, lddx, n)) {
        int64_t blocks = ceil_div64(B, 8 * 16);            // >= 16 samples per warp amortise the final atomics
        const int64_t cap = (int64_t)ctr_sm_count() * 2;
        if (blocks > cap) blocks = cap;
        if (blocks < 1) blocks = 1;
#define CV_BWD(NQ, LL) cross_vector_bwd_reg_kernel<NQ, LL><<<(unsigned)blocks, 256, 0, st>>>(x0, ldx, kernels, bias, n, s, dout, lddo, dx0, lddx, accumulate_dx, dkernels, dbias, B)
#define CV_BWD_L(NQ) switch (L) { case 1: CV_BWD(NQ, 1); break; case 2: CV_BWD(NQ, 2); break; case 3: CV_BWD(NQ, 3); break; default: CV_BWD(NQ, 4); break; }
        if (n <= 128) { CV_BWD_L(1) } else if (n <= 256) { CV_BWD_L(2) } else { CV_BWD_L(4) }
#undef CV_BWD_L
#undef CV_BWD
        CTR_LAUNCH_OK("cross_vector_bwd_reg_kernel");
        return 0;
    }
    int warps = 4;
    const size_t per_warp = (size_t)(3 * L + 2) * n * sizeof(float);
    while (warps > 1 && warps * per_warp > (size_t)kMaxSmemBytes) warps >>= 1;
    const size_t smem = warps * per_warp;
    CTR_ARG(smem <= (size_t)kMaxSmemBytes, "ctr_cross_vector_bwd: L=%d n=%d too large for shared memory", L, n);
    if (smem > 48 * 1024)
        CTR_CUDA(cudaFuncSetAttribute(cross_vector_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int64_t blocks = ceil_div64(B, (int64_t)warps * 16);  // >= 16 samples per warp amortise the final atomics
    const int64_t cap = (int64_t)ctr_sm_count() * 4;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    cross_vector_bwd_kernel<<<(unsigned)blocks, warps * 32, smem, st>>>(
        x0, ldx, kernels, bias, L, n, s, dout, lddo, dx0, lddx, accumulate_dx, dkernels, dbias, B);
    CTR_LAUNCH_OK("cross_vector_bwd_kernel");
    return 0;
}

extern "C" int ctr_cross_matrix_layer_fwd(const float* x0, int64_t ldx0, const float* xl,
                                          int64_t ldxl, const float* W, const float* bias, int n,
                                          float* U, int64_t ldu, float* xnext, int64_t ldn,
                                          int64_t B, void* stream) {
    CTR_ARG(x0 && xl && W && bias && U && xnext && n > 0 && B >= 0, "ctr_cross_matrix_layer_fwd: bad arguments");
    GemmArgs g = gemm_args_default();
    g.M = B; g.N = n; g.K = n;
    g.A = xl; g.sam = ldxl; g.sak = 1;
    g.B = W; g.sbn = n; g.sbk = 1;          // U[b,i] = sum_j W[i,j] xl[b,j]
    g.C = xnext; g.ldc = ldn;
    g.epilogue = EPI_CROSS; g.bias = bias;
    g.aux = x0; g.ldaux = ldx0; g.aux2 = xl; g.ldaux2 = ldxl;
    g.out2 = U; g.ldout2 = ldu;
    return launch_sgemm(g, as_stream(stream));
}

extern "C" int ctr_cross_matrix_layer_bwd(const float* x0, int64_t ldx0, const float* xl,
                                          int64_t ldxl, const float* W, const float* U, int64_t ldu,
                                          const float* g, int64_t ldg, int n, float* dU,
                                          int64_t lddu, float* dx0, int64_t lddx0, float* dW,
                                          float* db, float* gprev, int64_t ldgp, int64_t B,
                                          void* stream) {
    CTR_ARG(x0 && xl && W && U && g && dU && dx0 && dW && db && gprev && n > 0 && B >= 0,
            "ctr_cross_matrix_layer_bwd: bad arguments");
    cudaStream_t st = as_stream(stream);
    int rc;
    if (B > 0) {
        cross_matrix_bwd_elem_kernel<<<elem_grid(B * n), 256, 0, st>>>(x0, ldx0, U, ldu, g, ldg, n, dU, lddu,
                                                                       dx0, lddx0, gprev, ldgp, B);
        CTR_LAUNCH_OK("cross_matrix_bwd_elem_kernel");
    }
    {   // dW[i,j] = sum_b dU[b,i] xl[b,j]
        GemmArgs a = gemm_args_default();
        a.M = n; a.N = n; a.K = B;
        a.A = dU; a.sam = 1; a.sak = lddu;
        a.B = xl; a.sbn = 1; a.sbk = ldxl;
        a.C = dW; a.ldc = n; a.allow_split_k = 1;
        if ((rc = launch_sgemm(a, st)) != 0) return rc;
    }
    if ((rc = launch_colsum(dU, lddu, 1, nullptr, 0, 0, 0, nullptr, B, n, db, st)) != 0) return rc;
    {   // gprev[b,j] += sum_i dU[b,i] W[i,j]
        GemmArgs a = gemm_args_default();
        a.M = B; a.N = n; a.K = n;
        a.A = dU; a.sam = lddu; a.sak = 1;
        a.B = W; a.sbn = 1; a.sbk = n;
        a.C = gprev; a.ldc = ldgp; a.accumulate = 1;
        if ((rc = launch_sgemm(a, st)) != 0) return rc;
    }
    return 0;
}

extern "C" int ctr_cross_mix_fwd(const float* x0, int64_t ldx0, const float* xl, int64_t ldxl,
                                 const float* uv, int64_t ldu, const float* gate, const float* bias,
                                 int E, int n, float* xnext, int64_t ldn, int64_t B, void* stream) {
    CTR_ARG(x0 && xl && uv && gate && bias && xnext && n > 0 && B >= 0, "ctr_cross_mix_fwd: bad arguments");
    CTR_ARG(E > 0 && E <= kMaxExperts, "ctr_cross_mix_fwd: 1 <= num_experts <= %d", kMaxExperts);
    if (B == 0) return 0;
    cross_mix_fwd_kernel<<<elem_grid(B * n), 256, 0, as_stream(stream)>>>(x0, ldx0, xl, ldxl, uv, ldu, gate,
                                                                         bias, E, n, xnext, ldn, B);
    CTR_LAUNCH_OK("cross_mix_fwd_kernel");
    return 0;
}

extern "C" int ctr_cross_mix_bwd(const float* x0, int64_t ldx0, const float* uv, int64_t ldu,
                                 const float* gate, const float* bias, const float* g, int64_t ldg,
                                 int E, int n, float* duv, float* dgate, float* dx0, int64_t lddx0,
                                 float* dbias, int64_t B, void* stream) {
    CTR_ARG(x0 && uv && gate && bias && g && duv && dgate && dx0 && dbias && n > 0 && B >= 0,
            "ctr_cross_mix_bwd: bad arguments");
    CTR_ARG(E > 0 && E <= kMaxExperts, "ctr_cross_mix_bwd: 1 <= num_experts <= %d", kMaxExperts);
    cudaStream_t st = as_stream(stream);
    if (B > 0) {
        int64_t blocks = ceil_div64(B, 8);
        const int64_t cap = (int64_t)ctr_sm_count() * 8;
        if (blocks > cap) blocks = cap;
        cross_mix_bwd_kernel<<<(unsigned)blocks, 256, 0, st>>>(x0, ldx0, uv, ldu, gate, bias, g, ldg, E, n,
                                                               duv, dgate, dx0, lddx0, B);
        CTR_LAUNCH_OK("cross_mix_bwd_kernel");
    }
    // dbias[j] = sum_b sum_e duv_e[b,j]  (sum_e p_e = 1)
    return launch_colsum(duv, ldu, 1, nullptr, 0, 0, 0, nullptr, (int64_t)E * B, n, dbias, st);
}
