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
