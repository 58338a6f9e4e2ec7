// SENET layer and bilinear interaction (FiBiNET).
// Reference: layers/interaction.py:93-101 (SENETLayer.forward), :140-156 (BilinearInteraction).
#include <stdlib.h>

#include "gemm.cuh"

namespace {

// ---------------------------------------------------------------------------------------------
// SENET: one warp per sample; weights staged in shared memory; per-warp scratch Z[F], A1[R], A2[F]
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) senet_fwd_kernel(const float* __restrict__ E, int64_t se,
                                                        int F, int D, const float* __restrict__ W1,
                                                        const float* __restrict__ W2, int R,
                                                        float* V, int64_t sv, int64_t B) {
    extern __shared__ float smem[];
    float* s_w1 = smem;                 // [R][F]
    float* s_w2 = s_w1 + R * F;         // [F][R]
    const int nw = blockDim.x >> 5;
    float* s_scr = s_w2 + F * R;        // per warp: Z[F], A1[R], A2[F]
    for (int i = threadIdx.x; i < R * F; i += blockDim.x) {
        s_w1[i] = W1[i];
        s_w2[i] = W2[i];
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    float* z = s_scr + (size_t)wid * (2 * F + R);
    float* a1 = z + F;
    float* a2 = a1 + R;
    const float invD = 1.f / (float)D;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = (int64_t)gridDim.x * nw;
    for (int64_t b = warp0; b < B; b += nwarps) {
        const float* e = E + b * se;
        for (int f = lane; f < F; f += 32) {
            float acc = 0.f;
            for (int d = 0; d < D; ++d) acc += __ldg(e + f * D + d);
            z[f] = acc * invD;
        }
        __syncwarp();
        for (int r = lane; r < R; r += 32) {
            float acc = 0.f;
            for (int f = 0; f < F; ++f) acc = fmaf(s_w1[r * F + f], z[f], acc);
            a1[r] = fmaxf(acc, 0.f);
        }
        __syncwarp();
        for (int f = lane; f < F; f += 32) {
            float acc = 0.f;
            for (int r = 0; r < R; ++r) acc = fmaf(s_w2[f * R + r], a1[r], acc);
            a2[f] = fmaxf(acc, 0.f);
        }
        __syncwarp();
        float* v = V + b * sv;
        for (int i = lane; i < F * D; i += 32) v[i] = __ldg(e + i) * a2[i / D];
        __syncwarp();
    }
}

// per warp scratch: Z[F], A1[R], A2[F], dA2[F], dA1[R], dZ[F]; lane-private accumulators
// accW1[R*F], accW2[F*R] per warp.
__global__ void __launch_bounds__(128) senet_bwd_kernel(const float* __restrict__ E, int64_t se, int F,
                                                        int D, const float* __restrict__ W1,
                                                        const float* __restrict__ W2, int R,
                                                        const float* __restrict__ dV, int64_t sdv,
                                                        float* dE, int64_t sde, int accumulate_de,
                                                        float* dW1, float* dW2, int64_t B) {
    extern __shared__ float smem[];
    const int nw = blockDim.x >> 5;
    float* s_w1 = smem;
    float* s_w2 = s_w1 + R * F;
    float* s_rest = s_w2 + F * R;
    const size_t per_warp = (size_t)(4 * F + 2 * R) + 2 * (size_t)R * F;
    for (int i = threadIdx.x; i < R * F; i += blockDim.x) {
        s_w1[i] = W1[i];
        s_w2[i] = W2[i];
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    float* z = s_rest + wid * per_warp;
    float* a1 = z + F;
    float* a2 = a1 + R;
    float* da2 = a2 + F;
    float* da1 = da2 + F;
    float* dz = da1 + R;
    float* acc1 = dz + F;           // [R*F]
    float* acc2 = acc1 + R * F;     // [F*R]
    for (int i = lane; i < 2 * R * F; i += 32) acc1[i] = 0.f;
    __syncthreads();
    const float invD = 1.f / (float)D;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = (int64_t)gridDim.x * nw;
    for (int64_t b = warp0; b < B; b += nwarps) {
        const float* e = E + b * se;
        const float* dv = dV + b * sdv;
        for (int f = lane; f < F; f += 32) {
            float acc = 0.f, dot = 0.f;
            for (int d = 0; d < D; ++d) {
                const float ev = __ldg(e + f * D + d);
                acc += ev;
                dot = fmaf(__ldg(dv + f * D + d), ev, dot);
            }
            z[f] = acc * invD;
            da2[f] = dot;               // dL/dA2[f]
        }
        __syncwarp();
        for (int r = lane; r < R; r += 32) {
            float acc = 0.f;
            for (int f = 0; f < F; ++f) acc = fmaf(s_w1[r * F + f], z[f], acc);
            a1[r] = fmaxf(acc, 0.f);
        }
        __syncwarp();
        for (int f = lane; f < F; f += 32) {
            float acc = 0.f;
            for (int r = 0; r < R; ++r) acc = fmaf(s_w2[f * R + r], a1[r], acc);
            a2[f] = fmaxf(acc, 0.f);
            da2[f] = (acc > 0.f) ? da2[f] : 0.f;   // through the second relu
        }
        __syncwarp();
        for (int r = lane; r < R; r += 32) {
            float acc = 0.f;
            for (int f = 0; f < F; ++f) acc = fmaf(s_w2[f * R + r], da2[f], acc);
            da1[r] = (a1[r] > 0.f) ? acc : 0.f;    // through the first relu
        }
        __syncwarp();
        for (int f = lane; f < F; f += 32) {
            float acc = 0.f;
            for (int r = 0; r < R; ++r) acc = fmaf(s_w1[r * F + f], da1[r], acc);
            dz[f] = acc * invD;
        }
        for (int i = lane; i < R * F; i += 32) {
            acc1[i] += da1[i / F] * z[i % F];      // dW1[r,f]
            acc2[i] += da2[i / R] * a1[i % R];     // dW2[f,r]
        }
        __syncwarp();
        float* de = dE + b * sde;
        for (int i = lane; i < F * D; i += 32) {
            const int f = i / D;
            const float v = __ldg(dv + i) * a2[f] + dz[f];
            de[i] = accumulate_de ? de[i] + v : v;
        }
        __syncwarp();
    }
    __syncthreads();
    for (int i = threadIdx.x; i < R * F; i += blockDim.x) {
        float s1 = 0.f, s2 = 0.f;
        for (int w = 0; w < nw; ++w) {
            const float* base = s_rest + w * per_warp + (4 * F + 2 * R);
            s1 += base[i];
            s2 += base[R * F + i];
        }
        atomicAdd(dW1 + i, s1);
        atomicAdd(dW2 + i, s2);
    }
}

// ---------------------------------------------------------------------------------------------
// Bilinear: grid = (sample tiles, pairs).  thread = (sample lane sl, output coordinate d);
// W_p[d,:] lives in registers, E_i rows of the current pass are staged in shared memory.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void pair_from_index(int p, int F, int& i, int& j) {
    i = 0;
    int rem = p;
    while (rem >= F - 1 - i) {
        rem -= F - 1 - i;
        ++i;
    }
    j = i + 1 + rem;
}

constexpr int kBilinearTile = 256;   // samples per CTA

template <int D>
__global__ void __launch_bounds__(256) bilinear_fwd_kernel(const float* __restrict__ E, int64_t se,
                                                           int F, const float* __restrict__ W,
                                                           int wsel, float* out, int64_t so,
                                                           int64_t B) {
    constexpr int SP = 256 / D;  // samples per pass
    __shared__ float s_ei[SP][D];
    const int p = blockIdx.y;
    int fi, fj;
    pair_from_index(p, F, fi, fj);
    const float* Wp = W + (size_t)(wsel == 0 ? 0 : (wsel == 1 ? fi : p)) * D * D;
    const int d = threadIdx.x % D, sl = threadIdx.x / D;
    float wrow[D];
#pragma unroll
    for (int k = 0; k < D; ++k) wrow[k] = __ldg(Wp + d * D + k);
    const int64_t b0 = (int64_t)blockIdx.x * kBilinearTile;
    for (int pass = 0; pass < kBilinearTile / SP; ++pass) {
        const int64_t b = b0 + pass * SP + sl;
        __syncthreads();
        if (b < B) s_ei[sl][d] = __ldg(E + b * se + (int64_t)fi * D + d);
        __syncthreads();
        if (b < B) {
            float t = 0.f;
#pragma unroll
            for (int k = 0; k < D; ++k) t = fmaf(wrow[k], s_ei[sl][k], t);
            out[b * so + (int64_t)p * D + d] = t * __ldg(E + b * se + (int64_t)fj * D + d);
        }
    }
}

template <int D>
__global__ void __launch_bounds__(256) bilinear_bwd_kernel(const float* __restrict__ E, int64_t se,
                                                           int F, const float* __restrict__ W,
                                                           int wsel, const float* __restrict__ dout,
                                                           int64_t sdo, float* dE, int64_t sde,
                                                           float* dW, int64_t B) {
    constexpr int SP = 256 / D;
    __shared__ float s_ei[SP][D];
    __shared__ float s_dt[SP][D];
    __shared__ float s_red[SP][D + 1];
    const int p = blockIdx.y;
    int fi, fj;
    pair_from_index(p, F, fi, fj);
    const size_t woff = (size_t)(wsel == 0 ? 0 : (wsel == 1 ? fi : p)) * D * D;
    const float* Wp = W + woff;
    const int d = threadIdx.x % D, sl = threadIdx.x / D;
    float wrow[D], wcol[D], dwacc[D];
#pragma unroll
    for (int k = 0; k < D; ++k) {
        wrow[k] = __ldg(Wp + d * D + k);   // W[d, k]
        wcol[k] = __ldg(Wp + k * D + d);   // W[k, d]
        dwacc[k] = 0.f;
    }
    const int64_t b0 = (int64_t)blockIdx.x * kBilinearTile;
    for (int pass = 0; pass < kBilinearTile / SP; ++pass) {
        const int64_t b = b0 + pass * SP + sl;
        const bool ok = b < B;
        __syncthreads();
        float ei_d = 0.f, ej_d = 0.f, go = 0.f;
        if (ok) {
            ei_d = __ldg(E + b * se + (int64_t)fi * D + d);
            ej_d = __ldg(E + b * se + (int64_t)fj * D + d);
            go = __ldg(dout + b * sdo + (int64_t)p * D + d);
        }
        s_ei[sl][d] = ei_d;
        s_dt[sl][d] = go * ej_d;                       // dT[b, d]
        __syncthreads();
        if (ok) {
            float t = 0.f, dei = 0.f;
#pragma unroll
            for (int k = 0; k < D; ++k) {
                t = fmaf(wrow[k], s_ei[sl][k], t);      // T[b, d]
                dei = fmaf(s_dt[sl][k], wcol[k], dei);  // dE_i[b, d] = sum_k dT[b,k] W[k,d]
                dwacc[k] = fmaf(go * ej_d, s_ei[sl][k], dwacc[k]);  // dW[d,k] += dT[b,d] E_i[b,k]
            }
            atomicAdd(dE + b * sde + (int64_t)fj * D + d, go * t);
            atomicAdd(dE + b * sde + (int64_t)fi * D + d, dei);
        }
    }
    // reduce dwacc over the SP sample lanes, one k at a time, then one atomic per W entry
    for (int k = 0; k < D; ++k) {
        __syncthreads();
        s_red[sl][d] = dwacc[k];
        __syncthreads();
        if (sl == 0) {
            float s = 0.f;
            for (int q = 0; q < SP; ++q) s += s_red[q][d];
            atomicAdd(dW + woff + (size_t)d * D + k, s);
        }
    }
}

}  // namespace

extern "C" int ctr_senet_fwd(const float* E, int64_t se, int F, int D, const float* W1,
                             const float* W2, int R, float* V, int64_t sv, int64_t B, void* stream) {
    CTR_ARG(E && W1 && W2 && V && F > 0 && D > 0 && R > 0 && B >= 0, "ctr_senet_fwd: bad arguments");
    if (B == 0) return 0;
    const int nw = 8;
    const size_t smem = sizeof(float) * ((size_t)2 * R * F + (size_t)nw * (2 * F + R));
    CTR_ARG(smem <= 200 * 1024, "ctr_senet_fwd: F=%d R=%d too large for shared memory", F, R);
    if (smem > 48 * 1024)
        CTR_CUDA(cudaFuncSetAttribute(senet_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int64_t blocks = ceil_div64(B, nw);
    const int64_t cap = (int64_t)ctr_sm_count() * 8;
    if (blocks > cap) blocks = cap;
    senet_fwd_kernel<<<(unsigned)blocks, nw * 32, smem, as_stream(stream)>>>(E, se, F, D, W1, W2, R, V, sv, B);
    CTR_LAUNCH_OK("senet_fwd_kernel");
    return 0;
}

extern "C" int ctr_senet_bwd(const float* E, int64_t se, int F, int D, const float* W1,
                             const float* W2, int R, const float* dV, int64_t sdv, float* dE,
                             int64_t sde, int accumulate_de, float* dW1, float* dW2, int64_t B,
                             void* stream) {
    CTR_ARG(E && W1 && W2 && dV && dE && dW1 && dW2 && F > 0 && D > 0 && R > 0 && B >= 0,
            "ctr_senet_bwd: bad arguments");
    cudaStream_t st = as_stream(stream);
    CTR_CUDA(cudaMemsetAsync(dW1, 0, sizeof(float) * R * F, st));
    CTR_CUDA(cudaMemsetAsync(dW2, 0, sizeof(float) * R * F, st));
    if (B == 0) return 0;
    const int nw = 4;
    const size_t smem = sizeof(float) * ((size_t)2 * R * F + (size_t)nw * ((4 * F + 2 * R) + 2 * (size_t)R * F));
    CTR_ARG(smem <= 200 * 1024, "ctr_senet_bwd: F=%d R=%d too large for shared memory", F, R);
    if (smem > 48 * 1024)
        CTR_CUDA(cudaFuncSetAttribute(senet_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int64_t blocks = ceil_div64(B, (int64_t)nw * 16);
    const int64_t cap = (int64_t)ctr_sm_count() * 4;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    senet_bwd_kernel<<<(unsigned)blocks, nw * 32, smem, st>>>(E, se, F, D, W1, W2, R, dV, sdv, dE, sde,
                                                             accumulate_de, dW1, dW2, B);
    CTR_LAUNCH_OK("senet_bwd_kernel");
    return 0;
}

// ---------------------------------------------------------------------------------------------
// bilinear_type = "interaction" as GEMMs (reference layers/interaction.py:146-153): for a fixed left
// field i the pairs (i, j>i) are consecutive in combinations() order and so are their weights, hence
//   out[:, start_i*D : (start_i + n_i)*D] = (E_i @ Wblk_i^T) (.) E[:, (i+1)*D :]      Wblk_i = [n_i*D, D]
// is ONE GEMM (M = B, N = n_i*D, K = D) whose epilogue multiplies by the right-hand fields; the
// backward is three more GEMMs per i (dE_j: same product times dout; dE_i: (dout (.) E_j) @ Wblk_i;
// dWblk_i = (dout (.) E_j)^T @ E_i).  The per-pair kernels below did D shared-memory reads per output
// (21 of FiBiNET's 39 ms per step); here the cost is the 1.36 GB output stream.
// ---------------------------------------------------------------------------------------------
static bool bilinear_use_gemm(int wsel, int D, int64_t B) {
    const char* e = getenv("CTR_BILINEAR_GEMM");
    if (e && e[0] == '0') return false;
    return wsel == 2 && D % 4 == 0 && B >= 1024;
}

static int64_t pair_start(int F, int i) { return (int64_t)i * (F - 1) - (int64_t)i * (i - 1) / 2; }

static int bilinear_gemm_fwd(const float* E, int64_t se, int F, int D, const float* W, float* out, int64_t so, int64_t B,
                             cudaStream_t st) {
    for (int i = 0; i + 1 < F; ++i) {
        const int64_t ni = F - 1 - i, p0 = pair_start(F, i);
        GemmArgs g = gemm_args_default();
        g.M = B; g.N = ni * D; g.K = D;
        g.A = E + (int64_t)i * D; g.sam = se; g.sak = 1;
        g.B = W + p0 * D * D; g.sbn = D; g.sbk = 1;
        g.C = out + p0 * D; g.ldc = so;
        g.epilogue = EPI_MUL;
        g.aux = E + (int64_t)(i + 1) * D; g.ldaux = se;
        const int rc = launch_sgemm(g, st);
        if (rc) return rc;
    }
    return 0;
}

static int bilinear_gemm_bwd(const float* E, int64_t se, int F, int D, const float* W, const float* dout, int64_t sdo,
                             float* dE, int64_t sde, float* dW, int64_t B, cudaStream_t st) {
    for (int i = 0; i + 1 < F; ++i) {
        const int64_t ni = F - 1 - i, p0 = pair_start(F, i);
        int rc;
        {   // dE_j[b, d] += dout[b, p, d] * (E_i W_p^T)[b, d]        for all j > i at once
            GemmArgs g = gemm_args_default();
            g.M = B; g.N = ni * D; g.K = D;
            g.A = E + (int64_t)i * D; g.sam = se; g.sak = 1;
            g.B = W + p0 * D * D; g.sbn = D; g.sbk = 1;
            g.C = dE + (int64_t)(i + 1) * D; g.ldc = sde; g.accumulate = 1;
            g.epilogue = EPI_MUL;
            g.aux = dout + p0 * D; g.ldaux = sdo;
            if ((rc = launch_sgemm(g, st)) != 0) return rc;
        }
        {   // dE_i[b, k] += sum_{(j,d)} (dout (.) E_j)[b, (j,d)] * Wblk[(j,d), k]
            GemmArgs g = gemm_args_default();
            g.M = B; g.N = D; g.K = ni * D;
            g.A = dout + p0 * D; g.sam = sdo; g.sak = 1;
            g.amask = E + (int64_t)(i + 1) * D; g.smm = se; g.smk = 1; g.amask_act = CTR_ACT_MULPRO;
            g.B = W + p0 * D * D; g.sbn = 1; g.sbk = D;
            g.C = dE + (int64_t)i * D; g.ldc = sde; g.accumulate = 1;
            if ((rc = launch_sgemm(g, st)) != 0) return rc;
        }
        {   // dWblk[(j,d), k] = sum_b (dout (.) E_j)[b, (j,d)] * E_i[b, k]
            GemmArgs g = gemm_args_default();
            g.M = ni * D; g.N = D; g.K = B;
            g.A = dout + p0 * D; g.sam = 1; g.sak = sdo;
            g.amask = E + (int64_t)(i + 1) * D; g.smm = 1; g.smk = se; g.amask_act = CTR_ACT_MULPRO;
            g.B = E + (int64_t)i * D; g.sbn = 1; g.sbk = se;
            g.C = dW + p0 * D * D; g.ldc = D;
            g.allow_split_k = 1;
            if ((rc = launch_sgemm(g, st)) != 0) return rc;
        }
    }
    return 0;
}

static int n_bilinear_weights(int F, int wsel) {
    return wsel == 0 ? 1 : (wsel == 1 ? F : F * (F - 1) / 2);
}

extern "C" int ctr_bilinear_fwd(const float* E, int64_t se, int F, int D, const float* W, int wsel,
                                float* out, int64_t so, int64_t B, void* stream) {
    CTR_ARG(E && W && out && F >= 2 && B >= 0 && wsel >= 0 && wsel <= 2, "ctr_bilinear_fwd: bad arguments");
    if (B == 0) return 0;
    const int P = F * (F - 1) / 2;
    CTR_ARG(P <= 65535, "ctr_bilinear_fwd: too many field pairs");
    dim3 grid((unsigned)ceil_div64(B, kBilinearTile), (unsigned)P);
    cudaStream_t st = as_stream(stream);
    if (bilinear_use_gemm(wsel, D, B)) return bilinear_gemm_fwd(E, se, F, D, W, out, so, B, st);
    switch (D) {
        case 4: bilinear_fwd_kernel<4><<<grid, 256, 0, st>>>(E, se, F, W, wsel, out, so, B); break;
        case 8: bilinear_fwd_kernel<8><<<grid, 256, 0, st>>>(E, se, F, W, wsel, out, so, B); break;
        case 16: bilinear_fwd_kernel<16><<<grid, 256, 0, st>>>(E, se, F, W, wsel, out, so, B); break;
        case 32: bilinear_fwd_kernel<32><<<grid, 256, 0, st>>>(E, se, F, W, wsel, out, so, B); break;
        case 64: bilinear_fwd_kernel<64><<<grid, 256, 0, st>>>(E, se, F, W, wsel, out, so, B); break;
        default:
            ctr_set_error("ctr_bilinear_fwd: embedding dim %d unsupported (4, 8, 16, 32, 64)", D);
            return -2;
    }
    CTR_LAUNCH_OK("bilinear_fwd_kernel");
    return 0;
}

extern "C" int ctr_bilinear_bwd(const float* E, int64_t se, int F, int D, const float* W, int wsel,
                                const float* dout, int64_t sdo, float* dE, int64_t sde, float* dW,
                                int64_t B, void* stream) {
    CTR_ARG(E && W && dout && dE && dW && F >= 2 && B >= 0 && wsel >= 0 && wsel <= 2,
            "ctr_bilinear_bwd: bad arguments");
    cudaStream_t st = as_stream(stream);
    CTR_CUDA(cudaMemsetAsync(dW, 0, sizeof(float) * (size_t)n_bilinear_weights(F, wsel) * D * D, st));
    if (B == 0) return 0;
    const int P = F * (F - 1) / 2;
    CTR_ARG(P <= 65535, "ctr_bilinear_bwd: too many field pairs");
    dim3 grid((unsigned)ceil_div64(B, kBilinearTile), (unsigned)P);
    if (bilinear_use_gemm(wsel, D, B)) return bilinear_gemm_bwd(E, se, F, D, W, dout, sdo, dE, sde, dW, B, st);
    switch (D) {
        case 4: bilinear_bwd_kernel<4><<<grid, 256, 0, st>>>(E, se, F, W, wsel, dout, sdo, dE, sde, dW, B); break;
        case 8: bilinear_bwd_kernel<8><<<grid, 256, 0, st>>>(E, se, F, W, wsel, dout, sdo, dE, sde, dW, B); break;
        case 16: bilinear_bwd_kernel<16><<<grid, 256, 0, st>>>(E, se, F, W, wsel, dout, sdo, dE, sde, dW, B); break;
        case 32: bilinear_bwd_kernel<32><<<grid, 256, 0, st>>>(E, se, F, W, wsel, dout, sdo, dE, sde, dW, B); break;
        case 64: bilinear_bwd_kernel<64><<<grid, 256, 0, st>>>(E, se, F, W, wsel, dout, sdo, dE, sde, dW, B); break;
        default:
            ctr_set_error("ctr_bilinear_bwd: embedding dim %d unsupported (4, 8, 16, 32, 64)", D);
            return -2;
    }
    CTR_LAUNCH_OK("bilinear_bwd_kernel");
    return 0;
}
