// Interaction ops of the adjacent models: NFM's bi-interaction pooling, the input-aware
// re-weighting of IFM / DIFM and AFM's attention layer.  All HBM/latency-bound per-sample kernels on the
// [B, F, D] embedding block the fused gather already produced; fp32.
//   reference: layers/interaction.py:54-61 (BiInteractionPooling), :250-331 (AFMLayer),
//              models/ifm.py:69-91, models/difm.py:82-112, models/basemodel.py:63-92 (refine weight)
#include "common.cuh"

namespace {

// ---------------------------------------------------------------------------------------------
// bi-interaction pooling: out[b,d] = 0.5 * ((sum_f e)^2 - sum_f e^2)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bipool_fwd_kernel(const float* __restrict__ E, int64_t se, int F, int D, float* out,
                                                         int64_t so, int64_t B) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < B * D; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = i / D;
        const int d = (int)(i - b * D);
        const float* e = E + b * se + d;
        float s = 0.f, q = 0.f;
        for (int f = 0; f < F; ++f) {
            const float v = __ldg(e + f * D);
            s += v;
            q += v * v;
        }
        out[b * so + d] = 0.5f * (s * s - q);
    }
}

__global__ void __launch_bounds__(256) bipool_bwd_kernel(const float* __restrict__ E, int64_t se, int F, int D,
                                                         const float* __restrict__ g, int64_t sg, float* dE, int64_t sde,
                                                         int64_t B) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < B * D; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = i / D;
        const int d = (int)(i - b * D);
        const float* e = E + b * se + d;
        float s = 0.f;
        for (int f = 0; f < F; ++f) s += __ldg(e + f * D);
        const float gv = __ldg(g + b * sg + d);
        for (int f = 0; f < F; ++f) dE[b * sde + f * D + d] = gv * (s - __ldg(e + f * D));
    }
}

// ---------------------------------------------------------------------------------------------
// input-aware refinement (IFM: m = F * softmax(P); DIFM: m = P)
//   Er[b,f,:] = m[b,f] * E[b,f,:]      lin[b] = sum_f m[b,f] * L[b,f]
// one warp per sample
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) refine_fwd_kernel(const float* __restrict__ P, const float* __restrict__ E, int64_t se,
                                                         const float* __restrict__ L, int F, int D, int softmax, float* m,
                                                         float* Er, int64_t ser, float* lin, int64_t B) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t b = warp0; b < B; b += nwarps) {
        const float* p = P + b * F;
        float mx = -INFINITY, sum = 0.f;
        if (softmax) {
            for (int f = lane; f < F; f += 32) mx = fmaxf(mx, p[f]);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            for (int f = lane; f < F; f += 32) sum += expf(p[f] - mx);
            sum = warp_sum(sum);
        }
        float lp = 0.f;
        for (int f = lane; f < F; f += 32) {
            const float mv = softmax ? (float)F * (expf(p[f] - mx) / sum) : p[f];
            m[b * F + f] = mv;
            if (L) lp += mv * L[b * F + f];
        }
        __syncwarp();
        for (int i = lane; i < F * D; i += 32) Er[b * ser + i] = m[b * F + i / D] * E[b * se + i];
        if (lin) {
            lp = warp_sum(lp);
            if (lane == 0) lin[b] = lp;
        }
    }
}

__global__ void __launch_bounds__(128) refine_bwd_kernel(const float* __restrict__ m, const float* __restrict__ E, int64_t se,
                                                         const float* __restrict__ L, int F, int D, int softmax,
                                                         const float* __restrict__ dEr, int64_t sder,
                                                         const float* __restrict__ dlin, float* dP, float* dE, int64_t sde,
                                                         float* dL, int64_t B) {
    extern __shared__ float s_dm[];               // [warps per block][F]
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    float* dm = s_dm + (size_t)wid * F;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t b = warp0; b < B; b += nwarps) {
        const float gl = dlin ? dlin[b] : 0.f;
        float dot = 0.f;
        for (int f = lane; f < F; f += 32) {
            float acc = 0.f;
            if (dEr)
                for (int d = 0; d < D; ++d) acc += dEr[b * sder + f * D + d] * E[b * se + f * D + d];
            if (L) {
                acc += gl * L[b * F + f];
                if (dL) dL[b * F + f] = gl * m[b * F + f];
            }
            dm[f] = acc;
            dot += acc * m[b * F + f];
        }
        dot = warp_sum(dot) / (float)F;
        __syncwarp();
        for (int f = lane; f < F; f += 32) dP[b * F + f] = softmax ? m[b * F + f] * (dm[f] - dot) : dm[f];
        if (dE)
            for (int i = lane; i < F * D; i += 32) dE[b * sde + i] = dEr ? dEr[b * sder + i] * m[b * F + i / D] : 0.f;
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------
// AFM attention layer, one warp per sample.  Shared memory per block: the attention parameters and,
// per warp, the sample's E [F*D], pair scores [P] and (backward) dE [F*D].
// ---------------------------------------------------------------------------------------------
constexpr int AFM_MAX_A = 16;     // attention_factor
constexpr int AFM_MAX_D = 64;

struct AfmArgs {
    const float* E; int64_t se; int F, D, A;
    const float* W;    // [D, A]
    const float* bvec; // [A]
    const float* h;    // [A]
    float* out;        // [B, D] attention output (the projection p and the dropout stay outside)
    // backward
    const float* g;    // [B, D]
    float* dE; int64_t sde;
    float* dW; float* db; float* dh;
    int64_t B;
};

__device__ __forceinline__ void pair_of(int p, int F, int& i, int& j) {
    // combinations(range(F), 2) in lexicographic order
    int r = p;
    i = 0;
    while (r >= F - 1 - i) {
        r -= F - 1 - i;
        ++i;
    }
    j = i + 1 + r;
}

template <bool BWD>
__global__ void __launch_bounds__(128) afm_kernel(AfmArgs a) {
    extern __shared__ float smem[];
    const int F = a.F, D = a.D, A = a.A, P = F * (F - 1) / 2;
    float* sW = smem;                 // D*A
    float* sb = sW + D * A;           // A
    float* sh = sb + A;               // A
    float* gacc = sh + A;             // BWD: D*A + A + A block-level parameter-gradient accumulators
    const int n_par = D * A + 2 * A;
    float* per_warp = gacc + (BWD ? n_par : 0);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int per = F * D + P + (BWD ? F * D + D : 0);
    float* sE = per_warp + (size_t)wid * per;
    float* ss = sE + F * D;
    float* sdE = ss + P;
    float* sdo = sdE + F * D;         // BWD: d(attention output) of the sample
    for (int i = threadIdx.x; i < D * A; i += blockDim.x) sW[i] = a.W[i];
    for (int i = threadIdx.x; i < A; i += blockDim.x) {
        sb[i] = a.bvec[i];
        sh[i] = a.h[i];
    }
    if (BWD)
        for (int i = threadIdx.x; i < n_par; i += blockDim.x) gacc[i] = 0.f;
    __syncthreads();
    const int64_t warp0 = (int64_t)blockIdx.x * nw + wid;
    const int64_t nwarps = (int64_t)gridDim.x * nw;
    for (int64_t b = warp0; b < a.B; b += nwarps) {
        for (int i = lane; i < F * D; i += 32) {
            sE[i] = a.E[b * a.se + i];
            if (BWD) sdE[i] = 0.f;
        }
        __syncwarp();
        // scores s_p
        float mx = -INFINITY;
        for (int p = lane; p < P; p += 32) {
            int i, j;
            pair_of(p, F, i, j);
            float t[AFM_MAX_A];
#pragma unroll
            for (int k = 0; k < AFM_MAX_A; ++k) t[k] = (k < A) ? sb[k] : 0.f;
            for (int d = 0; d < D; ++d) {
                const float ip = sE[i * D + d] * sE[j * D + d];
#pragma unroll
                for (int k = 0; k < AFM_MAX_A; ++k)
                    if (k < A) t[k] += ip * sW[d * A + k];
            }
            float s = 0.f;
#pragma unroll
            for (int k = 0; k < AFM_MAX_A; ++k)
                if (k < A) s += fmaxf(t[k], 0.f) * sh[k];
            ss[p] = s;
            mx = fmaxf(mx, s);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        float sum = 0.f;
        for (int p = lane; p < P; p += 32) sum += expf(ss[p] - mx);
        sum = warp_sum(sum);
        __syncwarp();
        // attention output o[d] = sum_p alpha_p ip_p[d]  (lane-private partials over its pairs, then reduce)
        float dalpha_dot = 0.f;
        for (int d0 = 0; d0 < D; d0 += 32) {
            const int dn = D - d0 < 32 ? D - d0 : 32;
            float part[32];
#pragma unroll
            for (int d = 0; d < 32; ++d) part[d] = 0.f;
            for (int p = lane; p < P; p += 32) {
                int i, j;
                pair_of(p, F, i, j);
                const float al = expf(ss[p] - mx) / sum;
#pragma unroll
                for (int d = 0; d < 32; ++d)
                    if (d < dn) part[d] += al * sE[i * D + d0 + d] * sE[j * D + d0 + d];
            }
#pragma unroll
            for (int d = 0; d < 32; ++d) {
                if (d < dn) {
                    const float od = warp_sum(part[d]);
                    if (!BWD && lane == 0) a.out[b * D + d0 + d] = od;
                }
            }
        }
        if (!BWD) {
            __syncwarp();
            continue;
        }
        // ---------------- backward ----------------
        for (int d = lane; d < D; d += 32) sdo[d] = a.g[b * D + d];
        __syncwarp();
        // dalpha_p = sum_d do[d] ip_p[d];  dot = sum_q alpha_q dalpha_q
        for (int p = lane; p < P; p += 32) {
            int i, j;
            pair_of(p, F, i, j);
            float da = 0.f;
            for (int d = 0; d < D; ++d) da += sdo[d] * sE[i * D + d] * sE[j * D + d];
            dalpha_dot += (expf(ss[p] - mx) / sum) * da;
        }
        dalpha_dot = warp_sum(dalpha_dot);
        for (int p = lane; p < P; p += 32) {
            int i, j;
            pair_of(p, F, i, j);
            const float al = expf(ss[p] - mx) / sum;
            float da = 0.f;
            for (int d = 0; d < D; ++d) da += sdo[d] * sE[i * D + d] * sE[j * D + d];
            const float ds = al * (da - dalpha_dot);
            // recompute t_p
            float t[AFM_MAX_A];
#pragma unroll
            for (int k = 0; k < AFM_MAX_A; ++k) t[k] = (k < A) ? sb[k] : 0.f;
            for (int d = 0; d < D; ++d) {
                const float ip = sE[i * D + d] * sE[j * D + d];
#pragma unroll
                for (int k = 0; k < AFM_MAX_A; ++k)
                    if (k < A) t[k] += ip * sW[d * A + k];
            }
            float dt[AFM_MAX_A];
#pragma unroll
            for (int k = 0; k < AFM_MAX_A; ++k) {
                dt[k] = 0.f;
                if (k < A) {
                    atomicAdd(&gacc[D * A + A + k], ds * fmaxf(t[k], 0.f));                 // dh
                    dt[k] = t[k] > 0.f ? ds * sh[k] : 0.f;
                    atomicAdd(&gacc[D * A + k], dt[k]);                                    // db
                }
            }
            for (int d = 0; d < D; ++d) {
                const float ei = sE[i * D + d], ej = sE[j * D + d];
                float dip = al * sdo[d];
#pragma unroll
                for (int k = 0; k < AFM_MAX_A; ++k) {
                    if (k < A) {
                        dip += dt[k] * sW[d * A + k];
                        atomicAdd(&gacc[d * A + k], ei * ej * dt[k]);                      // dW
                    }
                }
                atomicAdd(&sdE[i * D + d], dip * ej);
                atomicAdd(&sdE[j * D + d], dip * ei);
            }
        }
        __syncwarp();
        for (int i = lane; i < F * D; i += 32) a.dE[b * a.sde + i] = sdE[i];
        __syncwarp();
    }
    if (BWD) {
        __syncthreads();
        for (int i = threadIdx.x; i < n_par; i += blockDim.x) {
            const float v = gacc[i];
            if (v != 0.f) {
                if (i < D * A) atomicAdd(a.dW + i, v);
                else if (i < D * A + A) atomicAdd(a.db + (i - D * A), v);
                else atomicAdd(a.dh + (i - D * A - A), v);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// multi-head self-attention over the fields of one sample (AutoInt's InteractingLayer as used by DIFM,
// reference layers/interaction.py:352-394) after the four projections: Q, K, V, R are [B, F, D] tensors
// (the projections themselves are GEMMs on the [B*F, D] view), H heads of dh = D / H coordinates.
//   Y[b,i,hd] = relu( sum_j softmax_j(Q_i . K_j * scale)[j] * V[b,j,hd] + R[b,i,hd] )
// one warp per sample; lanes iterate over (field i, head) pairs.
// ---------------------------------------------------------------------------------------------
constexpr int ATT_MAX_DH = 32;

template <bool BWD>
__global__ void __launch_bounds__(128) fieldattn_kernel(const float* __restrict__ Q, const float* __restrict__ K,
                                                        const float* __restrict__ V, const float* __restrict__ R, int F, int D,
                                                        int H, float scale, float* Y, const float* __restrict__ dY, float* dQ,
                                                        float* dK, float* dV, float* dR, int64_t B) {
    extern __shared__ float smem[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int FD = F * D, dh = D / H;
    float* base = smem + (size_t)wid * (BWD ? 6 : 3) * FD;
    float *sQ = base, *sK = base + FD, *sV = base + 2 * FD;
    float *sdQ = base + 3 * FD, *sdK = base + 4 * FD, *sdV = base + 5 * FD;
    const int64_t warp0 = (int64_t)blockIdx.x * nw + wid;
    const int64_t nwarps = (int64_t)gridDim.x * nw;
    for (int64_t b = warp0; b < B; b += nwarps) {
        for (int i = lane; i < FD; i += 32) {
            sQ[i] = Q[b * FD + i];
            sK[i] = K[b * FD + i];
            sV[i] = V[b * FD + i];
            if (BWD) sdQ[i] = sdK[i] = sdV[i] = 0.f;
        }
        __syncwarp();
        for (int pr = lane; pr < F * H; pr += 32) {
            const int i = pr / H, hh = pr - i * H;
            const float* qi = sQ + i * D + hh * dh;
            float mx = -INFINITY;
            for (int j = 0; j < F; ++j) {
                float s = 0.f;
                for (int d = 0; d < dh; ++d) s += qi[d] * sK[j * D + hh * dh + d];
                mx = fmaxf(mx, s * scale);
            }
            float sum = 0.f;
            for (int j = 0; j < F; ++j) {
                float s = 0.f;
                for (int d = 0; d < dh; ++d) s += qi[d] * sK[j * D + hh * dh + d];
                sum += expf(s * scale - mx);
            }
            float o[ATT_MAX_DH];
#pragma unroll
            for (int d = 0; d < ATT_MAX_DH; ++d) o[d] = 0.f;
            float dot = 0.f;              // BWD: sum_j p_ij dp_ij
            float go[ATT_MAX_DH];         // BWD: masked dY of this (i, head)
            if (BWD) {
#pragma unroll
                for (int d = 0; d < ATT_MAX_DH; ++d) {
                    go[d] = 0.f;
                    if (d < dh) {
                        const int64_t off = b * FD + i * D + hh * dh + d;
                        go[d] = Y[off] > 0.f ? dY[off] : 0.f;
                        dR[off] = go[d];
                    }
                }
            }
            for (int j = 0; j < F; ++j) {
                float s = 0.f;
                for (int d = 0; d < dh; ++d) s += qi[d] * sK[j * D + hh * dh + d];
                const float p = expf(s * scale - mx) / sum;
                if (!BWD) {
#pragma unroll
                    for (int d = 0; d < ATT_MAX_DH; ++d)
                        if (d < dh) o[d] += p * sV[j * D + hh * dh + d];
                } else {
                    float dp = 0.f;
#pragma unroll
                    for (int d = 0; d < ATT_MAX_DH; ++d)
                        if (d < dh) {
                            dp += go[d] * sV[j * D + hh * dh + d];
                            atomicAdd(&sdV[j * D + hh * dh + d], p * go[d]);
                        }
                    dot += p * dp;
                }
            }
            if (!BWD) {
#pragma unroll
                for (int d = 0; d < ATT_MAX_DH; ++d)
                    if (d < dh) {
                        const int64_t off = b * FD + i * D + hh * dh + d;
                        Y[off] = fmaxf(o[d] + R[off], 0.f);
                    }
            } else {
                for (int j = 0; j < F; ++j) {
                    float s = 0.f, dp = 0.f;
                    for (int d = 0; d < dh; ++d) {
                        s += qi[d] * sK[j * D + hh * dh + d];
                        dp += go[d < ATT_MAX_DH ? d : 0] * sV[j * D + hh * dh + d];
                    }
                    const float p = expf(s * scale - mx) / sum;
                    const float ds = p * (dp - dot) * scale;
                    for (int d = 0; d < dh; ++d) {
                        atomicAdd(&sdQ[i * D + hh * dh + d], ds * sK[j * D + hh * dh + d]);
                        atomicAdd(&sdK[j * D + hh * dh + d], ds * qi[d]);
                    }
                }
            }
        }
        __syncwarp();
        if (BWD)
            for (int i = lane; i < FD; i += 32) {
                dQ[b * FD + i] = sdQ[i];
                dK[b * FD + i] = sdK[i];
                dV[b * FD + i] = sdV[i];
            }
        __syncwarp();
    }
}

unsigned grid_for(int64_t work, int threads, int per_sm) {
    int64_t blocks = ceil_div64(work, threads);
    const int64_t cap = (int64_t)ctr_sm_count() * per_sm;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (unsigned)blocks;
}

}  // namespace

extern "C" int ctr_bipool_fwd(const float* E, int64_t se, int F, int D, float* out, int64_t so, int64_t B, void* stream) {
    CTR_ARG(E && out && F > 0 && D > 0 && B >= 0 && se >= (int64_t)F * D && so >= D, "ctr_bipool_fwd: bad arguments");
    if (B == 0) return 0;
    bipool_fwd_kernel<<<grid_for(B * D, 256, 8), 256, 0, as_stream(stream)>>>(E, se, F, D, out, so, B);
    CTR_LAUNCH_OK("bipool_fwd_kernel");
    return 0;
}

extern "C" int ctr_bipool_bwd(const float* E, int64_t se, int F, int D, const float* g, int64_t sg, float* dE, int64_t sde,
                              int64_t B, void* stream) {
    CTR_ARG(E && g && dE && F > 0 && D > 0 && B >= 0, "ctr_bipool_bwd: bad arguments");
    if (B == 0) return 0;
    bipool_bwd_kernel<<<grid_for(B * D, 256, 8), 256, 0, as_stream(stream)>>>(E, se, F, D, g, sg, dE, sde, B);
    CTR_LAUNCH_OK("bipool_bwd_kernel");
    return 0;
}

extern "C" int ctr_refine_fwd(const float* P, const float* E, int64_t se, const float* L, int F, int D, int softmax, float* m,
                              float* Er, int64_t ser, float* lin, int64_t B, void* stream) {
    CTR_ARG(P && E && m && Er && F > 0 && D > 0 && B >= 0, "ctr_refine_fwd: bad arguments");
    if (B == 0) return 0;
    refine_fwd_kernel<<<grid_for(B * 32, 256, 8), 256, 0, as_stream(stream)>>>(P, E, se, L, F, D, softmax, m, Er, ser, lin, B);
    CTR_LAUNCH_OK("refine_fwd_kernel");
    return 0;
}

extern "C" int ctr_refine_bwd(const float* m, const float* E, int64_t se, const float* L, int F, int D, int softmax,
                              const float* dEr, int64_t sder, const float* dlin, float* dP, float* dE, int64_t sde,
                              float* dL, int64_t B, void* stream) {
    CTR_ARG(m && E && dP && F > 0 && F <= 2048 && D > 0 && B >= 0, "ctr_refine_bwd: bad arguments");
    if (B == 0) return 0;
    refine_bwd_kernel<<<grid_for(B * 32, 128, 16), 128, (size_t)4 * F * sizeof(float), as_stream(stream)>>>(
        m, E, se, L, F, D, softmax, dEr, sder, dlin, dP, dE, sde, dL, B);
    CTR_LAUNCH_OK("refine_bwd_kernel");
    return 0;
}

static int afm_launch(bool bwd, AfmArgs a, void* stream) {
    const int P = a.F * (a.F - 1) / 2;
    const int n_par = a.D * a.A + 2 * a.A;
    const size_t smem = sizeof(float) * ((size_t)n_par + (bwd ? n_par : 0) + 4 * ((size_t)a.F * a.D + P + (bwd ? a.F * a.D + a.D : 0)));
    if (smem > 200 * 1024) {
        ctr_set_error("ctr_afm: %d fields x dim %d need %zu bytes of shared memory per block", a.F, a.D, smem);
        return -2;
    }
    if (bwd) {
        CTR_CUDA(cudaFuncSetAttribute(afm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        afm_kernel<true><<<grid_for(a.B * 32, 128, 4), 128, smem, as_stream(stream)>>>(a);
    } else {
        CTR_CUDA(cudaFuncSetAttribute(afm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        afm_kernel<false><<<grid_for(a.B * 32, 128, 4), 128, smem, as_stream(stream)>>>(a);
    }
    CTR_LAUNCH_OK("afm_kernel");
    return 0;
}

extern "C" int ctr_afm_fwd(const float* E, int64_t se, int F, int D, int A, const float* W, const float* b, const float* h,
                           float* out, int64_t B, void* stream) {
    CTR_ARG(E && W && b && h && out && F >= 2 && D > 0 && D <= AFM_MAX_D && A > 0 && A <= AFM_MAX_A && B >= 0,
            "ctr_afm_fwd: bad arguments (D <= %d, attention_factor <= %d)", AFM_MAX_D, AFM_MAX_A);
    if (B == 0) return 0;
    AfmArgs a{E, se, F, D, A, W, b, h, out, nullptr, nullptr, 0, nullptr, nullptr, nullptr, B};
    return afm_launch(false, a, stream);
}

extern "C" int ctr_afm_bwd(const float* E, int64_t se, int F, int D, int A, const float* W, const float* b, const float* h,
                           const float* g, float* dE, int64_t sde, float* dW, float* db, float* dh, int64_t B,
                           void* stream) {
    CTR_ARG(E && W && b && h && g && dE && dW && db && dh && F >= 2 && D > 0 && D <= AFM_MAX_D && A > 0 &&
                A <= AFM_MAX_A && B >= 0,
            "ctr_afm_bwd: bad arguments");
    cudaStream_t st = as_stream(stream);
    CTR_CUDA(cudaMemsetAsync(dW, 0, sizeof(float) * D * A, st));
    CTR_CUDA(cudaMemsetAsync(db, 0, sizeof(float) * A, st));
    CTR_CUDA(cudaMemsetAsync(dh, 0, sizeof(float) * A, st));
    if (B == 0) return 0;
    AfmArgs a{E, se, F, D, A, W, b, h, nullptr, g, dE, sde, dW, db, dh, B};
    return afm_launch(true, a, stream);
}

extern "C" int ctr_fieldattn_fwd(const float* Q, const float* K, const float* V, const float* R, int F, int D, int H,
                                 float scale, float* Y, int64_t B, void* stream) {
    CTR_ARG(Q && K && V && R && Y && F > 0 && D > 0 && H > 0 && D % H == 0 && D / H <= ATT_MAX_DH && B >= 0,
            "ctr_fieldattn_fwd: bad arguments (D %% heads == 0, D / heads <= %d)", ATT_MAX_DH);
    if (B == 0) return 0;
    const size_t smem = sizeof(float) * 4 * 3 * (size_t)F * D;
    CTR_ARG(smem <= 200 * 1024, "ctr_fieldattn_fwd: F*D too large for shared memory");
    CTR_CUDA(cudaFuncSetAttribute(fieldattn_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    fieldattn_kernel<false><<<grid_for(B * 32, 128, 4), 128, smem, as_stream(stream)>>>(Q, K, V, R, F, D, H, scale, Y, nullptr,
                                                                                       nullptr, nullptr, nullptr, nullptr, B);
    CTR_LAUNCH_OK("fieldattn_kernel");
    return 0;
}

extern "C" int ctr_fieldattn_bwd(const float* Q, const float* K, const float* V, const float* Y, const float* dY, int F, int D,
                                 int H, float scale, float* dQ, float* dK, float* dV, float* dR, int64_t B, void* stream) {
    CTR_ARG(Q && K && V && Y && dY && dQ && dK && dV && dR && F > 0 && D > 0 && H > 0 && D % H == 0 && D / H <= ATT_MAX_DH &&
                B >= 0,
            "ctr_fieldattn_bwd: bad arguments");
    if (B == 0) return 0;
    const size_t smem = sizeof(float) * 4 * 6 * (size_t)F * D;
    CTR_ARG(smem <= 200 * 1024, "ctr_fieldattn_bwd: F*D too large for shared memory");
    CTR_CUDA(cudaFuncSetAttribute(fieldattn_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    fieldattn_kernel<true><<<grid_for(B * 32, 128, 4), 128, smem, as_stream(stream)>>>(
        Q, K, V, nullptr, F, D, H, scale, const_cast<float*>(Y), dY, dQ, dK, dV, dR, B);
    CTR_LAUNCH_OK("fieldattn_kernel");
    return 0;
}
