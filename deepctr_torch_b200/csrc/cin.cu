// Compressed Interaction Network layer, forward and backward, WITHOUT materialising the
// [B, H*M, D] outer product (6.98 GB at BASELINE config #3 in the reference).
// Reference: layers/interaction.py:207-248 (einsum 'bhd,bmd->bhmd' -> reshape -> Conv1d(1x1)).
//
// Per layer:  Z[b,n,d] = sum_{h,m} W[n, h*M+m] * Xp[b,h,d] * X0[b,m,d] + bias[n]
// viewed as a GEMM with rows r = (b,d), K = h*M+m, N = channels.  The A operand
// P[r, hm] = Xp[b,h,d]*X0[b,m,d] is generated on the fly in shared memory from the two
// [TB, *, D] tiles of the CTA; FP32 FFMA accumulation (parity mode, see gemm.cu).
//
//   fwd   : rows tile = TB samples x D (<=128 rows) x 128 channels, K step 16
//   dW    : [128 channels] x [128 hm] tiles, K = (b,d) split across CTAs, fp32 atomics at the end
//   dX    : rows tile x (HT*M) hm columns per step with HT whole h-groups so that
//           dXp[b,h,d] = sum_m G*X0 is complete inside the tile (plain store) and
//           dX0[b,m,d] += sum_h G*Xp accumulates in CTA-private shared memory (no atomics).
#include <stdlib.h>

#include "common.cuh"

int launch_cin_tc_fwd(const float* Xp, int64_t sxp, int H, const float* X0, int64_t sx0, int M, int D,
                      const float* W, const float* bias, int N, int direct_start, int act, float* Y,
                      float* out, int64_t ld_out, int64_t B, cudaStream_t st);

int launch_cin_tc_bwd(const float* Xp, int64_t sxp, int H, const float* X0, int64_t sx0, int M, int D,
                      const float* W, int N, const float* dZ, float* dW, float* dXp, int64_t sdxp,
                      float* dX0, int64_t sdx0, int64_t B, int same, cudaStream_t st, int skip_dw);

// second-generation forward (cin_v2.cu): experimental, only with CTR_CIN_V2=1
int launch_cin_v2_fwd(const float* Xp, int64_t sxp, int H, const float* X0, int64_t sx0, int M, int D, const float* W,
                      const float* bias, int N, int direct_start, int act, float* Y, float* out, int64_t ld_out, int64_t B,
                      cudaStream_t st);
int launch_cin_v2_dw(const float* Xp, int64_t sxp, int H, const float* X0, int64_t sx0, int M, int D, int N, const float* dZ,
                     float* dW, int64_t B, cudaStream_t st);

namespace {

constexpr int BM = 128, BN = 128, BK = 16;
constexpr int NT = 256;

struct CinFwdArgs {
    const float* Xp; int64_t sxp; int H;
    const float* X0; int64_t sx0; int M;
    int D;
    const float* W; const float* bias; int N; int direct_start; int act;
    float* Y; float* out; int64_t ld_out; int64_t B;
    int TB;  // samples per CTA, TB*D <= BM
};

__global__ void __launch_bounds__(NT) cin_fwd_kernel(CinFwdArgs a) {
    extern __shared__ __align__(16) float smem[];
    float* As = smem;                               // [2][BK][BM+4]
    float* Bs = As + 2 * BK * (BM + 4);             // [2][BK][BN+4]
    float* sXp = Bs + 2 * BK * (BN + 4);            // [TB][H][D]
    float* sX0 = sXp + (size_t)a.TB * a.H * a.D;    // [TB][M][D]
    const int tid = threadIdx.x;
    const int tx = tid % 16, ty = tid / 16;
    const int D = a.D, H = a.H, M = a.M, N = a.N;
    const int HM = H * M;
    const int64_t b0 = (int64_t)blockIdx.x * a.TB;
    const int n0 = blockIdx.y * BN;
    const int tb = (int)((a.B - b0 < a.TB) ? a.B - b0 : a.TB);
    const int rows = tb * D;

    for (int i = tid; i < a.TB * H * D; i += NT) {
        const int bl = i / (H * D);
        sXp[i] = (bl < tb) ? __ldg(a.Xp + (b0 + bl) * a.sxp + (i - bl * H * D)) : 0.f;
    }
    for (int i = tid; i < a.TB * M * D; i += NT) {
        const int bl = i / (M * D);
        sX0[i] = (bl < tb) ? __ldg(a.X0 + (b0 + bl) * a.sx0 + (i - bl * M * D)) : 0.f;
    }
    __syncthreads();

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    float rb[BN * BK / NT];
    auto load_b = [&](int k0) {
#pragma unroll
        for (int i = 0; i < BN * BK / NT; ++i) {
            const int e = tid + i * NT;
            const int kk = e % BK, nn = e / BK;
            const int n = n0 + nn, k = k0 + kk;
            rb[i] = (n < N && k < HM) ? __ldg(a.W + (size_t)n * HM + k) : 0.f;
        }
    };
    auto store_tiles = [&](int buf, int k0) {
#pragma unroll
        for (int i = 0; i < BN * BK / NT; ++i) {
            const int e = tid + i * NT;
            Bs[(buf * BK + e % BK) * (BN + 4) + e / BK] = rb[i];
        }
#pragma unroll
        for (int i = 0; i < BM * BK / NT; ++i) {
            const int e = tid + i * NT;
            const int r = e % BM, kk = e / BM;
            const int k = k0 + kk;
            float v = 0.f;
            if (r < rows && k < HM) {
                const int bl = r / D, d = r - bl * D;
                const int h = k / M, m = k - h * M;
                v = sXp[(bl * H + h) * D + d] * sX0[(bl * M + m) * D + d];
            }
            As[(buf * BK + kk) * (BM + 4) + r] = v;
        }
    };

    load_b(0);
    store_tiles(0, 0);
    __syncthreads();
    int buf = 0;
    for (int k0 = 0; k0 < HM; k0 += BK) {
        const bool has_next = k0 + BK < HM;
        if (has_next) load_b(k0 + BK);
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[(buf * BK + kk) * (BM + 4) + ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[(buf * BK + kk) * (BM + 4) + 64 + ty * 4]);
            const float4 c0 = *reinterpret_cast<const float4*>(&Bs[(buf * BK + kk) * (BN + 4) + tx * 4]);
            const float4 c1 = *reinterpret_cast<const float4*>(&Bs[(buf * BK + kk) * (BN + 4) + 64 + tx * 4]);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float bv[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        if (has_next) store_tiles(buf ^ 1, k0 + BK);
        __syncthreads();
        buf ^= 1;
    }

    // epilogue: bias + activation, store Y, reduce the direct-connect channels over d
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int n = n0 + (j / 4) * 64 + tx * 4 + (j % 4);
        if (n >= N) continue;
        const float bn = a.bias ? __ldg(a.bias + n) : 0.f;
        float part = 0.f;
        int part_b = -1;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int r = (i / 4) * 64 + ty * 4 + (i % 4);
            if (r >= rows) continue;
            const int bl = r / D, d = r - bl * D;
            const float y = act_apply(a.act, acc[i][j] + bn);
            a.Y[((b0 + bl) * N + n) * D + d] = y;
            if (n >= a.direct_start && a.out) {
                if (bl != part_b) {
                    if (part_b >= 0) atomicAdd(a.out + (b0 + part_b) * a.ld_out + (n - a.direct_start), part);
                    part = 0.f;
                    part_b = bl;
                }
                part += y;
            }
        }
        if (part_b >= 0) atomicAdd(a.out + (b0 + part_b) * a.ld_out + (n - a.direct_start), part);
    }
}

// dZ[b,n,d] = upstream(b,n,d) * act'(Y[b,n,d])
__global__ void __launch_bounds__(256) cin_dz_kernel(const float* __restrict__ Y,
                                                     const float* __restrict__ dYh, int64_t sdyh,
                                                     const float* __restrict__ dout, int64_t ld_dout,
                                                     int N, int D, int n_hidden, int direct_start,
                                                     int act, float* dZ, int64_t B) {
    const int64_t total = B * N * D;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = i / ((int64_t)N * D);
        const int rem = (int)(i - b * (int64_t)N * D);
        const int n = rem / D, d = rem - n * D;
        float up = 0.f;
        if (n < n_hidden && dYh) up += dYh[b * sdyh + n * D + d];
        if (n >= direct_start && dout) up += dout[b * ld_dout + (n - direct_start)];
        dZ[i] = up * act_grad_from_y(act, Y[i]);
    }
}

// dbias[n] = sum_{b,d} dZ[b,n,d]
__global__ void __launch_bounds__(256) cin_dbias_kernel(const float* __restrict__ dZ, int N, int D,
                                                        int64_t B, int64_t b_per_block, float* dbias) {
    const int64_t bbeg = (int64_t)blockIdx.x * b_per_block;
    const int64_t bend = (bbeg + b_per_block < B) ? bbeg + b_per_block : B;
    const int ND = N * D;
    for (int j = threadIdx.x; j < ND; j += blockDim.x) {
        float acc = 0.f;
        for (int64_t b = bbeg; b < bend; ++b) acc += __ldg(dZ + b * ND + j);
        atomicAdd(dbias + j / D, acc);
    }
}

struct CinDwArgs {
    const float* Xp; int64_t sxp; int H;
    const float* X0; int64_t sx0; int M;
    int D; int N;
    const float* dZ; float* dW; int64_t B; int64_t b_per_block;
};

// dW[n, hm] = sum_{b,d} dZ[b,n,d] * Xp[b,h,d] * X0[b,m,d];  grid = (hm tiles, n tiles, b chunks)
__global__ void __launch_bounds__(NT) cin_dw_kernel(CinDwArgs a) {
    __shared__ __align__(16) float As[2][BK][BM + 4];   // [k=(b,d)][n]
    __shared__ __align__(16) float Bs[2][BK][BN + 4];   // [k=(b,d)][hm]
    const int tid = threadIdx.x;
    const int tx = tid % 16, ty = tid / 16;
    const int D = a.D, H = a.H, M = a.M, N = a.N;
    const int HM = H * M;
    const int hm0 = blockIdx.x * BN, n0 = blockIdx.y * BM;
    const int64_t qbeg = (int64_t)blockIdx.z * a.b_per_block * D;
    int64_t qend = qbeg + a.b_per_block * D;
    if (qend > a.B * D) qend = a.B * D;

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    float ra[BM * BK / NT], rb[BN * BK / NT];
    auto load_tiles = [&](int64_t q0) {
#pragma unroll
        for (int i = 0; i < BM * BK / NT; ++i) {
            const int e = tid + i * NT;
            const int kk = e % BK, nn = e / BK;
            const int64_t q = q0 + kk;
            const int n = n0 + nn;
            float v = 0.f;
            if (q < qend && n < N) {
                const int64_t b = q / D;
                const int d = (int)(q - b * D);
                v = __ldg(a.dZ + (b * N + n) * D + d);
            }
            ra[i] = v;
        }
#pragma unroll
        for (int i = 0; i < BN * BK / NT; ++i) {
            const int e = tid + i * NT;
            const int kk = e % BK, cc = e / BK;
            const int64_t q = q0 + kk;
            const int hm = hm0 + cc;
            float v = 0.f;
            if (q < qend && hm < HM) {
                const int64_t b = q / D;
                const int d = (int)(q - b * D);
                const int h = hm / M, m = hm - h * M;
                v = __ldg(a.Xp + b * a.sxp + h * D + d) * __ldg(a.X0 + b * a.sx0 + m * D + d);
            }
            rb[i] = v;
        }
    };
    auto store_tiles = [&](int buf) {
#pragma unroll
        for (int i = 0; i < BM * BK / NT; ++i) {
            const int e = tid + i * NT;
            As[buf][e % BK][e / BK] = ra[i];
        }
#pragma unroll
        for (int i = 0; i < BN * BK / NT; ++i) {
            const int e = tid + i * NT;
            Bs[buf][e % BK][e / BK] = rb[i];
        }
    };

    if (qbeg < qend) {
        load_tiles(qbeg);
        store_tiles(0);
    }
    __syncthreads();
    int buf = 0;
    for (int64_t q0 = qbeg; q0 < qend; q0 += BK) {
        const bool has_next = q0 + BK < qend;
        if (has_next) load_tiles(q0 + BK);
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][64 + ty * 4]);
            const float4 c0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
            const float4 c1 = *reinterpret_cast<const float4*>(&Bs[buf][kk][64 + tx * 4]);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float bv[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        if (has_next) store_tiles(buf ^ 1);
        __syncthreads();
        buf ^= 1;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int n = n0 + (i / 4) * 64 + ty * 4 + (i % 4);
        if (n >= N) continue;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int hm = hm0 + (j / 4) * 64 + tx * 4 + (j % 4);
            if (hm < HM) atomicAdd(a.dW + (size_t)n * HM + hm, acc[i][j]);
        }
    }
}

struct CinDxArgs {
    const float* Xp; int64_t sxp; int H;
    const float* X0; int64_t sx0; int M;
    int D; int N;
    const float* W; const float* dZ;
    float* dXp; int64_t sdxp; float* dX0; int64_t sdx0;
    int64_t B; int TB; int HT;   // HT whole h-groups per column tile, HT*M <= BN
    int same;                    // Xp aliases X0 (layer 0): dXp contributions also go to dX0
};

// G[r, hm] = sum_n dZ[b,n,d] W[n,hm] for a column tile of HT h-groups, then
// dXp[b,h,d] = sum_m G X0 (complete -> store) and dX0[b,m,d] += sum_h G Xp (CTA-private).
__global__ void __launch_bounds__(NT) cin_dx_kernel(CinDxArgs a) {
    extern __shared__ __align__(16) float smem[];
    float* As = smem;                                // [2][BK][BM+4]   k = n
    float* Bs = As + 2 * BK * (BM + 4);              // [2][BK][BN+4]
    float* sG = Bs + 2 * BK * (BN + 4);              // [BM][BN+1]
    float* sXp = sG + BM * (BN + 1);                 // [TB][H][D]
    float* sX0 = sXp + (size_t)a.TB * a.H * a.D;     // [TB][M][D]
    float* sdX0 = sX0 + (size_t)a.TB * a.M * a.D;    // [TB][M][D]
    const int tid = threadIdx.x;
    const int tx = tid % 16, ty = tid / 16;
    const int D = a.D, H = a.H, M = a.M, N = a.N;
    const int HM = H * M;
    const int64_t b0 = (int64_t)blockIdx.x * a.TB;
    const int tb = (int)((a.B - b0 < a.TB) ? a.B - b0 : a.TB);
    const int rows = tb * D;
    const int cols_per_tile = a.HT * M;

    for (int i = tid; i < a.TB * H * D; i += NT) {
        const int bl = i / (H * D);
        sXp[i] = (bl < tb) ? __ldg(a.Xp + (b0 + bl) * a.sxp + (i - bl * H * D)) : 0.f;
    }
    for (int i = tid; i < a.TB * M * D; i += NT) {
        const int bl = i / (M * D);
        sX0[i] = (bl < tb) ? __ldg(a.X0 + (b0 + bl) * a.sx0 + (i - bl * M * D)) : 0.f;
        sdX0[i] = 0.f;
    }
    __syncthreads();

    float ra[BM * BK / NT], rb[BN * BK / NT];
    for (int h0 = 0; h0 < H; h0 += a.HT) {
        const int c0 = h0 * M;                                  // first hm column of this tile
        int ncols = cols_per_tile;
        if (c0 + ncols > HM) ncols = HM - c0;
        float acc[8][8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

        auto load_tiles = [&](int k0) {
#pragma unroll
            for (int i = 0; i < BM * BK / NT; ++i) {
                const int e = tid + i * NT;
                const int r = e % BM, kk = e / BM;
                const int n = k0 + kk;
                float v = 0.f;
                if (r < rows && n < N) {
                    const int bl = r / D, d = r - bl * D;
                    v = __ldg(a.dZ + ((b0 + bl) * N + n) * D + d);
                }
                ra[i] = v;
            }
#pragma unroll
            for (int i = 0; i < BN * BK / NT; ++i) {
                const int e = tid + i * NT;
                const int cc = e % BN, kk = e / BN;
                const int n = k0 + kk;
                rb[i] = (cc < ncols && n < N) ? __ldg(a.W + (size_t)n * HM + c0 + cc) : 0.f;
            }
        };
        auto store_tiles = [&](int buf) {
#pragma unroll
            for (int i = 0; i < BM * BK / NT; ++i) {
                const int e = tid + i * NT;
                As[(buf * BK + e / BM) * (BM + 4) + e % BM] = ra[i];
            }
#pragma unroll
            for (int i = 0; i < BN * BK / NT; ++i) {
                const int e = tid + i * NT;
                Bs[(buf * BK + e / BN) * (BN + 4) + e % BN] = rb[i];
            }
        };
        __syncthreads();   // previous tile's phase A/B readers are done with sG/As/Bs
        load_tiles(0);
        store_tiles(0);
        __syncthreads();
        int buf = 0;
        for (int k0 = 0; k0 < N; k0 += BK) {
            const bool has_next = k0 + BK < N;
            if (has_next) load_tiles(k0 + BK);
#pragma unroll
            for (int kk = 0; kk < BK; ++kk) {
                const float4 a0 = *reinterpret_cast<const float4*>(&As[(buf * BK + kk) * (BM + 4) + ty * 4]);
                const float4 a1 = *reinterpret_cast<const float4*>(&As[(buf * BK + kk) * (BM + 4) + 64 + ty * 4]);
                const float4 c0v = *reinterpret_cast<const float4*>(&Bs[(buf * BK + kk) * (BN + 4) + tx * 4]);
                const float4 c1v = *reinterpret_cast<const float4*>(&Bs[(buf * BK + kk) * (BN + 4) + 64 + tx * 4]);
                const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                const float bv[8] = {c0v.x, c0v.y, c0v.z, c0v.w, c1v.x, c1v.y, c1v.z, c1v.w};
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
            }
            if (has_next) store_tiles(buf ^ 1);
            __syncthreads();
            buf ^= 1;
        }
        // G tile -> shared memory
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int r = (i / 4) * 64 + ty * 4 + (i % 4);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int c = (j / 4) * 64 + tx * 4 + (j % 4);
                sG[r * (BN + 1) + c] = acc[i][j];
            }
        }
        __syncthreads();
        const int nh = ncols / M;   // whole h-groups in this tile
        // phase A: dXp[b, h0+hh, d] = sum_m G[r, hh*M+m] * X0[b,m,d]
        for (int idx = tid; idx < rows * nh; idx += NT) {
            const int r = idx % rows, hh = idx / rows;
            const int bl = r / D, d = r - bl * D;
            float s = 0.f;
            for (int m = 0; m < M; ++m) s = fmaf(sG[r * (BN + 1) + hh * M + m], sX0[(bl * M + m) * D + d], s);
            if (a.same) sdX0[(bl * M + (h0 + hh)) * D + d] += s;   // layer 0: Xp is X0 (H == M)
            else a.dXp[(b0 + bl) * a.sdxp + (h0 + hh) * D + d] = s;
        }
        if (a.same) __syncthreads();
        // phase B: dX0[b,m,d] += sum_hh G[r, hh*M+m] * Xp[b,h0+hh,d]   (entry owned by one thread)
        for (int idx = tid; idx < rows * M; idx += NT) {
            const int r = idx % rows, m = idx / rows;
            const int bl = r / D, d = r - bl * D;
            float s = 0.f;
            for (int hh = 0; hh < nh; ++hh)
                s = fmaf(sG[r * (BN + 1) + hh * M + m], sXp[(bl * H + h0 + hh) * D + d], s);
            sdX0[(bl * M + m) * D + d] += s;
        }
    }
    __syncthreads();
    for (int i = tid; i < tb * M * D; i += NT) {
        const int bl = i / (M * D);
        a.dX0[(b0 + bl) * a.sdx0 + (i - bl * M * D)] += sdX0[i];
    }
}

int samples_per_cta(int D) {
    int tb = BM / D;
    return tb < 1 ? 0 : tb;
}

}  // namespace

extern "C" int ctr_cin_layer_fwd(const float* Xp, int64_t sxp, int H, const float* X0, int64_t sx0,
                                 int M, int D, const float* W, const float* bias, int N,
                                 int direct_start, int act, float* Y, float* out, int64_t ld_out,
                                 int64_t B, void* stream) {
    CTR_ARG(Xp && X0 && W && Y && H > 0 && M > 0 && D > 0 && N > 0 && B >= 0, "ctr_cin_layer_fwd: bad arguments");
    CTR_ARG(direct_start >= 0 && direct_start <= N, "ctr_cin_layer_fwd: direct_start out of range");
    CTR_ARG(direct_start == N || out, "ctr_cin_layer_fwd: out required for direct-connect channels");
    CTR_ARG(D <= BM, "ctr_cin_layer_fwd: embedding dim %d > %d unsupported", D, BM);
    if (B == 0) return 0;
    cudaStream_t st = as_stream(stream);
    {   // tensor-core path (cin_tc.cu) unless CTR_GEMM=simt or the shape is unsupported
        const char* e = getenv("CTR_GEMM");
        if (!(e && e[0] == 's')) {
            const int rc2 = launch_cin_v2_fwd(Xp, sxp, H, X0, sx0, M, D, W, bias, N, direct_start, act, Y, out, ld_out, B, st);
            if (rc2 == 1) return 0;
            if (rc2 != 0) return rc2;
            const int rc = launch_cin_tc_fwd(Xp, sxp, H, X0, sx0, M, D, W, bias, N, direct_start, act, Y, out,
                                             ld_out, B, st);
            if (rc == 1) return 0;
            if (rc != 0) return rc;
        }
    }
    const int n_direct = N - direct_start;
    if (n_direct > 0)
        CTR_CUDA(cudaMemset2DAsync(out, ld_out * sizeof(float), 0, n_direct * sizeof(float), B, st));
    CinFwdArgs a{Xp, sxp, H, X0, sx0, M, D, W, bias, N, direct_start, act, Y, out, ld_out, B, samples_per_cta(D)};
    const size_t smem = sizeof(float) * ((size_t)2 * BK * (BM + 4) + 2 * BK * (BN + 4) +
                                         (size_t)a.TB * (H + M) * D);
    CTR_ARG(smem <= 220 * 1024, "ctr_cin_layer_fwd: H=%d M=%d D=%d need %zu B of shared memory", H, M, D, smem);
    CTR_CUDA(cudaFuncSetAttribute(cin_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)ceil_div64(B, a.TB), (unsigned)ceil_div64(N, BN));
    cin_fwd_kernel<<<grid, NT, smem, st>>>(a);
    CTR_LAUNCH_OK("cin_fwd_kernel");
    return 0;
}

extern "C" int ctr_cin_layer_bwd(const float* Xp, int64_t sxp, int H, const float* X0, int64_t sx0,
                                 int M, int D, const float* W, int N, int n_hidden,
                                 int direct_start, int act, const float* Y, const float* dYh, int64_t sdyh, const float* dout,
                                 int64_t ld_dout, float* dZ, float* dW, float* dbias, float* dXp,
                                 int64_t sdxp, float* dX0, int64_t sdx0, int64_t B, void* stream) {
    CTR_ARG(Xp && X0 && W && Y && dZ && dW && dbias && dX0 && H > 0 && M > 0 && D > 0 && N > 0 && B >= 0,
            "ctr_cin_layer_bwd: bad arguments");
    CTR_ARG(D <= BM && M <= BN, "ctr_cin_layer_bwd: D=%d / M=%d exceed the tile (%d / %d)", D, M, BM, BN);
    const int same = (Xp == X0) ? 1 : 0;
    CTR_ARG(same || dXp, "ctr_cin_layer_bwd: dXp required when Xp is not X0");
    CTR_ARG(!same || H == M, "ctr_cin_layer_bwd: Xp aliases X0 but H != M");
    cudaStream_t st = as_stream(stream);
    const int HM = H * M;
    CTR_CUDA(cudaMemsetAsync(dW, 0, sizeof(float) * (size_t)N * HM, st));
    CTR_CUDA(cudaMemsetAsync(dbias, 0, sizeof(float) * N, st));
    if (B == 0) return 0;
    const int sms = ctr_sm_count();
    {
        int64_t blocks = ceil_div64(B * N * D, 256 * 4);
        if (blocks > (int64_t)sms * 8) blocks = (int64_t)sms * 8;
        cin_dz_kernel<<<(unsigned)blocks, 256, 0, st>>>(Y, dYh, sdyh, dout, ld_dout, N, D, n_hidden, direct_start, act, dZ, B);
        CTR_LAUNCH_OK("cin_dz_kernel");
    }
    {
        int64_t blocks = 2LL * sms;
        if (blocks > B) blocks = B;
        const int64_t bpb = ceil_div64(B, blocks);
        blocks = ceil_div64(B, bpb);
        cin_dbias_kernel<<<(unsigned)blocks, 256, 0, st>>>(dZ, N, D, B, bpb, dbias);
        CTR_LAUNCH_OK("cin_dbias_kernel");
    }
    bool dw_done = false;
    {   // tensor-core dW + dX (cin_tc.cu) unless CTR_GEMM=simt or the shape is unsupported
        const char* e = getenv("CTR_GEMM");
        const char* e2 = getenv("CTR_CIN_TC_BWD");
        if (!(e && e[0] == 's') && !(e2 && e2[0] == '0')) {
            const int rcw = launch_cin_v2_dw(Xp, sxp, H, X0, sx0, M, D, N, dZ, dW, B, st);     // experimental, off by default
            if (rcw != 0 && rcw != 1) return rcw;
            dw_done = (rcw == 1);
            const int rc = launch_cin_tc_bwd(Xp, sxp, H, X0, sx0, M, D, W, N, dZ, dW, dXp, sdxp, dX0, sdx0, B, same, st,
                                             rcw == 1);
            if (rc == 1) return 0;
            if (rc != 0) return rc;
        }
    }
    if (!dw_done) {
        const int64_t tiles = ceil_div64(HM, BN) * ceil_div64(N, BM);
        int64_t splits = ceil_div64(2LL * sms, tiles);
        if (splits > B) splits = B;
        if (splits < 1) splits = 1;
        const int64_t bpb = ceil_div64(B, splits);
        splits = ceil_div64(B, bpb);
        CTR_ARG(splits <= 65535, "ctr_cin_layer_bwd: too many K splits");
        CinDwArgs a{Xp, sxp, H, X0, sx0, M, D, N, dZ, dW, B, bpb};
        dim3 grid((unsigned)ceil_div64(HM, BN), (unsigned)ceil_div64(N, BM), (unsigned)splits);
        cin_dw_kernel<<<grid, NT, 0, st>>>(a);
        CTR_LAUNCH_OK("cin_dw_kernel");
    }
    {
        CinDxArgs a{Xp, sxp, H, X0, sx0, M, D, N, W, dZ, dXp, sdxp, dX0, sdx0, B, samples_per_cta(D), BN / M, same};
        if (a.HT > H) a.HT = H;
        const size_t smem = sizeof(float) * ((size_t)2 * BK * (BM + 4) + 2 * BK * (BN + 4) + (size_t)BM * (BN + 1) +
                                             (size_t)a.TB * (H + 2 * M) * D);
        CTR_ARG(smem <= 220 * 1024, "ctr_cin_layer_bwd: H=%d M=%d D=%d need %zu B of shared memory", H, M, D, smem);
        CTR_CUDA(cudaFuncSetAttribute(cin_dx_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        cin_dx_kernel<<<(unsigned)ceil_div64(B, a.TB), NT, smem, st>>>(a);
        CTR_LAUNCH_OK("cin_dx_kernel");
    }
    return 0;
}
