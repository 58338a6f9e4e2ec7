// Summed binary cross-entropy on the predictions (reference models/basemodel.py:254:
// F.binary_cross_entropy(y_pred, y, reduction='sum')), same clamping as ATen (log terms >= -100,
// gradient denominator >= 1e-12).  Two tiny HBM-bound kernels instead of three ATen launches inside the step.
#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256) bce_sum_fwd_kernel(const float* __restrict__ p, const float* __restrict__ y, int64_t B,
                                                          float* out) {
    __shared__ float s_part[8];
    float acc = 0.f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < B; i += (int64_t)gridDim.x * blockDim.x) {
        const float pv = __ldg(p + i), yv = __ldg(y + i);
        acc += (yv - 1.f) * fmaxf(logf(1.f - pv), -100.f) - yv * fmaxf(logf(pv), -100.f);
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float v = 0.f;
        for (int w = 0; w < 8; ++w) v += s_part[w];
        atomicAdd(out, v);
    }
}

__global__ void __launch_bounds__(256) bce_sum_bwd_kernel(const float* __restrict__ p, const float* __restrict__ y,
                                                          const float* __restrict__ g, int64_t B, float* dp) {
    const float gv = __ldg(g);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < B; i += (int64_t)gridDim.x * blockDim.x) {
        const float pv = __ldg(p + i), yv = __ldg(y + i);
        dp[i] = gv * (pv - yv) / fmaxf((1.f - pv) * pv, 1e-12f);
    }
}

unsigned loss_grid(int64_t B) {
    int64_t blocks = ceil_div64(B, 256 * 4);
    const int64_t cap = (int64_t)ctr_sm_count() * 4;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (unsigned)blocks;
}

}  // namespace

extern "C" int ctr_bce_sum_fwd(const float* y_pred, const float* y, int64_t B, float* out, void* stream) {
    CTR_ARG(y_pred && y && out && B >= 0, "ctr_bce_sum_fwd: bad arguments");
    cudaStream_t st = as_stream(stream);
    CTR_CUDA(cudaMemsetAsync(out, 0, sizeof(float), st));
    if (B == 0) return 0;
    bce_sum_fwd_kernel<<<loss_grid(B), 256, 0, st>>>(y_pred, y, B, out);
    CTR_LAUNCH_OK("bce_sum_fwd_kernel");
    return 0;
}

extern "C" int ctr_bce_sum_bwd(const float* y_pred, const float* y, const float* g, int64_t B, float* d_pred, void* stream) {
    CTR_ARG(y_pred && y && g && d_pred && B >= 0, "ctr_bce_sum_bwd: bad arguments");
    if (B == 0) return 0;
    bce_sum_bwd_kernel<<<loss_grid(B), 256, 0, as_stream(stream)>>>(y_pred, y, g, B, d_pred);
    CTR_LAUNCH_OK("bce_sum_bwd_kernel");
    return 0;
}

// ---------------------------------------------------------------------------------------------
// PReLU (reference layers/activation.py:61-62: nn.PReLU(), one learnable slope shared by all units):
//   y = max(z, 0) + alpha * min(z, 0);   dz = dy * (z > 0 ? 1 : alpha);   dalpha = sum dy * min(z, 0)
// ---------------------------------------------------------------------------------------------
namespace {

__global__ void __launch_bounds__(256) prelu_fwd_kernel(const float* __restrict__ z, const float* __restrict__ alpha,
                                                        int64_t n, float* y) {
    const float a = __ldg(alpha);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float v = __ldg(z + i);
        y[i] = v > 0.f ? v : a * v;
    }
}

__global__ void __launch_bounds__(256) prelu_bwd_kernel(const float* __restrict__ z, const float* __restrict__ alpha,
                                                        const float* __restrict__ dy, int64_t n, float* dz, float* dalpha) {
    __shared__ float s_part[8];
    const float a = __ldg(alpha);
    float acc = 0.f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float v = __ldg(z + i), g = __ldg(dy + i);
        dz[i] = v > 0.f ? g : a * g;
        acc += v > 0.f ? 0.f : g * v;
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float v = 0.f;
        for (int w = 0; w < 8; ++w) v += s_part[w];
        atomicAdd(dalpha, v);
    }
}

}  // namespace

extern "C" int ctr_prelu_fwd(const float* z, const float* alpha, int64_t n, float* y, void* stream) {
    CTR_ARG(z && alpha && y && n >= 0, "ctr_prelu_fwd: bad arguments");
    if (n == 0) return 0;
    prelu_fwd_kernel<<<loss_grid(n), 256, 0, as_stream(stream)>>>(z, alpha, n, y);
    CTR_LAUNCH_OK("prelu_fwd_kernel");
    return 0;
}

extern "C" int ctr_prelu_bwd(const float* z, const float* alpha, const float* dy, int64_t n, float* dz, float* dalpha,
                             void* stream) {
    CTR_ARG(z && alpha && dy && dz && dalpha && n >= 0, "ctr_prelu_bwd: bad arguments");
    cudaStream_t st = as_stream(stream);
    CTR_CUDA(cudaMemsetAsync(dalpha, 0, sizeof(float), st));
    if (n == 0) return 0;
    prelu_bwd_kernel<<<loss_grid(n), 256, 0, st>>>(z, alpha, dy, n, dz, dalpha);
    CTR_LAUNCH_OK("prelu_bwd_kernel");
    return 0;
}
