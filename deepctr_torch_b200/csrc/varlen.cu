// VarLenSparseFeat pooled lookup: gather T rows per sample and pool them (sum | mean | max) in
// registers — the [B,T,D] sequence tensor of the reference is never written.
// Reference: inputs.py:141-155,213-227 (lookup + mask construction), layers/sequence.py:49-77.
#include "common.cuh"

namespace {

// id_mode 0: fp32-encoded integers (`.long()` truncation); 1: the cell holds the int32 itself
__device__ __forceinline__ int64_t cell_int(const float* p, int id_mode) {
    const float v = __ldg(p);
    return id_mode ? (int64_t)__float_as_int(v) : (int64_t)v;
}

__device__ __forceinline__ int64_t id_at(const float* xrow, int col, int vocab, int32_t* err_flag, int id_mode) {
    int64_t id = cell_int(xrow + col, id_mode);
    if (id < 0 || id >= (int64_t)vocab) {
        if (err_flag) atomicOr(err_flag, 1);
        id = 0;
    }
    return id;
}

// mask of position t: id != 0 (len_col < 0) or t < length (length column)
__device__ __forceinline__ bool pos_valid(const float* xrow, int col, int t, int len_col, int id_mode) {
    if (len_col < 0) return cell_int(xrow + col + t, id_mode) != 0;
    return (int64_t)t < cell_int(xrow + len_col, id_mode);
}

__global__ void __launch_bounds__(256) varlen_pool_fwd_kernel(const float* __restrict__ X, int64_t ldx,
                                                              int64_t B, int col, int T, int len_col,
                                                              const float* __restrict__ table,
                                                              int vocab, int D, int mode, float* dst,
                                                              int64_t ld, int32_t* err_flag, int id_mode) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t b = warp0; b < B; b += nwarps) {
        const float* xrow = X + b * ldx;
        int count = 0;
        for (int t = 0; t < T; ++t) count += pos_valid(xrow, col, t, len_col, id_mode) ? 1 : 0;
        for (int d = lane; d < D; d += 32) {
            float acc = (mode == 2) ? -INFINITY : 0.f;
            for (int t = 0; t < T; ++t) {
                const int64_t id = id_at(xrow, col + t, vocab, err_flag, id_mode);
                const float v = __ldg(table + id * D + d);
                const float m = pos_valid(xrow, col, t, len_col, id_mode) ? 1.f : 0.f;
                if (mode == 2) acc = fmaxf(acc, v - (1.f - m) * 1e9f);
                else acc += v * m;
            }
            if (mode == 1) acc = acc / ((float)count + 1e-8f);
            dst[b * ld + d] = acc;
        }
    }
}

__global__ void __launch_bounds__(256) varlen_pool_bwd_kernel(const float* __restrict__ X, int64_t ldx,
                                                              int64_t B, int col, int T, int len_col,
                                                              const float* __restrict__ table,
                                                              int vocab, int D, int mode,
                                                              const float* __restrict__ ddst,
                                                              int64_t ld, float* dtable, int id_mode) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t b = warp0; b < B; b += nwarps) {
        const float* xrow = X + b * ldx;
        int count = 0;
        for (int t = 0; t < T; ++t) count += pos_valid(xrow, col, t, len_col, id_mode) ? 1 : 0;
        for (int d = lane; d < D; d += 32) {
            const float g = __ldg(ddst + b * ld + d);
            if (mode == 2) {
                float best = -INFINITY;
                int64_t best_id = 0;
                for (int t = 0; t < T; ++t) {
                    const int64_t id = id_at(xrow, col + t, vocab, nullptr, id_mode);
                    const float m = pos_valid(xrow, col, t, len_col, id_mode) ? 1.f : 0.f;
                    const float v = __ldg(table + id * D + d) - (1.f - m) * 1e9f;
                    if (v > best) {   // first maximum wins, like torch.max(dim)
                        best = v;
                        best_id = id;
                    }
                }
                atomicAdd(dtable + best_id * D + d, g);
            } else {
                const float scale = (mode == 1) ? 1.f / ((float)count + 1e-8f) : 1.f;
                for (int t = 0; t < T; ++t) {
                    if (!pos_valid(xrow, col, t, len_col, id_mode)) continue;
                    const int64_t id = id_at(xrow, col + t, vocab, nullptr, id_mode);
                    atomicAdd(dtable + id * D + d, g * scale);
                }
            }
        }
    }
}

unsigned warp_grid(int64_t B) {
    int64_t blocks = ceil_div64(B, 8);
    const int64_t cap = (int64_t)ctr_sm_count() * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (unsigned)blocks;
}

}  // namespace

extern "C" int ctr_varlen_pool_fwd(const float* X, int64_t ldx, int64_t B, int col, int T,
                                   int len_col, const float* table, int vocab, int D, int mode,
                                   float* dst, int64_t ld, int32_t* err_flag, int id_mode, void* stream) {
    CTR_ARG(X && table && dst && T > 0 && D > 0 && vocab > 0 && B >= 0 && col >= 0, "ctr_varlen_pool_fwd: bad arguments");
    CTR_ARG(mode >= 0 && mode <= 2, "ctr_varlen_pool_fwd: mode must be 0 (sum), 1 (mean) or 2 (max)");
    if (B == 0) return 0;
    varlen_pool_fwd_kernel<<<warp_grid(B), 256, 0, as_stream(stream)>>>(X, ldx, B, col, T, len_col, table, vocab,
                                                                        D, mode, dst, ld, err_flag, id_mode);
    CTR_LAUNCH_OK("varlen_pool_fwd_kernel");
    return 0;
}

extern "C" int ctr_varlen_pool_bwd(const float* X, int64_t ldx, int64_t B, int col, int T,
                                   int len_col, const float* table, int vocab, int D, int mode,
                                   const float* ddst, int64_t ld, float* dtable, int id_mode, void* stream) {
    CTR_ARG(X && table && ddst && dtable && T > 0 && D > 0 && vocab > 0 && B >= 0 && col >= 0,
            "ctr_varlen_pool_bwd: bad arguments");
    CTR_ARG(mode >= 0 && mode <= 2, "ctr_varlen_pool_bwd: mode must be 0 (sum), 1 (mean) or 2 (max)");
    if (B == 0) return 0;
    varlen_pool_bwd_kernel<<<warp_grid(B), 256, 0, as_stream(stream)>>>(X, ldx, B, col, T, len_col, table, vocab,
                                                                        D, mode, ddst, ld, dtable, id_mode);
    CTR_LAUNCH_OK("varlen_pool_bwd_kernel");
    return 0;
}
