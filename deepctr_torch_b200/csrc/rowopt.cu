// Fused row-wise optimizer on the row-gradient stream and the owner-side combine of
// the sharded backward.
//
// Replaces, for the embedding / linear tables, `optim.step()` and the table part of
// `get_regularization_loss().backward()` (reference models/basemodel.py:262, 412-428, 447-461): the
// reference materialises a dense [V,D] gradient per table and lets a torch optimizer sweep the whole
// table (1.77 GB x {w, g, m, v} at BASELINE config #2).  Here the (uniq, rowgrad) pairs written by
// ctr_scatter_bwd_rowwise are consumed in place: one lane group per touched row, 128-bit accesses,
// every touched row read and written exactly once.  HBM-bound: per touched row 2 x D x 4 B (w) +
// D x 4 B (g) + 2 x D x 4 B per state tensor.
#include "common.cuh"

namespace {

struct OptArgs {
    int kind;
    int64_t cap;
    const int32_t* n_uniq;
    const int32_t* uniq;
    int n_fields, D;
    const float* rowgrad;
    int64_t rg_stride;
    const int32_t* plan_col;
    float* const* tables;
    float* const* s1;
    float* const* s2;
    const float* hp;
    float l2x2;
};

__device__ __forceinline__ float opt_update(int kind, float w, float g, float& a, float& b, float lr, float b1, float b2,
                                            float eps, float bc1, float bc2_sqrt) {
    if (kind == CTR_OPT_ADAGRAD) {
        a += g * g;
        return w - lr * g / (sqrtf(a) + eps);
    }
    if (kind == CTR_OPT_ADAM) {
        a = b1 * a + (1.f - b1) * g;
        b = b2 * b + (1.f - b2) * g * g;
        const float denom = sqrtf(b) / bc2_sqrt + eps;
        return w - (lr / bc1) * (a / denom);
    }
    if (kind == CTR_OPT_RMSPROP) {
        a = b1 * a + (1.f - b1) * g * g;
        return w - lr * g / (sqrtf(a) + eps);
    }
    return w - lr * g;
}

// VEC: D % 4 == 0 — D/4 lanes per row, float4 accesses; otherwise one thread per element
template <bool VEC>
__global__ void __launch_bounds__(256) rowopt_kernel(OptArgs a) {
    const int f = blockIdx.y;
    const int pc = a.plan_col[f];
    const int64_t nu = a.n_uniq[pc];
    const float lr = a.hp[1], b1 = a.hp[2], b2 = a.hp[3], eps = a.hp[4], bc1 = a.hp[5], bc2s = sqrtf(a.hp[6]);
    float* W = a.tables[f];
    float* S1 = a.s1 ? a.s1[f] : nullptr;
    float* S2 = a.s2 ? a.s2[f] : nullptr;
    const int per_row = VEC ? (a.D >> 2) : a.D;
    const int64_t total = nu * per_row;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t u = i / per_row;
        const int sub = (int)(i - u * per_row);
        const int64_t id = a.uniq[(int64_t)pc * a.cap + u];
        const float* gp = a.rowgrad + f * a.rg_stride + u * a.D;
        if (VEC) {
            const int64_t off = id * a.D + sub * 4;
            float4 w = *reinterpret_cast<const float4*>(W + off);
            float4 g = ld_stream4(gp + sub * 4);
            g.x += a.l2x2 * w.x; g.y += a.l2x2 * w.y; g.z += a.l2x2 * w.z; g.w += a.l2x2 * w.w;
            float4 m = make_float4(0.f, 0.f, 0.f, 0.f), v = m;
            if (S1) m = *reinterpret_cast<const float4*>(S1 + off);
            if (S2) v = *reinterpret_cast<const float4*>(S2 + off);
            w.x = opt_update(a.kind, w.x, g.x, m.x, v.x, lr, b1, b2, eps, bc1, bc2s);
            w.y = opt_update(a.kind, w.y, g.y, m.y, v.y, lr, b1, b2, eps, bc1, bc2s);
            w.z = opt_update(a.kind, w.z, g.z, m.z, v.z, lr, b1, b2, eps, bc1, bc2s);
            w.w = opt_update(a.kind, w.w, g.w, m.w, v.w, lr, b1, b2, eps, bc1, bc2s);
            *reinterpret_cast<float4*>(W + off) = w;
            if (S1) *reinterpret_cast<float4*>(S1 + off) = m;
            if (S2) *reinterpret_cast<float4*>(S2 + off) = v;
        } else {
            const int64_t off = id * a.D + sub;
            float w = W[off];
            const float g = gp[sub] + a.l2x2 * w;
            float m = S1 ? S1[off] : 0.f, v = S2 ? S2[off] : 0.f;
            w = opt_update(a.kind, w, g, m, v, lr, b1, b2, eps, bc1, bc2s);
            W[off] = w;
            if (S1) S1[off] = m;
            if (S2) S2[off] = v;
        }
    }
}

__global__ void rowopt_tick_kernel(float* hp) {
    const float step = hp[0] + 1.f;
    hp[0] = step;
    hp[5] = 1.f - powf(hp[2], step);
    hp[6] = 1.f - powf(hp[3], step);
}

// out[f][inv[b, pc(f)]] += recv[f][b], b < count[pc(f)]: rows of the same id meet in one output row
// (receive lists belong to the id columns of the plan; every field of a column shares the slots)
template <bool VEC>
__global__ void __launch_bounds__(256) combine_kernel(int64_t cap, int D, const int32_t* __restrict__ count,
                                                      const int32_t* __restrict__ inv, int n_plan,
                                                      const int32_t* __restrict__ plan_col,
                                                      const float* __restrict__ recv, int64_t recv_stride,
                                                      float* out, int64_t out_stride) {
    const int f = blockIdx.y;
    const int pc = plan_col[f];
    const int64_t n = count[pc] < cap ? count[pc] : cap;
    const int per_row = VEC ? (D >> 2) : D;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n * per_row; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = i / per_row;
        const int sub = (int)(i - b * per_row);
        const int64_t u = inv[b * n_plan + pc];
        if (VEC) red_add4(out + f * out_stride + u * D + sub * 4, ld_stream4(recv + f * recv_stride + b * D + sub * 4));
        else atomicAdd(out + f * out_stride + u * D + sub, recv[f * recv_stride + b * D + sub]);
    }
}

// rows u < n_uniq[pc(f)] of the combined buffer are zeroed before the reductions land in them
template <bool VEC>
__global__ void __launch_bounds__(256) combine_zero_kernel(int64_t cap, int D, const int32_t* __restrict__ n_uniq,
                                                           const int32_t* __restrict__ plan_col, float* out,
                                                           int64_t out_stride) {
    const int f = blockIdx.y;
    const int64_t nu = n_uniq[plan_col[f]];
    const int per_row = VEC ? (D >> 2) : D;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nu * per_row; i += (int64_t)gridDim.x * blockDim.x) {
        if (VEC) *reinterpret_cast<float4*>(out + f * out_stride + i * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
        else out[f * out_stride + i] = 0.f;
    }
}

struct PtrPack {
    const void* p[64];
};
__global__ void write_ptrs_kernel(PtrPack pk, int n, void** out) {
    if (threadIdx.x < n) out[threadIdx.x] = const_cast<void*>(pk.p[threadIdx.x]);
}

}  // namespace

extern "C" int ctr_rowopt_tick(float* hp, void* stream) {
    CTR_ARG(hp, "ctr_rowopt_tick: hp missing");
    rowopt_tick_kernel<<<1, 1, 0, as_stream(stream)>>>(hp);
    CTR_LAUNCH_OK("rowopt_tick_kernel");
    return 0;
}

extern "C" int ctr_rowopt_step(int kind, int64_t cap, const int32_t* n_uniq, const int32_t* uniq, int n_fields, int D,
                               const float* rowgrad, int64_t rowgrad_stride, const int32_t* plan_col,
                               float* const* tables, float* const* state1, float* const* state2, const float* hp,
                               float l2x2, void* stream) {
    CTR_ARG(kind >= CTR_OPT_SGD && kind <= CTR_OPT_RMSPROP, "ctr_rowopt_step: unknown optimizer kind %d", kind);
    CTR_ARG(cap >= 0 && n_fields >= 0 && D > 0 && hp, "ctr_rowopt_step: bad sizes");
    if (n_fields == 0 || cap == 0) return 0;
    CTR_ARG(n_uniq && uniq && rowgrad && plan_col && tables && rowgrad_stride >= cap * D, "ctr_rowopt_step: null argument");
    CTR_ARG(kind == CTR_OPT_SGD || state1, "ctr_rowopt_step: optimizer state missing");
    CTR_ARG(kind != CTR_OPT_ADAM || state2, "ctr_rowopt_step: second Adam moment missing");
    OptArgs a{kind, cap, n_uniq, uniq, n_fields, D, rowgrad, rowgrad_stride, plan_col, tables,
              kind == CTR_OPT_SGD ? nullptr : state1, kind == CTR_OPT_ADAM ? state2 : nullptr, hp, l2x2};
    const bool vec = (D % 4 == 0) && (rowgrad_stride % 4 == 0) && ((reinterpret_cast<uintptr_t>(rowgrad) & 15) == 0);
    int64_t bx = ceil_div64(cap * (vec ? D / 4 : D), 256);
    const int64_t limit = ceil_div64((int64_t)ctr_sm_count() * 8, n_fields);
    if (bx > limit) bx = limit;
    if (bx < 1) bx = 1;
    dim3 grid((unsigned)bx, (unsigned)n_fields);
    if (vec) rowopt_kernel<true><<<grid, 256, 0, as_stream(stream)>>>(a);
    else rowopt_kernel<false><<<grid, 256, 0, as_stream(stream)>>>(a);
    CTR_LAUNCH_OK("rowopt_kernel");
    return 0;
}

extern "C" int ctr_rowgrad_combine(int64_t cap, int n_fields, int D, const int32_t* count, const int32_t* n_uniq,
                                   const int32_t* inv, int n_plan_cols, const int32_t* plan_col, const float* recv,
                                   int64_t recv_stride, float* out, int64_t out_stride, void* stream) {
    CTR_ARG(cap >= 0 && n_fields >= 0 && D > 0 && n_plan_cols > 0, "ctr_rowgrad_combine: bad sizes");
    if (n_fields == 0 || cap == 0) return 0;
    CTR_ARG(count && n_uniq && inv && plan_col && recv && out, "ctr_rowgrad_combine: null argument");
    const bool vec = (D % 4 == 0) && (recv_stride % 4 == 0) && (out_stride % 4 == 0) &&
                     ((reinterpret_cast<uintptr_t>(recv) & 15) == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
    int64_t bx = ceil_div64(cap * (vec ? D / 4 : D), 256);
    const int64_t limit = ceil_div64((int64_t)ctr_sm_count() * 8, n_fields);
    if (bx > limit) bx = limit;
    if (bx < 1) bx = 1;
    dim3 grid((unsigned)bx, (unsigned)n_fields);
    if (vec) combine_zero_kernel<true><<<grid, 256, 0, as_stream(stream)>>>(cap, D, n_uniq, plan_col, out, out_stride);
    else combine_zero_kernel<false><<<grid, 256, 0, as_stream(stream)>>>(cap, D, n_uniq, plan_col, out, out_stride);
    CTR_LAUNCH_OK("combine_zero_kernel");
    if (vec) combine_kernel<true><<<grid, 256, 0, as_stream(stream)>>>(cap, D, count, inv, n_plan_cols, plan_col, recv, recv_stride, out, out_stride);
    else combine_kernel<false><<<grid, 256, 0, as_stream(stream)>>>(cap, D, count, inv, n_plan_cols, plan_col, recv, recv_stride, out, out_stride);
    CTR_LAUNCH_OK("combine_kernel");
    return 0;
}

extern "C" int ctr_write_ptrs(const void* const* host_ptrs, int n, void** dev_out, void* stream) {
    CTR_ARG(n >= 0 && n <= 4096 && (n == 0 || (host_ptrs && dev_out)), "ctr_write_ptrs: bad arguments");
    for (int i0 = 0; i0 < n; i0 += 64) {
        PtrPack pk{};
        const int m = n - i0 < 64 ? n - i0 : 64;
        for (int i = 0; i < m; ++i) pk.p[i] = host_ptrs[i0 + i];
        write_ptrs_kernel<<<1, 64, 0, as_stream(stream)>>>(pk, m, dev_out + i0);
        CTR_LAUNCH_OK("write_ptrs_kernel");
    }
    return 0;
}
