// Row-sharded embedding tables across the GPUs of one node, exchanged over NVLink peer memory.
//
// Replaces the reference's single-process torch.nn.DataParallel (models/basemodel.py:206-209),
// which re-broadcasts every table to every GPU each step.  Here table f is split row-wise:
// GPU s owns the rows {id : id % G == s}, stored at local index id / G.  There is no NCCL
// collective on the sparse data path:
//   forward : the fused gather kernel (gather.cu) loads remote rows directly through peer
//             pointers (NVLink P2P loads), so "all-to-all of ids + all-to-all of rows" collapses
//             into the gather itself;
//   backward: after the local duplicate-free row gradients are formed, ctr_rowgrad_push appends
//             each (local row, gradient row) to its owner's receive list with a remote atomic slot
//             claim + 128-bit P2P stores.
// Peer pointers come from CUDA IPC: ctr_p2p_alloc/export/open below.
#include "common.cuh"

namespace {

struct PushArgs {
    int64_t B;
    int G;
    const int32_t* n_uniq;
    const int32_t* uniq;
    int n_emb, D;
    const float* emb_rowgrad;
    int64_t emb_rg_stride;
    const int32_t* emb_plan_col;
    int n_lin;
    const float* lin_rowgrad;
    int64_t lin_rg_stride;
    const int32_t* lin_plan_col;
    int32_t* const* recv_count;
    int32_t* const* recv_ids;
    float* const* recv_emb_rows;
    float* const* recv_lin_rows;
    int64_t cap;
    int32_t* err_flag;
};

// one thread per (field, unique row)
__global__ void __launch_bounds__(256) rowgrad_push_kernel(PushArgs a) {
    const int64_t total = (int64_t)(a.n_emb + a.n_lin) * a.B;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int f = (int)(i / a.B);
        const int64_t u = i - (int64_t)f * a.B;
        const bool is_emb = f < a.n_emb;
        const int pc = is_emb ? a.emb_plan_col[f] : a.lin_plan_col[f - a.n_emb];
        if (u >= a.n_uniq[pc]) continue;
        const int32_t id = a.uniq[(int64_t)pc * a.B + u];
        const int owner = id % a.G;
        const int32_t local = id / a.G;
        const int32_t slot = atomicAdd(a.recv_count[owner] + f, 1);   // remote atomic over NVLink
        if (slot >= a.cap) {
            atomicOr(a.err_flag, 2);
            continue;
        }
        a.recv_ids[owner][(int64_t)f * a.cap + slot] = local;
        if (is_emb) {
            const float* src = a.emb_rowgrad + f * a.emb_rg_stride + u * a.D;
            float* dst = a.recv_emb_rows[owner] + ((int64_t)f * a.cap + slot) * a.D;
            if ((a.D & 3) == 0) {
                for (int d = 0; d < a.D; d += 4)
                    *reinterpret_cast<float4*>(dst + d) = *reinterpret_cast<const float4*>(src + d);
            } else {
                for (int d = 0; d < a.D; ++d) dst[d] = src[d];
            }
        } else {
            const int fl = f - a.n_emb;
            a.recv_lin_rows[owner][(int64_t)fl * a.cap + slot] = a.lin_rowgrad[fl * a.lin_rg_stride + u];
        }
    }
}

}  // namespace

extern "C" int ctr_p2p_alloc(int64_t bytes, void** ptr) {
    CTR_ARG(ptr && bytes > 0, "ctr_p2p_alloc: bad arguments");
    CTR_CUDA(cudaMalloc(ptr, (size_t)bytes));
    CTR_CUDA(cudaMemset(*ptr, 0, (size_t)bytes));
    return 0;
}

extern "C" int ctr_p2p_free(void* ptr) {
    if (ptr) CTR_CUDA(cudaFree(ptr));
    return 0;
}

extern "C" int ctr_p2p_export(void* ptr, unsigned char* handle64) {
    CTR_ARG(ptr && handle64, "ctr_p2p_export: bad arguments");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle is 64 bytes");
    cudaIpcMemHandle_t h;
    CTR_CUDA(cudaIpcGetMemHandle(&h, ptr));
    memcpy(handle64, &h, 64);
    return 0;
}

extern "C" int ctr_p2p_open(const unsigned char* handle64, void** peer_ptr) {
    CTR_ARG(handle64 && peer_ptr, "ctr_p2p_open: bad arguments");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    CTR_CUDA(cudaIpcOpenMemHandle(peer_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return 0;
}

extern "C" int ctr_p2p_close(void* peer_ptr) {
    if (peer_ptr) CTR_CUDA(cudaIpcCloseMemHandle(peer_ptr));
    return 0;
}

extern "C" int ctr_rowgrad_push(int64_t B, int n_shards, const int32_t* n_uniq, const int32_t* uniq,
                                int n_emb, int D, const float* emb_rowgrad, int64_t emb_rowgrad_stride,
                                const int32_t* emb_plan_col, int n_lin, const float* lin_rowgrad,
                                int64_t lin_rowgrad_stride, const int32_t* lin_plan_col,
                                int32_t* const* recv_count, int32_t* const* recv_ids,
                                float* const* recv_emb_rows, float* const* recv_lin_rows,
                                int64_t cap, int32_t* err_flag, void* stream) {
    CTR_ARG(B >= 0 && n_shards >= 1 && n_uniq && uniq && recv_count && recv_ids && err_flag && cap > 0,
            "ctr_rowgrad_push: bad arguments");
    CTR_ARG(n_emb == 0 || (D > 0 && emb_rowgrad && emb_plan_col && recv_emb_rows), "ctr_rowgrad_push: embedding arrays missing");
    CTR_ARG(n_lin == 0 || (lin_rowgrad && lin_plan_col && recv_lin_rows), "ctr_rowgrad_push: linear arrays missing");
    if (B == 0 || n_emb + n_lin == 0) return 0;
    PushArgs a{B, n_shards, n_uniq, uniq, n_emb, D, emb_rowgrad, emb_rowgrad_stride, emb_plan_col, n_lin,
               lin_rowgrad, lin_rowgrad_stride, lin_plan_col, recv_count, recv_ids, recv_emb_rows, recv_lin_rows, cap, err_flag};
    int64_t blocks = ceil_div64((int64_t)(n_emb + n_lin) * B, 256);
    const int64_t limit = (int64_t)ctr_sm_count() * 8;
    if (blocks > limit) blocks = limit;
    rowgrad_push_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(a);
    CTR_LAUNCH_OK("rowgrad_push_kernel");
    return 0;
}
