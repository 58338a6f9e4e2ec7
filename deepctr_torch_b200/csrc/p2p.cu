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

// One block = one id column of the plan x a chunk of PUSH_CHUNK unique rows.  A receive list belongs to an id
// COLUMN, not to a field: every field that reads the column (DeepFM: one embedding table + one linear table)
// delivers its row into the same slot, so the owner builds ONE duplicate-free plan per column (round 1 kept a list
// per field: twice the ids, twice the slot claims and twice the owner-side plan).
// The slot claim in the owner's receive list is aggregated per block: positions inside the block come from
// shared-memory atomics, and ONE remote atomicAdd per (block, owner) reserves the range — a remote atomic that
// returns a value is a full NVLink round trip, and the first version issued one per row (3.4 M per step: 4.5 ms at G = 2).
constexpr int PUSH_CHUNK = 1024;
constexpr int PUSH_MAX_G = 16;
__global__ void __launch_bounds__(256) rowgrad_push_kernel(PushArgs a) {
    __shared__ int s_cnt[PUSH_MAX_G];
    __shared__ int s_base[PUSH_MAX_G];
    __shared__ int s_dst[PUSH_CHUNK];
    const int pc = blockIdx.y;
    const int64_t nu = a.n_uniq[pc];
    const bool vec_rows = (a.D & 3) == 0 && a.D <= 128 && 256 % (a.D >> 2) == 0;
    for (int64_t u0 = (int64_t)blockIdx.x * PUSH_CHUNK; u0 < nu; u0 += (int64_t)gridDim.x * PUSH_CHUNK) {
        if (threadIdx.x < PUSH_MAX_G) s_cnt[threadIdx.x] = 0;
        __syncthreads();
        int owner[PUSH_CHUNK / 256], pos[PUSH_CHUNK / 256];
        int32_t local[PUSH_CHUNK / 256];
#pragma unroll
        for (int j = 0; j < PUSH_CHUNK / 256; ++j) {
            const int64_t u = u0 + threadIdx.x + 256 * j;
            owner[j] = -1;
            if (u < nu) {
                const int32_t id = a.uniq[(int64_t)pc * a.B + u];
                owner[j] = id % a.G;
                local[j] = id / a.G;
                pos[j] = atomicAdd(&s_cnt[owner[j]], 1);
            }
        }
        __syncthreads();
        if (threadIdx.x < a.G && s_cnt[threadIdx.x] > 0)
            s_base[threadIdx.x] = atomicAdd(a.recv_count[threadIdx.x] + pc, s_cnt[threadIdx.x]);   // remote, once per owner
        __syncthreads();
#pragma unroll
        for (int j = 0; j < PUSH_CHUNK / 256; ++j) {     // destinations into shared memory ...
            int dst = -1;
            if (owner[j] >= 0) {
                const int64_t slot = (int64_t)s_base[owner[j]] + pos[j];
                if (slot >= a.cap) atomicOr(a.err_flag, 2);
                else {
                    dst = (owner[j] << 26) | (int)slot;
                    a.recv_ids[owner[j]][(int64_t)pc * a.cap + slot] = local[j];
                }
            }
            s_dst[threadIdx.x + 256 * j] = dst;
        }
        __syncthreads();
        // ... so that D/4 lanes move one row: a warp writes 8 consecutive source rows with 128-bit
        // stores (the first version let one thread write its whole 64-byte row: 0.40 ms per step)
        for (int f = 0; f < a.n_emb; ++f) {
            if (a.emb_plan_col[f] != pc) continue;       // block-uniform
            if (vec_rows) {
                const int lpr = a.D >> 2, sub = threadIdx.x % lpr;
                for (int it = threadIdx.x / lpr; it < PUSH_CHUNK; it += 256 / lpr) {
                    const int dst = s_dst[it];
                    if (dst < 0) continue;
                    const int64_t u = u0 + it;
                    const float4 v = *reinterpret_cast<const float4*>(a.emb_rowgrad + f * a.emb_rg_stride + u * a.D + sub * 4);
                    *reinterpret_cast<float4*>(a.recv_emb_rows[dst >> 26] + ((int64_t)f * a.cap + (dst & ((1 << 26) - 1))) * a.D + sub * 4) = v;
                }
            } else {
                for (int it = threadIdx.x; it < PUSH_CHUNK; it += 256) {
                    const int dst = s_dst[it];
                    if (dst < 0) continue;
                    const float* src = a.emb_rowgrad + f * a.emb_rg_stride + (u0 + it) * a.D;
                    float* d2 = a.recv_emb_rows[dst >> 26] + ((int64_t)f * a.cap + (dst & ((1 << 26) - 1))) * a.D;
                    for (int d = 0; d < a.D; ++d) d2[d] = src[d];
                }
            }
        }
        for (int fl = 0; fl < a.n_lin; ++fl) {
            if (a.lin_plan_col[fl] != pc) continue;      // block-uniform
            for (int it = threadIdx.x; it < PUSH_CHUNK; it += 256) {
                const int dst = s_dst[it];
                if (dst < 0) continue;
                a.recv_lin_rows[dst >> 26][(int64_t)fl * a.cap + (dst & ((1 << 26) - 1))] = a.lin_rowgrad[fl * a.lin_rg_stride + u0 + it];
            }
        }
        __syncthreads();
    }
}


// ---------------------------------------------------------------------------------------------
// Forward exchange of cross-shard rows ("all-to-all" over NVLink peer memory, owner computes).
// Random 64-byte loads out of a peer's 2.5 GB table arena ran at 570 rows/us (profiles/, run 24:
// 19x below the NVLink bandwidth, independent of the loads in flight), so rows are not fetched
// remotely.  Instead:
//   shard_request_kernel : every (sample, id column) whose row lives on another rank appends
//                          (column, local row) to that owner's inbox (remote 8-byte stores; the slot
//                          comes from a LOCAL counter — only this rank writes its slice of the inbox)
//                          and remembers (owner, slot) in where[b, c]; local ids get where = -1.
//   [barrier]
//   shard_serve_kernel   : the owner gathers the requested rows from its LOCAL tables and streams them
//                          into the requester's response buffer (contiguous remote 128-bit stores).
//   [barrier]
//   gather (gather.cu)   : reads local rows from the tables and remote rows from its response buffer.
// ---------------------------------------------------------------------------------------------
struct RequestArgs {
    const float* X; int64_t ldx; int64_t B;
    int n_cols; const int32_t* cols; const int32_t* vocab;     // distinct id columns (the unique-plan columns)
    int G, me;
    int32_t* cnt_to;                 // [G] local counters: requests issued to each owner this step
    int32_t* const* inbox_req;       // [G] owner o's request list for THIS rank (peer pointer): int2 (col, local row)
    int32_t* where;                  // [B, n_cols]
    int64_t cap;
    int32_t* err_flag;
    int id_mode;
};

constexpr int REQ_CHUNK = 1024;
__global__ void __launch_bounds__(256) shard_request_kernel(RequestArgs a) {
    __shared__ int s_cnt[PUSH_MAX_G];
    __shared__ int s_base[PUSH_MAX_G];
    const int c = blockIdx.y;
    const int col = a.cols[c], vocab = a.vocab[c];
    for (int64_t b0 = (int64_t)blockIdx.x * REQ_CHUNK; b0 < a.B; b0 += (int64_t)gridDim.x * REQ_CHUNK) {
        if (threadIdx.x < PUSH_MAX_G) s_cnt[threadIdx.x] = 0;
        __syncthreads();
        int owner[REQ_CHUNK / 256], pos[REQ_CHUNK / 256];
        int32_t local[REQ_CHUNK / 256];
#pragma unroll
        for (int j = 0; j < REQ_CHUNK / 256; ++j) {
            const int64_t b = b0 + threadIdx.x + 256 * j;
            owner[j] = -1;
            if (b < a.B) {
                const float xv = __ldg(a.X + b * a.ldx + col);
                int id = a.id_mode ? __float_as_int(xv) : __float2int_rz(xv);
                if ((unsigned)id >= (unsigned)vocab) {
                    atomicOr(a.err_flag, 1);
                    id = 0;
                }
                const int o = id % a.G;
                local[j] = id / a.G;
                if (o == a.me) {
                    a.where[b * a.n_cols + c] = -1;
                } else {
                    owner[j] = o;
                    pos[j] = atomicAdd(&s_cnt[o], 1);
                }
            }
        }
        __syncthreads();
        if (threadIdx.x < a.G && s_cnt[threadIdx.x] > 0) s_base[threadIdx.x] = atomicAdd(a.cnt_to + threadIdx.x, s_cnt[threadIdx.x]);
        __syncthreads();
#pragma unroll
        for (int j = 0; j < REQ_CHUNK / 256; ++j) {
            if (owner[j] < 0) continue;
            const int64_t b = b0 + threadIdx.x + 256 * j;
            const int o = owner[j];
            const int64_t slot = (int64_t)s_base[o] + pos[j];
            if (slot >= a.cap) {
                atomicOr(a.err_flag, 2);
                a.where[b * a.n_cols + c] = -1;
                continue;
            }
            reinterpret_cast<int2*>(a.inbox_req[o])[slot] = make_int2(c, local[j]);     // remote 8-byte store
            a.where[b * a.n_cols + c] = (o << 26) | (int)slot;
        }
        __syncthreads();
    }
}

// publish the request counts to the owners' inboxes (after shard_request_kernel): inbox_cnt[o][me] = cnt_to[o]
__global__ void shard_publish_kernel(const int32_t* cnt_to, int32_t* const* inbox_cnt, int G, int me) {
    const int o = threadIdx.x;
    if (o < G && o != me) inbox_cnt[o][me] = cnt_to[o];
}

struct ServeArgs {
    int G, me, D;
    const int32_t* req_cnt;              // [G] requests received from each rank
    const int32_t* req;                  // [G][cap] int2 (col, local row)
    int64_t cap;
    const float* const* emb_of_col;      // [n_cols] LOCAL table of the column (or NULL)
    const float* const* lin_of_col;      // [n_cols]
    float* const* resp_emb;              // [G] requester q's response rows for THIS owner (peer pointer) [cap, D]
    float* const* resp_lin;              // [G] [cap]
};

// one group of D/4 lanes per request: 128-bit local row load -> 128-bit remote store (consecutive
// requests are consecutive rows of the response buffer: a warp writes 512 contiguous bytes at D = 16)
__global__ void __launch_bounds__(256) shard_serve_kernel(ServeArgs a) {
    const int q = blockIdx.y;
    if (q == a.me) return;
    const int n = a.req_cnt[q];
    const int lpr = a.D >> 2;                                  // lanes per row (D % 4 == 0, D/4 <= 32 checked on the host)
    const int64_t groups = ((int64_t)gridDim.x * blockDim.x) / lpr;
    const int64_t gid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / lpr;
    const int sub = threadIdx.x % lpr;
    const int2* req = reinterpret_cast<const int2*>(a.req) + (int64_t)q * a.cap;
    float* oemb = a.resp_emb[q];
    float* olin = a.resp_lin[q];
    for (int64_t k = gid; k < n; k += groups) {
        const int2 r = req[k];
        const float* tab = a.emb_of_col[r.x];
        if (tab) {
            const float4 v = ld_stream4(tab + (size_t)r.y * a.D + sub * 4);
            *reinterpret_cast<float4*>(oemb + k * a.D + sub * 4) = v;
        }
        if (sub == 0) {
            const float* lt = a.lin_of_col[r.x];
            if (lt) olin[k] = __ldg(lt + r.y);
        }
    }
}

// Device-side barrier over peer memory: rank r bumps its private epoch, stores it into slot [r] of every
// peer's flag array (after a system-scope fence, so everything this stream wrote to peer memory before is
// visible first) and spins until all of its own slots reached the epoch.  One block, one thread per peer.
// Replaces the NCCL all-reduce "barriers" of the row exchange (two per forward in round 1).  The spin is
// bounded: a missing peer sets bit 2 of err_flag instead of hanging the GPU.
__global__ void p2p_barrier_kernel(int32_t* const* peer_flags, int32_t* epoch_ctr, int G, int me, int32_t* err_flag) {
    __shared__ int32_t s_epoch;
    if (threadIdx.x == 0) {
        s_epoch = *epoch_ctr + 1;
        *epoch_ctr = s_epoch;
    }
    __syncthreads();
    const int32_t e = s_epoch;
    const int r = threadIdx.x;
    if (r < G) {
        __threadfence_system();
        volatile int32_t* remote = peer_flags[r] + me;
        *remote = e;
        __threadfence_system();
        volatile int32_t* mine = peer_flags[me] + r;
        long long spins = 0;
        while (*mine - e < 0) {
            if (++spins > (1ll << 26)) {      // tens of seconds: a peer died or never launched
                atomicOr(err_flag, 4);
                break;
            }
        }
        __threadfence_system();
    }
}

}  // namespace

extern "C" int ctr_p2p_barrier(int32_t* const* peer_flags, int32_t* epoch_ctr, int n_shards, int rank, int32_t* err_flag,
                               void* stream) {
    CTR_ARG(peer_flags && epoch_ctr && err_flag && n_shards >= 1 && n_shards <= PUSH_MAX_G && rank >= 0 && rank < n_shards,
            "ctr_p2p_barrier: bad arguments");
    p2p_barrier_kernel<<<1, 32, 0, as_stream(stream)>>>(peer_flags, epoch_ctr, n_shards, rank, err_flag);
    CTR_LAUNCH_OK("p2p_barrier_kernel");
    return 0;
}

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

extern "C" int ctr_rowgrad_push(int64_t B, int n_shards, int n_plan_cols, const int32_t* n_uniq, const int32_t* uniq,
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
    CTR_ARG(n_plan_cols > 0 && n_plan_cols <= 65535, "ctr_rowgrad_push: bad number of plan columns");
    if (B == 0 || n_emb + n_lin == 0) return 0;
    PushArgs a{B, n_shards, n_uniq, uniq, n_emb, D, emb_rowgrad, emb_rowgrad_stride, emb_plan_col, n_lin,
               lin_rowgrad, lin_rowgrad_stride, lin_plan_col, recv_count, recv_ids, recv_emb_rows, recv_lin_rows, cap, err_flag};
    CTR_ARG(n_shards <= PUSH_MAX_G, "ctr_rowgrad_push: at most %d shards", PUSH_MAX_G);
    CTR_ARG(cap < (1 << 26), "ctr_rowgrad_push: receive-list capacity must be below 2^26 rows");
    int64_t bx = ceil_div64(B, PUSH_CHUNK);
    const int64_t limit = ceil_div64((int64_t)ctr_sm_count() * 8, n_plan_cols);
    if (bx > limit) bx = limit;
    if (bx < 1) bx = 1;
    dim3 grid((unsigned)bx, (unsigned)n_plan_cols);
    rowgrad_push_kernel<<<grid, 256, 0, as_stream(stream)>>>(a);
    CTR_LAUNCH_OK("rowgrad_push_kernel");
    return 0;
}

extern "C" int ctr_shard_request(const float* X, int64_t ldx, int64_t B, int n_cols, const int32_t* cols,
                                 const int32_t* vocab, int n_shards, int rank, int32_t* cnt_to,
                                 int32_t* const* inbox_req, int32_t* const* inbox_cnt, int32_t* where,
                                 int64_t cap, int32_t* err_flag, int id_mode, void* stream) {
    CTR_ARG(X && cols && vocab && cnt_to && inbox_req && inbox_cnt && where && err_flag, "ctr_shard_request: null argument");
    CTR_ARG(n_cols > 0 && B >= 0 && n_shards >= 1 && n_shards <= PUSH_MAX_G && rank >= 0 && rank < n_shards && cap > 0 &&
                cap < (1 << 26),
            "ctr_shard_request: bad sizes");
    cudaStream_t st = as_stream(stream);
    CTR_CUDA(cudaMemsetAsync(cnt_to, 0, sizeof(int32_t) * n_shards, st));
    if (B > 0) {
        RequestArgs a{X, ldx, B, n_cols, cols, vocab, n_shards, rank, cnt_to, inbox_req, where, cap, err_flag, id_mode};
        int64_t bx = ceil_div64(B, REQ_CHUNK);
        const int64_t limit = ceil_div64((int64_t)ctr_sm_count() * 8, n_cols);
        if (bx > limit) bx = limit;
        if (bx < 1) bx = 1;
        dim3 grid((unsigned)bx, (unsigned)n_cols);
        shard_request_kernel<<<grid, 256, 0, st>>>(a);
        CTR_LAUNCH_OK("shard_request_kernel");
    }
    shard_publish_kernel<<<1, 32, 0, st>>>(cnt_to, inbox_cnt, n_shards, rank);
    CTR_LAUNCH_OK("shard_publish_kernel");
    return 0;
}

extern "C" int ctr_shard_serve(int n_shards, int rank, int D, const int32_t* req_cnt, const int32_t* req, int64_t cap,
                               const float* const* emb_of_col, const float* const* lin_of_col,
                               float* const* resp_emb, float* const* resp_lin, void* stream) {
    CTR_ARG(req_cnt && req && emb_of_col && lin_of_col && resp_emb && resp_lin, "ctr_shard_serve: null argument");
    CTR_ARG(n_shards >= 1 && n_shards <= PUSH_MAX_G && rank >= 0 && rank < n_shards && cap > 0, "ctr_shard_serve: bad sizes");
    CTR_ARG(D > 0 && D % 4 == 0 && D / 4 <= 32 && 256 % (D / 4) == 0, "ctr_shard_serve: embedding dim %d unsupported", D);
    if (n_shards == 1) return 0;
    ServeArgs a{n_shards, rank, D, req_cnt, req, cap, emb_of_col, lin_of_col, resp_emb, resp_lin};
    int64_t bx = (int64_t)ctr_sm_count() * 8 / n_shards;
    if (bx < 1) bx = 1;
    dim3 grid((unsigned)bx, (unsigned)n_shards);
    shard_serve_kernel<<<grid, 256, 0, as_stream(stream)>>>(a);
    CTR_LAUNCH_OK("shard_serve_kernel");
    return 0;
}
