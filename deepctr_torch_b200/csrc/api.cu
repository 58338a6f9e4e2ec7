// Library-level entry points: version, thread-local error string, device queries.
#include <stdarg.h>

#include "common.cuh"

static thread_local char g_err[512] = "";

void ctr_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int ctr_sm_count() {
    static int cached = 0;
    if (cached == 0) {
        int dev = 0, n = 0;
        if (cudaGetDevice(&dev) == cudaSuccess &&
            cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
            cached = n;
        else
            cached = 148;
    }
    return cached;
}

#include <atomic>
static std::atomic<long long> g_launches{0};
void ctr_count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
extern "C" int64_t ctr_launch_count(void) { return (int64_t)g_launches.load(); }

extern "C" int ctr_version(void) { return 2; }
extern "C" const char* ctr_last_error(void) { return g_err; }

// ---------------------------------------------------------------------------------------------
// a16: streaming sum of squares (whole-table L2 term), out[0] += scale * sum(w^2)
// ---------------------------------------------------------------------------------------------
__global__ void sumsq_kernel(const float* __restrict__ w, int64_t n, float scale, float* out) {
    float acc = 0.f;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t n4 = ((reinterpret_cast<uintptr_t>(w) & 15) == 0) ? n / 4 : 0;
    const float4* w4 = reinterpret_cast<const float4*>(w);
    for (int64_t j = i; j < n4; j += stride) {
        float4 v = w4[j];
        acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
    for (int64_t j = n4 * 4 + i; j < n; j += stride) acc += w[j] * w[j];
    acc = warp_sum(acc);
    __shared__ float part[32];
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) part[wid] = acc;
    __syncthreads();
    if (wid == 0) {
        acc = lane < (blockDim.x >> 5) ? part[lane] : 0.f;
        acc = warp_sum(acc);
        if (lane == 0) atomicAdd(out, acc * scale);
    }
}

extern "C" int ctr_sumsq_acc(const float* w, int64_t n, float scale, float* out, void* stream) {
    CTR_ARG(w && out && n >= 0, "ctr_sumsq_acc: bad arguments");
    if (n == 0) return 0;
    int64_t blocks = ceil_div64(n, 256 * 16);
    int64_t cap = (int64_t)ctr_sm_count() * 8;
    if (blocks > cap) blocks = cap;
    sumsq_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(w, n, scale, out);
    CTR_LAUNCH_OK("sumsq_kernel");
    return 0;
}
