// Hardware probe for the tcgen05 features the tower GEMM wants to use next (run on the B200 box):
//   mode 0  SS, K-major A and B (the layout gemm_pk.cu uses today)                      -> reference
//   mode 1  TS: A written into TMEM with tcgen05.st (lane = row, column = k), B from smem
//   mode 2  SS, B MN-major (b_major bit of the instruction descriptor)  -> no transposing converter for wgrad
//   mode 3  SS, A MN-major
//   mode 4  how kind::tf32 narrows fp32 inputs that are NOT tf32-exact (truncate vs round)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o scripts/probe_umma scripts/probe_umma.cu
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../deepctr_torch_b200/csrc/tc_common.cuh"

namespace {

__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
        "r"(tmem_a), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]),
                 "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// A [128 x 8], B [128 x 8] fp32 row-major in global; D [128 x 128]
__global__ void __launch_bounds__(128, 1) probe_kernel(const float* A, const float* B, float* D, int mode) {
    __shared__ __align__(1024) float sA[128 * 8];
    __shared__ __align__(1024) float sB[128 * 8];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
    // K-major canonical, no swizzle: 16-byte chunk c (k = 4c..4c+3) of row r at (c*128 + r)*16 bytes
    // MN-major canonical, no swizzle: element (r, k) at ((r/4)*8 + k)*16 bytes + (r%4)*4
    for (int i = tid; i < 128 * 8; i += 128) {
        const int r = i >> 3, k = i & 7;
        const int kmaj = ((k >> 2) * 128 + r) * 4 + (k & 3);
        const int mnmaj = ((r >> 2) * 8 + k) * 4 + (r & 3);
        sA[mode == 3 ? mnmaj : kmaj] = A[i];
        sB[mode == 2 ? mnmaj : kmaj] = B[i];
    }
    if (tid == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (wid == 0) tmem_alloc_warp(&tmem_slot, 256);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = tmem_slot;
    const uint32_t a_col = 128;   // TMEM columns 128.. hold the A operand in TS mode
    if (mode == 1) {
        uint32_t v[8];
        for (int k = 0; k < 8; ++k) v[k] = __float_as_uint(A[tid * 8 + k]);
        tmem_st8(tbase + ((uint32_t)(wid * 32) << 16) + a_col, v);
        tmem_st_wait();
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
    }
    if (tid == 0) {
        uint32_t idesc = tf32_idesc(128);
        if (mode == 2) idesc |= 1u << 16;
        if (mode == 3) idesc |= 1u << 15;
        const uint64_t da = (mode == 3) ? make_smem_desc(smem_u32(sA), 4096, 128) : make_smem_desc(smem_u32(sA), 2048, 128);
        const uint64_t db = (mode == 2) ? make_smem_desc(smem_u32(sB), 4096, 128) : make_smem_desc(smem_u32(sB), 2048, 128);
        if (mode == 1) umma_tf32_ts(tbase, tbase + a_col, db, idesc, 0u);
        else umma_tf32(tbase, da, db, idesc, 0u);
        umma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
              "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
              "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
              "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
            : "r"(tbase + ((uint32_t)(wid * 32) << 16) + (uint32_t)(c * 32)));
        tmem_ld_wait();
        for (int j = 0; j < 32; ++j) D[(size_t)tid * 128 + c * 32 + j] = __uint_as_float(v[j]);
    }
    tc_fence_before();
    __syncthreads();
    if (wid == 0) {
        tc_fence_after();
        tmem_dealloc_warp(tbase, 256);
    }
}

// Timing: one thread issues `n_mma` back-to-back kind::tf32 MMAs (M = 128, N columns) and commits; cycles from the
// first issue to the completion mbarrier.  variant bit0: A from TMEM (TS) instead of shared memory; bit1: alternate between two
// accumulators; bit2: B (and A) described as SWIZZLE_128B K-major (timing only: the data is not laid out for it).
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t addr) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;                       // LBO (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;             // SBO: 8 rows x 128 B
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;                       // SWIZZLE_128B
    return d;
}
__global__ void __launch_bounds__(128, 1) time_kernel(int N, int n_mma, int variant, long long* out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, wid = tid >> 5;
    for (int i = tid; i < (128 + 256) * 32; i += 128) reinterpret_cast<float*>(smem)[i] = 0.001f * (i & 63);
    if (tid == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (wid == 0) tmem_alloc_warp(&tmem_slot, 512);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = tmem_slot;
    if (tid == 0) {
        const uint32_t idesc = tf32_idesc(N);
        const uint32_t sa = smem_u32(smem), sb = smem_u32(smem + 128 * 128);
        const bool ts = variant & 1, alt = variant & 2, sw = variant & 4;
        const long long t0 = clock64();
        for (int i = 0; i < n_mma; ++i) {
            const uint32_t k = (uint32_t)(i & 3);
            const uint64_t da = sw ? make_desc_sw128(sa + k * 32u) : make_smem_desc(sa + k * 4096u, 2048, 128);
            const uint64_t db = sw ? make_desc_sw128(sb + k * 32u) : make_smem_desc(sb + k * 2u * (uint32_t)N * 16u, (uint32_t)N * 16u, 128);
            const uint32_t d = tbase + ((alt && (i & 1)) ? 256u : 0u);
            if (ts) umma_tf32_ts(d, tbase + 480u + 8u * (k & 1), db, idesc, 1u);
            else umma_tf32(d, da, db, idesc, 1u);
        }
        const long long t1 = clock64();
        umma_commit(&bar);
        mbar_wait(&bar, 0);
        const long long t2 = clock64();
        out[0] = t1 - t0;
        out[1] = t2 - t0;
    }
    __syncthreads();
    if (wid == 0) {
        tc_fence_after();
        tmem_dealloc_warp(tbase, 512);
    }
}

// Interference: the TS N=256 MMA stream of time_kernel while the other warps of the CTA (a) spin on an mbarrier that never
// completes, (b) stream LDS.128 over 32 KB, (c) stream STS.128, (d) one thread keeps 32 KB cp.async.bulk copies in flight.
__global__ void __launch_bounds__(576, 1) interfere_kernel(int n_mma, int what, const float* gsrc, long long* out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar, never, tbar[4];
    __shared__ uint32_t tmem_slot;
    __shared__ volatile int stop;
    const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
    for (int i = tid; i < 48 * 1024 / 4; i += 576) reinterpret_cast<float*>(smem)[i] = 0.001f * (i & 63);
    if (tid == 0) {
        mbar_init(&bar, 1);
        mbar_init(&never, 1);
        for (int i = 0; i < 4; ++i) mbar_init(&tbar[i], 1);
        stop = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (wid == 0) tmem_alloc_warp(&tmem_slot, 512);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = tmem_slot;
    if (wid == 0) {
        if (lane == 0) {
            const uint32_t idesc = tf32_idesc(256);
            const uint32_t sb = smem_u32(smem);
            const long long t0 = clock64();
            for (int i = 0; i < n_mma; ++i) {
                const uint32_t k = (uint32_t)(i & 3);
                const uint64_t db = make_smem_desc(sb + k * 2u * 4096u, 4096u, 128);
                umma_tf32_ts(tbase, tbase + 480u + 8u * (k & 1), db, idesc, 1u);
            }
            umma_commit(&bar);
            mbar_wait(&bar, 0);
            out[what] = clock64() - t0;
            stop = 1;
        }
    } else if (wid == 1 && what == 4) {
        if (lane == 0) {      // keep 4 x 32 KB bulk copies global -> smem (region 48..176 KB) in flight
            uint32_t ph[4] = {0, 0, 0, 0};
            for (int q = 0; q < 4; ++q) {
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&tbar[q])), "r"(32768u) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                 smem_u32(smem + 49152 + q * 32768)), "l"(gsrc + q * 8192), "r"(32768u), "r"(smem_u32(&tbar[q])) : "memory");
            }
            int q = 0;
            while (!stop) {
                mbar_wait(&tbar[q], ph[q]);
                ph[q] ^= 1u;
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&tbar[q])), "r"(32768u) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                 smem_u32(smem + 49152 + q * 32768)), "l"(gsrc + q * 8192), "r"(32768u), "r"(smem_u32(&tbar[q])) : "memory");
                q = (q + 1) & 3;
            }
            for (int k = 0; k < 4; ++k) mbar_wait(&tbar[k], ph[k]);
        }
    } else if (wid >= 2) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        float4* reg = reinterpret_cast<float4*>(smem + 49152 + (wid - 2) * 8192);     // 8 KB per warp
        if (what == 1) {
            while (!stop) (void)mbar_try_wait(&never, 0);
        } else if (what == 2) {
            while (!stop)
                for (int j = 0; j < 16; ++j) {
                    const float4 v = reg[j * 32 + lane];
                    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
                }
        } else if (what == 3) {
            while (!stop)
                for (int j = 0; j < 16; ++j) reg[j * 32 + lane] = acc;
        }
        if (acc.x == 123.456f) out[7] = 1;
    }
    __syncthreads();
    if (wid == 0) {
        tc_fence_after();
        tmem_dealloc_warp(tbase, 512);
    }
}

float rn_tf32_h(float x) {
    uint32_t u;
    memcpy(&u, &x, 4);
    u = (u + 0x1000u) & 0xFFFFE000u;
    memcpy(&x, &u, 4);
    return x;
}
float tr_tf32_h(float x) {
    uint32_t u;
    memcpy(&u, &x, 4);
    u &= 0xFFFFE000u;
    memcpy(&x, &u, 4);
    return x;
}

}  // namespace

#define CK(x)                                                                     \
    do {                                                                          \
        cudaError_t e = (x);                                                      \
        if (e != cudaSuccess) {                                                   \
            printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); \
            return 2;                                                             \
        }                                                                         \
    } while (0)

int main() {
    const int M = 128, N = 128, K = 8;
    float hA[M * K], hB[N * K], *dA, *dB, *dD;
    static float hD[5][M * N];
    CK(cudaMalloc(&dA, sizeof(hA)));
    CK(cudaMalloc(&dB, sizeof(hB)));
    CK(cudaMalloc(&dD, sizeof(float) * M * N));
    srand(1);
    for (int i = 0; i < M * K; ++i) hA[i] = rn_tf32_h((float)rand() / RAND_MAX - 0.5f);
    for (int i = 0; i < N * K; ++i) hB[i] = rn_tf32_h((float)rand() / RAND_MAX - 0.5f);
    CK(cudaMemcpy(dA, hA, sizeof(hA), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, hB, sizeof(hB), cudaMemcpyHostToDevice));
    for (int mode = 0; mode < 4; ++mode) {
        CK(cudaMemset(dD, 0xFF, sizeof(float) * M * N));
        probe_kernel<<<1, 128>>>(dA, dB, dD, mode);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(hD[mode], dD, sizeof(float) * M * N, cudaMemcpyDeviceToHost));
        double worst = 0, worst0 = 0;
        for (int m = 0; m < M; ++m)
            for (int n = 0; n < N; ++n) {
                double ref = 0;
                for (int k = 0; k < K; ++k) ref += (double)hA[m * K + k] * hB[n * K + k];
                worst = fmax(worst, fabs(ref - hD[mode][m * N + n]));
                worst0 = fmax(worst0, fabs((double)hD[0][m * N + n] - hD[mode][m * N + n]));
            }
        printf("mode %d: max |D - fp64 ref| = %.3e   max |D - D(mode 0)| = %.3e   %s\n", mode, worst, worst0,
               worst < 1e-6 ? "OK" : "MISMATCH");
    }
    // mode 4: inputs with all 23 mantissa bits populated, one non-zero k per row -> D = a*b for one product
    for (int i = 0; i < M * K; ++i) hA[i] = 0.f;
    for (int i = 0; i < N * K; ++i) hB[i] = 0.f;
    for (int m = 0; m < M; ++m) hA[m * K] = 1.0f + (float)(rand() & 0x7FFFFF) / 8388608.0f;
    for (int n = 0; n < N; ++n) hB[n * K] = 1.0f + (float)(rand() & 0x7FFFFF) / 8388608.0f;
    CK(cudaMemcpy(dA, hA, sizeof(hA), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, hB, sizeof(hB), cudaMemcpyHostToDevice));
    probe_kernel<<<1, 128>>>(dA, dB, dD, 0);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(hD[4], dD, sizeof(float) * M * N, cudaMemcpyDeviceToHost));
    int n_tr = 0, n_rn = 0, n_other = 0;
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
            const float got = hD[4][m * N + n];
            const float t = tr_tf32_h(hA[m * K]) * tr_tf32_h(hB[n * K]);
            const float r = rn_tf32_h(hA[m * K]) * rn_tf32_h(hB[n * K]);
            if (got == t) ++n_tr;
            else if (got == r) ++n_rn;
            else ++n_other;
        }
    printf("mode 4 (fp32 inputs that are not tf32-exact): matches truncation %d, round-to-nearest %d, neither %d of %d\n",
           n_tr, n_rn, n_other, M * N);
    long long* dt;
    CK(cudaMalloc(&dt, 64));
    CK(cudaFuncSetAttribute(time_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    const char* names[8] = {"SS", "TS", "SS 2 accumulators", "TS 2 accumulators", "SS sw128", "TS sw128", "SS sw128 2 acc", "TS sw128 2 acc"};
    for (int Nn = 128; Nn <= 256; Nn += 128)
        for (int v = 0; v < 8; ++v) {
            long long h[2];
            for (int rep = 0; rep < 2; ++rep) {
                time_kernel<<<1, 128, 196 * 1024>>>(Nn, 96, v, dt);
                CK(cudaDeviceSynchronize());
            }
            CK(cudaMemcpy(h, dt, 16, cudaMemcpyDeviceToHost));
            printf("timing N=%d %-22s: issue %5.1f clk/MMA, issue->complete %6.1f clk/MMA (ideal %d)\n", Nn, names[v], h[0] / 96.0,
                   h[1] / 96.0, Nn / 2);
        }
    {
        float* gsrc;
        CK(cudaMalloc(&gsrc, 4 * 32768));
        CK(cudaMemset(gsrc, 0, 4 * 32768));
        CK(cudaFuncSetAttribute(interfere_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        const char* wn[5] = {"alone", "16 warps spinning on an mbarrier", "16 warps LDS.128 streaming", "16 warps STS.128 streaming",
                             "4 x 32 KB bulk copies in flight"};
        for (int w = 0; w < 5; ++w) {
            for (int rep = 0; rep < 2; ++rep) {
                interfere_kernel<<<1, 576, 196 * 1024>>>(192, w, gsrc, dt);
                CK(cudaDeviceSynchronize());
            }
            long long h[8];
            CK(cudaMemcpy(h, dt, 16, cudaMemcpyDeviceToHost));
            long long v;
            CK(cudaMemcpy(&v, dt + w, 8, cudaMemcpyDeviceToHost));
            printf("interference TS N=256, %-34s: %6.1f clk/MMA\n", wn[w], v / 192.0);
        }
    }
    return 0;
}
