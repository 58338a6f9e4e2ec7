// Internal interface of the fp32 GEMM engine (gemm.cu) shared by the layer kernels.
#pragma once
#include "common.cuh"

constexpr int CTR_ACT_MULPRO = 100;    // prologue "activation" code: operand *= mask (no derivative involved)

enum GemmEpilogue {
    EPI_STORE = 0,       // C = acc            (or C += acc when accumulate)
    EPI_BIAS_ACT = 1,    // C = act(acc + bias[n])
    EPI_MUL_ACTGRAD = 2, // C (+)= acc * act'(aux[m,n])     (dgrad feeding a previous activation)
    EPI_CROSS = 3,       // U = acc + bias[n]; C = aux[m,n] * U + aux2[m,n]   (CrossNet matrix)
    EPI_MUL = 4,         // C (+)= acc * aux[m,n]               (bilinear interaction: (v_i W^T) (.) v_j)
};

struct GemmArgs {
    int64_t M, N, K;
    const float* A; int64_t sam, sak;      // A(m,k)  = A[m*sam + k*sak]
    const float* B; int64_t sbn, sbk;      // B(n,k)  = B[n*sbn + k*sbk]
    float* C; int64_t ldc;
    int accumulate;                        // C += result
    int epilogue;
    int act;
    const float* bias;                     // [N]
    const float* aux; int64_t ldaux;       // epilogue operand [M,N]
    const float* aux2; int64_t ldaux2;     // second epilogue operand [M,N]
    float* out2; int64_t ldout2;           // secondary output (U of EPI_CROSS)
    // optional A prologue: A(m,k) *= act'(amask[m*smm + k*smk])
    const float* amask; int64_t smm, smk; int amask_act;
    // optional B prologue: B(n,k) *= act'(bmask[n*sbmn + k*sbmk])
    const float* bmask; int64_t sbmn, sbmk; int bmask_act;
    int allow_split_k;                     // K may be split across CTAs (atomic accumulation)
};

inline GemmArgs gemm_args_default() {
    GemmArgs g{};
    g.epilogue = EPI_STORE;
    return g;
}

// returns 0 or an error code (ctr_set_error already called); dispatches between the tcgen05
// 3xTF32 kernel (gemm_tc.cu) and the FP32 FFMA tiles (gemm.cu)
int launch_sgemm(const GemmArgs& g, cudaStream_t st);
int launch_gemm_tc(const GemmArgs& g, cudaStream_t st);
int ctr_gemm_passes();                           // 3 = 3xTF32 (parity), 1 = single-pass TF32 (fast, non-parity)
unsigned long long* ctr_debug_buffer();          // buffer registered with ctr_debug_set_buffer (or NULL)
// packed-operand engine (gemm_pk.cu): needs the scratch registered with ctr_set_scratch
int launch_gemm_pk(const GemmArgs& g, cudaStream_t st);
// weight-gradient engine (gemm_pk.cu): C[M,N] = A^T B over a long K (the batch); db (may be NULL) = column sums of the
// masked A operand.  Returns -3 when the shape does not qualify (caller falls back to launch_sgemm + launch_colsum).
int launch_gemm_tsw(const GemmArgs& g, float* db, cudaStream_t st);
bool gemm_tsw_eligible(const GemmArgs& g);
int64_t gemm_pk_scratch_bytes(int64_t M, int64_t N, int64_t K);
bool gemm_pk_has_scratch(int64_t M, int64_t N, int64_t K);
void* gemm_scratch_ptr(int64_t need_bytes);      // registered scratch if it holds need_bytes, else NULL
int gemm_pack_operand(const float* P, int64_t s_row, int64_t s_k, int64_t n_rows, int64_t K, int R, int64_t nkb,
                      float* out, cudaStream_t st);

// out[n] = sum_m A[m*sam + n*san] * (mask ? act'(mask[m*smm + n*smn]) : 1) * (w ? w[m] : 1)
int launch_colsum(const float* A, int64_t sam, int64_t san, const float* mask, int64_t smm,
                  int64_t smn, int mask_act, const float* w, int64_t M, int64_t N, float* out,
                  cudaStream_t st);
