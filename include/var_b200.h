/* var_b200 — C ABI of the B200-native VAR next-scale-prediction hot path.
 *
 * The reference (culiver/VAR, pure Python/PyTorch) has no FFI layer: the hot path sits behind nn.Module
 * methods. This header is the boundary a maintainer binds instead of the ATen calls those
 * methods make; every entry cites the reference lines it replaces. Conventions:
 *   - plain pointers and sizes, all device pointers unless noted "host";
 *   - the last argument is a cudaStream_t passed as void* (NULL = legacy default stream);
 *   - return 0 on success, a negative VAR_B200_ERR_* code otherwise; var_b200_last_error() gives the message;
 *   - no allocation inside: workspaces are caller-provided (sizes from the *_workspace_bytes helpers);
 *   - bf16 tensors are raw 16-bit storage (torch.bfloat16), fp32 otherwise; indices are int64 where the
 *     reference returns LongTensors, int32 where noted.
 */
#ifndef VAR_B200_H
#define VAR_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(VAR_B200_BUILD)
#define VAR_B200_API __attribute__((visibility("default")))
#else
#define VAR_B200_API
#endif

#define VAR_B200_OK 0
#define VAR_B200_ERR_ARG (-1)
#define VAR_B200_ERR_CUDA (-2)
#define VAR_B200_ERR_DRIVER (-3)
#define VAR_B200_ERR_WORKSPACE (-4)

/* Message describing the most recent error on the calling thread (never NULL). */
VAR_B200_API const char* var_b200_last_error(void);

/* ------------------------------------------------------------------------------------------------
 * GEMM family:  D[M,N] = A[M,K] * W[N,K]^T, bf16 operands, fp32 accumulation on tcgen05/TMEM.
 * Replaces F.linear / nn.Linear calls of the transformer: models/basic_var.py:93 (QKV), :119 (proj),
 * :52 (fc1/fc2), models/var.py:124 (head), basic_var.py:156,173 (adaLN linears).
 * ---------------------------------------------------------------------------------------------- */
enum {
  VAR_B200_EPI_BIAS_F32 = 0,   /* out fp32 = acc + bias */
  VAR_B200_EPI_BIAS_BF16 = 1,  /* out bf16 = acc + bias */
  VAR_B200_EPI_GELU_BF16 = 2,  /* out bf16 = gelu_tanh(acc + bias)              basic_var.py:40,52 */
  VAR_B200_EPI_GATE_RESID = 3, /* out fp32 = resid + gate[m / rows_per_seq] * (acc + bias)   basic_var.py:157-158 */
  VAR_B200_EPI_QKV = 4,        /* +[q_bias,0,v_bias]; q,k L2-normalised per head; q *= scale; scatter into q buffer
                                  and the preallocated K/V cache                  basic_var.py:93-109 */
  VAR_B200_EPI_SCORE = 5       /* per-row partial log-sum-exp and ground-truth logit eval_prob.py:446-452 */
};

typedef struct var_b200_gemm_args {
  const void* A; /* [M,K] bf16 row-major */
  const void* W; /* [N,K] bf16 row-major (nn.Linear weight) */
  int M, N, K;
  int epilogue;      /* VAR_B200_EPI_* */
  int force_bn;      /* 0 = auto tile width, else 128 / 192 / 256 */
  const float* bias; /* [N] or NULL (required for QKV / SCORE) */
  void* out;         /* [M,N] fp32 or bf16 */
  /* GATE_RESID */
  const float* resid; /* [M,N] fp32, may alias out */
  const float* gate;  /* gate[(m / rows_per_seq) * gate_ld + n] */
  int rows_per_seq;
  int gate_ld;
  /* QKV (N == 3*C, head_dim 64) */
  void* q_out;          /* bf16 [n_seq, H, rows_per_seq, 64] */
  void* k_cache;        /* bf16 [n_seq, H, Lmax, 64], rows written at pos0 + t */
  void* v_cache;        /* bf16 [n_seq, H, Lmax, 64] */
  const float* q_scale; /* [H] exp(min(scale_mul, ln 100)) */
  int C, H, pos0, Lmax;
  /* SCORE */
  const int32_t* gt; /* [M] */
  float* part;       /* [M, ceil(N / var_b200_gemm_tile_n(N)), 2] (max, sumexp) */
  float* gt_logit;   /* [M] */
} var_b200_gemm_args_t;

VAR_B200_API int var_b200_gemm_bf16(const var_b200_gemm_args_t* args, void* stream);
/* Tile width (128/192/256) the GEMM uses for a given N. */
VAR_B200_API int var_b200_gemm_tile_n(int N);

/* Hardware probe used by the test-suite to pin UMMA shared-memory descriptor encodings:
 * D[128,N] = A[128,64] * B, B = [N,64] (K-major) or [64,N] (MN-major). */
VAR_B200_API int var_b200_umma_probe(const void* A, const void* B, float* D, int N, int b_mn_major, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VAR_B200_H */
