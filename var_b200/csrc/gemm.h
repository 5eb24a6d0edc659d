// GEMM family interface shared by the host-side drivers (no device code here).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>

namespace vb {

// Epilogue warps per TMEM lane quarter: the 32-column chunks of a tile are dealt round-robin to them (EPI_SCORE also
// writes one partial per row, tile and such warp). 2 -> eight epilogue warps, 320 threads. Three (twelve warps) was
// measured in a same-box A/B inside both bench steps: d30 sampling 274.5 vs 276.6 img/s, d16 scoring 3.03 vs 3.06 (the
// 64-column QKV chunks deal unevenly over three warps), although the GELU epilogue alone gains 2-3 %.
constexpr int GEMM_EPI_SUB = 2;

enum EpiMode : int {
  EPI_BIAS_F32 = 0,    // out fp32 = acc + bias                                  (head: models/var.py:124)
  EPI_BIAS_BF16 = 1,   // out bf16 = acc + bias
  EPI_GELU_BF16 = 2,   // out bf16 = gelu_tanh(acc + bias)                        (fc1: models/basic_var.py:52)
  EPI_GATE_RESID = 3,  // out fp32 = resid + gate[seq] * (acc + bias)             (proj/fc2: basic_var.py:157-158)
  EPI_QKV = 4,         // q,k L2-norm per head, q*scale, scatter to q / K-cache / V-cache (basic_var.py:93-109)
  EPI_SCORE = 5,       // per-row partial log-sum-exp + logit at the ground-truth token (eval_prob.py:446-452)
};

struct GemmParams {
  int M, N, K;
  const float* bias;  // [N] (may be null)
  void* out;          // [M,N] fp32 or bf16, row stride N
  // EPI_GATE_RESID
  const float* resid;  // [M,N] fp32 (may alias out)
  const float* gate;   // gate[(m / rows_per_seq) * gate_ld + n]
  int rows_per_seq;
  int gate_ld;
  // EPI_QKV  (N == 3*C): rows are (seq, t) with t in [0, rows_per_seq)
  __nv_bfloat16* q_out;    // [n_seq, H, rows_per_seq, 64]
  __nv_bfloat16* k_cache;  // [n_seq, H, Lmax, 64]   written at position pos0 + t
  __nv_bfloat16* v_cache;  // [n_seq, H, Lmax, 64]
  const float* q_scale;    // [H] = exp(min(scale_mul, ln 100))
  int C, H, pos0, Lmax;
  int no_l2norm;           // attn_l2_norm=False (basic_var.py:72): q *= q_scale[head] only, k untouched
  // EPI_SCORE
  const int* gt;    // ground-truth token of row m = gt[m % gt_mod]
  int gt_mod;
  float2* part;     // [M, GEMM_EPI_SUB*n_tiles] (max, sum exp(x - max)) per row, tile and epilogue sub-warp
  float* gt_logit;  // [M]
  // EPI_BIAS_BF16: optional bf16 [M,N] added to the rounded result (ResnetBlock shortcut, models/basic_vae.py:60)
  const __nv_bfloat16* resid_bf16;
  // Implicit-GEMM 3x3 convolution, stride 1, zero padding 1 (models/basic_vae.py:45-46,52-59): A is the NHWC bf16
  // activation tensor [B, conv_H, conv_W, Cin] seen through a 4-D tensor map, row m = pixel ((b*H + y)*W + x);
  // K = 9 taps x conv_kpt 64-channel blocks (weights packed [N, 9, conv_kpt*64], zero padded). 0 = plain GEMM.
  int conv_kpt, conv_H, conv_W;
  // conv_stride 2 = the reference's Downsample2x (basic_vae.py:31-37): zero pad (0,1,0,1), 3x3, stride 2, no padding.
  // conv_H / conv_W are then the OUTPUT extent (rows m = output pixels), the input is 2*conv_H x 2*conv_W and tap
  // (dy,dx) of output (y,x) reads input (2y+dy, 2x+dx): the same TMA box with a traversal stride of 2 in W and H.
  int conv_stride;
  // Deferred LayerNorm (gemm_sm100.cuh, LNF). Producer side (EPI_GATE_RESID, set ln_a_out):
  __nv_bfloat16* ln_a_out;  // [M,N] bf16 = x_new * (1 + ln_scale[seq]): the A operand of the next QKV / fc1 GEMM
  const float* ln_scale;    // the NEXT LayerNorm's adaLN scale, ln_scale[(m / rows_per_seq) * gate_ld + n]
  float2* ln_part_out;      // [M, gemm_ln_parts(M,N)] partial (sum, sum of squares) of x_new per row, tile and sub-warp
  // Consumer side (EPI_QKV / EPI_GELU_BF16, set ln_part_in): out = rstd*(acc - mean*U[label]) + V[label] replaces acc + bias
  const float2* ln_part_in;
  int ln_parts, ln_C;       // partials per row, row length the statistics are taken over
  float ln_eps;
  const float* ln_u;        // [n_classes, N]: W (1 + scale_class)
  const float* ln_v;        // [n_classes, N]: W shift_class + bias
  const int* ln_labels;     // [M / rows_per_seq] class of every sequence
  int conv_cin;  // true channel count of the activation tensor (the tensor map's innermost extent; tails read as zeros)
  // Operand leading dimensions in elements (0 = K): a column slice of a wider row-major matrix can be an operand.
  int lda, ldw;
  // L2 eviction priority of the two operand streams (TMA cache hints; 0 = normal; VAR_B200_L2HINT for experiments).
  // Measured (profiles/r02_gemm_dram_probe.txt): evict-last weights save 0.2-0.7 GB of DRAM reads per d30 fc1 / fc2
  // launch but no step time; evict-first activations cost 2-5x the reads (the N tiles of an M block share them through
  // L2); evict-first epilogue traffic and a rotated K order per tile do not help either. Default: no hints.
  unsigned long long l2_hint_a, l2_hint_w;
  // EPI_BIAS_BF16, conv mode: per-channel (sum, sum of squares) of the stored bf16 outputs over every 128-row block
  // (= 128 pixels of one image): gn_part[(m / 128) * N + n] as float2, fixed summation order. The GroupNorm that
  // follows the convolution (models/basic_vae.py:57-58) takes its statistics from these instead of re-reading the tensor.
  float2* gn_part;
  int a_cols;  // true column count of A when it is narrower than K (K padded to 64; the tail reads as zeros), 0 = K
  // Block-diagonal batching (VQVAE AttnBlock, models/basic_vae.py:74-87: one bmm per image). Rows [j*bd_rows,
  // (j+1)*bd_rows) of A use the W rows shifted by j*bd_w_row and the W columns shifted by j*bd_w_k; bd_rows must be a
  // multiple of the M tile. w_rows / w_cols: full extent of the W matrix behind the tensor map (0 = N / K).
  int bd_rows, bd_w_row, bd_w_k, w_rows, w_cols;
};

// Tile width the launcher will use for a given N (needed to size EPI_SCORE partials: n_tiles = ceil(N / bn)).
int gemm_pick_bn(int N);
// Partials per row a deferred-LayerNorm producer launch (EPI_GATE_RESID with ln_a_out) of this shape writes.
int gemm_ln_parts(int M, int N);
// A: [M,K] bf16 row-major, W: [N,K] bf16 row-major (nn.Linear layout). force_bn: 0 = auto, else 128/192/256,
// optionally | 0x10000 to force the 1-CTA kernel (the default is the CTA-pair kernel whenever M > 128).
int gemm_launch(const void* A, const void* W, const GemmParams& p, int epi, cudaStream_t st, int force_bn = 0);
// 3x3 / stride 1 / pad 1 convolution on NHWC bf16 through the same kernels: x [B,H,W,Cin], w_packed [Cout, 9*kpt*64]
// (tap-major, channels zero-padded to kpt*64), out [B,H,W,Cout] bf16 = conv + bias (+ resid). Requires Cout % 32 == 0,
// Cin % 8 == 0, W | 128 or 128 | W, (H*W) % 128 == 0.
int conv3x3_launch(const void* x, const void* w_packed, const float* bias, const void* resid, void* out, int B, int H,
                   int W, int Cin, int Cout, cudaStream_t st, int stride = 1, float2* gn_part = nullptr);

}  // namespace vb
