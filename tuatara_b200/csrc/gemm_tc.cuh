// Host-side interface of the one tensor-core kernel both networks run on:
//   out[M x N] = epilogue( A[M x K] * Wt[N x K]^T )      bf16 x bf16 -> fp32 (TMEM) -> bf16/fp32
// A is either a plain row-major matrix (PARSeq linears) or an NHWC activation read as an
// implicit GEMM (CRAFT convs: per-tap shifted TMA boxes, OOB zero fill == zero padding,
// optional second source == channel concat without materialising it).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace tt {

enum Act : int { ACT_NONE = 0, ACT_RELU = 1, ACT_GELU = 2 };
enum ResType : int { RES_NONE = 0, RES_BF16 = 1, RES_F32 = 2, RES_SPLIT = 3 };
enum OutType : int { OUT_BF16 = 0, OUT_F32 = 1, OUT_CLS_TAIL = 2, OUT_SPLIT = 3 };
// RES_SPLIT / OUT_SPLIT: an fp32 tensor stored as two bf16 tensors hi = bf16(x), lo = bf16(x - hi) (same bytes as fp32,
// x = hi + lo to ~2^-17 relative).  The PARSeq encoder keeps its residual stream this way: `hi` IS the bf16 operand the
// next GEMM reads, so the LayerNorm-fused path needs no separate bf16 copy of x.

struct Epilogue {
  const float* bias = nullptr;      // [N] fp32 (BatchNorm already folded in)
  int act = ACT_NONE;
  const void* residual = nullptr;   // added after bias, before the activation; same element type as `out`
  int res_type = RES_NONE;
  int ldr = 0;                      // residual row pitch (elements)
  int res_mod = 0;                  // >0: residual row = m % res_mod (positional table)
  void* out = nullptr;
  int out_type = OUT_BF16;
  int ldc = 0;                      // output row pitch (elements)
  // OUT_CLS_TAIL: after bias+ReLU on the 16 accumulators, two 1x1 convs in registers
  // (16->16 ReLU, 16->2) and an fp32 [pixel][2] store.  tail = {w4[16][16], b4[16], w5[2][16], b5[2]}
  const float* tail = nullptr;
  // ---- 2x2 max-pool fused into a conv's bf16 epilogue (CRAFT: vgg16_bn features 6/13/23/33 follow conv1_2, conv2_2,
  // conv3_3, conv4_3).  pool_mode 1: only the pooled tensor [B][H/2][W/2][Cout] is written (out may be null);
  // 2: both (conv2_2's un-pooled output is a U-net skip tensor).  Pooling the bf16-rounded values equals rounding the
  // pooled fp32 values (rounding is monotonic), so the bits are those of the separate pooling kernel.
  __nv_bfloat16* pool_out = nullptr;
  int pool_mode = 0;
  // ---- LayerNorm fused away (PARSeq encoder: x -> LN -> Linear).  With W' = W * gamma (folded at export),
  //   Linear(LN(x))[n] = rstd * (x . W'[n]) - rstd * mean * c1[n] + c0[n],  c1[n] = sum_k W'[n][k], c0[n] = b[n] + beta . W[n]
  // so the GEMM reads the UN-normalised row as bf16 and its epilogue applies the row's (mean, rstd).
  // Producer (RES_SPLIT in, OUT_SPLIT out, TMA epilogue: `residual` / `out` are the hi tensors, `residual_lo` / `out_lo`
  // the lo tensors, all bf16 with pitches ldr / ldc): x = hi + lo + acc + bias in fp32, written back split, and per row
  // and N tile (sum x, sum x^2).
  const void* residual_lo = nullptr;
  void* out_lo = nullptr;
  float* ln_stats_out = nullptr;       // [M][ln_parts][2] fp32 partial sums; ln_parts = N / BN is written to *ln_parts_out
  int* ln_parts_out = nullptr;
  // Consumer (bf16 out): `bias` holds c0, ln_c1 holds c1, ln_stats_in the producer's partial sums over ln_dim columns.
  const float* ln_stats_in = nullptr;
  const float* ln_c1 = nullptr;
  int ln_parts = 0;
  int ln_dim = 0;
  float ln_eps = 1e-6f;
};

struct ConvSrc {
  const __nv_bfloat16* ptr = nullptr;  // NHWC
  int C = 0;                           // channels used from this source
  int pitch = 0;                       // channel pitch of the buffer (>= C)
};

struct ConvProblem {
  int batch = 1, H = 0, W = 0;    // stride-1 "same" conv: input and output share H x W
  ConvSrc src[2];
  int nsrc = 1;
  int taps = 9;                   // 9 (3x3) or 1 (1x1)
  int dil = 1;
  const __nv_bfloat16* weight = nullptr;  // [Cout][taps][C0 + C1], K-major
  int Cout = 0;
  int BN = 0;                     // 0 = choose
  int resident = -1;              // weight-resident schedule: -1 auto, 0 never, 1 force (with BN given)
  int pair = -1;                  // CTA-pair (cta_group::2, 256-row tiles): -1 auto, 0 never, 1 force
  int algo_k = 0;                 // true K for FLOP accounting when the stored K is padded (conv1_1: 27)
};

struct LinearProblem {
  const __nv_bfloat16* A = nullptr;  // [M][lda]
  int lda = 0;
  int M = 0, K = 0;
  const __nv_bfloat16* W = nullptr;  // [N][K]
  int N = 0;
  int BN = 0;
  int resident = -1;                 // weight-resident schedule: -1 auto, 0 never, 1 force (with BN given)
  int pair = -1;                     // CTA-pair (cta_group::2, 256-row tiles): -1 auto, 0 never, 1 force
  int algo_n = 0;                    // true N for FLOP accounting when N is padded (head: 95)
};

// Both return cudaSuccess or the first error (also recorded via tt::set_error).
cudaError_t conv_forward(const ConvProblem& p, const Epilogue& e, cudaStream_t s);
cudaError_t linear_forward(const LinearProblem& p, const Epilogue& e, cudaStream_t s);

// cuTensorMapEncodeTiled for a bf16 tensor (dims/strides innermost first; strides for dims 1..rank-1 in
// bytes); row_bytes = box[0]*2 selects the swizzle (128 -> SWIZZLE_128B, 64 -> SWIZZLE_64B).
bool make_tmap_bf16(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                    const cuuint32_t* box, int row_bytes);


}  // namespace tt
