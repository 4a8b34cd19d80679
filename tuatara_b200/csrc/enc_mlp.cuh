// Fused MLP block of the PARSeq encoder (timm Block: x + fc2(GELU(fc1(LN2(x)))), reached by the reference through
// TorchScript at tuatara.cpp:307): one kernel per layer instead of the fc1 and fc2 GEMM launches, the hidden
// activations never leave the SM.  See enc_mlp.cu for the layout.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

namespace tt {

struct EncMlpWeights {
  const __nv_bfloat16* w1 = nullptr;   // [mlp][D]  fc1 with LayerNorm's gamma folded in (weights.py: "<block>.fc1.wf")
  const float* c0 = nullptr;           // [mlp]     fc1 bias + beta W   ("<block>.fc1.c0")
  const float* c1 = nullptr;           // [mlp]     row sums of the rounded folded weights ("<block>.fc1.c1")
  const __nv_bfloat16* w2 = nullptr;   // [D][mlp]  fc2
  const float* b2 = nullptr;           // [D]
};

// The attention block's output projection in front of the MLP: x1 = x + att Wp^T + bp is formed inside the kernel (fp32 in
// TMEM, where fc2 accumulates onto it; bf16 in smem as fc1's operand; its LayerNorm sums in registers) instead of
// travelling to HBM and back between two kernels.
struct EncProj {
  const __nv_bfloat16* att = nullptr;  // [M][D] attention output
  const __nv_bfloat16* wp = nullptr;   // [D][D]
  const float* bp = nullptr;           // [D]
};

bool enc_mlp_supported(int D, int mlp);

// x = hi + lo (split bf16 residual stream, [M][D] each, updated in place); stats: per row `parts_in` partial
// (sum x, sum x^2) float2 on entry, 2 partials on return (columns [0, D/2) and [D/2, D)).
// proj != null: x = x + att Wp^T + bp first (stats are not read then, only written).
cudaError_t enc_mlp_forward(const EncMlpWeights& w, __nv_bfloat16* hi, __nv_bfloat16* lo, float* stats, int parts_in,
                            long long M, int D, int mlp, float eps, cudaStream_t s, const EncProj* proj = nullptr);

}  // namespace tt
