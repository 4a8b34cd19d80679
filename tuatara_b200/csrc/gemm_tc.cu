// The tensor-core kernel of the OCR path: a persistent, warp-specialised tcgen05 GEMM.
//
//   warp 0      TMA producer   (one lane): per k-block, A tile (128 rows) + W tile (BN rows) -> smem ring
//   warp 1      MMA issuer     (one lane): tcgen05.mma cta_group::1, M=128, N=BN, K=16 per instruction,
//                                          fp32 accumulators in TMEM, double-buffered (2 x 256 columns)
//   warps 2..9  epilogue       (8 warps) : tcgen05.ld -> bias / ReLU / GELU / residual / cls tail -> global
//
// Replaces the ATen conv2d/batch_norm/relu/linear calls the reference reaches through
// TorchScript at tuatara.cpp:376 (CRAFT) and tuatara.cpp:307 (PARSeq).
//
// Implicit-GEMM convolution: a CTA's 128 output pixels are a TH x TW rectangle of one image; for
// tap (dy,dx) and channel block c the A tile is the TMA box {BK ch, TW, TH, 1} at
// (c*BK, x0+dx*dil, y0+dy*dil, n) of the NHWC tensor -- out-of-bounds elements are zero-filled by
// the TMA unit, which *is* the conv's zero padding.  A second source tensor continues the K loop
// (channel concat fused away).  K order of the weights: [tap][src0 channels | src1 channels].
#include "gemm_tc.cuh"

#include <cuda.h>

#include <algorithm>
#include <cstdio>
#include <mutex>

#include "common.h"
#include "ptx.cuh"

namespace tt {

namespace {

constexpr int kBlockM = 128;
constexpr int kThreads = 320;       // 10 warps
constexpr int kEpiWarps = 8;
constexpr int kMaxStages = 8;
constexpr int kAccStride = 256;     // TMEM columns per accumulator stage
constexpr int kTmemCols = 512;
constexpr int kSmemBudget = 222 * 1024;
constexpr int kResidentMax = 144 * 1024;  // largest weight slice kept in smem (BN=192 x K=384 bf16)

struct KParams {
  CUtensorMap tmA[2];
  CUtensorMap tmB;
  int mode;  // 0 plain rows, 1 conv tiles
  int M, N, BN, BK;
  int kb_src[2];
  int taps, dil;
  int H, W, TH, TW, tiles_x, tiles_y;
  int num_m_tiles, num_n_tiles;
  int stages;
  int b_resident;  // 1: the CTA's [BN x K] weight slice is loaded once and stays in smem; only A streams
  Epilogue epi;
};

struct SmemCtl {
  uint64_t full[kMaxStages];
  uint64_t empty[kMaxStages];
  uint64_t acc_full[2];
  uint64_t acc_empty[2];
  uint64_t b_full;
  uint32_t tmem_base;
  float tail[16 * 16 + 16 + 2 * 16 + 2];
};

// Tile schedule shared by the three warp roles.  Streaming mode: tiles round-robin over CTAs with the
// N index fastest (neighbouring CTAs share the A tile in L2).  Weight-resident mode: a CTA owns one
// N slice for its whole life and walks M tiles with stride `groups`.
struct TileIter {
  int m, n, step, limit;
  bool resident;
  __device__ TileIter(const KParams& p) {
    resident = p.b_resident != 0;
    if (resident) {
      const int groups = gridDim.x / p.num_n_tiles;
      n = blockIdx.x % p.num_n_tiles;
      m = blockIdx.x / p.num_n_tiles;
      step = groups;
      limit = p.num_m_tiles;
    } else {
      m = blockIdx.x;  // linear tile index in this mode
      n = 0;
      step = gridDim.x;
      limit = p.num_m_tiles * p.num_n_tiles;
    }
  }
  __device__ bool valid() const { return m < limit; }
  __device__ void next() { m += step; }
  __device__ int m_tile(const KParams& p) const { return resident ? m : m / p.num_n_tiles; }
  __device__ int n_tile(const KParams& p) const { return resident ? n : m % p.num_n_tiles; }
};

// GELU(x) = x * Phi(x) evaluated as 0.5 x (1 + tanh(x (a + b x^2 + c x^4))) with (a, b, c) fitted to
// the erf form: |error| <= 2.6e-5 on the real line (the textbook tanh form is off by 4.7e-4), plus
// tanh.approx.f32's 2^-11 relative error -- both far below the bf16 rounding applied to the result.
// 8 issue slots + 1 MUFU per element: the exact erf costs > 20 and made fc1's epilogue the bottleneck.
__device__ __forceinline__ float gelu_fast(float x) {
  const float x2 = fminf(x * x, 36.0f);  // tanh is saturated beyond |x| = 6; keeps the quartic monotone
  float p = fmaf(x2, -0.00035151678866f, 0.037005646023f);
  p = fmaf(x2, p, 0.797507884285f);
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(x * p));
  const float hx = 0.5f * x;
  return fmaf(hx, t, hx);
}

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

struct ResRegs { uint4 u[4]; };  // 16 fp32 or 16 bf16 (first two) residual values of one chunk

__global__ void __launch_bounds__(kThreads, 1) gemm_tc_kernel(const __grid_constant__ KParams p) {
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B tiles need 1024-byte alignment
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int row_bytes = p.BK * 2;
  const int a_bytes = kBlockM * row_bytes;
  const int b_bytes = p.BN * row_bytes;
  const int kb_per_tap = p.kb_src[0] + p.kb_src[1];
  const int num_kb = p.taps * kb_per_tap;
  const bool resident = p.b_resident != 0;
  const int stage_bytes = resident ? a_bytes : a_bytes + b_bytes;
  uint8_t* sBres = smem;                                        // [num_kb][BN rows] when resident
  uint8_t* ring = smem + (resident ? num_kb * b_bytes : 0);
  SmemCtl* ctl = reinterpret_cast<SmemCtl*>(ring + p.stages * stage_bytes);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    ptx::prefetch_tmap(&p.tmA[0]);
    ptx::prefetch_tmap(&p.tmB);
    if (p.kb_src[1] > 0) ptx::prefetch_tmap(&p.tmA[1]);
    for (int s = 0; s < p.stages; ++s) {
      ptx::mbar_init(&ctl->full[s], 1);
      ptx::mbar_init(&ctl->empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(&ctl->acc_full[s], 1);
      ptx::mbar_init(&ctl->acc_empty[s], kEpiWarps);
    }
    ptx::mbar_init(&ctl->b_full, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc(&ctl->tmem_base, kTmemCols);
  if (p.epi.out_type == OUT_CLS_TAIL) {
    for (int i = threadIdx.x; i < 16 * 16 + 16 + 2 * 16 + 2; i += kThreads) ctl->tail[i] = p.epi.tail[i];
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = ctl->tmem_base;

  if (warp == 0) {
    // ------------------------------------------------------------- TMA producer
    if (lane == 0) {
      TileIter it(p);
      if (resident && it.valid()) {
        ptx::mbar_arrive_expect_tx(&ctl->b_full, static_cast<uint32_t>(num_kb * b_bytes));
        for (int kb = 0; kb < num_kb; ++kb)
          ptx::tma_load_2d(sBres + kb * b_bytes, &p.tmB, &ctl->b_full, kb * p.BK, it.n_tile(p) * p.BN);
      }
      int stage = 0;
      uint32_t phase = 0;
      for (; it.valid(); it.next()) {
        const int n_tile = it.n_tile(p), m_tile = it.m_tile(p);
        int img = 0, y0 = 0, x0 = 0;
        if (p.mode == 1) {
          const int per_img = p.tiles_x * p.tiles_y;
          img = m_tile / per_img;
          const int t = m_tile - img * per_img;
          y0 = (t / p.tiles_x) * p.TH;
          x0 = (t % p.tiles_x) * p.TW;
        }
        int kb = 0;
        for (int tap = 0; tap < p.taps; ++tap) {
          const int dy = (p.taps == 9) ? (tap / 3 - 1) * p.dil : 0;
          const int dx = (p.taps == 9) ? (tap % 3 - 1) * p.dil : 0;
          for (int src = 0; src < 2; ++src) {
            for (int cb = 0; cb < p.kb_src[src]; ++cb, ++kb) {
              ptx::mbar_wait(&ctl->empty[stage], phase ^ 1);
              uint8_t* sA = ring + stage * stage_bytes;
              ptx::mbar_arrive_expect_tx(&ctl->full[stage], static_cast<uint32_t>(stage_bytes));
              if (p.mode == 1)
                ptx::tma_load_4d(sA, &p.tmA[src], &ctl->full[stage], cb * p.BK, x0 + dx, y0 + dy, img);
              else
                ptx::tma_load_2d(sA, &p.tmA[src], &ctl->full[stage], cb * p.BK, m_tile * kBlockM);
              if (!resident) ptx::tma_load_2d(sA + a_bytes, &p.tmB, &ctl->full[stage], kb * p.BK, n_tile * p.BN);
              if (++stage == p.stages) { stage = 0; phase ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // --------------------------------------------------------------- MMA issuer
    if (lane == 0) {
      const uint32_t idesc = ptx::make_idesc_bf16(kBlockM, p.BN);
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      TileIter it(p);
      if (resident && it.valid()) ptx::mbar_wait(&ctl->b_full, 0);
      const int ksteps = p.BK / 16;
      for (; it.valid(); it.next()) {
        ptx::mbar_wait(&ctl->acc_empty[as], aphase ^ 1);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * kAccStride;
        for (int kb = 0; kb < num_kb; ++kb) {
          ptx::mbar_wait(&ctl->full[stage], phase);
          ptx::tc_fence_after();
          const uint32_t a_addr = ptx::smem_u32(ring + stage * stage_bytes);
          const uint32_t b_addr = resident ? ptx::smem_u32(sBres + kb * b_bytes) : a_addr + a_bytes;
          for (int k = 0; k < ksteps; ++k) {
            const uint64_t da = ptx::make_smem_desc(a_addr + k * 32, row_bytes);
            const uint64_t db = ptx::make_smem_desc(b_addr + k * 32, row_bytes);
            ptx::mma_bf16(d_tmem, da, db, idesc, (kb | k) != 0);
          }
          ptx::mma_commit(&ctl->empty[stage]);
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        ptx::mma_commit(&ctl->acc_full[as]);
        if (++as == 2) { as = 0; aphase ^= 1; }
      }
    }
  } else {
    // ----------------------------------------------------------------- epilogue
    const int q = warp & 3;             // TMEM lane quadrant this warp may read
    const int half = (warp - 2) >> 2;   // two warps share a quadrant, interleaving 16-column chunks
    const int r = q * 32 + lane;        // tile row == TMEM lane
    const Epilogue& e = p.epi;
    int as = 0;
    uint32_t aphase = 0;
    const int chunks = p.BN / 16;
    for (TileIter it(p); it.valid(); it.next()) {
      const int n_tile = it.n_tile(p), m_tile = it.m_tile(p);
      long long orow;  // output row (pixel index or matrix row)
      bool valid;
      if (p.mode == 1) {
        const int per_img = p.tiles_x * p.tiles_y;
        const int img = m_tile / per_img;
        const int t = m_tile - img * per_img;
        const int y = (t / p.tiles_x) * p.TH + r / p.TW;
        const int x = (t % p.tiles_x) * p.TW + r % p.TW;
        valid = (y < p.H) && (x < p.W);
        orow = (static_cast<long long>(img) * p.H + y) * p.W + x;
      } else {
        orow = static_cast<long long>(m_tile) * kBlockM + r;
        valid = orow < p.M;
      }
      const int n0 = n_tile * p.BN;
      // residual of this thread's row: fetched one chunk ahead of its use so the global-load latency
      // hides behind the TMEM load / math / stores of the previous chunk
      const bool has_res = e.res_type != RES_NONE && valid;
      const char* res_row = nullptr;
      if (has_res) {
        const long long rrow = e.res_mod > 0 ? (orow % e.res_mod) : orow;
        res_row = static_cast<const char*>(e.residual) + (rrow * e.ldr + n0) * (e.res_type == RES_F32 ? 4 : 2);
      }
      auto load_res = [&](int ch, ResRegs& rr) {
        if (!has_res || n0 + ch * 16 >= p.N) return;
        if (e.res_type == RES_F32) {
          const uint4* src = reinterpret_cast<const uint4*>(res_row + ch * 64);
#pragma unroll
          for (int i = 0; i < 4; ++i) rr.u[i] = src[i];
        } else {
          const uint4* src = reinterpret_cast<const uint4*>(res_row + ch * 32);
          rr.u[0] = src[0];
          rr.u[1] = src[1];
        }
      };
      ResRegs res_next;
      load_res(half, res_next);
      ptx::mbar_wait(&ctl->acc_full[as], aphase);
      ptx::tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * kAccStride;
      for (int ch = half; ch < chunks; ch += 2) {
        uint32_t raw[16];
        __syncwarp();  // tcgen05.ld is warp-collective: reconverge after the masked stores below
        ptx::tmem_ld16(t_row + ch * 16, raw);
        const ResRegs res = res_next;
        if (ch + 2 < chunks) load_res(ch + 2, res_next);
        const int col0 = n0 + ch * 16;
        float bias[16];
        if (e.bias != nullptr && col0 < p.N) {
          const float4* b4 = reinterpret_cast<const float4*>(e.bias + col0);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 b = __ldg(b4 + i);
            bias[4 * i + 0] = b.x; bias[4 * i + 1] = b.y; bias[4 * i + 2] = b.z; bias[4 * i + 3] = b.w;
          }
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) bias[i] = 0.f;
        }
        ptx::tmem_ld_wait();
        if (col0 >= p.N) continue;  // warp-uniform
        float v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(raw[i]) + bias[i];
        if (has_res) {
          if (e.res_type == RES_F32) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              v[4 * i + 0] += __uint_as_float(res.u[i].x); v[4 * i + 1] += __uint_as_float(res.u[i].y);
              v[4 * i + 2] += __uint_as_float(res.u[i].z); v[4 * i + 3] += __uint_as_float(res.u[i].w);
            }
          } else {
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              const uint32_t w[4] = {res.u[i].x, res.u[i].y, res.u[i].z, res.u[i].w};
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&w[j]);
                v[8 * i + 2 * j + 0] += __low2float(h);
                v[8 * i + 2 * j + 1] += __high2float(h);
              }
            }
          }
        }
        if (e.act == ACT_RELU) {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.0f);
        } else if (e.act == ACT_GELU) {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = gelu_fast(v[i]);
        }
        if (!valid) {
          // masked row (tile overhangs the image / matrix): nothing to store
        } else if (e.out_type == OUT_BF16) {
          uint4 o0, o1;
          o0.x = pack_bf16(v[0], v[1]);   o0.y = pack_bf16(v[2], v[3]);
          o0.z = pack_bf16(v[4], v[5]);   o0.w = pack_bf16(v[6], v[7]);
          o1.x = pack_bf16(v[8], v[9]);   o1.y = pack_bf16(v[10], v[11]);
          o1.z = pack_bf16(v[12], v[13]); o1.w = pack_bf16(v[14], v[15]);
          uint4* dst = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(e.out) + orow * e.ldc + col0);
          dst[0] = o0;
          dst[1] = o1;
          if (e.out2 != nullptr) {
            uint4* dst2 = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(e.out2) + orow * e.ldc + col0);
            dst2[0] = o0;
            dst2[1] = o1;
          }
        } else if (e.out_type == OUT_F32) {
          float* dst = static_cast<float*>(e.out) + orow * e.ldc + col0;
          if (col0 + 16 <= p.N) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
              reinterpret_cast<float4*>(dst)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
          } else {
            for (int i = 0; i < 16; ++i)
              if (col0 + i < p.N) dst[i] = v[i];
          }
        } else {  // OUT_CLS_TAIL: v = relu(conv 32->16); two 1x1 convs in registers
          const float* w4 = ctl->tail;
          const float* b4 = w4 + 256;
          const float* w5 = b4 + 16;
          const float* b5 = w5 + 32;
          float h[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            float acc = b4[j];
#pragma unroll
            for (int i = 0; i < 16; ++i) acc = fmaf(w4[j * 16 + i], v[i], acc);
            h[j] = fmaxf(acc, 0.0f);
          }
          float o0 = b5[0], o1 = b5[1];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            o0 = fmaf(w5[j], h[j], o0);
            o1 = fmaf(w5[16 + j], h[j], o1);
          }
          reinterpret_cast<float2*>(e.out)[orow] = make_float2(o0, o1);
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&ctl->acc_empty[as]);
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ------------------------------------------------------------------------ host side
using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

}  // namespace

bool make_tmap_bf16(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                    const cuuint32_t* box, int row_bytes) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
    return false;
  }
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  const CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(base), dims, strides_bytes, box,
                        estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult " + std::to_string(static_cast<int>(r)));
    return false;
  }
  return true;
}

namespace {

inline bool make_map(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                     const cuuint32_t* box, int row_bytes) {
  return make_tmap_bf16(m, base, rank, dims, strides_bytes, box, row_bytes);
}

int num_sms() {
  static int n = 0;  // all GPUs of one box are the same part
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  }
  return n;
}

// Cost model behind the tile-width / schedule choice (cycles per SM, see DESIGN.md "GEMM schedule"):
//   one k-block of a 128 x BN tile costs max(MMA, operand traffic): MMA = 128*BN*BK/4096 cycles
//   (tcgen05 cta_group::1 rate), traffic = bytes / ~40 B/cycle/SM (L2 -> SM share with all SMs busy).
//   Streaming reloads the weight tile for every M tile; weight-resident keeps the CTA's [BN x K]
//   slice in smem and streams only A.
struct Plan { int BN; int resident; double cost; };

Plan plan_tiles(int N, int Ktot, int BK, long long m_tiles, bool allow_resident) {
  static const int cand[] = {256, 192, 128, 96, 64, 48, 32, 16};
  const double kL2 = 40.0;
  const int sms = num_sms();
  const int kblocks = Ktot / BK;
  Plan best{0, 0, 1e300};
  for (int c : cand) {
    if (N % c != 0) continue;
    const double mma = 128.0 * c * BK / 4096.0;
    const double a_bytes = 128.0 * BK * 2, b_bytes = static_cast<double>(c) * BK * 2;
    const long long n_tiles = N / c;
    {  // streaming
      const double tile = kblocks * std::max(mma, (a_bytes + b_bytes) / kL2) + 600.0;
      const double waves = static_cast<double>((m_tiles * n_tiles + sms - 1) / sms);
      const double cost = waves * tile;
      if (cost < best.cost) best = Plan{c, 0, cost};
    }
    const long long groups = std::min<long long>(sms / n_tiles, m_tiles);
    if (allow_resident && n_tiles <= sms && static_cast<long long>(c) * Ktot * 2 <= kResidentMax && groups >= 1 &&
        m_tiles >= 4 * groups) {
      const int stages = (kSmemBudget - static_cast<int>(sizeof(SmemCtl)) - 1024 - c * Ktot * 2) / (128 * BK * 2);
      if (stages >= 3) {
        const double tile = kblocks * std::max(mma, a_bytes / kL2) + 600.0;
        const double cost = static_cast<double>((m_tiles + groups - 1) / groups) * tile + b_bytes * kblocks / kL2;
        if (cost < best.cost) best = Plan{c, 1, cost};
      }
    }
  }
  if (best.BN == 0) best = Plan{((N + 15) / 16) * 16 <= 256 ? ((N + 15) / 16) * 16 : 128, 0, 0.0};
  return best;
}

cudaError_t launch(KParams& kp, cudaStream_t s, double flops) {
  const int row_bytes = kp.BK * 2;
  const int num_kb = kp.taps * (kp.kb_src[0] + kp.kb_src[1]);
  if (kp.b_resident && (num_kb * kp.BN * row_bytes > kResidentMax || kp.num_n_tiles > num_sms())) kp.b_resident = 0;
  const int a_bytes = kBlockM * row_bytes, b_bytes = kp.BN * row_bytes;
  const int res_bytes = kp.b_resident ? num_kb * b_bytes : 0;
  const int stage_bytes = kp.b_resident ? a_bytes : a_bytes + b_bytes;
  kp.stages = std::max(2, std::min(kMaxStages, (kSmemBudget - static_cast<int>(sizeof(SmemCtl)) - 1024 - res_bytes) / stage_bytes));
  const size_t smem = static_cast<size_t>(res_bytes) + static_cast<size_t>(kp.stages) * stage_bytes + sizeof(SmemCtl) + 1024;
  TT_CUDA_TRY(ensure_dynamic_smem(reinterpret_cast<const void*>(gemm_tc_kernel), 227 * 1024));
  int grid;
  if (kp.b_resident) {
    const int groups = std::min(num_sms() / kp.num_n_tiles, kp.num_m_tiles);
    grid = groups * kp.num_n_tiles;
  } else {
    grid = std::min(kp.num_m_tiles * kp.num_n_tiles, num_sms());
  }
  char tag[112];
  if (prof_enabled()) {
    std::snprintf(tag, sizeof(tag), "%s M%d N%d K%d BN%d BK%d st%d grid%d %s", kp.mode ? "conv" : "lin", kp.M, kp.N,
                  num_kb * kp.BK, kp.BN, kp.BK, kp.stages, grid, kp.b_resident ? "resident" : "stream");
    prof_record(s, true, 0, 0);
  }
  gemm_tc_kernel<<<grid, kThreads, smem, s>>>(kp);
  prof_record(s, false, flops, 0, tag);
  TT_LAUNCH_CHECK();
  return cudaSuccess;
}

cudaError_t check_epilogue(const Epilogue& e, int N, int BN) {
  if (e.out == nullptr) { set_error("gemm: null output"); return cudaErrorInvalidValue; }
  if (e.out_type == OUT_CLS_TAIL && (BN != 16 || N != 16 || e.tail == nullptr)) {
    set_error("gemm: cls tail needs N == BN == 16 and tail weights");
    return cudaErrorInvalidValue;
  }
  if (e.out_type == OUT_BF16 && (e.ldc % 8 != 0 || N % 16 != 0)) {
    set_error("gemm: bf16 output needs ldc % 8 == 0 and N % 16 == 0");
    return cudaErrorInvalidValue;
  }
  return cudaSuccess;
}

}  // namespace

cudaError_t conv_forward(const ConvProblem& c, const Epilogue& e, cudaStream_t s) {
  KParams kp{};
  const int ctot = c.src[0].C + (c.nsrc > 1 ? c.src[1].C : 0);
  kp.BK = (c.src[0].C % 64 == 0 && (c.nsrc == 1 || c.src[1].C % 64 == 0)) ? 64 : 32;
  for (int i = 0; i < c.nsrc; ++i) {
    if (c.src[i].C % kp.BK != 0 || c.src[i].pitch % 8 != 0) {
      set_error("conv: channel count must be a multiple of 32 and pitch of 8");
      return cudaErrorInvalidValue;
    }
  }
  if (c.taps != 1 && c.taps != 9) { set_error("conv: taps must be 1 or 9"); return cudaErrorInvalidValue; }
  kp.mode = 1;
  kp.H = c.H; kp.W = c.W;
  // 128 output pixels per tile as a TH x TW rectangle; wider-than-tall keeps TMA rows long
  kp.TW = c.W >= 16 ? 16 : 8;
  kp.TH = kBlockM / kp.TW;
  kp.tiles_x = (c.W + kp.TW - 1) / kp.TW;
  kp.tiles_y = (c.H + kp.TH - 1) / kp.TH;
  kp.num_m_tiles = c.batch * kp.tiles_x * kp.tiles_y;
  kp.N = c.Cout;
  kp.M = c.batch * c.H * c.W;
  {
    const Plan pl = plan_tiles(c.Cout, c.taps * ctot, kp.BK, kp.num_m_tiles, c.resident != 0);
    kp.BN = c.BN ? c.BN : pl.BN;
    kp.b_resident = c.BN ? (c.resident == 1) : pl.resident;
  }
  kp.num_n_tiles = (c.Cout + kp.BN - 1) / kp.BN;
  kp.taps = c.taps; kp.dil = c.dil;
  kp.kb_src[0] = c.src[0].C / kp.BK;
  kp.kb_src[1] = c.nsrc > 1 ? c.src[1].C / kp.BK : 0;
  kp.epi = e;
  if (cudaError_t err = check_epilogue(e, c.Cout, kp.BN)) return err;
  const int row_bytes = kp.BK * 2;
  for (int i = 0; i < c.nsrc; ++i) {
    const cuuint64_t dims[4] = {static_cast<cuuint64_t>(c.src[i].C), static_cast<cuuint64_t>(c.W),
                                static_cast<cuuint64_t>(c.H), static_cast<cuuint64_t>(c.batch)};
    const cuuint64_t pitch = static_cast<cuuint64_t>(c.src[i].pitch) * 2;
    const cuuint64_t strides[3] = {pitch, pitch * c.W, pitch * c.W * c.H};
    const cuuint32_t box[4] = {static_cast<cuuint32_t>(kp.BK), static_cast<cuuint32_t>(kp.TW),
                               static_cast<cuuint32_t>(kp.TH), 1};
    if (!make_map(&kp.tmA[i], c.src[i].ptr, 4, dims, strides, box, row_bytes)) return cudaErrorInvalidValue;
  }
  {
    const cuuint64_t ktot = static_cast<cuuint64_t>(c.taps) * ctot;
    const cuuint64_t dims[2] = {ktot, static_cast<cuuint64_t>(c.Cout)};
    const cuuint64_t strides[1] = {ktot * 2};
    const cuuint32_t box[2] = {static_cast<cuuint32_t>(kp.BK), static_cast<cuuint32_t>(kp.BN)};
    if (!make_map(&kp.tmB, c.weight, 2, dims, strides, box, row_bytes)) return cudaErrorInvalidValue;
  }
  const double k_algo = c.algo_k > 0 ? c.algo_k : static_cast<double>(c.taps) * ctot;
  return launch(kp, s, 2.0 * c.batch * c.H * c.W * c.Cout * k_algo);
}

cudaError_t linear_forward(const LinearProblem& l, const Epilogue& e, cudaStream_t s) {
  KParams kp{};
  kp.BK = (l.K % 64 == 0) ? 64 : 32;
  if (l.K % kp.BK != 0 || l.lda % 8 != 0) {
    set_error("linear: K must be a multiple of 32 and lda of 8");
    return cudaErrorInvalidValue;
  }
  kp.mode = 0;
  kp.M = l.M; kp.N = l.N;
  kp.num_m_tiles = (l.M + kBlockM - 1) / kBlockM;
  {
    const Plan pl = plan_tiles(l.N, l.K, kp.BK, kp.num_m_tiles, l.resident != 0);
    kp.BN = l.BN ? l.BN : pl.BN;
    kp.b_resident = l.BN ? (l.resident == 1) : pl.resident;
  }
  kp.num_n_tiles = (l.N + kp.BN - 1) / kp.BN;
  kp.taps = 1; kp.dil = 1;
  kp.kb_src[0] = l.K / kp.BK;
  kp.kb_src[1] = 0;
  kp.TW = 1; kp.TH = 1; kp.tiles_x = 1; kp.tiles_y = 1;
  kp.epi = e;
  if (cudaError_t err = check_epilogue(e, l.N, kp.BN)) return err;
  const int row_bytes = kp.BK * 2;
  {
    const cuuint64_t dims[2] = {static_cast<cuuint64_t>(l.K), static_cast<cuuint64_t>(l.M)};
    const cuuint64_t strides[1] = {static_cast<cuuint64_t>(l.lda) * 2};
    const cuuint32_t box[2] = {static_cast<cuuint32_t>(kp.BK), kBlockM};
    if (!make_map(&kp.tmA[0], l.A, 2, dims, strides, box, row_bytes)) return cudaErrorInvalidValue;
  }
  {
    const cuuint64_t dims[2] = {static_cast<cuuint64_t>(l.K), static_cast<cuuint64_t>(l.N)};
    const cuuint64_t strides[1] = {static_cast<cuuint64_t>(l.K) * 2};
    const cuuint32_t box[2] = {static_cast<cuuint32_t>(kp.BK), static_cast<cuuint32_t>(kp.BN)};
    if (!make_map(&kp.tmB, l.W, 2, dims, strides, box, row_bytes)) return cudaErrorInvalidValue;
  }
  return launch(kp, s, 2.0 * l.M * (l.algo_n > 0 ? l.algo_n : l.N) * l.K);
}

}  // namespace tt
