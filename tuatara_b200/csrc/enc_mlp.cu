// Second half of a PARSeq encoder block in ONE tcgen05 kernel (see enc_mlp.cuh):
//     x1 = x + att Wp^T + bp            (the attention output projection; template PROJ)
//     x' = x1 + fc2(GELU(fc1(LN2(x1))))
//
// Why: as GEMM launches (gemm_tc.cu: proj, fc1, fc2) x1 makes an HBM round trip between proj and the MLP and the
// [rows][1536] bf16 hidden tensor a 1.9 GB one per layer and 2400 crops, and the GEMMs are paced by the shared-memory port
// (DESIGN.md section 6): the hidden tile is written to staging buffers, read back by a TMA store, and brought in again by
// TMA as fc2's A operand.  Here a CTA pair owns 256 rows for the whole chain: GELU(fc1 chunk) goes from registers straight
// into the smem tile that fc2's MMA reads, x1 lives in TMEM.  At the 1000 W cap the removed traffic also shows as SM clock.
//
// A CTA pair (cta_group::2, one TPC) per 256-row tile, each CTA 128 rows = 128 TMEM lanes; per CTA
//   warp 0      TMA producer : the tile's A rows -> sX (PROJ: att, else x_hi; 6 k-blocks, resident for the tile); through a
//                              ring of 24 KB slots this CTA's half of every weight tile (PROJ: Wp, then the residual rows
//                              hi / lo block by block; then per hidden chunk 64 of its 128 fc1 rows and 192 of fc2's 384);
//                              L2 prefetches of the next tile's rows, one box per chunk
//   warp 1      MMA issuer   : (leader CTA only)  PROJ: acc2 = att Wp^T, then acc2 += hi + lo: each staged residual block
//                              is the A operand of an MMA whose B operand is a 64 x 64 identity (exact in fp32, no thread
//                              touches the residual).  acc1[128 cols] = x1 W1_c^T, acc2[384 cols] += GELU(h_c) W2_c^T,
//                              software pipelined: fc1 of chunk c+1 is issued before fc2 of chunk c, so the tensor pipe
//                              works while the row owners evaluate GELU(c)
//   warps 2..9  row owners   : thread = row; two warps per TMEM lane quadrant.  PROJ: bf16(acc2 + bp) -> sX (fc1's A
//                              operand, over the att tile) and the row's LayerNorm sums (the two halves meet in smem).
//                              Per chunk: LayerNorm applied algebraically (rstd * acc - rstd * mean * c1 + c0, as
//                              gemm_tc.cu's consumer epilogue), GELU, bf16 -> sH (K-major SWIZZLE_128B).  Final phase:
//                              x' = acc2 + b2 (+ bp) [+ hi + lo without PROJ: x_lo staged in sH and the idle ring] in fp32,
//                              hi' = bf16(x'), lo' = bf16(x' - hi') written over the staged tiles, the row's
//                              (sum, sum of squares) emitted for the next layer's LayerNorm
//   warp 10     store warp   : hands each finished 64-column block of hi' / lo' to TMA stores and releases its smem
// TMEM: columns [0, 384) = acc2 (the block's output rows), [384, 512) = acc1 (one hidden chunk).
// Every wait is bounded (ptx::mbar_wait traps after 4 s).
#include "enc_mlp.cuh"

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#include "common.h"
#include "epi_math.cuh"
#include "gemm_tc.cuh"
#include "ptx.cuh"

namespace tt {

namespace {

constexpr int kD = 384, kMlp = 1536;
constexpr int kKB = kD / 64;             // k-blocks of x
constexpr int kChunk = 128;              // hidden columns per chunk
constexpr int kNC = kMlp / kChunk;
constexpr int kUnit = 16384;             // [128 rows][64 bf16], SWIZZLE_128B
constexpr int kSlot = 24576;             // ring slot: fc1 = 3 x [64 rows][64 k], fc2 = [128 rows][64 k] + [64 rows][64 k]
constexpr int kStages = 3;
constexpr int kEpiWarps = 8;
constexpr int kRowThreads = 32 * kEpiWarps;
constexpr int kStoreWarp = 2 + kEpiWarps;   // warp 10: hands finished output blocks to TMA stores
constexpr int kThreads = 64 + kRowThreads + 32;
constexpr int kVec = 2 * kMlp + 2 * kD;  // c0 | c1 | b2 | proj bias
constexpr int kXch = 2 * 128 * 2;        // floats: the row-owner halves exchange their LayerNorm partial sums
constexpr int kTmemCols = 512;
constexpr uint32_t kAcc1 = 384;          // TMEM column of the hidden-chunk accumulator

struct MlpParams {
  CUtensorMap tm_hi, tm_lo;              // [M][D] bf16, box {64, 128}
  CUtensorMap tm_att;                    // PROJ: attention output [M][D] bf16, box {64, 128}
  CUtensorMap tm_wpa, tm_wpb;            // PROJ: proj weight [D][D], boxes {64 k, 128 rows} and {64 k, 64 rows}
  const __nv_bfloat16 *hi_g, *lo_g;      // PROJ: the residual stream, read by the row owners themselves
  const float* bp;                       // PROJ: proj bias
  CUtensorMap tm_w1;                     // [mlp][D], box {64 k, 64 rows}
  CUtensorMap tm_w2a, tm_w2b;            // [D][mlp], boxes {64 k, 128 rows} and {64 k, 64 rows}
  const float *c0, *c1, *b2;
  float* stats;                          // [M][2] float2 partial (sum, sum of squares): read at tile start, rewritten at its end
  long long M;
  int n_tiles;                           // 256-row tiles
  float eps;
  unsigned long long* dbg;               // TT_MLP_DEBUG=1 (development): role cycle counters of CTA 0
};

struct alignas(16) Ctl {
  uint64_t w_full[kStages], w_empty[kStages];
  uint64_t x_full, x_empty[kKB], vec_full;   // x_empty[j]: the store of block j of hi' has read sX block j
  uint64_t a1_full, a1_free, h_full, h_free, acc2_full, acc2_free;
  uint64_t lo_full[kKB];                 // x_lo block j has landed (sH halves for j < 2, the idle weight ring for the rest)
  uint64_t proj_full, x1_ready;          // PROJ: att Wp^T landed in acc2; x1 (fp32 in acc2, bf16 in sX) and the row statistics are ready
  uint64_t blk_done[kKB], fin_done;      // row owners -> store warp (output block j complete in smem); store warp -> row owners (per tile)
  uint32_t tmem_base;
};

constexpr int kIdent = 4096;             // PROJ: this CTA's 32 rows of a 64 x 64 bf16 identity (B operand that adds a staged tile into acc2)
constexpr int kSmem = kKB * kUnit + 2 * kUnit + kStages * kSlot + kIdent + (kVec + kXch) * 4 + static_cast<int>(sizeof(Ctl)) + 1024;
static_assert(kSmem <= 227 * 1024, "shared memory budget");

// arrive on a barrier of the leader CTA; release at cluster scope: the arriving CTA's smem writes (already fenced to the
// async proxy) are what the leader's next tcgen05.mma makes this CTA's tensor core read
__device__ __forceinline__ void arrive_leader(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ bool try_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(ptx::smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void wait_cluster(uint64_t* bar, uint32_t parity) {
  if (try_wait_cluster(bar, parity)) return;
  const uint64_t t0 = ptx::globaltimer_ns();
#pragma unroll 1
  for (;;) {
#pragma unroll 1
    for (int i = 0; i < 64; ++i)
      if (try_wait_cluster(bar, parity)) return;
    if (ptx::globaltimer_ns() - t0 > 4000000000ull) __trap();
  }
}

// TMA prefetch of one box into L2 (warp converged, one elected lane issues)
__device__ __forceinline__ void ptx_prefetch_2d_e(const CUtensorMap* m, int c0, int c1) {
  asm volatile("{\n\t.reg .pred q;\n\t" TT_ELECT_PRED "@q cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];\n\t}" ::"l"(
                   reinterpret_cast<uint64_t>(m)), "r"(c0), "r"(c1) : "memory");
}

// PROJ: the attention block's output projection runs in front of the MLP (see the file header)
template <bool PROJ>
__global__ void __launch_bounds__(kThreads, 1) k_enc_mlp(const __grid_constant__ MlpParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sX = smem;
  uint8_t* sH = sX + kKB * kUnit;
  uint8_t* ring = sH + 2 * kUnit;
  uint8_t* sI = ring + kStages * kSlot;
  float* vec = reinterpret_cast<float*>(sI + kIdent);
  float* xch = vec + kVec;
  Ctl* ctl = reinterpret_cast<Ctl*>(xch + kXch);
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const uint32_t rank = __shfl_sync(0xffffffffu, ptx::cluster_ctarank(), 0);   // 0 = leader (issues the MMAs)
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;

  if (threadIdx.x == 0) {
    ptx::prefetch_tmap(&p.tm_hi);
    ptx::prefetch_tmap(&p.tm_lo);
    ptx::prefetch_tmap(&p.tm_w1);
    ptx::prefetch_tmap(&p.tm_w2a);
    ptx::prefetch_tmap(&p.tm_w2b);
    for (int s = 0; s < kStages; ++s) { ptx::mbar_init(&ctl->w_full[s], 1); ptx::mbar_init(&ctl->w_empty[s], 1); }
    ptx::mbar_init(&ctl->x_full, 1);
    for (int j = 0; j < kKB; ++j) ptx::mbar_init(&ctl->x_empty[j], 1);
    ptx::mbar_init(&ctl->vec_full, 1);
    ptx::mbar_init(&ctl->a1_full, 1);
    ptx::mbar_init(&ctl->a1_free, 2 * kEpiWarps);     // both CTAs' row owners release the leader's
    ptx::mbar_init(&ctl->h_full, 2 * kEpiWarps);
    ptx::mbar_init(&ctl->h_free, 1);
    ptx::mbar_init(&ctl->acc2_full, 1);
    ptx::mbar_init(&ctl->acc2_free, 2 * kEpiWarps);
    for (int j = 0; j < kKB; ++j) ptx::mbar_init(&ctl->lo_full[j], 1);
    for (int j = 0; j < kKB; ++j) ptx::mbar_init(&ctl->blk_done[j], kRowThreads);
    ptx::mbar_init(&ctl->fin_done, 1);
    ptx::mbar_init(&ctl->proj_full, 1);
    ptx::mbar_init(&ctl->x1_ready, 2 * kEpiWarps);
    ptx::fence_barrier_init();
    ptx::mbar_arrive_expect_tx(&ctl->vec_full, (PROJ ? kVec : kVec - kD) * 4);
    ptx::bulk_load(vec, p.c0, kMlp * 4, &ctl->vec_full);
    ptx::bulk_load(vec + kMlp, p.c1, kMlp * 4, &ctl->vec_full);
    ptx::bulk_load(vec + 2 * kMlp, p.b2, kD * 4, &ctl->vec_full);
    if constexpr (PROJ) ptx::bulk_load(vec + 2 * kMlp + kD, p.bp, kD * 4, &ctl->vec_full);
  }
  if constexpr (PROJ) {
    // I[n][k] = (k == n), rows n = rank 32 .. + 32, K-major SWIZZLE_128B: D[:, n] += A[:, n] adds a staged bf16 tile to acc2 exactly
    for (int i = threadIdx.x; i < kIdent / 16; i += kThreads) reinterpret_cast<uint4*>(sI)[i] = make_uint4(0u, 0u, 0u, 0u);
    __syncthreads();
    if (threadIdx.x < 32) {
      const int i = threadIdx.x, k = static_cast<int>(rank) * 32 + i;
      *reinterpret_cast<uint16_t*>(sI + i * 128 + (((k >> 3) ^ (i & 7)) << 4) + (k & 7) * 2) = 0x3F80;   // bf16 1.0
    }
    ptx::fence_proxy_async();
  }
  // cluster barrier first (publishes the peer's barrier initialisation and makes sure the peer runs), then the
  // pair-collective TMEM allocation: see gemm_tc.cu / profiles/r2_hang_root_cause.md
  ptx::cluster_sync_all();
  if (warp == 1) ptx::tmem_alloc_pair(&ctl->tmem_base, kTmemCols);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, ctl->tmem_base, 0);

  if (warp == 0) {
    // ------------------------------------------------------------------------------ TMA producer (both CTAs)
    const uint32_t xfull_c = __shfl_sync(0xffffffffu, ptx::mapa(ptx::smem_u32(&ctl->x_full), 0), 0);
    const uint32_t wfull0_c = __shfl_sync(0xffffffffu, ptx::mapa(ptx::smem_u32(&ctl->w_full[0]), 0), 0);
    int stage = 0;
    uint32_t phase = 0, it_par = 0, n_it = 0;
    auto slot_begin = [&](uint32_t bytes = kSlot) -> uint8_t* {
      ptx::mbar_wait(&ctl->w_empty[stage], phase ^ 1); __syncwarp();
      if (rank == 0) ptx::mbar_arrive_expect_tx_e(&ctl->w_full[stage], 2u * bytes);   // the leader's barrier counts both CTAs' bytes
      return ring + stage * kSlot;
    };
    auto slot_end = [&]() { if (++stage == kStages) { stage = 0; phase ^= 1; } };
    // fc1 rows of chunk c: this CTA's 64 of the 128, k-blocks [3 part, 3 part + 3)
    auto load_w1 = [&](int c, int part) {
      uint8_t* dst = slot_begin();
      const uint32_t bar = wfull0_c + stage * 8;
      for (int i = 0; i < 3; ++i)
        ptx::tma_load_2d_pair_e(dst + i * 8192, &p.tm_w1, bar, (3 * part + i) * 64, c * kChunk + static_cast<int>(rank) * 64);
      slot_end();
    };
    // fc2, hidden columns [c 128 + kb2 64, + 64): this CTA's output rows [rank 128, + 128) of the N = 256 MMA and
    // 256 + [rank 64, + 64) of the N = 128 MMA
    auto load_w2 = [&](int c, int kb2) {
      uint8_t* dst = slot_begin();
      const uint32_t bar = wfull0_c + stage * 8;
      ptx::tma_load_2d_pair_e(dst, &p.tm_w2a, bar, c * kChunk + kb2 * 64, static_cast<int>(rank) * 128);
      ptx::tma_load_2d_pair_e(dst + kUnit, &p.tm_w2b, bar, c * kChunk + kb2 * 64, 256 + static_cast<int>(rank) * 64);
      slot_end();
    };
    for (int t = pair; t < p.n_tiles; t += n_pairs, it_par ^= 1) {
      const int m0 = (t * 2 + static_cast<int>(rank)) * 128;   // past M for the odd last tile's second half: zero rows in, nothing out
      // sX block by block, as the previous tile's hi' stores release it; then the ring (it staged x_lo blocks at the
      // end of the previous tile)
      if (rank == 0) ptx::mbar_arrive_expect_tx_e(&ctl->x_full, 2u * kKB * kUnit);
      for (int kb = 0; kb < kKB; ++kb) {
        ptx::mbar_wait(&ctl->x_empty[kb], it_par ^ 1); __syncwarp();
        ptx::tma_load_2d_pair_e(sX + kb * kUnit, PROJ ? &p.tm_att : &p.tm_hi, xfull_c, kb * 64, m0);
      }
      if (n_it > 0) { ptx::mbar_wait(&ctl->fin_done, (n_it - 1) & 1); __syncwarp(); }
      ++n_it;
      const bool more = t + n_pairs < p.n_tiles;
      if constexpr (PROJ) {   // the output projection's weights: 6 k-blocks, the slot layout of fc2
        for (int kb = 0; kb < kKB; ++kb) {
          uint8_t* dst = slot_begin();
          const uint32_t bar = wfull0_c + stage * 8;
          ptx::tma_load_2d_pair_e(dst, &p.tm_wpa, bar, kb * 64, static_cast<int>(rank) * 128);
          ptx::tma_load_2d_pair_e(dst + kUnit, &p.tm_wpb, bar, kb * 64, 256 + static_cast<int>(rank) * 64);
          slot_end();
        }
        // the residual rows x = hi + lo, block by block: the MMA issuer adds them into acc2 through the identity operand
        for (int j2 = 0; j2 < 2 * kKB; ++j2) {
          uint8_t* dst = slot_begin(kUnit);
          const uint32_t bar = wfull0_c + stage * 8;
          ptx::tma_load_2d_pair_e(dst, j2 < kKB ? &p.tm_hi : &p.tm_lo, bar, (j2 % kKB) * 64, m0);
          slot_end();
        }
      }
      // the MMA issuer's order: W1(0), then per chunk W1(c) before W2(c - 1), W2(last)
      load_w1(0, 0); load_w1(0, 1);
      for (int c = 1; c < kNC; ++c) {
        load_w1(c, 0); load_w1(c, 1);
        load_w2(c - 1, 0); load_w2(c - 1, 1);
        // this tile's x_lo blocks and the next tile's x_hi rows towards L2, one box per chunk (all at once they queue in
        // front of the weight loads): both are needed at the tile boundary, where their latency is exposed
        if constexpr (PROJ) {   // the next tile's attention rows and residual rows
          if (more) {
            if (c <= kKB) ptx_prefetch_2d_e(&p.tm_att, (c - 1) * 64, m0 + n_pairs * 256);
            else ptx_prefetch_2d_e(&p.tm_hi, (c - 1 - kKB) * 64, m0 + n_pairs * 256);
          }
        } else {
          if (c <= kKB) ptx_prefetch_2d_e(&p.tm_lo, (c - 1) * 64, m0);
          else if (more) ptx_prefetch_2d_e(&p.tm_hi, (c - 1 - kKB) * 64, m0 + n_pairs * 256);
        }
      }
      load_w2(kNC - 1, 0); load_w2(kNC - 1, 1);
      if (more) ptx_prefetch_2d_e(&p.tm_hi, (kKB - 1) * 64, m0 + n_pairs * 256);
      if constexpr (PROJ) {
        if (more)
          for (int j = 0; j < kKB; ++j) ptx_prefetch_2d_e(&p.tm_lo, j * 64, m0 + n_pairs * 256);
        continue;   // no x_lo staging: the residual was folded into acc2 at the start of the tile
      }
      // every MMA of the tile has completed: sH and the whole weight ring are idle, together they take the tile's six
      // x_lo blocks at once (one HBM round trip instead of one per block)
      ptx::mbar_wait(&ctl->acc2_full, it_par); __syncwarp();
      for (int j = 0; j < kKB; ++j) {
        ptx::mbar_arrive_expect_tx_e(&ctl->lo_full[j], kUnit);
        ptx::tma_load_2d_e(j < 2 ? sH + j * kUnit : ring + (j - 2) * kUnit, &p.tm_lo, &ctl->lo_full[j], j * 64, m0);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------------------- MMA issuer (leader CTA)
    if (rank == 0) {
      const uint32_t idesc1 = ptx::make_idesc_bf16(256, 128), idesc2a = ptx::make_idesc_bf16(256, 256), idesc2b = ptx::make_idesc_bf16(256, 128);
      const uint32_t acc1 = tmem_base + kAcc1, acc2 = tmem_base;
      const uint32_t x_addr = ptx::smem_u32(sX), h_addr = ptx::smem_u32(sH);
      int stage = 0;
      uint32_t phase = 0, n1 = 0, n2 = 0, it_par = 0;
      const bool dbg = p.dbg != nullptr && blockIdx.x == 0;
      long long w_x = 0, w_a1 = 0, w_w = 0, w_h = 0, w_acc2 = 0;
      const long long t_start = dbg ? clock64() : 0;
#define TT_TIMED(acc, stmt) do { const long long _t0 = dbg ? clock64() : 0; stmt; acc += dbg ? clock64() - _t0 : 0; } while (0)
      auto slot_wait = [&]() -> uint32_t {
        TT_TIMED(w_w, ptx::mbar_wait(&ctl->w_full[stage], phase)); __syncwarp();
        ptx::tc_fence_after();
        return ptx::smem_u32(ring + stage * kSlot);
      };
      auto slot_done = [&]() {
        ptx::mma_commit_pair_e(&ctl->w_empty[stage], 3);   // frees the slot in both CTAs
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      };
      auto mma1 = [&]() {   // acc1 = x W1_c^T  (overwrites: the row owners have drained the previous chunk)
        TT_TIMED(w_a1, ptx::mbar_wait(&ctl->a1_free, (n1 & 1) ^ 1)); __syncwarp();
        ptx::tc_fence_after();
        ++n1;
        for (int part = 0; part < 2; ++part) {
          const uint32_t b_addr = slot_wait();
          for (int i = 0; i < 3; ++i) {
            const int kb = 3 * part + i;
#pragma unroll
            for (int k = 0; k < 4; ++k)
              ptx::mma_bf16_pair_e(acc1, ptx::make_smem_desc(x_addr + kb * kUnit + k * 32, 128),
                                   ptx::make_smem_desc(b_addr + i * 8192 + k * 32, 128), idesc1, (kb | k) != 0 ? 1u : 0u);
          }
          slot_done();
        }
        ptx::mma_commit_pair_e(&ctl->a1_full, 3);
      };
      auto mma2 = [&](int c) {   // acc2 (+)= GELU(h_c) W2_c^T
        TT_TIMED(w_h, wait_cluster(&ctl->h_full, n2 & 1)); __syncwarp();
        ++n2;
        if (!PROJ && c == 0) { TT_TIMED(w_acc2, ptx::mbar_wait(&ctl->acc2_free, it_par ^ 1)); __syncwarp(); }   // the previous tile's rows have been read out
        ptx::tc_fence_after();
        for (int kb2 = 0; kb2 < 2; ++kb2) {
          const uint32_t b_addr = slot_wait();
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t da = ptx::make_smem_desc(h_addr + kb2 * kUnit + k * 32, 128);
            const uint32_t acc = (PROJ || (c | kb2 | k) != 0) ? 1u : 0u;   // PROJ: onto x1, which the row owners wrote into acc2
            ptx::mma_bf16_pair_e(acc2, da, ptx::make_smem_desc(b_addr + k * 32, 128), idesc2a, acc);
            ptx::mma_bf16_pair_e(acc2 + 256, da, ptx::make_smem_desc(b_addr + kUnit + k * 32, 128), idesc2b, acc);
          }
          slot_done();
        }
        ptx::mma_commit_pair_e(&ctl->h_free, 3);
      };
      for (int t = pair; t < p.n_tiles; t += n_pairs, it_par ^= 1) {
        TT_TIMED(w_x, ptx::mbar_wait(&ctl->x_full, it_par)); __syncwarp();
        ptx::tc_fence_after();
        if constexpr (PROJ) {
          // acc2 = att Wp^T, then the row owners turn it into x1 = x + acc2 + bias (fp32 back into acc2, bf16 into sX)
          TT_TIMED(w_acc2, ptx::mbar_wait(&ctl->acc2_free, it_par ^ 1)); __syncwarp();
          ptx::tc_fence_after();
          for (int kb = 0; kb < kKB; ++kb) {
            const uint32_t b_addr = slot_wait();
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t da = ptx::make_smem_desc(x_addr + kb * kUnit + k * 32, 128);
              const uint32_t acc = (kb | k) != 0 ? 1u : 0u;
              ptx::mma_bf16_pair_e(acc2, da, ptx::make_smem_desc(b_addr + k * 32, 128), idesc2a, acc);
              ptx::mma_bf16_pair_e(acc2 + 256, da, ptx::make_smem_desc(b_addr + kUnit + k * 32, 128), idesc2b, acc);
            }
            slot_done();
          }
          const uint32_t idesc3 = ptx::make_idesc_bf16(256, 64), i_addr = ptx::smem_u32(sI);
          for (int j2 = 0; j2 < 2 * kKB; ++j2) {   // acc2[:, 64 j .. + 64) += hi / lo block j
            const uint32_t a_addr = slot_wait();
#pragma unroll
            for (int k = 0; k < 4; ++k)
              ptx::mma_bf16_pair_e(acc2 + (j2 % kKB) * 64, ptx::make_smem_desc(a_addr + k * 32, 128), ptx::make_smem_desc(i_addr + k * 32, 128), idesc3, 1u);
            slot_done();
          }
          ptx::mma_commit_pair_e(&ctl->proj_full, 3);
          TT_TIMED(w_x, wait_cluster(&ctl->x1_ready, it_par)); __syncwarp();
          ptx::tc_fence_after();
        }
        mma1();
        for (int c = 1; c < kNC; ++c) { mma1(); mma2(c - 1); }
        mma2(kNC - 1);
        ptx::mma_commit_pair_e(&ctl->acc2_full, 3);
      }
      if (dbg && lane == 0) { p.dbg[0] = w_x; p.dbg[1] = w_a1; p.dbg[2] = w_w; p.dbg[3] = w_h; p.dbg[4] = w_acc2; p.dbg[5] = clock64() - t_start; }
    }
  } else if (warp == kStoreWarp) {
    // ------------------------------------------------------------------------------- store warp (both CTAs)
    uint32_t it_par = 0;
    for (int t = pair; t < p.n_tiles; t += n_pairs, it_par ^= 1) {
      const long long m0 = static_cast<long long>(t * 2 + static_cast<int>(rank)) * 128;
      for (int j = 0; j < kKB; ++j) {
        ptx::mbar_wait(&ctl->blk_done[j], it_par); __syncwarp();
        if (m0 < p.M) {
          ptx::tma_store_2d_e(&p.tm_hi, sX + j * kUnit, j * 64, static_cast<int>(m0));
          ptx::tma_store_2d_e(&p.tm_lo, j < 2 ? sH + j * kUnit : ring + (j - 2) * kUnit, j * 64, static_cast<int>(m0));
        }
        ptx::bulk_commit_e();
        if (ptx::elect_one()) {
          ptx::bulk_wait_read<0>();              // the block's stores have read their smem:
          ptx::mbar_arrive(&ctl->x_empty[j]);    //   sX block j may take the next tile's rows
          if (j == kKB - 1) ptx::mbar_arrive(&ctl->fin_done);   // and sH / the ring their hidden chunks / weights
        }
        __syncwarp();
      }
    }
    if (ptx::elect_one()) ptx::bulk_wait<0>();   // global writes done before the CTA exits
    __syncwarp();
  } else {
    // --------------------------------------------------- row owners (thread = row; two warps per TMEM lane quadrant)
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int r = q * 32 + lane;
    const uint32_t tl = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const uint32_t a1_free_c = ptx::mapa(ptx::smem_u32(&ctl->a1_free), 0);
    const uint32_t h_full_c = ptx::mapa(ptx::smem_u32(&ctl->h_full), 0);
    const uint32_t acc2_free_c = ptx::mapa(ptx::smem_u32(&ctl->acc2_free), 0);
    const uint32_t vec_s = ptx::smem_u32(vec);
    const uint32_t sw = static_cast<uint32_t>(r & 7);
    ptx::mbar_wait(&ctl->vec_full, 0);
    uint32_t n_chunk = 0, it_par = 0, n_tile = 0;
    const bool dbg = p.dbg != nullptr && blockIdx.x == 0 && warp == 2;
    long long e_a1 = 0, e_hfree = 0, e_acc2 = 0, e_final = 0, e_tiles = 0;
    const long long e_start = dbg ? clock64() : 0;
    for (int t = pair; t < p.n_tiles; t += n_pairs, it_par ^= 1) {
      const long long m0 = static_cast<long long>(t * 2 + static_cast<int>(rank)) * 128;
      const long long row = m0 + r;
      const bool row_ok = row < p.M;
      float s1 = 0.f, s2 = 0.f;
      if constexpr (PROJ) {
        // ---- acc2 = x + att Wp^T (the MMA issuer added the residual rows through the identity operand); x1 = acc2 + bias:
        //      bf16(x1) into sX (fc1's A operand) and the row's LayerNorm sums.  acc2 itself stays as it is: fc2 accumulates
        //      onto it and the final phase adds both biases.
        const uint32_t x1_ready_c = ptx::mapa(ptx::smem_u32(&ctl->x1_ready), 0);
        ptx::mbar_wait(&ctl->proj_full, it_par);
        ptx::tc_fence_after();
        uint64_t q1 = 0ull, q2 = 0ull;
        uint32_t rb0[2][32];
        ptx::tmem_ld<32>(tl + half * 32, rb0[0]);
#pragma unroll
        for (int j = 0; j < kKB; ++j) {
          uint32_t (&raw)[32] = rb0[j & 1];
          ptx::tmem_ld_wait(raw);
          if (j + 1 < kKB) ptx::tmem_ld<32>(tl + (j + 1) * 64 + half * 32, rb0[(j + 1) & 1]);
          const uint32_t xrow = ptx::smem_u32(sX) + j * kUnit + r * 128;
          const uint32_t bp_s = vec_s + (2 * kMlp + kD + j * 64 + half * 32) * 4;
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const uint4 b0 = ptx::lds128(bp_s + g * 32), b1 = ptx::lds128(bp_s + g * 32 + 16);
            const uint32_t bw[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
            uint32_t ho[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const uint64_t v = add2(pk2u(raw[8 * g + 2 * e], raw[8 * g + 2 * e + 1]), pk2u(bw[2 * e], bw[2 * e + 1]));
              q1 = add2(q1, v);
              q2 = fma2(v, v, q2);
              float x0, x1;
              upk2(v, x0, x1);
              ho[e] = pack_bf16(x0, x1);
            }
            ptx::sts128(xrow + ((static_cast<uint32_t>(half * 4 + g) ^ sw) << 4), make_uint4(ho[0], ho[1], ho[2], ho[3]));
          }
        }
        {
          float a0, a1, b0, b1;
          upk2(q1, a0, a1);
          upk2(q2, b0, b1);
          s1 = a0 + a1; s2 = b0 + b1;
          xch[(half * 128 + r) * 2] = s1;
          xch[(half * 128 + r) * 2 + 1] = s2;
          ptx::named_bar_sync<2, kRowThreads>();
          s1 += xch[((half ^ 1) * 128 + r) * 2];
          s2 += xch[((half ^ 1) * 128 + r) * 2 + 1];
        }
        ptx::fence_proxy_async();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) arrive_leader(x1_ready_c);
      } else if (row_ok) {   // LayerNorm of this row from the producer's two partial sums
        const float2* st = reinterpret_cast<const float2*>(p.stats) + row * 2;
        const float2 a = __ldg(st), b = __ldg(st + 1);
        s1 = a.x + b.x; s2 = a.y + b.y;
      }
      const float mu = s1 * (1.f / kD);
      const float rstd = rsqrtf(fmaxf(s2 * (1.f / kD) - mu * mu, 0.f) + p.eps);
      const uint64_t ln_a = pk2(rstd, rstd), ln_b = pk2(-rstd * mu, -rstd * mu);
#pragma unroll 1
      for (int c = 0; c < kNC; ++c, ++n_chunk) {
        TT_TIMED(e_a1, ptx::mbar_wait(&ctl->a1_full, n_chunk & 1));
        ptx::tc_fence_after();
        uint32_t raw[2][32];
        ptx::tmem_ld<32>(tl + kAcc1 + half * 64, raw[0]);
        ptx::tmem_ld<32>(tl + kAcc1 + half * 64 + 32, raw[1]);
        ptx::tmem_ld_wait(raw[0]);
        ptx::tmem_ld_wait(raw[1]);
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) arrive_leader(a1_free_c);   // acc1 drained: the next chunk's fc1 MMAs may overwrite it
        uint4 o[8];
        const uint32_t c0_s = vec_s + (c * kChunk + half * 64) * 4, c1_s = c0_s + kMlp * 4;
#pragma unroll
        for (int b = 0; b < 2; ++b) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {   // 8 columns -> one 16-byte unit
            uint32_t w[4];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int col = b * 32 + g * 8 + e * 4;
              const uint4 cc = ptx::lds128(c1_s + col * 4), bb = ptx::lds128(c0_s + col * 4);
              const uint64_t v0 = gelu_fast2(fma2(pk2u(raw[b][g * 8 + e * 4 + 0], raw[b][g * 8 + e * 4 + 1]), ln_a, fma2(ln_b, pk2u(cc.x, cc.y), pk2u(bb.x, bb.y))));
              const uint64_t v1 = gelu_fast2(fma2(pk2u(raw[b][g * 8 + e * 4 + 2], raw[b][g * 8 + e * 4 + 3]), ln_a, fma2(ln_b, pk2u(cc.z, cc.w), pk2u(bb.z, bb.w))));
              float y0, y1, y2, y3;
              upk2(v0, y0, y1);
              upk2(v1, y2, y3);
              w[2 * e] = pack_bf16(y0, y1);
              w[2 * e + 1] = pack_bf16(y2, y3);
            }
            o[b * 4 + g] = make_uint4(w[0], w[1], w[2], w[3]);
          }
        }
        TT_TIMED(e_hfree, ptx::mbar_wait(&ctl->h_free, (n_chunk & 1) ^ 1));   // fc2 of the previous chunk has read sH
        if (c == 0 && n_tile > 0) ptx::mbar_wait(&ctl->fin_done, (n_tile - 1) & 1);   // and the previous tile's lo' stores have
        const uint32_t hrow = ptx::smem_u32(sH) + half * kUnit + r * 128;
#pragma unroll
        for (int j = 0; j < 8; ++j) ptx::sts128(hrow + ((static_cast<uint32_t>(j) ^ sw) << 4), o[j]);
        ptx::fence_proxy_async();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) arrive_leader(h_full_c);
      }
      // ---- x' = acc2 + b2 + hi + lo, written back split; 64-column blocks, this thread 32 columns of each
      TT_TIMED(e_acc2, ptx::mbar_wait(&ctl->acc2_full, it_par));
      ptx::tc_fence_after();
      const long long t_fin = dbg ? clock64() : 0;
      uint64_t ps1 = 0ull, ps2 = 0ull;
      uint32_t rbuf[2][32];
      ptx::tmem_ld<32>(tl + half * 32, rbuf[0]);
#pragma unroll
      for (int j = 0; j < kKB; ++j) {
        uint32_t (&raw)[32] = rbuf[j & 1];
        if constexpr (!PROJ) ptx::mbar_wait(&ctl->lo_full[j], it_par);
        ptx::tmem_ld_wait(raw);
        if (j + 1 < kKB) ptx::tmem_ld<32>(tl + (j + 1) * 64 + half * 32, rbuf[(j + 1) & 1]);   // in flight during this block's math
        const uint32_t xrow = ptx::smem_u32(sX) + j * kUnit + r * 128;
        const uint32_t lrow = ptx::smem_u32(j < 2 ? sH + j * kUnit : ring + (j - 2) * kUnit) + r * 128;
        const uint32_t b2_s = vec_s + (2 * kMlp + j * 64 + half * 32) * 4;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const uint32_t off = ((static_cast<uint32_t>(half * 4 + g) ^ sw) << 4);
          const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);
          const uint4 hv = PROJ ? zero4 : ptx::lds128(xrow + off), lv = PROJ ? zero4 : ptx::lds128(lrow + off);   // PROJ: acc2 already holds x1
          const uint4 b0 = ptx::lds128(b2_s + g * 32), b1 = ptx::lds128(b2_s + g * 32 + 16);
          const uint32_t hw[4] = {hv.x, hv.y, hv.z, hv.w}, lw[4] = {lv.x, lv.y, lv.z, lv.w};
          uint32_t bw[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
          if constexpr (PROJ) {   // + the projection's bias: acc2 holds x + att Wp^T + the MLP update
            const uint4 p0 = ptx::lds128(b2_s + kD * 4 + g * 32), p1 = ptx::lds128(b2_s + kD * 4 + g * 32 + 16);
            const uint32_t pw[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i) bw[i] = __float_as_uint(__uint_as_float(bw[i]) + __uint_as_float(pw[i]));
          }
          uint32_t ho[4], lo[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {   // two columns per 32-bit word: low half = even column
            const uint64_t res = add2(pk2u(hw[e] << 16, hw[e] & 0xffff0000u), pk2u(lw[e] << 16, lw[e] & 0xffff0000u));
            const uint64_t v = add2(add2(pk2u(raw[8 * g + 2 * e], raw[8 * g + 2 * e + 1]), pk2u(bw[2 * e], bw[2 * e + 1])), res);
            ps1 = add2(ps1, v);
            ps2 = fma2(v, v, ps2);
            float x0, x1;
            upk2(v, x0, x1);
            ho[e] = pack_bf16(x0, x1);
            lo[e] = pack_bf16(x0 - __uint_as_float(ho[e] << 16), x1 - __uint_as_float(ho[e] & 0xffff0000u));
          }
          ptx::sts128(xrow + off, make_uint4(ho[0], ho[1], ho[2], ho[3]));
          ptx::sts128(lrow + off, make_uint4(lo[0], lo[1], lo[2], lo[3]));
        }
        ptx::fence_proxy_async();
        ptx::mbar_arrive(&ctl->blk_done[j]);   // this thread's part of block j of hi' / lo' is in smem: the store warp takes it from there
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) arrive_leader(acc2_free_c);   // acc2 read out: the next tile's fc2 MMAs may overwrite it
      if (row_ok) {
        float a0, a1, b0, b1;
        upk2(ps1, a0, a1);
        upk2(ps2, b0, b1);
        reinterpret_cast<float2*>(p.stats)[row * 2 + half] = make_float2(a0 + a1, b0 + b1);
      }
      e_final += dbg ? clock64() - t_fin : 0;
      ++e_tiles;
      ++n_tile;
    }
    if (dbg && lane == 0) { p.dbg[8] = e_a1; p.dbg[9] = e_hfree; p.dbg[10] = e_acc2; p.dbg[11] = e_final; p.dbg[12] = clock64() - e_start; p.dbg[13] = e_tiles; }
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync_all();   // neither CTA exits (or frees TMEM) while its peer still works
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc_pair(tmem_base, kTmemCols);
  }
}

bool make_map(CUtensorMap* m, const void* base, long long rows, int cols, int box_rows) {
  const cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  const cuuint64_t strides[1] = {static_cast<cuuint64_t>(cols) * 2};
  const cuuint32_t box[2] = {64, static_cast<cuuint32_t>(box_rows)};
  return make_tmap_bf16(m, base, 2, dims, strides, box, 128);
}

}  // namespace

bool enc_mlp_supported(int D, int mlp) { return D == kD && mlp == kMlp; }

cudaError_t enc_mlp_forward(const EncMlpWeights& w, __nv_bfloat16* hi, __nv_bfloat16* lo, float* stats, int parts_in,
                            long long M, int D, int mlp, float eps, cudaStream_t s, const EncProj* proj) {
  if (M <= 0) return cudaSuccess;
  if (!enc_mlp_supported(D, mlp) || (proj == nullptr && parts_in != 2)) {
    set_error("enc_mlp: built for PARSeq-base (embed 384, MLP 1536) with two LayerNorm partials per row");
    return cudaErrorInvalidValue;
  }
  if ((reinterpret_cast<uintptr_t>(w.c0) | reinterpret_cast<uintptr_t>(w.c1) | reinterpret_cast<uintptr_t>(w.b2)) & 15) {
    set_error("enc_mlp: per-column vectors must be 16-byte aligned");
    return cudaErrorInvalidValue;
  }
  MlpParams p{};
  if (!make_map(&p.tm_hi, hi, M, D, 128) || !make_map(&p.tm_lo, lo, M, D, 128) || !make_map(&p.tm_w1, w.w1, mlp, D, 64) ||
      !make_map(&p.tm_w2a, w.w2, D, mlp, 128) || !make_map(&p.tm_w2b, w.w2, D, mlp, 64))
    return cudaErrorInvalidValue;
  p.c0 = w.c0; p.c1 = w.c1; p.b2 = w.b2;
  if (proj) {
    if (reinterpret_cast<uintptr_t>(proj->bp) & 15) { set_error("enc_mlp: proj bias must be 16-byte aligned"); return cudaErrorInvalidValue; }
    if (!make_map(&p.tm_att, proj->att, M, D, 128) || !make_map(&p.tm_wpa, proj->wp, D, D, 128) || !make_map(&p.tm_wpb, proj->wp, D, D, 64))
      return cudaErrorInvalidValue;
    p.bp = proj->bp;
    p.hi_g = hi; p.lo_g = lo;
  }
  p.stats = stats;
  p.M = M;
  p.n_tiles = static_cast<int>((M + 255) / 256);
  p.eps = eps;
  static const bool dbg_on = std::getenv("TT_MLP_DEBUG") && std::atoi(std::getenv("TT_MLP_DEBUG")) != 0;
  static unsigned long long* dbg_buf = nullptr;
  if (dbg_on && !dbg_buf) cudaMalloc(&dbg_buf, 16 * sizeof(unsigned long long));
  if (dbg_on) cudaMemsetAsync(dbg_buf, 0, 16 * sizeof(unsigned long long), s);
  p.dbg = dbg_on ? dbg_buf : nullptr;
  const void* fn = proj ? reinterpret_cast<const void*>(k_enc_mlp<true>) : reinterpret_cast<const void*>(k_enc_mlp<false>);
  TT_CUDA_TRY(ensure_dynamic_smem(fn, kSmem));
  // co-resident CTA pairs (one per TPC: 74 on a full B200; all GPUs of a box are the same part), queried once
  static const int max_pairs = [&]() {
    cudaLaunchConfig_t qc{};
    qc.gridDim = dim3(2); qc.blockDim = dim3(kThreads); qc.dynamicSmemBytes = kSmem;
    cudaLaunchAttribute qa[1];
    qa[0].id = cudaLaunchAttributeClusterDimension;
    qa[0].val.clusterDim.x = 2; qa[0].val.clusterDim.y = 1; qa[0].val.clusterDim.z = 1;
    qc.attrs = qa; qc.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, fn, &qc) != cudaSuccess || n <= 0) n = 64;   // fn: its dynamic smem limit is set
    return n;
  }();
  const int pairs = std::min(p.n_tiles, max_pairs);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * pairs);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = kSmem;
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  char tag[96];
  if (prof_enabled()) {
    std::snprintf(tag, sizeof(tag), "mlp M%lld D%d H%d grid%d pair fused %sfc1+gelu+fc2", M, D, mlp, 2 * pairs, proj ? "proj+" : "");
    prof_record(s, true, 0, 0);
  }
  if (proj) TT_CUDA_TRY(cudaLaunchKernelEx(&cfg, k_enc_mlp<true>, p));
  else TT_CUDA_TRY(cudaLaunchKernelEx(&cfg, k_enc_mlp<false>, p));
  prof_record(s, false, 4.0 * static_cast<double>(M) * D * mlp + (proj ? 2.0 * static_cast<double>(M) * D * D : 0.0), 0, tag);   // fc1 + fc2 (+ proj)
  TT_LAUNCH_CHECK();
  if (dbg_on) {
    unsigned long long h[16];
    cudaStreamSynchronize(s);
    cudaMemcpy(h, dbg_buf, sizeof(h), cudaMemcpyDeviceToHost);
    std::fprintf(stderr, "[mlp dbg] M %lld tiles/pair %llu | mma: wait x %llu a1_free %llu weights %llu h_full %llu acc2_free %llu of %llu | "
                 "rows: wait a1_full %llu h_free %llu acc2_full %llu final %llu of %llu\n", M, h[13], h[0], h[1], h[2], h[3], h[4], h[5],
                 h[8], h[9], h[10], h[11], h[12]);
  }
  return cudaSuccess;
}

}  // namespace tt
