// The tensor-core kernel of the OCR path: a persistent, warp-specialised tcgen05 GEMM.
//
//   warp 0      TMA producer   (one lane): per k-block, A tile (128 rows) + W tile (BN rows) -> smem ring
//   warp 1      MMA issuer     (one lane): tcgen05.mma cta_group::1, M=128, N=BN, K=16 per instruction,
//                                          fp32 accumulators in TMEM, double-buffered (2 x 256 columns)
//   warps 2..9  epilogue       (8 warps) : tcgen05.ld -> bias / ReLU / GELU / residual / cls tail -> global
//
// PAIR mode (cta_group::2): two CTAs of a 2-CTA cluster share one 256 x BN tile.  Each CTA stages its own
// 128 rows of A and BN/2 rows of the weights, the even CTA issues tcgen05.mma.cta_group::2 for both, each
// CTA's TMEM holds (and its epilogue drains) its own 128 accumulator rows.  Per-CTA L2->SM operand traffic
// per MMA cycle drops from (128 + BN) to (128 + BN/2) rows, or to the A rows alone when the CTA's weight
// half-slice stays resident in smem -- the K=384 PARSeq GEMMs were bound by exactly that traffic.
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
#include <cstdlib>
#include <mutex>

#include "common.h"
#include "epi_math.cuh"
#include "ptx.cuh"
#include "trace.h"

namespace tt {

namespace {

constexpr int kBlockM = 128;
constexpr int kMaxEpiWarps = 16;    // epilogue warps: 8 (fp32 / residual outputs) or 16 (bf16 outputs), template parameter EW
constexpr int kMaxStages = 8;
constexpr int kAccStride = 256;     // TMEM columns per accumulator stage
constexpr int kTmemCols = 512;
constexpr int kSmemBudget = 227 * 1024;
constexpr int kResidentMax = 64 * 1024;   // largest weight slice kept resident in smem (leaves >= 8 A stages)
constexpr int kResidentMaxPair = 100 * 1024;  // PAIR: half of a 256 x 384 slice (96 KB) + 6 A stages
constexpr int kStagingBytes = kMaxEpiWarps * 2048;  // epilogue transpose tiles, 2 KB per warp (upper bound, used for planning)
// TMA epilogue (fp32 residual GEMMs): ring of [128 rows x 32 fp32] SWIZZLE_128B chunks that the residual is loaded
// into, updated in place by the epilogue warps and stored from
constexpr int kHaloPitch = 16;                               // pixels per staged halo row (tile width 8 + 2, padded: SBO = 2048)
constexpr int kHaloRows = 18;                                // tile height 16 + 2
constexpr int kHaloPixels = kHaloRows * kHaloPitch;           // x 128 B (64 channels) = 36 KB, x 64 B (32 channels) = 18 KB
constexpr int kHaloResidentMax = 82 * 1024;                  // weights that leave room for 3 halo stages in one CTA
constexpr int kResSlots = 4;
constexpr int kResSlotBytes = 128 * 128;    // fp32 chunk [128 rows][32 cols]

struct KParams {
  CUtensorMap tmA[2];
  CUtensorMap tmB;
  CUtensorMap tmR, tmC;  // TMA epilogue: fp32 residual (load) and output (store), box {32 cols, 128 rows}
  CUtensorMap tmP;       // fused 2x2 max-pool: the pooled output, box {32 ch, TW/2, 16/TW, 1}, SWIZZLE_64B
  CUtensorMap tmR2, tmC2;  // split residual stream (RES_SPLIT / OUT_SPLIT): the lo tensors; tmR / tmC are then the hi tensors,
                           // all four bf16 with box {32 cols, 128 rows}, SWIZZLE_64B
  int mode;  // 0 plain rows, 1 conv tiles
  int M, N, BN, BK;
  int kb_src[2];
  int taps, dil;
  int H, W, TH, TW, tiles_x, tiles_y;
  int num_m_tiles, num_n_tiles;  // num_m_tiles counts scheduling units: 128-row tiles, or 256-row pairs in PAIR mode
  int m_tiles_total;             // 128-row tiles that exist (PAIR: the last pair may have only one)
  int pair;
  int halo;        // conv, 3x3 dil 1, Cin 64/128, weights resident: one (TH+2) x 16-pixel halo tile per 64-channel block
                   // feeds all 9 taps (A operand = the staged tile read at a pixel offset); 0 = one TMA box per tap
  int stages;
  int b_resident;  // 1: the CTA's [BN x K] weight slice is loaded once and stays in smem; only A streams
  int debug;       // TT_GEMM_DEBUG (development only): 1 = skip epilogue stores, 2 = skip TMEM loads + math + stores, 8 = legacy epilogue preamble (A/B)
  unsigned long long* dbg_out;  // TT_GEMM_DEBUG & 4: per-role wait/busy cycle counters of CTA 0
  uint32_t* trace;              // TT_TRACE=1 (trace.h): host-mapped progress buffer, nullptr otherwise
  uint32_t serial;              //   launch serial inside it
  int alloc_sync;               // PAIR: cluster barrier before the TMEM allocation (1; 0 = the round-1 order, regression probe only)
  Epilogue epi;
};

struct alignas(16) SmemCtl {
  uint64_t full[kMaxStages];
  uint64_t empty[kMaxStages];
  uint64_t acc_full[2];
  uint64_t acc_empty[2];
  uint64_t b_full;
  uint64_t res_full[kResSlots], res_empty[kResSlots], chunk_done[kResSlots];  // TMA epilogue ring
  uint32_t tmem_base;
  float tail[16 * 16 + 16 + 2 * 16 + 2];
  alignas(16) float bias[2][256];  // the tile's bias row, staged per accumulator stage (broadcast reads in the epilogue)
  alignas(16) float c1[2][256];    // LayerNorm consumer: the tile's c1 row (see Epilogue::ln_c1)
};

// Tile schedule shared by the three warp roles.  Streaming mode: tiles round-robin over CTAs with the
// N index fastest (neighbouring CTAs share the A tile in L2).  Weight-resident mode: a CTA owns one
// N slice for its whole life and walks M tiles with stride `groups`.
struct TileIter {
  int m, n, step, limit;
  bool resident;
  __device__ TileIter(const KParams& p) {
    resident = p.b_resident != 0;
    const int cl = p.pair ? blockIdx.x >> 1 : blockIdx.x;   // scheduling slot: CTA, or CTA pair
    const int ncl = p.pair ? gridDim.x >> 1 : gridDim.x;
    if (resident) {
      const int groups = ncl / p.num_n_tiles;
      n = cl % p.num_n_tiles;
      m = cl / p.num_n_tiles;
      step = groups;
      limit = p.num_m_tiles;
    } else {
      m = cl;  // linear tile index in this mode
      n = 0;
      step = ncl;
      limit = p.num_m_tiles * p.num_n_tiles;
    }
  }
  __device__ bool valid() const { return m < limit; }
  __device__ void next() { m += step; }
  __device__ int m_tile(const KParams& p) const { return resident ? m : m / p.num_n_tiles; }
  __device__ int n_tile(const KParams& p) const { return resident ? n : m % p.num_n_tiles; }
};

// role counters (TT_GEMM_DEBUG & 4) read the clock only when they are on
__device__ __forceinline__ long long dbg_clock(bool on) { return on ? clock64() : 0; }

// OUT: OUT_BF16 / OUT_F32 / OUT_CLS_TAIL; ACT: ACT_*; RES: fp32 residual added (OUT_F32 only).  The epilogue is
// specialised at compile time: with these as runtime flags only ~1/4 of its executed instructions were
// useful work (ncu opcode histogram, profiles/r1b_gemm_roles.md) and it, not the MMA, set the pace.
// TE ("TMA epilogue", fp32 output + fp32 residual only): warp 2 streams the residual tile through a smem ring with
// TMA loads that run ahead of the MMAs (the register path had one 2 KB segment per warp in flight and the
// epilogue took 2.6x the mainloop: profiles/r1b_gemm_roles.md), 4 epilogue warps add accumulator + bias to their
// own rows in place, warp 3 writes the chunk back with a TMA store.  No per-thread global access at all.
// TS ("TMA store", bf16 outputs): an epilogue warp writes its 32 rows x 64 B of a segment into a warp-private
// SWIZZLE_64B smem tile and one lane hands it to a TMA store -- no read-back, no per-thread global stores (the LSU
// store path cost ~1000 cycles per 128 x 256 tile, profiles/r1b_gemm_roles.md).  Image edges / short last tiles
// are clipped by the TMA unit.
// LN: LayerNorm fused away (Epilogue::ln_*).  With TE the kernel is the producer (row statistics + bf16 copy of its
// output), with bf16 outputs it is the consumer (normalisation applied algebraically to the accumulators).
// POOL (TS conv epilogue only): 1 = write the 2x2 max-pooled tile instead of the tile, 2 = both.
template <int OUT, int ACT, bool RES, bool PAIR, int EW, bool TE = false, bool TS = false, bool LN = false, int POOL = 0>
__global__ void __launch_bounds__(64 + 32 * EW + (TE ? 64 : 0), 1) gemm_tc_kernel(const __grid_constant__ KParams p) {
  static_assert(POOL == 0 || (TS && EW == 8 && !LN), "fused max-pool: bf16 TMA-store epilogue with 8 warps");
  static_assert(!TE || (OUT == OUT_F32 && RES && EW == 4), "TMA epilogue: fp32 out + residual, 4 epilogue warps");
  static_assert(!LN || TE || OUT == OUT_BF16, "LayerNorm fusion: TMA-epilogue producer or bf16-output consumer");
  constexpr int kSlot = kResSlotBytes;   // stride of the TMA-epilogue ring
  static_assert(!TS || (OUT == OUT_BF16 && !RES && !TE), "TMA store epilogue: bf16 outputs");
  constexpr int kTsBufs = EW == 16 ? 1 : 2;         // 2 KB store tiles per warp (ping-pong with 8 warps)
  constexpr int kThreads = 64 + 32 * EW + (TE ? 64 : 0);
  constexpr int kEpiWarp0 = TE ? 4 : 2;   // first epilogue warp (a multiple of 4 apart from 2: TMEM quadrant = warp & 3)
  constexpr int kEpiWarps = EW;
  constexpr int kParts = EW / 4;   // warps sharing a TMEM lane quadrant split the tile's columns
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B tiles need 1024-byte alignment
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int row_bytes = p.BK * 2;
  const int a_bytes = kBlockM * row_bytes;
  const int b_bytes = (PAIR ? p.BN / 2 : p.BN) * row_bytes;   // this CTA's share of the weight tile
  // warp-uniform by construction (shfl from lane 0): the producer / MMA warps must stay in the uniform datapath
  const uint32_t rank = PAIR ? __shfl_sync(0xffffffffu, ptx::cluster_ctarank(), 0) : 0;    // 0 = leader (issues the MMAs)
  const int kb_per_tap = p.kb_src[0] + p.kb_src[1];
  const int num_kb = p.taps * kb_per_tap;
  const bool resident = p.b_resident != 0;
  const int stage_bytes = p.halo ? kHaloPixels * row_bytes : resident ? a_bytes : a_bytes + b_bytes;
  uint8_t* sBres = smem;                                        // [num_kb][BN rows] when resident
  uint8_t* ring = smem + (resident ? num_kb * b_bytes : 0);
  uint8_t* res_ring = ring + p.stages * stage_bytes;            // TE only
  SmemCtl* ctl = reinterpret_cast<SmemCtl*>(res_ring + (TE ? kResSlots * kSlot : TS ? EW * kTsBufs * 2048 + (POOL ? EW * kTsBufs * 512 : 0) : 0));

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const bool dbgt = (p.debug & 4) != 0;

  if (threadIdx.x == 0) {
    trace_mark(p.trace, p.serial, blockIdx.x, TR_ENTER);
    ptx::prefetch_tmap(&p.tmA[0]);
    ptx::prefetch_tmap(&p.tmB);
    if (p.kb_src[1] > 0) ptx::prefetch_tmap(&p.tmA[1]);
    for (int s = 0; s < p.stages; ++s) {
      ptx::mbar_init(&ctl->full[s], 1);
      ptx::mbar_init(&ctl->empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(&ctl->acc_full[s], 1);
      ptx::mbar_init(&ctl->acc_empty[s], PAIR ? 2 * kEpiWarps : kEpiWarps);  // PAIR: both CTAs' epilogues free the leader's
    }
    ptx::mbar_init(&ctl->b_full, 1);
    if constexpr (TE) {
      ptx::prefetch_tmap(&p.tmR);
      ptx::prefetch_tmap(&p.tmC);
      if constexpr (LN) { ptx::prefetch_tmap(&p.tmR2); ptx::prefetch_tmap(&p.tmC2); }
      for (int s = 0; s < kResSlots; ++s) {
        ptx::mbar_init(&ctl->res_full[s], 1);
        ptx::mbar_init(&ctl->res_empty[s], 1);
        ptx::mbar_init(&ctl->chunk_done[s], 32 * EW);
      }
    }
    ptx::fence_barrier_init();
  }
  // PAIR: tcgen05.alloc.cta_group::2 is a collective of the two CTAs.  Issued while the peer CTA has not started on its
  // SM yet, the peer's own alloc never returns (no trap, the cluster hangs in its first barrier).  A cluster's CTAs
  // are co-scheduled but do not start in the same cycle, and next to kernels of another stream the skew gets large
  // enough to hit this: root cause of the round-1 "two slots" hang (profiles/r2_hang_root_cause.md).
  // So: cluster barrier FIRST (it also publishes the peer's mbarrier initialisation), then the allocation.
  // TT_PAIR_ALLOC_SYNC=0 restores the old order (alloc, then the barrier) for the regression probe.
  if constexpr (PAIR) {
    if (p.alloc_sync) ptx::cluster_sync_all();
  }
  if (warp == 1) {
    if constexpr (PAIR) ptx::tmem_alloc_pair(&ctl->tmem_base, kTmemCols);
    else ptx::tmem_alloc(&ctl->tmem_base, kTmemCols);
    if (lane == 0) { trace_role_done(p.trace, p.serial, blockIdx.x, kTraceAllocByte); trace_tmem_event(p.trace, p.serial, blockIdx.x, 1); }
  }
  if constexpr (OUT == OUT_CLS_TAIL) {
    for (int i = threadIdx.x; i < 16 * 16 + 16 + 2 * 16 + 2; i += kThreads) ctl->tail[i] = p.epi.tail[i];
  }
  ptx::tc_fence_before();
  __syncthreads();
  if constexpr (PAIR) {
    if (!p.alloc_sync) ptx::cluster_sync_all();
  }
  ptx::tc_fence_after();
  if (threadIdx.x == 0) trace_mark(p.trace, p.serial, blockIdx.x, TR_SYNC0);
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, ctl->tmem_base, 0);

  if (warp == 0) {
    // ------------------------------------------------------------- TMA producer
    {  // all 32 lanes run the loops converged; one elected lane issues (see ptx.cuh "_e")
      TileIter it(p);
      const int b_rows = PAIR ? p.BN / 2 : p.BN;
      const int b_row0 = static_cast<int>(rank) * b_rows;           // this CTA's rows inside the BN-row weight tile
      const uint32_t tx_mult = PAIR ? 2u : 1u;                      // the leader's barrier counts both CTAs' bytes
      uint32_t bfull_c = 0, full0_c = 0;                            // PAIR: the leader's barriers as shared::cluster addresses
      if constexpr (PAIR) {
        bfull_c = __shfl_sync(0xffffffffu, ptx::mapa(ptx::smem_u32(&ctl->b_full), 0), 0);
        full0_c = __shfl_sync(0xffffffffu, ptx::mapa(ptx::smem_u32(&ctl->full[0]), 0), 0);
      }
      if (resident && it.valid()) {
        if (rank == 0) ptx::mbar_arrive_expect_tx_e(&ctl->b_full, tx_mult * static_cast<uint32_t>(num_kb * b_bytes));
        for (int kb = 0; kb < num_kb; ++kb) {
          if constexpr (PAIR) ptx::tma_load_2d_pair_e(sBres + kb * b_bytes, &p.tmB, bfull_c, kb * p.BK, it.n_tile(p) * p.BN + b_row0);
          else ptx::tma_load_2d_e(sBres + kb * b_bytes, &p.tmB, &ctl->b_full, kb * p.BK, it.n_tile(p) * p.BN);
        }
      }
      int stage = 0;
      uint32_t phase = 0;
      long long dbg_prod_wait = 0;
      const long long dbg_t_start = dbg_clock(dbgt);
      for (; it.valid(); it.next()) {
        const int n_tile = it.n_tile(p), m_tile = PAIR ? it.m_tile(p) * 2 + static_cast<int>(rank) : it.m_tile(p);
        int img = 0, y0 = 0, x0 = 0;
        if (p.mode == 1) {
          const int per_img = p.tiles_x * p.tiles_y;
          img = m_tile / per_img;
          const int t = m_tile - img * per_img;
          y0 = (t / p.tiles_x) * p.TH;
          x0 = (t % p.tiles_x) * p.TW;
        }
        if (p.halo) {
          // one halo block per 64 input channels: pixels (x0-1 .. x0+14, y0-1 .. y0+16), out-of-image = zero padding
          for (int cb = 0; cb < p.kb_src[0]; ++cb) {
            ptx::mbar_wait(&ctl->empty[stage], phase ^ 1); __syncwarp();
            uint8_t* sA = ring + stage * stage_bytes;
            if (rank == 0) ptx::mbar_arrive_expect_tx_e(&ctl->full[stage], tx_mult * static_cast<uint32_t>(stage_bytes));
            if constexpr (PAIR) ptx::tma_load_4d_pair_e(sA, &p.tmA[0], full0_c + stage * 8, cb * p.BK, x0 - 1, y0 - 1, img);
            else ptx::tma_load_4d_e(sA, &p.tmA[0], &ctl->full[stage], cb * p.BK, x0 - 1, y0 - 1, img);
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
          }
          continue;
        }
        int kb = 0;
        for (int tap = 0; tap < p.taps; ++tap) {
          const int dy = (p.taps == 9) ? (tap / 3 - 1) * p.dil : 0;
          const int dx = (p.taps == 9) ? (tap % 3 - 1) * p.dil : 0;
          for (int src = 0; src < 2; ++src) {
            for (int cb = 0; cb < p.kb_src[src]; ++cb, ++kb) {
              { const long long t0 = dbg_clock(dbgt); ptx::mbar_wait(&ctl->empty[stage], phase ^ 1); __syncwarp(); dbg_prod_wait += dbg_clock(dbgt) - t0; }
              uint8_t* sA = ring + stage * stage_bytes;
              if (rank == 0) ptx::mbar_arrive_expect_tx_e(&ctl->full[stage], tx_mult * static_cast<uint32_t>(stage_bytes));
              if constexpr (PAIR) {
                const uint32_t full_bar = full0_c + stage * 8;
                // a pair's second tile may not exist (odd tile count): its box is out of bounds and arrives as zeros
                if (p.mode == 1)
                  ptx::tma_load_4d_pair_e(sA, &p.tmA[src], full_bar, cb * p.BK, x0 + dx, y0 + dy, img);
                else
                  ptx::tma_load_2d_pair_e(sA, &p.tmA[src], full_bar, cb * p.BK, m_tile * kBlockM);
                if (!resident) ptx::tma_load_2d_pair_e(sA + a_bytes, &p.tmB, full_bar, kb * p.BK, n_tile * p.BN + b_row0);
              } else {
                if (p.mode == 1)
                  ptx::tma_load_4d_e(sA, &p.tmA[src], &ctl->full[stage], cb * p.BK, x0 + dx, y0 + dy, img);
                else
                  ptx::tma_load_2d_e(sA, &p.tmA[src], &ctl->full[stage], cb * p.BK, m_tile * kBlockM);
                if (!resident) ptx::tma_load_2d_e(sA + a_bytes, &p.tmB, &ctl->full[stage], kb * p.BK, n_tile * p.BN);
              }
              if (++stage == p.stages) { stage = 0; phase ^= 1; }
            }
          }
        }
      }
      if constexpr (PAIR) {
        // drain: the leader's last commits still arrive on this CTA's empty barriers; do not exit before they landed
        for (int i = 0; i < p.stages; ++i) {
          ptx::mbar_wait(&ctl->empty[stage], phase ^ 1); __syncwarp();
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
      if ((p.debug & 4) && blockIdx.x == 0 && lane == 0) { p.dbg_out[0] = dbg_prod_wait; p.dbg_out[1] = dbg_clock(dbgt) - dbg_t_start; }
    }
  } else if (warp == 1) {
    // --------------------------------------------------------------- MMA issuer
    if (rank == 0) {  // all 32 lanes converged, one elected lane issues
      const uint32_t idesc = ptx::make_idesc_bf16(PAIR ? 2 * kBlockM : kBlockM, p.BN);
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      TileIter it(p);
      if (resident && it.valid()) ptx::mbar_wait(&ctl->b_full, 0);
      const int ksteps = p.BK / 16;
      long long dbg_w_acc = 0, dbg_w_full = 0;
      const long long dbg_t_start = dbg_clock(dbgt);
      for (; it.valid(); it.next()) {
        { const long long t0 = dbg_clock(dbgt); ptx::mbar_wait(&ctl->acc_empty[as], aphase ^ 1); dbg_w_acc += dbg_clock(dbgt) - t0; }
        __syncwarp(); ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * kAccStride;
        if (p.halo) {
          for (int cb = 0; cb < p.kb_src[0]; ++cb) {
            { const long long t0 = dbg_clock(dbgt); ptx::mbar_wait(&ctl->full[stage], phase); dbg_w_full += dbg_clock(dbgt) - t0; }
            __syncwarp(); ptx::tc_fence_after();
            const uint32_t h_addr = ptx::smem_u32(ring + stage * stage_bytes);
            for (int tap = 0; tap < 9; ++tap) {
              const int ty = tap / 3, tx = tap - 3 * ty;
              const uint32_t a_addr = h_addr + (ty * kHaloPitch + tx) * row_bytes;   // output pixel (y, x) reads halo pixel (y + ty, x + tx)
              const uint32_t b_addr = ptx::smem_u32(sBres + (tap * p.kb_src[0] + cb) * b_bytes);
              for (int k = 0; k < ksteps; ++k) {
                const uint64_t da = ptx::make_smem_desc_sbo(a_addr + k * 32, row_bytes, kHaloPitch * row_bytes);
                const uint64_t db = ptx::make_smem_desc(b_addr + k * 32, row_bytes);
                if constexpr (PAIR) ptx::mma_bf16_pair_e(d_tmem, da, db, idesc, (cb | tap | k) != 0);
                else ptx::mma_bf16_e(d_tmem, da, db, idesc, (cb | tap | k) != 0);
              }
            }
            if constexpr (PAIR) ptx::mma_commit_pair_e(&ctl->empty[stage], 3);
            else ptx::mma_commit_e(&ctl->empty[stage]);
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
          }
          if constexpr (PAIR) ptx::mma_commit_pair_e(&ctl->acc_full[as], 3);
          else ptx::mma_commit_e(&ctl->acc_full[as]);
          if (++as == 2) { as = 0; aphase ^= 1; }
          continue;
        }
        for (int kb = 0; kb < num_kb; ++kb) {
          { const long long t0 = dbg_clock(dbgt); ptx::mbar_wait(&ctl->full[stage], phase); dbg_w_full += dbg_clock(dbgt) - t0; }
          __syncwarp(); ptx::tc_fence_after();
          const uint32_t a_addr = ptx::smem_u32(ring + stage * stage_bytes);
          const uint32_t b_addr = resident ? ptx::smem_u32(sBres + kb * b_bytes) : a_addr + a_bytes;
          for (int k = 0; k < ksteps; ++k) {
            const uint64_t da = ptx::make_smem_desc(a_addr + k * 32, row_bytes);
            const uint64_t db = ptx::make_smem_desc(b_addr + k * 32, row_bytes);
            if constexpr (PAIR) ptx::mma_bf16_pair_e(d_tmem, da, db, idesc, (kb | k) != 0);
            else ptx::mma_bf16_e(d_tmem, da, db, idesc, (kb | k) != 0);
          }
          if constexpr (PAIR) ptx::mma_commit_pair_e(&ctl->empty[stage], 3);   // frees the stage in both CTAs
          else ptx::mma_commit_e(&ctl->empty[stage]);
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        if constexpr (PAIR) ptx::mma_commit_pair_e(&ctl->acc_full[as], 3);      // both CTAs' epilogues may drain
        else ptx::mma_commit_e(&ctl->acc_full[as]);
        if (++as == 2) { as = 0; aphase ^= 1; }
      }
      if ((p.debug & 4) && blockIdx.x == 0 && lane == 0) { p.dbg_out[2] = dbg_w_acc; p.dbg_out[3] = dbg_w_full; p.dbg_out[4] = dbg_clock(dbgt) - dbg_t_start; }
    }
  } else if (TE && warp == 2) {
    // ------------------------------------------------- residual loader (TMA epilogue)
    {
      int slot = 0;
      uint32_t ph = 0;
      for (TileIter it(p); it.valid(); it.next()) {
        const int m_tile = PAIR ? it.m_tile(p) * 2 + static_cast<int>(rank) : it.m_tile(p);
        const int n0 = it.n_tile(p) * p.BN, m0 = m_tile * kBlockM;
        const int rrow = p.epi.res_mod > 0 ? m0 % p.epi.res_mod : m0;
        for (int c = 0; c < p.BN / 32; ++c) {
          ptx::mbar_wait(&ctl->res_empty[slot], ph ^ 1); __syncwarp();
          ptx::mbar_arrive_expect_tx_e(&ctl->res_full[slot], kResSlotBytes);
          // a pair's second tile may not exist, columns may end before the tile does: out-of-bounds parts arrive as zeros
          ptx::tma_load_2d_e(res_ring + slot * kSlot, &p.tmR, &ctl->res_full[slot], n0 + c * 32, rrow);
          if constexpr (LN)   // split residual: hi tile [128 rows x 64 B] then lo tile, 8 KB each
            ptx::tma_load_2d_e(res_ring + slot * kSlot + kResSlotBytes / 2, &p.tmR2, &ctl->res_full[slot], n0 + c * 32, rrow);
          if (++slot == kResSlots) { slot = 0; ph ^= 1; }
        }
      }
    }
  } else if (TE && warp == 3) {
    // --------------------------------------------------- output storer (TMA epilogue)
    {
      int slot = 0, prev = -1;
      uint32_t ph = 0;
      for (TileIter it(p); it.valid(); it.next()) {
        const int m_tile = PAIR ? it.m_tile(p) * 2 + static_cast<int>(rank) : it.m_tile(p);
        const int n0 = it.n_tile(p) * p.BN, m0 = m_tile * kBlockM;
        for (int c = 0; c < p.BN / 32; ++c) {
          ptx::mbar_wait(&ctl->chunk_done[slot], ph); __syncwarp();
          if ((p.debug & 3) == 0) {
            ptx::tma_store_2d_e(&p.tmC, res_ring + slot * kSlot, n0 + c * 32, m0);  // rows/cols past the tensor are clipped
            if constexpr (LN) ptx::tma_store_2d_e(&p.tmC2, res_ring + slot * kSlot + kResSlotBytes / 2, n0 + c * 32, m0);
          }
          ptx::bulk_commit_e();
          if (prev >= 0) {
            if (ptx::elect_one()) ptx::bulk_wait_read<1>();   // the previous chunk's store has finished reading its slot
            __syncwarp();
            ptx::mbar_arrive_e(&ctl->res_empty[prev]);
          }
          prev = slot;
          if (++slot == kResSlots) { slot = 0; ph ^= 1; }
        }
      }
      if (ptx::elect_one()) ptx::bulk_wait<0>();
      __syncwarp();
    }
  } else if (warp >= kEpiWarp0) {
    // ----------------------------------------------------------------- epilogue
    // Thread t of a warp owns accumulator row (TMEM lane) q*32 + t, but a row-per-thread global store
    // touches 32 different cache lines per instruction (measured: it halved the kernel's throughput).
    // So every 64-byte row segment goes through a warp-private 2 KB smem tile (32 rows x 64 B, 16-byte
    // units XOR-swizzled): rows are written/read by their owner thread, global memory is accessed with
    // 4 lanes per row segment, 8 rows per instruction.  Residual reads take the same path backwards.
    constexpr bool kF32 = OUT == OUT_F32;
    constexpr int kEsize = kF32 ? 4 : 2;
    constexpr int kSegChunks = kF32 ? 1 : 2;     // 16-column chunks per 64-byte output segment
    const int q = warp & 3;                      // TMEM lane quadrant this warp may read
    const int part = (warp - kEpiWarp0) >> 2;    // kParts warps share a quadrant and split the tile's columns
    const int r = q * 32 + lane;                 // tile row == TMEM lane
    const uint32_t stg = TS ? ptx::smem_u32(res_ring) + (warp - kEpiWarp0) * kTsBufs * 2048
                            : ptx::smem_u32(reinterpret_cast<uint8_t*>(ctl + 1)) + (warp - kEpiWarp0) * 2048;
    int ts_buf = 0;
    (void)ts_buf;
    // fused 2x2 max-pool: a warp's 32 pixels are 32/TW rows of TW pixels; partners are lane ^ 1 (x) and lane ^ TW (y),
    // the even/even lane writes pooled pixel j of the warp's 8 into a 512-byte SWIZZLE_64B tile
    const int pool_tw = p.TW;
    const bool pool_writer = POOL != 0 && (lane & 1) == 0 && (lane & pool_tw) == 0;
    const int pool_j = pool_tw == 8 ? ((lane >> 4) * 4 + ((lane & 7) >> 1)) : ((lane & 15) >> 1);
    const uint32_t pool_stg = POOL != 0 ? ptx::smem_u32(res_ring) + EW * kTsBufs * 2048 + (warp - kEpiWarp0) * kTsBufs * 512 : 0u;
    const uint32_t pool_row = pool_stg + pool_j * 64;
    const int pool_sw = (pool_j >> 1) & 3;
    (void)pool_writer; (void)pool_row; (void)pool_sw;
    const uint32_t bias_s = ptx::smem_u32(&ctl->bias[0][0]);
    const uint32_t own_row = stg + lane * 64;
    const int own_sw = (lane >> 1) & 3;          // swizzle of this thread's own row
    const int co_row = lane >> 2, co_unit = lane & 3;  // cooperative access: 4 lanes per row, 8 rows per pass
    uint32_t co_addr[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int row = j * 8 + co_row;
      co_addr[j] = stg + row * 64 + ((co_unit ^ ((row >> 1) & 3)) << 4);
    }
    // hoisted parameters
    const int BN = p.BN, N = p.N, M = p.M, mode = p.mode, dbg = p.debug & 3;
    const int H = p.H, W = p.W, TH = p.TH, TW = p.TW, tiles_x = p.tiles_x, per_img = p.tiles_x * p.tiles_y;
    const long long ldc = p.epi.ldc, ldr = p.epi.ldr;
    const int res_mod = p.epi.res_mod;
    char* const out_base = static_cast<char*>(p.epi.out);
    const char* const res_base = static_cast<const char*>(p.epi.residual);
    const float* const bias_g = p.epi.bias;
    const int chunks = BN / 16;
    // TS splits the columns in whole 32-column segments (a TMA store box), the register path in 16-column chunks
    const int c_begin = TS ? 2 * ((chunks / 2) * part / kParts) : (chunks * part + kParts - 1) / kParts;
    const int c_end = TS ? 2 * ((chunks / 2) * (part + 1) / kParts) : (chunks * (part + 1) + kParts - 1) / kParts;
    int as = 0;
    uint32_t aphase = 0;
    long long dbg_e_wait = 0, dbg_e_busy = 0, dbg_tiles = 0;
    const int m_tiles_total = p.m_tiles_total;
    const uint32_t acc_empty_leader = PAIR ? ptx::mapa(ptx::smem_u32(&ctl->acc_empty[0]), 0) : 0;
    auto row_to_out = [&](int m_tile, int rr, long long& orow) -> bool {
      if (PAIR && m_tile >= m_tiles_total) { orow = 0; return false; }
      if (mode == 1) {
        const int img = m_tile / per_img;
        const int t = m_tile - img * per_img;
        const int y = (t / tiles_x) * TH + rr / TW;
        const int x = (t % tiles_x) * TW + rr % TW;
        orow = (static_cast<long long>(img) * H + y) * W + x;
        return (y < H) && (x < W);
      }
      orow = static_cast<long long>(m_tile) * kBlockM + rr;
      return orow < M;
    };
    // bias of a tile's columns -> smem (a per-chunk LDG of it was 44% of all stall samples); the value for the
    // NEXT tile is fetched while this one is processed (its LDG latency was 9% of the samples after that).
    const int bias_t = threadIdx.x - 32 * kEpiWarp0;  // 0 .. 32*EW-1 over the epilogue warps
    constexpr int kBiasPer = (256 + 32 * EW - 1) / (32 * EW);  // bias columns per epilogue thread (BN <= 256)
    constexpr bool kLNA = LN && !TE;   // LayerNorm consumer
    const float* const c1_g = p.epi.ln_c1;
    auto bias_fetch = [&](const TileIter& t, float (&b)[kBiasPer], float (&c)[kBiasPer]) {
#pragma unroll
      for (int i = 0; i < kBiasPer; ++i) {
        const int tc = bias_t + i * 32 * EW, col = t.n_tile(p) * BN + tc;
        const bool ok = t.valid() && tc < BN && col < N;
        b[i] = (ok && bias_g != nullptr) ? __ldg(bias_g + col) : 0.f;
        if constexpr (kLNA) c[i] = ok ? __ldg(c1_g + col) : 0.f;
      }
    };
    TileIter it(p);
    float bias_cur[kBiasPer], c1_cur[kBiasPer];
    bias_fetch(it, bias_cur, c1_cur);
    // LayerNorm consumer: the producer's (sum, sum of squares) partials of this thread's row in tile t.  They are requested
    // one tile ahead and only LOADED here -- the sums are formed where they are used, one tile later: an add right behind
    // the load parks the (in-order) warp on the L2 round trip before it reaches the accumulator wait, every tile.
    constexpr int kLnMaxParts = 4;   // the producer has <= 4 N tiles (launch())
    float2 ln_raw[kLnMaxParts];
#pragma unroll
    for (int pp = 0; pp < kLnMaxParts; ++pp) ln_raw[pp] = make_float2(0.f, 0.f);
    auto ln_fetch = [&](const TileIter& t) {
      if constexpr (kLNA) {
        bool ok = false;
        const float2* st = nullptr;
        if (t.valid()) {
          const int mt = PAIR ? t.m_tile(p) * 2 + static_cast<int>(rank) : t.m_tile(p);
          const long long mrow = static_cast<long long>(mt) * kBlockM + r;
          ok = mrow < M && (!PAIR || mt < p.m_tiles_total);
          st = reinterpret_cast<const float2*>(p.epi.ln_stats_in) + mrow * p.epi.ln_parts;
        }
#pragma unroll
        for (int pp = 0; pp < kLnMaxParts; ++pp)
          ln_raw[pp] = (ok && pp < p.epi.ln_parts) ? __ldg(st + pp) : make_float2(0.f, 0.f);
        if (p.debug & 8) asm volatile("" ::"f"(ln_raw[0].x), "f"(ln_raw[1].x));   // A/B: wait for the loads here, as the eager sums did
      }
    };
    ln_fetch(it);
    (void)ln_raw;
    const uint32_t c1_s = ptx::smem_u32(&ctl->c1[0][0]);
    (void)c1_s;
    int te_slot = 0;
    uint32_t te_ph = 0;
    (void)te_slot; (void)te_ph;
    int staged_n0 = -1, staged_n1 = -1;   // N tile whose bias / c1 vectors sit in smem slot 0 / 1
    for (; it.valid(); it.next()) {
      const int n_tile = it.n_tile(p), m_tile = PAIR ? it.m_tile(p) * 2 + static_cast<int>(rank) : it.m_tile(p);
      const int n0 = n_tile * BN;
      {
        // A weight-resident CTA keeps one N slice for its whole life: after the first two tiles both slots already hold
        // this slice's vectors, and the staging + the barrier over all epilogue warps (a rendezvous per tile) are skipped.
        // The branch is uniform over the CTA: every warp walks the same tile sequence.
        if ((as ? staged_n1 : staged_n0) != n_tile || (p.debug & 8)) {
#pragma unroll
          for (int i = 0; i < kBiasPer; ++i)
            if (bias_t + i * 32 * EW < BN) {
              ptx::sts32(bias_s + (as * 256 + bias_t + i * 32 * EW) * 4, __float_as_uint(bias_cur[i]));
              if constexpr (kLNA) ptx::sts32(c1_s + (as * 256 + bias_t + i * 32 * EW) * 4, __float_as_uint(c1_cur[i]));
            }
          asm volatile("bar.sync 1, %0;" ::"n"(32 * EW) : "memory");  // epilogue warps only
          if (as) staged_n1 = n_tile; else staged_n0 = n_tile;
        }
        TileIter nx = it;
        nx.next();
        if (nx.valid() && nx.n_tile(p) != (as ? staged_n0 : staged_n1)) bias_fetch(nx, bias_cur, c1_cur);
      }
      uint64_t ln_a = 0ull, ln_b = 0ull;   // LayerNorm consumer: (rstd, rstd) and (-rstd * mean, -rstd * mean) of this thread's row
      (void)ln_a; (void)ln_b;
      if constexpr (kLNA) {
        // this tile's sums were requested one tile ago (ln_s1n / ln_s2n): their latency hides behind the previous tile
        const float inv_d = 1.f / static_cast<float>(p.epi.ln_dim);
        float ln_s1n = 0.f, ln_s2n = 0.f;   // same summation order as before: partial 0, 1, ...
#pragma unroll
        for (int pp = 0; pp < kLnMaxParts; ++pp) { ln_s1n += ln_raw[pp].x; ln_s2n += ln_raw[pp].y; }
        const float mu = ln_s1n * inv_d;
        const float rstd = rsqrtf(fmaxf(ln_s2n * inv_d - mu * mu, 0.f) + p.epi.ln_eps);
        ln_a = pk2(rstd, rstd);
        ln_b = pk2(-rstd * mu, -rstd * mu);
        TileIter nx2 = it;
        nx2.next();
        ln_fetch(nx2);
      }
      if constexpr (TE) {
        // out = residual + (acc + bias), chunk by chunk: this thread's row of the chunk sits at r*128 in the slot,
        // its 16-byte units XOR-swizzled by (r & 7) (SWIZZLE_128B) -- the same layout the TMA store reads back.
        { const long long t0 = dbg_clock(dbgt); ptx::mbar_wait(&ctl->acc_full[as], aphase); dbg_e_wait += dbg_clock(dbgt) - t0; }
        const long long dbg_t_busy0 = dbg_clock(dbgt);
        ptx::tc_fence_after();
        const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * kAccStride;
        const uint32_t bias_row = bias_s + as * 1024;
        const uint32_t ring_s = ptx::smem_u32(res_ring);
        uint64_t ln_s1 = 0ull, ln_s2 = 0ull;   // packed (even, odd column) partial sums of x and x^2 over this tile's columns
        (void)ln_s1; (void)ln_s2;
        const long long ln_mrow = static_cast<long long>(m_tile) * kBlockM + r;
        const bool ln_row_ok = LN && ln_mrow < M && (!PAIR || m_tile < m_tiles_total);
        (void)ln_row_ok;
        for (int c = 0; c < BN / 32; ++c) {
          uint32_t raw[32];
          ptx::tmem_ld<32>(t_row + c * 32, raw);
          ptx::mbar_wait(&ctl->res_full[te_slot], te_ph);
          ptx::tmem_ld_wait(raw);
          if (dbg != 2) {
            if constexpr (LN) {
              // split residual stream: this thread's row of the hi tile at r*64, of the lo tile 8 KB behind it, 16-byte
              // units (8 bf16) XOR-swizzled by (r >> 1) & 3 (SWIZZLE_64B).  x = hi + lo + acc + bias in fp32; written back
              // as hi' = bf16(x), lo' = bf16(x - hi') in place -- hi' is the next GEMM's A operand.
              const uint32_t hrow = ring_s + te_slot * kSlot + r * 64;
              const int sw = (r >> 1) & 3;
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const uint32_t a = hrow + ((j ^ sw) << 4);
                const uint4 h = ptx::lds128(a), l = ptx::lds128(a + kResSlotBytes / 2);
                const uint4 b0 = ptx::lds128(bias_row + (c * 32 + 8 * j) * 4), b1 = ptx::lds128(bias_row + (c * 32 + 8 * j + 4) * 4);
                const uint32_t hw[4] = {h.x, h.y, h.z, h.w}, lw[4] = {l.x, l.y, l.z, l.w};
                const uint32_t bw[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
                uint32_t ho[4], lo[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {   // two columns per 32-bit word: low half = even column
                  const uint64_t res = add2(pk2u(hw[e] << 16, hw[e] & 0xffff0000u), pk2u(lw[e] << 16, lw[e] & 0xffff0000u));
                  const uint64_t v = add2(add2(pk2u(raw[8 * j + 2 * e], raw[8 * j + 2 * e + 1]), pk2u(bw[2 * e], bw[2 * e + 1])), res);
                  ln_s1 = add2(ln_s1, v);
                  ln_s2 = fma2(v, v, ln_s2);
                  float x0, x1;
                  upk2(v, x0, x1);
                  ho[e] = pack_bf16(x0, x1);
                  lo[e] = pack_bf16(x0 - __uint_as_float(ho[e] << 16), x1 - __uint_as_float(ho[e] & 0xffff0000u));
                }
                ptx::sts128(a, make_uint4(ho[0], ho[1], ho[2], ho[3]));
                ptx::sts128(a + kResSlotBytes / 2, make_uint4(lo[0], lo[1], lo[2], lo[3]));
              }
            } else {
              const uint32_t rowaddr = ring_s + te_slot * kSlot + r * 128;
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                const uint32_t a = rowaddr + ((k ^ (r & 7)) << 4);
                const uint4 u = ptx::lds128(a);
                const uint4 b = ptx::lds128(bias_row + (c * 32 + 4 * k) * 4);
                const uint64_t v0 = add2(add2(pk2u(raw[4 * k + 0], raw[4 * k + 1]), pk2u(b.x, b.y)), pk2u(u.x, u.y));
                const uint64_t v1 = add2(add2(pk2u(raw[4 * k + 2], raw[4 * k + 3]), pk2u(b.z, b.w)), pk2u(u.z, u.w));
                uint4 o;
                upk2u(v0, o.x, o.y);
                upk2u(v1, o.z, o.w);
                ptx::sts128(a, o);
              }
            }
          }
          ptx::fence_proxy_async();
          ptx::mbar_arrive(&ctl->chunk_done[te_slot]);
          if (++te_slot == kResSlots) { te_slot = 0; te_ph ^= 1; }
        }
        if constexpr (LN) {
          // this row's partial (sum x, sum x^2) over the tile's BN columns: slot n_tile of the row's ln_parts partials
          if (ln_row_ok) {
            float a0, a1, b0, b1;
            upk2(ln_s1, a0, a1);
            upk2(ln_s2, b0, b1);
            reinterpret_cast<float2*>(p.epi.ln_stats_out)[ln_mrow * p.num_n_tiles + n_tile] = make_float2(a0 + a1, b0 + b1);
          }
        }
        dbg_e_busy += dbg_clock(dbgt) - dbg_t_busy0;
      } else if constexpr (OUT == OUT_CLS_TAIL) {
        // BN == 16: v = relu(conv 32->16 + b); two 1x1 convs in registers; fp32 [pixel][2] store
        long long orow;
        const bool valid = row_to_out(m_tile, r, orow);
        { const long long t0 = dbg_clock(dbgt); ptx::mbar_wait(&ctl->acc_full[as], aphase); dbg_e_wait += dbg_clock(dbgt) - t0; }
        ptx::tc_fence_after();
        const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * kAccStride;
        if (part == 0) {
          uint32_t raw[16];
          ptx::tmem_ld16(t_row, raw);
          ptx::tmem_ld_wait(raw);
          float v[16];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const uint4 b = ptx::lds128(bias_s + (as * 256 + 4 * i) * 4);
            v[4 * i + 0] = fmaxf(__uint_as_float(raw[4 * i + 0]) + __uint_as_float(b.x), 0.f);
            v[4 * i + 1] = fmaxf(__uint_as_float(raw[4 * i + 1]) + __uint_as_float(b.y), 0.f);
            v[4 * i + 2] = fmaxf(__uint_as_float(raw[4 * i + 2]) + __uint_as_float(b.z), 0.f);
            v[4 * i + 3] = fmaxf(__uint_as_float(raw[4 * i + 3]) + __uint_as_float(b.w), 0.f);
          }
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
          if (valid) reinterpret_cast<float2*>(out_base)[orow] = make_float2(o0, o1);
        }
      } else {
        // the 4 rows this lane touches in the cooperative (coalesced) passes
        char* co_out[4];
        const char* co_res[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          long long ro;
          const bool v = row_to_out(m_tile, q * 32 + j * 8 + co_row, ro);
          co_out[j] = v ? out_base + (ro * ldc + n0) * kEsize + co_unit * 16 : nullptr;
          if constexpr (RES) {
            const long long rrow = res_mod > 0 ? (ro % res_mod) : ro;
            co_res[j] = v ? res_base + (rrow * ldr + n0) * kEsize + co_unit * 16 : nullptr;
          }
        }
        // columns this lane's 16-byte unit covers inside a segment starting at chunk c0: valid if < N and < c_end
        const int unit_col = (co_unit * 16) / kEsize;
        const int unit_chunk = kF32 ? 0 : (co_unit >> 1);
        uint4 res_next[4];
        auto load_res = [&](int c0) {
          const bool ok = (n0 + c0 * 16 + unit_col < N) && (c0 + unit_chunk < c_end);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            res_next[j] = (ok && co_res[j] != nullptr) ? *reinterpret_cast<const uint4*>(co_res[j] + c0 * 16 * kEsize)
                                                       : make_uint4(0u, 0u, 0u, 0u);
        };
        if constexpr (RES) { if (c_begin < c_end) load_res(c_begin); }
        // TMA store coordinates of this warp's 32 rows: linear {col, row}; conv {channel, x, y, image}
        int ts_c1 = 0, ts_c2 = 0, ts_c3 = 0;
        if constexpr (TS) {
          if (mode == 1) {
            const int img = m_tile / per_img, t = m_tile - img * per_img;
            ts_c1 = (t % tiles_x) * TW;
            ts_c2 = (t / tiles_x) * TH + q * (32 / TW);
            ts_c3 = img;
          } else {
            ts_c1 = m_tile * kBlockM + q * 32;
          }
        }
        (void)ts_c1; (void)ts_c2; (void)ts_c3;
        { const long long t0 = dbg_clock(dbgt); ptx::mbar_wait(&ctl->acc_full[as], aphase); dbg_e_wait += dbg_clock(dbgt) - t0; }
        const long long dbg_t_busy0 = dbg_clock(dbgt);
        ptx::tc_fence_after();
        const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * kAccStride;
        const uint32_t bias_row = bias_s + as * 1024;
        const uint32_t c1_row = c1_s + as * 1024;
        (void)c1_row;
        if (dbg != 2) {
          // One 64-byte output segment (kSegCols accumulator columns) per step.  Two register sets
          // alternate so the TMEM load of the next segment is in flight while this one is processed.
          constexpr int kSegCols = 16 * kSegChunks;
          auto segment = [&](int c0, uint32_t (&raw)[kSegCols], uint32_t (&nxt)[kSegCols], bool prefetch) {
            ptx::tmem_ld_wait(raw);
            if (prefetch && c0 + kSegChunks < c_end) ptx::tmem_ld<kSegCols>(t_row + (c0 + kSegChunks) * 16, nxt);
            float resv[16];
            if constexpr (RES) {  // fp32 residual of this 16-column segment: registers -> smem -> own row
#pragma unroll
              for (int j = 0; j < 4; ++j) ptx::sts128(co_addr[j], res_next[j]);
              __syncwarp();
              if (c0 + kSegChunks < c_end) load_res(c0 + kSegChunks);  // next segment, in flight during the math
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const uint4 u = ptx::lds128(own_row + ((k ^ own_sw) << 4));
                resv[4 * k + 0] = __uint_as_float(u.x); resv[4 * k + 1] = __uint_as_float(u.y);
                resv[4 * k + 2] = __uint_as_float(u.z); resv[4 * k + 3] = __uint_as_float(u.w);
              }
              __syncwarp();
            }
#pragma unroll
            for (int sc = 0; sc < kSegChunks; ++sc) {
              uint64_t v2[8];  // 16 columns as fp32 pairs (packed FADD2/FFMA2: half the issue slots)
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const uint4 b = ptx::lds128(bias_row + ((c0 + sc) * 16 + 4 * i) * 4);
                if constexpr (kLNA) {   // rstd * acc - rstd * mean * c1 + c0
                  const uint4 cc = ptx::lds128(c1_row + ((c0 + sc) * 16 + 4 * i) * 4);
                  v2[2 * i] = fma2(pk2u(raw[sc * 16 + 4 * i + 0], raw[sc * 16 + 4 * i + 1]), ln_a, fma2(ln_b, pk2u(cc.x, cc.y), pk2u(b.x, b.y)));
                  v2[2 * i + 1] = fma2(pk2u(raw[sc * 16 + 4 * i + 2], raw[sc * 16 + 4 * i + 3]), ln_a, fma2(ln_b, pk2u(cc.z, cc.w), pk2u(b.z, b.w)));
                } else {
                  v2[2 * i] = add2(pk2u(raw[sc * 16 + 4 * i + 0], raw[sc * 16 + 4 * i + 1]), pk2u(b.x, b.y));
                  v2[2 * i + 1] = add2(pk2u(raw[sc * 16 + 4 * i + 2], raw[sc * 16 + 4 * i + 3]), pk2u(b.z, b.w));
                }
              }
              if constexpr (RES) {
#pragma unroll
                for (int i = 0; i < 8; ++i) v2[i] = add2(v2[i], pk2(resv[2 * i], resv[2 * i + 1]));
              }
              if constexpr (ACT == ACT_GELU) {
#pragma unroll
                for (int i = 0; i < 8; ++i) v2[i] = gelu_fast2(v2[i]);
              }
              float v[16];
#pragma unroll
              for (int i = 0; i < 8; ++i) upk2(v2[i], v[2 * i], v[2 * i + 1]);
              if constexpr (ACT == ACT_RELU) {
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.0f);
              }
              if constexpr (kF32) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  ptx::sts128(own_row + ((k ^ own_sw) << 4),
                              make_uint4(__float_as_uint(v[4 * k]), __float_as_uint(v[4 * k + 1]),
                                         __float_as_uint(v[4 * k + 2]), __float_as_uint(v[4 * k + 3])));
              } else {
                if constexpr (TS) {
                  if (sc == 0) {  // the store issued from this buffer kTsBufs segments ago has finished reading it
                    if (lane == 0) ptx::bulk_wait_read<kTsBufs - 1>();
                    __syncwarp();
                  }
                }
                const uint32_t dst_row = TS ? own_row + ts_buf * 2048 : own_row;
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                  uint4 o;
                  o.x = pack_bf16(v[8 * k + 0], v[8 * k + 1]); o.y = pack_bf16(v[8 * k + 2], v[8 * k + 3]);
                  o.z = pack_bf16(v[8 * k + 4], v[8 * k + 5]); o.w = pack_bf16(v[8 * k + 6], v[8 * k + 7]);
                  if constexpr (POOL != 1) ptx::sts128(dst_row + (((sc * 2 + k) ^ own_sw) << 4), o);
                  if constexpr (POOL != 0) {
                    uint32_t w[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                      uint32_t t = __shfl_xor_sync(0xffffffffu, w[e], 1);
                      __nv_bfloat162 a = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&w[e]), *reinterpret_cast<__nv_bfloat162*>(&t));
                      w[e] = *reinterpret_cast<uint32_t*>(&a);
                      t = __shfl_xor_sync(0xffffffffu, w[e], pool_tw);
                      a = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&w[e]), *reinterpret_cast<__nv_bfloat162*>(&t));
                      w[e] = *reinterpret_cast<uint32_t*>(&a);
                    }
                    if (pool_writer) ptx::sts128(pool_row + ts_buf * 512 + (((sc * 2 + k) ^ pool_sw) << 4), make_uint4(w[0], w[1], w[2], w[3]));
                  }
                }
              }
            }
            if constexpr (TS) {
              ptx::fence_proxy_async();
              __syncwarp();
              if (lane == 0) {
                if (dbg != 1) {
                  if constexpr (POOL != 1) {
                    if (mode == 1) ptx::tma_store_4d_s(&p.tmC, stg + ts_buf * 2048, n0 + c0 * 16, ts_c1, ts_c2, ts_c3);
                    else ptx::tma_store_2d_s(&p.tmC, stg + ts_buf * 2048, n0 + c0 * 16, ts_c1);
                  }
                  if constexpr (POOL != 0) ptx::tma_store_4d_s(&p.tmP, pool_stg + ts_buf * 512, n0 + c0 * 16, ts_c1 >> 1, ts_c2 >> 1, ts_c3);
                }
                ptx::bulk_commit();
              }
              ts_buf ^= kTsBufs - 1;
              return;
            }
            __syncwarp();
            // coalesced store: 4 lanes x 16 B per row, 8 rows per pass.  (With an odd chunk count the last
            // bf16 segment's second half holds columns of the neighbouring range: masked by unit_chunk.)
            const bool ok = (n0 + c0 * 16 + unit_col < N) && (c0 + unit_chunk < c_end) && dbg != 1;
            const int byte0 = c0 * 16 * kEsize;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint4 o = ptx::lds128(co_addr[j]);
              if (ok && co_out[j] != nullptr) *reinterpret_cast<uint4*>(co_out[j] + byte0) = o;
            }
            __syncwarp();
          };
          if constexpr (EW == 8) {
            uint32_t ra[kSegCols], rb[kSegCols];
            if (c_begin < c_end) ptx::tmem_ld<kSegCols>(t_row + c_begin * 16, ra);
            for (int c0 = c_begin; c0 < c_end; c0 += 2 * kSegChunks) {
              segment(c0, ra, rb, true);
              if (c0 + kSegChunks < c_end) segment(c0 + kSegChunks, rb, ra, true);
            }
          } else {
            // 4 warps per scheduler hide the TMEM load latency; one register set keeps the 576-thread CTA under 112 registers
            uint32_t ra[kSegCols];
            for (int c0 = c_begin; c0 < c_end; c0 += kSegChunks) {
              ptx::tmem_ld<kSegCols>(t_row + c0 * 16, ra);
              segment(c0, ra, ra, false);
            }
          }
        }
        dbg_e_busy += dbg_clock(dbgt) - dbg_t_busy0;
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (PAIR) ptx::mbar_arrive_cluster(acc_empty_leader + as * 8);
        else ptx::mbar_arrive(&ctl->acc_empty[as]);
      }
      ++dbg_tiles;
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
    if constexpr (TS) { if (lane == 0) ptx::bulk_wait<0>(); }
    if ((p.debug & 4) && blockIdx.x == 0 && lane == 0 && (warp == kEpiWarp0 || warp == kEpiWarp0 + 4)) {
      const int o = (warp == kEpiWarp0 + 4) * 3;
      p.dbg_out[5 + o] = dbg_e_wait; p.dbg_out[6 + o] = dbg_e_busy; p.dbg_out[7 + o] = dbg_tiles;
    }
  }

  if (lane == 0) trace_role_done(p.trace, p.serial, blockIdx.x, warp);
  ptx::tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) trace_mark(p.trace, p.serial, blockIdx.x, TR_SYNC1);
  if constexpr (PAIR) ptx::cluster_sync_all();  // neither CTA exits (or frees TMEM) while its peer still works
  if (threadIdx.x == 0) trace_mark(p.trace, p.serial, blockIdx.x, TR_CSYNC1);
  if (warp == 1) {
    ptx::tc_fence_after();
    if constexpr (PAIR) ptx::tmem_dealloc_pair(tmem_base, kTmemCols);
    else ptx::tmem_dealloc(tmem_base, kTmemCols);
    if (lane == 0) { trace_role_done(p.trace, p.serial, blockIdx.x, kTraceFreeByte); trace_tmem_event(p.trace, p.serial, blockIdx.x, 2); }
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
//   one k-block of a tile costs max(MMA, smem operand reads, L2->SM operand traffic):
//     MMA      = 128*BN*BK/4096 cycles per CTA (tcgen05 rate; a CTA pair computes 256 x BN in the same time),
//     smem     = operand bytes the CTA's tensor core reads / 128 B/cycle,
//     traffic  = operand bytes the CTA loads / ~40 B/cycle/SM (L2 -> SM share with all SMs busy:
//                ~6300 B/cycle chip-wide, B300_MICROARCH.md "LTS throughput cap").
//   Streaming reloads the weight tile for every M tile; weight-resident keeps the CTA's [rows x K] weight
//   slice in smem and streams only A.  In PAIR mode a CTA stages BN/2 weight rows instead of BN.
struct Plan { int BN; int resident; int pair; double cost; };

int env_int(const char* name, int dflt) {
  const char* v = std::getenv(name);
  return v ? std::atoi(v) : dflt;
}

Plan plan_tiles(int N, int Ktot, int BK, long long m_tiles, bool allow_resident, int want_pair, int min_bn = 0, int bn_mult = 1) {
  static const int cand[] = {256, 192, 128, 96, 64, 48, 32, 16};
  static const int pair_env = env_int("TT_GEMM_PAIR", 1);  // 0 never, 1 cost model, 2 whenever legal (development)
  const double kL2 = 40.0, kSmem = 128.0;
  const int sms = num_sms();
  const int kblocks = Ktot / BK;
  Plan best{0, 0, 0, 1e300};
  for (int pair = 0; pair < 2; ++pair) {
    if (pair && (want_pair == 0 || pair_env == 0)) continue;
    if (!pair && (want_pair == 1 || pair_env == 2) && N % 16 == 0) continue;
    const long long units = pair ? (m_tiles + 1) / 2 : m_tiles;
    const int slots = pair ? sms / 2 : sms;
    const int res_max = pair ? kResidentMaxPair : kResidentMax;
    for (int c : cand) {
      if (N % c != 0 || c < min_bn || c % bn_mult != 0) continue;
      const int b_rows = pair ? c / 2 : c;
      const double mma = 128.0 * c * BK / 4096.0;
      const double a_bytes = 128.0 * BK * 2, b_bytes = static_cast<double>(b_rows) * BK * 2;
      const double smem_rd = (a_bytes + b_bytes) / kSmem;
      const double fixed = pair ? 800.0 : 600.0;
      const long long n_tiles = N / c;
      {  // streaming
        const double tile = kblocks * std::max({mma, smem_rd, (a_bytes + b_bytes) / kL2}) + fixed;
        const double waves = static_cast<double>((units * n_tiles + slots - 1) / slots);
        const double cost = waves * tile;
        if (cost < best.cost) best = Plan{c, 0, pair, cost};
      }
      const long long groups = std::min<long long>(slots / n_tiles, units);
      if (allow_resident && n_tiles <= slots && static_cast<long long>(b_rows) * Ktot * 2 <= res_max && groups >= 1 &&
          units >= 4 * groups) {
        const int stages = (kSmemBudget - static_cast<int>(sizeof(SmemCtl)) - kStagingBytes - 1024 - b_rows * Ktot * 2) / (128 * BK * 2);
        if (stages >= 3) {
          const double tile = kblocks * std::max({mma, smem_rd, a_bytes / kL2}) + fixed;
          const double cost = static_cast<double>((units + groups - 1) / groups) * tile + b_bytes * kblocks / kL2;
          if (cost < best.cost) best = Plan{c, 1, pair, cost};
        }
      }
    }
  }
  if (best.BN == 0) best = Plan{((N + 15) / 16) * 16 <= 256 ? ((N + 15) / 16) * 16 : 128, 0, 0, 0.0};
  return best;
}

// Co-resident CTA pairs the device can hold for this kernel (74 on a full B200: one pair per TPC).
int max_pairs(const void* fn, size_t smem, int threads) {
  static std::mutex mu;
  static int cached = 0;
  std::lock_guard<std::mutex> lock(mu);
  if (cached) return cached;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * num_sms());
  cfg.blockDim = dim3(threads);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, fn, &cfg) != cudaSuccess || n <= 0) { cudaGetLastError(); n = num_sms() / 2; }
  cached = std::min(n, num_sms() / 2);
  return cached;
}

// Epilogue warps per CTA.  Measured (tools/gpu_dbg.sh, profiles/r1b_gemm_roles.md): 16 warps only pay for the GELU
// epilogue on 256-wide tiles (MUFU-bound, +5 %); elsewhere the epilogue is bound by shared per-SM resources (TMEM
// read port, store path), 16 warps cost a smem stage and split 192/64-wide tiles into half-filled 64-byte segments.
// TT_GEMM_EW=8|16 forces one value (development).
int epi_warps(const Epilogue& e, int BN) {
  static const int ew_env = env_int("TT_GEMM_EW", 0);
  if (e.out_type == OUT_CLS_TAIL) return 8;
  if (ew_env == 8) return 8;
  if (ew_env == 16) return (e.out_type == OUT_BF16 && e.act == ACT_GELU && BN % 128 == 0) ? 16 : 8;
  // 16 warps for the MUFU-bound GELU tiles: +1 % end to end (profiles/r1c_ab_switches.md)
  return (e.out_type == OUT_BF16 && e.act == ACT_GELU && BN % 128 == 0) ? 16 : 8;
}

template <bool PAIR>
void (*select_kernel_ts(const Epilogue& e, int ew))(const KParams) {
  if (e.pool_mode == 1) return gemm_tc_kernel<OUT_BF16, ACT_RELU, false, PAIR, 8, false, true, false, 1>;
  if (e.pool_mode == 2) return gemm_tc_kernel<OUT_BF16, ACT_RELU, false, PAIR, 8, false, true, false, 2>;
  if (e.ln_stats_in != nullptr) {   // LayerNorm consumer: plain or GELU bf16 outputs
    if (ew == 8) return e.act == ACT_GELU ? gemm_tc_kernel<OUT_BF16, ACT_GELU, false, PAIR, 8, false, true, true>
                                          : gemm_tc_kernel<OUT_BF16, ACT_NONE, false, PAIR, 8, false, true, true>;
    return e.act == ACT_GELU ? gemm_tc_kernel<OUT_BF16, ACT_GELU, false, PAIR, 16, false, true, true>
                             : gemm_tc_kernel<OUT_BF16, ACT_NONE, false, PAIR, 16, false, true, true>;
  }
  if (ew == 8) {
    if (e.act == ACT_RELU) return gemm_tc_kernel<OUT_BF16, ACT_RELU, false, PAIR, 8, false, true>;
    if (e.act == ACT_GELU) return gemm_tc_kernel<OUT_BF16, ACT_GELU, false, PAIR, 8, false, true>;
    return gemm_tc_kernel<OUT_BF16, ACT_NONE, false, PAIR, 8, false, true>;
  }
  if (e.act == ACT_RELU) return gemm_tc_kernel<OUT_BF16, ACT_RELU, false, PAIR, 16, false, true>;
  if (e.act == ACT_GELU) return gemm_tc_kernel<OUT_BF16, ACT_GELU, false, PAIR, 16, false, true>;
  return gemm_tc_kernel<OUT_BF16, ACT_NONE, false, PAIR, 16, false, true>;
}

template <bool PAIR>
void (*select_kernel(const Epilogue& e, int ew))(const KParams) {
  if (e.out_type == OUT_CLS_TAIL) return gemm_tc_kernel<OUT_CLS_TAIL, ACT_RELU, false, PAIR, 8>;
  if (e.out_type == OUT_F32 && ew == 16)
    return e.res_type == RES_F32 ? gemm_tc_kernel<OUT_F32, ACT_NONE, true, PAIR, 16> : gemm_tc_kernel<OUT_F32, ACT_NONE, false, PAIR, 16>;
  if (e.out_type == OUT_F32)
    return e.res_type == RES_F32 ? gemm_tc_kernel<OUT_F32, ACT_NONE, true, PAIR, 8> : gemm_tc_kernel<OUT_F32, ACT_NONE, false, PAIR, 8>;
  if (ew == 8) {
    if (e.act == ACT_RELU) return gemm_tc_kernel<OUT_BF16, ACT_RELU, false, PAIR, 8>;
    if (e.act == ACT_GELU) return gemm_tc_kernel<OUT_BF16, ACT_GELU, false, PAIR, 8>;
    return gemm_tc_kernel<OUT_BF16, ACT_NONE, false, PAIR, 8>;
  }
  if (e.act == ACT_RELU) return gemm_tc_kernel<OUT_BF16, ACT_RELU, false, PAIR, 16>;
  if (e.act == ACT_GELU) return gemm_tc_kernel<OUT_BF16, ACT_GELU, false, PAIR, 16>;
  return gemm_tc_kernel<OUT_BF16, ACT_NONE, false, PAIR, 16>;
}

// fp32 [rows][cols] tensor (row pitch ld elements) as a TMA map with a {32 cols, 128 rows} SWIZZLE_128B box
bool make_tmap_f32_chunk(CUtensorMap* m, const void* base, long long rows, int cols, long long ld) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled unavailable (no CUDA driver?)"); return false; }
  const cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  const cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 4};
  const cuuint32_t box[2] = {32, 128};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (fp32 chunk) failed with CUresult " + std::to_string(static_cast<int>(r))); return false; }
  return true;
}

// The TMA epilogue handles plain row-major fp32 out = residual + acc + bias whose residual rows are the output rows
// (or a table whose period is a multiple of the 128-row tile: the patch embedding's pos_embed).
bool tma_epilogue_ok(const KParams& kp) {
  static const int te_env = env_int("TT_GEMM_TE", 1);
  const Epilogue& e = kp.epi;
  // Only for launches that fill the GPU: small ones are latency bound and gain nothing from the extra two warps and the
  // 64 KB ring (TT_GEMM_TE=2 forces it everywhere: the regime of the round-1 hang, kept for the regression probe).
  if (e.out_type == OUT_SPLIT)   // the split residual stream (LayerNorm-statistics producer) exists only as a TMA epilogue
    return kp.mode == 0 && e.res_type == RES_SPLIT && e.act == ACT_NONE && kp.BN % 32 == 0 && kp.N % 8 == 0 && e.ldc % 8 == 0 &&
           e.ldr % 8 == 0 && (e.res_mod == 0 || e.res_mod % kBlockM == 0) && e.out_lo != nullptr && e.residual_lo != nullptr &&
           reinterpret_cast<uintptr_t>(e.out) % 16 == 0 && reinterpret_cast<uintptr_t>(e.residual) % 16 == 0 &&
           reinterpret_cast<uintptr_t>(e.out_lo) % 16 == 0 && reinterpret_cast<uintptr_t>(e.residual_lo) % 16 == 0;
  if (te_env != 2 && kp.m_tiles_total < 2 * num_sms()) return false;
  return te_env != 0 && kp.mode == 0 && e.out_type == OUT_F32 && e.res_type == RES_F32 && e.act == ACT_NONE && kp.BN % 32 == 0 &&
         kp.N % 4 == 0 && e.ldc % 4 == 0 && e.ldr % 4 == 0 && (e.res_mod == 0 || e.res_mod % kBlockM == 0) &&
         reinterpret_cast<uintptr_t>(e.out) % 16 == 0 && reinterpret_cast<uintptr_t>(e.residual) % 16 == 0;
}

// kp.num_m_tiles holds the 128-row tile count on entry; PAIR mode turns it into the pair count.
cudaError_t launch(KParams& kp, cudaStream_t s, double flops) {
  const int row_bytes = kp.BK * 2;
  const int num_kb = kp.taps * (kp.kb_src[0] + kp.kb_src[1]);
  kp.m_tiles_total = kp.num_m_tiles;
  if (kp.pair && ((kp.BN / 2) % 8 != 0 || kp.epi.out_type == OUT_CLS_TAIL)) kp.pair = 0;
  if (kp.pair) kp.num_m_tiles = (kp.num_m_tiles + 1) / 2;
  const int b_rows = kp.pair ? kp.BN / 2 : kp.BN;
  const int a_bytes = kBlockM * row_bytes, b_bytes = b_rows * row_bytes;
  using KernelFn = void (*)(const KParams);
  bool te = tma_epilogue_ok(kp);
  {  // development switch for the open concurrency issue: TT_TE_PAIR=0 keeps the TMA epilogue off CTA-pair launches
    static const int te_pair_env = env_int("TT_TE_PAIR", 1);
    if (te && kp.pair && te_pair_env == 0) te = false;
  }
  static const int ts_env = env_int("TT_GEMM_TS", 1);
  const bool ts = !te && (ts_env != 0 || kp.epi.pool_mode != 0) && kp.epi.out_type == OUT_BF16 && kp.BN % 32 == 0 && kp.epi.ldc % 8 == 0 &&
                  reinterpret_cast<uintptr_t>(kp.epi.out) % 16 == 0;
  const int ew = te ? 4 : epi_warps(kp.epi, kp.BN);
  if (kp.epi.pool_mode != 0 && (!ts || kp.mode != 1 || ew != 8 || kp.epi.act != ACT_RELU || kp.epi.pool_out == nullptr || kp.H % 2 || kp.W % 2 ||
                                kp.N % 32 != 0 || reinterpret_cast<uintptr_t>(kp.epi.pool_out) % 16 != 0)) {
    set_error("gemm: the fused 2x2 max-pool needs a ReLU conv with bf16 TMA-store output, even H and W, Cout % 32 == 0");
    return cudaErrorInvalidValue;
  }
  KernelFn fn;
  const bool ln_prod = kp.epi.ln_stats_out != nullptr, ln_cons = kp.epi.ln_stats_in != nullptr;
  if ((ln_prod || kp.epi.out_type == OUT_SPLIT) && (!te || !ln_prod || kp.epi.out_type != OUT_SPLIT || kp.num_n_tiles > 4)) {
    set_error("gemm: the LayerNorm-statistics producer is the split (hi/lo bf16) residual GEMM with a TMA epilogue and <= 4 N tiles");
    return cudaErrorInvalidValue;
  }
  if (ln_cons && (!ts || kp.epi.ln_c1 == nullptr || kp.epi.ln_parts <= 0 || kp.epi.ln_dim <= 0)) {
    set_error("gemm: the LayerNorm consumer needs the TMA-store bf16 epilogue, c1 and the producer's statistics");
    return cudaErrorInvalidValue;
  }
  if (te) {
    if (ln_prod)
      fn = kp.pair ? gemm_tc_kernel<OUT_F32, ACT_NONE, true, true, 4, true, false, true> : gemm_tc_kernel<OUT_F32, ACT_NONE, true, false, 4, true, false, true>;
    else
      fn = kp.pair ? gemm_tc_kernel<OUT_F32, ACT_NONE, true, true, 4, true> : gemm_tc_kernel<OUT_F32, ACT_NONE, true, false, 4, true>;
    const long long res_rows = kp.epi.res_mod > 0 ? kp.epi.res_mod : kp.M;
    if (ln_prod) {
      auto split_map = [&](CUtensorMap* m, const void* base, long long rows, int ld) {
        const cuuint64_t dims[2] = {static_cast<cuuint64_t>(kp.N), static_cast<cuuint64_t>(rows)};
        const cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 2};
        const cuuint32_t box[2] = {32, 128};
        return make_tmap_bf16(m, base, 2, dims, strides, box, 64);
      };
      if (!split_map(&kp.tmR, kp.epi.residual, res_rows, kp.epi.ldr) || !split_map(&kp.tmR2, kp.epi.residual_lo, res_rows, kp.epi.ldr) ||
          !split_map(&kp.tmC, kp.epi.out, kp.M, kp.epi.ldc) || !split_map(&kp.tmC2, kp.epi.out_lo, kp.M, kp.epi.ldc))
        return cudaErrorInvalidValue;
      if (kp.epi.ln_parts_out) *kp.epi.ln_parts_out = kp.num_n_tiles;
    } else {
      if (!make_tmap_f32_chunk(&kp.tmR, kp.epi.residual, res_rows, kp.N, kp.epi.ldr)) return cudaErrorInvalidValue;
      if (!make_tmap_f32_chunk(&kp.tmC, kp.epi.out, kp.M, kp.N, kp.epi.ldc)) return cudaErrorInvalidValue;
    }
  } else if (ts) {
    fn = kp.pair ? select_kernel_ts<true>(kp.epi, ew) : select_kernel_ts<false>(kp.epi, ew);
    if (kp.epi.pool_mode != 0) {
      const cuuint64_t dims[4] = {static_cast<cuuint64_t>(kp.N), static_cast<cuuint64_t>(kp.W / 2), static_cast<cuuint64_t>(kp.H / 2),
                                  static_cast<cuuint64_t>(kp.M / (kp.H * kp.W))};
      const cuuint64_t pitch = static_cast<cuuint64_t>(kp.N) * 2;
      const cuuint64_t strides[3] = {pitch, pitch * (kp.W / 2), pitch * (kp.W / 2) * (kp.H / 2)};
      const cuuint32_t box[4] = {32, static_cast<cuuint32_t>(kp.TW / 2), static_cast<cuuint32_t>(16 / kp.TW), 1};
      if (!make_tmap_bf16(&kp.tmP, kp.epi.pool_out, 4, dims, strides, box, 64)) return cudaErrorInvalidValue;
    }
    if (kp.epi.pool_mode == 1) {
      // pooled output only: no un-pooled store map
    } else if (kp.mode == 1) {
      const cuuint64_t dims[4] = {static_cast<cuuint64_t>(kp.N), static_cast<cuuint64_t>(kp.W), static_cast<cuuint64_t>(kp.H),
                                  static_cast<cuuint64_t>(kp.M / (kp.H * kp.W))};
      const cuuint64_t pitch = static_cast<cuuint64_t>(kp.epi.ldc) * 2;
      const cuuint64_t strides[3] = {pitch, pitch * kp.W, pitch * kp.W * kp.H};
      const cuuint32_t box[4] = {32, static_cast<cuuint32_t>(kp.TW), static_cast<cuuint32_t>(32 / kp.TW), 1};
      if (!make_tmap_bf16(&kp.tmC, kp.epi.out, 4, dims, strides, box, 64)) return cudaErrorInvalidValue;
    } else {
      const cuuint64_t dims[2] = {static_cast<cuuint64_t>(kp.N), static_cast<cuuint64_t>(kp.M)};
      const cuuint64_t strides[1] = {static_cast<cuuint64_t>(kp.epi.ldc) * 2};
      const cuuint32_t box[2] = {32, 32};
      if (!make_tmap_bf16(&kp.tmC, kp.epi.out, 2, dims, strides, box, 64)) return cudaErrorInvalidValue;
    }
  } else {
    fn = kp.pair ? select_kernel<true>(kp.epi, ew) : select_kernel<false>(kp.epi, ew);
  }
  TT_CUDA_TRY(ensure_dynamic_smem(reinterpret_cast<const void*>(fn), 227 * 1024));
  const int threads = 64 + 32 * ew + (te ? 64 : 0);
  const int staging = te ? kResSlots * kResSlotBytes
                         : ts ? ew * (ew == 16 ? 1 : 2) * (2048 + (kp.epi.pool_mode ? 512 : 0)) : ew * 2048;
  const int slots = kp.pair ? max_pairs(reinterpret_cast<const void*>(fn), 200 * 1024, threads) : num_sms();
  if (kp.b_resident && (num_kb * b_bytes > (kp.pair ? kResidentMaxPair : kp.halo ? kHaloResidentMax : kResidentMax) || kp.num_n_tiles > slots)) kp.b_resident = 0;
  const int res_bytes = kp.b_resident ? num_kb * b_bytes : 0;
  if (kp.halo && !kp.b_resident) { set_error("conv: halo mode needs resident weights"); return cudaErrorInvalidValue; }
  const int stage_bytes = kp.halo ? kHaloPixels * row_bytes : kp.b_resident ? a_bytes : a_bytes + b_bytes;
  kp.stages = std::max(2, std::min(kMaxStages, (kSmemBudget - static_cast<int>(sizeof(SmemCtl)) - staging - 1024 - res_bytes) / stage_bytes));
  const size_t smem = static_cast<size_t>(res_bytes) + static_cast<size_t>(kp.stages) * stage_bytes + sizeof(SmemCtl) + staging + 1024;
  int grid;  // in scheduling slots: CTAs, or CTA pairs
  if (kp.b_resident) {
    const int groups = std::min(slots / kp.num_n_tiles, kp.num_m_tiles);
    grid = groups * kp.num_n_tiles;
  } else {
    grid = std::min(kp.num_m_tiles * kp.num_n_tiles, slots);
  }
  if (kp.pair) grid *= 2;
  static const int dbg = std::getenv("TT_GEMM_DEBUG") ? std::atoi(std::getenv("TT_GEMM_DEBUG")) : 0;
  kp.debug = dbg;
  static unsigned long long* dbg_buf = nullptr;
  if ((dbg & 4) && !dbg_buf) cudaMalloc(&dbg_buf, 16 * sizeof(unsigned long long));
  kp.dbg_out = dbg_buf;
  if (dbg & 4) cudaMemset(dbg_buf, 0, 16 * sizeof(unsigned long long));
  char tag[128];
  static const int alloc_sync_env = env_int("TT_PAIR_ALLOC_SYNC", 1);
  kp.alloc_sync = alloc_sync_env;
  kp.trace = trace_dev();
  if (kp.trace) {
    std::snprintf(tag, sizeof(tag), "%s M%d N%d K%d BN%d st%d %s%s%s ew%d", kp.mode ? "conv" : "lin", kp.M, kp.N, num_kb * kp.BK, kp.BN,
                  kp.stages, kp.b_resident ? "resident" : "stream", kp.pair ? " pair" : "", kp.halo ? " halo" : te ? " tma-epi" : ts ? " tma-store" : "", ew);
    kp.serial = trace_launch(tag, grid, threads, smem, s);
  }
  if (prof_enabled()) {
    std::snprintf(tag, sizeof(tag), "%s M%d N%d K%d BN%d BK%d st%d grid%d %s%s%s", kp.mode ? "conv" : "lin", kp.M, kp.N,
                  num_kb * kp.BK, kp.BN, kp.BK, kp.stages, grid, kp.b_resident ? "resident" : "stream", kp.pair ? " pair" : "",
                  kp.halo ? " halo" : te ? " tma-epi" : ts ? " tma-store" : "");
    prof_record(s, true, 0, 0);
  }
  if (kp.pair) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    TT_CUDA_TRY(cudaLaunchKernelEx(&cfg, fn, kp));
  } else {
    fn<<<grid, threads, smem, s>>>(kp);
  }
  prof_record(s, false, flops, 0, tag);
  TT_LAUNCH_CHECK();
  if (dbg & 4) {
    unsigned long long h[16];
    cudaStreamSynchronize(s);
    cudaMemcpy(h, dbg_buf, sizeof(h), cudaMemcpyDeviceToHost);
    std::fprintf(stderr, "[gemm dbg] M%d N%d K%d BN%d st%d grid%d | producer wait %llu of %llu | mma wait acc %llu full %llu of %llu | "
                 "epi w2 wait %llu busy %llu tiles %llu | epi w6 wait %llu busy %llu\n", kp.M, kp.N, num_kb * kp.BK, kp.BN, kp.stages,
                 grid, h[0], h[1], h[2], h[3], h[4], h[5], h[6], h[7], h[8], h[9]);
  }
  return cudaSuccess;
}

cudaError_t check_epilogue(const Epilogue& e, int N, int BN) {
  if (e.out == nullptr && e.pool_mode != 1) { set_error("gemm: null output"); return cudaErrorInvalidValue; }
  if (e.out_type == OUT_CLS_TAIL && (BN != 16 || N != 16 || e.tail == nullptr)) {
    set_error("gemm: cls tail needs N == BN == 16 and tail weights");
    return cudaErrorInvalidValue;
  }
  if (e.out_type == OUT_BF16 && (e.ldc % 8 != 0 || N % 16 != 0)) {
    set_error("gemm: bf16 output needs ldc % 8 == 0 and N % 16 == 0");
    return cudaErrorInvalidValue;
  }
  if (e.out_type == OUT_F32 && (e.ldc % 4 != 0 || N % 16 != 0)) {
    set_error("gemm: fp32 output needs ldc % 4 == 0 and N % 16 == 0");
    return cudaErrorInvalidValue;
  }
  if (e.out_type == OUT_SPLIT || e.res_type == RES_SPLIT) {
    if (e.out_type != OUT_SPLIT || e.res_type != RES_SPLIT || e.ln_stats_out == nullptr || e.act != ACT_NONE || N % 16 != 0) {
      set_error("gemm: the split (hi/lo bf16) residual stream is RES_SPLIT in, OUT_SPLIT out, with LayerNorm statistics");
      return cudaErrorInvalidValue;
    }
    return cudaSuccess;
  }
  if (e.res_type != RES_NONE && (e.res_type != RES_F32 || e.out_type != OUT_F32 || e.ldr % 4 != 0)) {
    set_error("gemm: the residual path is fp32 in / fp32 out with a 16-byte aligned pitch");
    return cudaErrorInvalidValue;
  }
  if (e.out_type == OUT_F32 && e.act != ACT_NONE) {
    set_error("gemm: fp32 output has no fused activation");
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
  // Halo mode (3x3, dilation 1, one source of 64 or 128 channels, weights small enough to stay resident): the
  // per-tap boxes re-read every input pixel 9 times from L2 and these narrow layers were bound by exactly that
  // traffic (c1_2: 361 TF).  The tile becomes 16 x 8 pixels so that every 8-row MMA group is one image row.
  static const int halo_env = env_int("TT_CONV_HALO", 1);
  {
    const int cout_pad = ((c.Cout + 15) / 16) * 16;
    const long long w_bytes = static_cast<long long>(cout_pad) * c.taps * ctot * 2;
    const bool pair_ok = c.pair != 0 && (cout_pad / 2) % 8 == 0 && e.out_type != OUT_CLS_TAIL;
    const bool fits = w_bytes <= kHaloResidentMax || (pair_ok && w_bytes / 2 <= kResidentMaxPair);
    kp.halo = (halo_env != 0 && c.taps == 9 && c.dil == 1 && c.nsrc == 1 && ctot <= (kp.BK == 64 ? 128 : 32) && cout_pad <= 256 &&
               c.Cout % 16 == 0 && c.BN == 0 && c.resident != 0 && fits && c.W >= 8 && c.H >= 16)
                  ? 1 : 0;
  }
  // 128 output pixels per tile as a TH x TW rectangle; wider-than-tall keeps TMA rows long
  kp.TW = (c.W >= 16 && !kp.halo) ? 16 : 8;
  kp.TH = kBlockM / kp.TW;
  kp.tiles_x = (c.W + kp.TW - 1) / kp.TW;
  kp.tiles_y = (c.H + kp.TH - 1) / kp.TH;
  kp.num_m_tiles = c.batch * kp.tiles_x * kp.tiles_y;
  kp.N = c.Cout;
  kp.M = c.batch * c.H * c.W;
  {
    const Plan pl = plan_tiles(c.Cout, c.taps * ctot, kp.BK, kp.num_m_tiles, c.resident != 0, c.pair);
    kp.BN = c.BN ? c.BN : pl.BN;
    kp.b_resident = c.BN ? (c.resident == 1) : pl.resident;
    kp.pair = c.BN ? (c.pair == 1) : pl.pair;
    if (kp.pair && ((kp.BN / 2) % 8 != 0 || e.out_type == OUT_CLS_TAIL)) kp.pair = 0;
    if (kp.halo) {  // the whole Cout in one resident tile; a CTA pair when the slice only fits halved
      kp.BN = c.Cout;
      kp.b_resident = 1;
      const long long w_bytes = static_cast<long long>(c.Cout) * c.taps * ctot * 2;
      const bool pair_ok = c.pair != 0 && (c.Cout / 2) % 8 == 0 && e.out_type != OUT_CLS_TAIL;
      // a cta_group::2 MMA takes >= ~110 cycles whatever its N (measured), so narrow layers only pair up when
      // the weights do not fit one CTA
      kp.pair = (w_bytes > kHaloResidentMax && pair_ok) ? 1 : 0;
    }
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
    const cuuint32_t box[4] = {static_cast<cuuint32_t>(kp.BK), static_cast<cuuint32_t>(kp.halo ? kHaloPitch : kp.TW),
                               static_cast<cuuint32_t>(kp.halo ? kHaloRows : kp.TH), 1};
    if (!make_map(&kp.tmA[i], c.src[i].ptr, 4, dims, strides, box, row_bytes)) return cudaErrorInvalidValue;
  }
  {
    const cuuint64_t ktot = static_cast<cuuint64_t>(c.taps) * ctot;
    const cuuint64_t dims[2] = {ktot, static_cast<cuuint64_t>(c.Cout)};
    const cuuint64_t strides[1] = {ktot * 2};
    const cuuint32_t box[2] = {static_cast<cuuint32_t>(kp.BK), static_cast<cuuint32_t>(kp.pair ? kp.BN / 2 : kp.BN)};
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
    // LayerNorm producer: at most 4 N tiles (statistics slots); both LayerNorm roles need whole 32-column TMA boxes
    const bool ln_any = e.ln_stats_out != nullptr || e.ln_stats_in != nullptr;
    // The producer's N tile is a function of N alone (the widest legal one): the partial sums' grouping, and with it every
    // bit downstream, must not depend on the batch size (test_parseq_full_batch_is_batch_invariant).
    int ln_bn = 0;
    if (e.ln_stats_out)
      for (int c : {256, 192, 128, 96, 64, 32})
        if (l.N % c == 0 && l.N / c <= 4) { ln_bn = c; break; }
    const Plan pl = plan_tiles(l.N, l.K, kp.BK, kp.num_m_tiles, l.resident != 0, l.pair, ln_bn, ln_any ? 32 : 1);
    kp.BN = l.BN ? l.BN : pl.BN;
    kp.b_resident = l.BN ? (l.resident == 1) : pl.resident;
    kp.pair = l.BN ? (l.pair == 1) : pl.pair;
    if (kp.pair && ((kp.BN / 2) % 8 != 0 || e.out_type == OUT_CLS_TAIL)) kp.pair = 0;
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
    const cuuint32_t box[2] = {static_cast<cuuint32_t>(kp.BK), static_cast<cuuint32_t>(kp.pair ? kp.BN / 2 : kp.BN)};
    if (!make_map(&kp.tmB, l.W, 2, dims, strides, box, row_bytes)) return cudaErrorInvalidValue;
  }
  return launch(kp, s, 2.0 * l.M * (l.algo_n > 0 ? l.algo_n : l.N) * l.K);
}

}  // namespace tt
