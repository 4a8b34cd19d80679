// Hang diagnosis aid (TT_TRACE=1, off by default): every TMEM-allocating kernel writes its progress into a device
// buffer that a host watchdog can snapshot (async copy on a private stream) while the kernel is still stuck on the
// GPU.  Nothing here runs unless the environment variable is set.
//
// Layout (uint32 words):
//   GEMM region: a ring of kTraceLaunches launches x kTraceCtas CTAs x 8 words
//     word 0: launch serial   word 1: (smid << 8) | state (written by thread 0)   words 2..7: one byte per warp,
//     1 = its role loop finished; the last two bytes flag "tcgen05.alloc returned" / "TMEM freed" (written by the allocating warp)
//   small-CTA region (encoder attention: thousands of CTAs per launch): kTraceSms SMs x 64 resident-warp slots,
//     one word per CTA keyed by the hardware slot of its first warp: (serial << 8) | state, 0 = slot free
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

namespace tt {

constexpr int kTraceLaunches = 4096;   // must exceed the launches a host thread can enqueue ahead of the GPU
constexpr int kTraceCtas = 160;
constexpr int kTraceSms = 192;
constexpr int kTraceEntry = 8;
constexpr int kTraceAllocByte = 23;   // role bytes 0..21: warp w finished its role; 23: tcgen05.alloc returned; 22: TMEM freed
constexpr int kTraceFreeByte = 22;
constexpr size_t kTraceGemmWords = static_cast<size_t>(kTraceLaunches) * kTraceCtas * kTraceEntry;
constexpr size_t kTraceSmallWords = static_cast<size_t>(kTraceSms) * 64;
// per-SM history of TMEM events: kTraceSms cursors, then kTraceSms x kTraceHist entries of 2 words
//   word 0: launch serial   word 1: (sequence number on that SM << 16) | ((cta & 0xfff) << 4) | event (1 = alloc returned, 2 = dealloc issued)
constexpr int kTraceHist = 64;
constexpr size_t kTraceHistBase = kTraceGemmWords + kTraceSmallWords;
constexpr size_t kTraceWords = kTraceHistBase + kTraceSms + static_cast<size_t>(kTraceSms) * kTraceHist * 2;

enum TraceState : uint32_t {
  TR_ENTER = 1,       // CTA started
  TR_ALLOC = 2,       // tcgen05.alloc returned (small-CTA region only)
  TR_SYNC0 = 3,       // prologue __syncthreads / cluster barrier passed
  TR_SYNC1 = 5,       // final __syncthreads passed
  TR_CSYNC1 = 6,      // final cluster barrier passed
  TR_DONE = 7,        // TMEM freed
};

// Flag bytes carry the ring generation of their launch, so stale flags of the slot's previous occupant never match.
__host__ __device__ inline uint8_t trace_flag(uint32_t serial) { return static_cast<uint8_t>(0x80u | ((serial / kTraceLaunches) & 0x7fu)); }

// Host side.  trace_dev() is the device pointer of the mapped buffer (nullptr when tracing is off).
uint32_t* trace_dev();
// Registers a launch and returns its serial (0 when tracing is off).
uint32_t trace_launch(const char* tag, int grid, int threads, size_t smem, cudaStream_t s);
// Every CTA of the recorded launches that has not reached TR_DONE, plus the resident small CTAs.
std::string trace_report();

#ifdef __CUDACC__
__device__ __forceinline__ uint32_t trace_smid() {
  uint32_t s;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(s));
  return s;
}
__device__ __forceinline__ volatile uint32_t* trace_entry(uint32_t* tr, uint32_t serial, uint32_t cta) {
  return tr + (static_cast<size_t>(serial % kTraceLaunches) * kTraceCtas + cta) * kTraceEntry;
}
__device__ __forceinline__ void trace_mark(uint32_t* tr, uint32_t serial, uint32_t cta, uint32_t state) {
  if (tr == nullptr || cta >= kTraceCtas) return;
  volatile uint32_t* e = trace_entry(tr, serial, cta);
  e[0] = serial;
  e[1] = (trace_smid() << 8) | state;
  __threadfence();
}
__device__ __forceinline__ void trace_role_done(uint32_t* tr, uint32_t serial, uint32_t cta, uint32_t warp) {
  if (tr == nullptr || cta >= kTraceCtas || warp >= 4 * (kTraceEntry - 2)) return;
  volatile uint8_t* b = reinterpret_cast<volatile uint8_t*>(trace_entry(tr, serial, cta) + 2);
  b[warp] = trace_flag(serial);
  __threadfence();
}
// TMEM allocator history of the SM this thread runs on (ev: 1 = alloc returned, 2 = dealloc issued)
__device__ __forceinline__ void trace_tmem_event(uint32_t* tr, uint32_t serial, uint32_t cta, uint32_t ev) {
  if (tr == nullptr) return;
  const uint32_t sm = trace_smid() % kTraceSms;
  const uint32_t seq = atomicAdd(tr + kTraceHistBase + sm, 1u);
  volatile uint32_t* e = tr + kTraceHistBase + kTraceSms + (static_cast<size_t>(sm) * kTraceHist + seq % kTraceHist) * 2;
  e[0] = serial;
  e[1] = (seq << 16) | ((cta & 0xfffu) << 4) | ev;
  __threadfence();
}
// small CTAs: call from the first thread of the CTA
__device__ __forceinline__ void trace_small(uint32_t* tr, uint32_t serial, uint32_t state) {
  if (tr == nullptr) return;
  uint32_t wid;
  asm volatile("mov.u32 %0, %%warpid;" : "=r"(wid));
  volatile uint32_t* e = tr + kTraceGemmWords + static_cast<size_t>(trace_smid() % kTraceSms) * 64 + (wid & 63);
  *e = state == TR_DONE ? 0u : ((serial << 8) | state);
  __threadfence();
}
#endif

}  // namespace tt
