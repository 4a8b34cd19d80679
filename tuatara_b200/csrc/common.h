// Shared host-side helpers: thread-local error string (behind tt_last_error), CUDA status
// macros, a global kernel-launch counter (bench.py reports it as gpu_launches).
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdio>
#include <string>

namespace tt {

void set_error(const std::string& msg);
const char* last_error();
extern std::atomic<unsigned long long> g_launches;
inline void count_launch(unsigned n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }
extern std::atomic<unsigned long long> g_h2d_bytes, g_d2h_bytes;  // engine-level host<->device traffic
// Optional per-launch CUDA-event timing of the tensor-core kernel (bench.py's roofline): when on,
// every gemm_tc launch is bracketed by two events on its stream; prof_collect() resolves them.
void prof_enable(bool on);
bool prof_enabled();
void prof_record(cudaStream_t s, bool begin, double flops, double bytes, const char* tag = nullptr);
// dump_path != nullptr: also append one CSV line per launch (tag, flops, ms) to that file.
void prof_collect(double* total_ms, double* total_flops, double* total_bytes, unsigned long long* launches,
                  const char* dump_path = nullptr);
// Stage-level timing for bench.py's per-stage rooflines (same switch as the per-launch profile): a stage is a run
// of kernels on one stream bracketed by two events; flops / bytes are the stage's ALGORITHMIC work.
void stage_begin(cudaStream_t s);
void stage_end(cudaStream_t s, const char* name, double flops, double bytes);
// "name,count,ms,flops,bytes\n" per stage (summed over the recorded intervals); clears the log.
std::string stage_collect();
// cudaStreamSynchronize with a host-side watchdog: polls the stream and gives up after TT_WATCHDOG_S seconds (default
// 120, 0 = wait forever), so that a stuck kernel becomes an error return (cudaErrorLaunchTimeout + tt_last_error with
// the device-side trace when TT_TRACE=1) instead of a caller blocked forever.  The stream itself stays stuck.
cudaError_t stream_sync(cudaStream_t s);
// cudaFuncAttributeMaxDynamicSharedMemorySize is per device: set it once per (kernel, device).
cudaError_t ensure_dynamic_smem(const void* func, int bytes);  // raises the attribute when a later call asks for more

}  // namespace tt

#define TT_CUDA_TRY(expr)                                                                     \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess) {                                                                  \
      tt::set_error(std::string(#expr) + " failed: " + cudaGetErrorString(_e) + " (" + __FILE__ + ":" + \
                    std::to_string(__LINE__) + ")");                                          \
      return _e;                                                                              \
    }                                                                                         \
  } while (0)

#define TT_LAUNCH_CHECK()                                          \
  do {                                                             \
    tt::count_launch();                                            \
    TT_CUDA_TRY(cudaGetLastError());                               \
  } while (0)
