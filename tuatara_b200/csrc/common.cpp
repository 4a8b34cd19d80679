#include "common.h"

#include <chrono>
#include <cstdlib>
#include <thread>

#include "trace.h"

#include <map>
#include <mutex>
#include <utility>
#include <vector>

namespace tt {

static thread_local std::string t_last_error;
std::atomic<unsigned long long> g_launches{0};

void set_error(const std::string& msg) { t_last_error = msg; }
const char* last_error() { return t_last_error.c_str(); }

std::atomic<unsigned long long> g_h2d_bytes{0}, g_d2h_bytes{0};

namespace {
struct ProfRec { cudaEvent_t a, b; double flops, bytes; std::string tag; };
std::mutex g_prof_mu;
std::vector<ProfRec> g_prof;
std::atomic<bool> g_prof_on{false};
thread_local cudaEvent_t t_open = nullptr;
}  // namespace

void prof_enable(bool on) { g_prof_on.store(on); }
bool prof_enabled() { return g_prof_on.load(std::memory_order_relaxed); }

void prof_record(cudaStream_t s, bool begin, double flops, double bytes, const char* tag) {
  if (!prof_enabled()) return;
  cudaEvent_t ev;
  if (cudaEventCreate(&ev) != cudaSuccess) return;
  cudaEventRecord(ev, s);
  if (begin) { t_open = ev; return; }
  std::lock_guard<std::mutex> lock(g_prof_mu);
  g_prof.push_back(ProfRec{t_open, ev, flops, bytes, tag ? tag : ""});
  t_open = nullptr;
}

void prof_collect(double* total_ms, double* total_flops, double* total_bytes, unsigned long long* launches,
                  const char* dump_path) {
  std::lock_guard<std::mutex> lock(g_prof_mu);
  double ms = 0, fl = 0, by = 0;
  FILE* f = dump_path ? std::fopen(dump_path, "a") : nullptr;
  for (ProfRec& r : g_prof) {
    cudaEventSynchronize(r.b);
    float t = 0.f;
    if (cudaEventElapsedTime(&t, r.a, r.b) == cudaSuccess) { ms += t; fl += r.flops; by += r.bytes; }
    if (f) std::fprintf(f, "%s,%.0f,%.6f\n", r.tag.c_str(), r.flops, t);
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  if (total_ms) *total_ms = ms;
  if (total_flops) *total_flops = fl;
  if (total_bytes) *total_bytes = by;
  if (f) std::fclose(f);
  if (launches) *launches = g_prof.size();
  g_prof.clear();
}

namespace {
std::vector<ProfRec> g_stage;
thread_local cudaEvent_t t_stage_open = nullptr;
}  // namespace

void stage_begin(cudaStream_t s) {
  if (!prof_enabled()) return;
  if (t_stage_open != nullptr) { cudaEventDestroy(t_stage_open); t_stage_open = nullptr; }   // left open by an error path
  cudaEvent_t ev;
  if (cudaEventCreate(&ev) != cudaSuccess) return;
  cudaEventRecord(ev, s);
  t_stage_open = ev;
}

void stage_end(cudaStream_t s, const char* name, double flops, double bytes) {
  if (!prof_enabled() || t_stage_open == nullptr) return;
  cudaEvent_t ev;
  if (cudaEventCreate(&ev) != cudaSuccess) return;
  cudaEventRecord(ev, s);
  std::lock_guard<std::mutex> lock(g_prof_mu);
  g_stage.push_back(ProfRec{t_stage_open, ev, flops, bytes, name});
  t_stage_open = nullptr;
}

std::string stage_collect() {
  std::lock_guard<std::mutex> lock(g_prof_mu);
  struct Agg { int n = 0; double ms = 0, fl = 0, by = 0; };
  std::vector<std::pair<std::string, Agg>> agg;
  for (ProfRec& r : g_stage) {
    cudaEventSynchronize(r.b);
    float t = 0.f;
    const bool ok = cudaEventElapsedTime(&t, r.a, r.b) == cudaSuccess;
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
    if (!ok) continue;
    auto it = agg.begin();
    for (; it != agg.end(); ++it)
      if (it->first == r.tag) break;
    if (it == agg.end()) { agg.emplace_back(r.tag, Agg{}); it = agg.end() - 1; }
    it->second.n += 1; it->second.ms += t; it->second.fl += r.flops; it->second.by += r.bytes;
  }
  g_stage.clear();
  std::string out;
  char line[256];
  for (auto& kv : agg) {
    std::snprintf(line, sizeof(line), "%s,%d,%.6f,%.0f,%.0f\n", kv.first.c_str(), kv.second.n, kv.second.ms, kv.second.fl, kv.second.by);
    out += line;
  }
  return out;
}

cudaError_t ensure_dynamic_smem(const void* func, int bytes) {
  static std::mutex mu;
  static std::map<std::pair<const void*, int>, int> done;   // largest size set so far per (kernel, device)
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lock(mu);
  auto key = std::make_pair(func, dev);
  auto it = done.find(key);
  if (it != done.end() && it->second >= bytes) return cudaSuccess;
  const cudaError_t e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) {
    set_error(std::string("cudaFuncSetAttribute failed: ") + cudaGetErrorString(e));
    return e;
  }
  done[key] = bytes;
  return cudaSuccess;
}

cudaError_t stream_sync(cudaStream_t s) {
  static const double limit_s = std::getenv("TT_WATCHDOG_S") ? std::atof(std::getenv("TT_WATCHDOG_S")) : 120.0;
  if (limit_s <= 0) return cudaStreamSynchronize(s);
  using clock = std::chrono::steady_clock;
  const auto t0 = clock::now();
  for (long long spin = 0;; ++spin) {
    const cudaError_t q = cudaStreamQuery(s);
    if (q == cudaSuccess) return cudaSuccess;
    if (q != cudaErrorNotReady) {
      set_error(std::string("stream failed: ") + cudaGetErrorString(q));
      return q;
    }
    if ((spin & 63) == 63) {
      const double el = std::chrono::duration<double>(clock::now() - t0).count();
      if (el > limit_s) {
        set_error("GPU stall: the stream did not drain within " + std::to_string(static_cast<int>(limit_s)) +
                  " s (TT_WATCHDOG_S); the device may need a reset\n" + trace_report());
        return cudaErrorLaunchTimeout;
      }
      if (el > 2e-3) std::this_thread::sleep_for(std::chrono::microseconds(20));   // busy-poll the first 2 ms, then back off
    }
  }
}

}  // namespace tt
