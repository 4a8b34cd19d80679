#include "common.h"

#include <map>
#include <mutex>
#include <utility>

namespace tt {

static thread_local std::string t_last_error;
std::atomic<unsigned long long> g_launches{0};

void set_error(const std::string& msg) { t_last_error = msg; }
const char* last_error() { return t_last_error.c_str(); }

cudaError_t ensure_dynamic_smem(const void* func, int bytes) {
  static std::mutex mu;
  static std::map<std::pair<const void*, int>, cudaError_t> done;
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lock(mu);
  auto key = std::make_pair(func, dev);
  auto it = done.find(key);
  if (it != done.end()) return it->second;
  const cudaError_t e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) set_error(std::string("cudaFuncSetAttribute failed: ") + cudaGetErrorString(e));
  done[key] = e;
  return e;
}

}  // namespace tt
