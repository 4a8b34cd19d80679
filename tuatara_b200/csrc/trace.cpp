// Host half of the hang-diagnosis trace (trace.h).  Development aid: active only with TT_TRACE=1.
#include "trace.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <time.h>

namespace tt {

namespace {

struct LaunchRec {
  uint32_t serial = 0;
  char tag[160] = {0};
  int grid = 0, threads = 0;
  size_t smem = 0;
  const void* stream = nullptr;
};

std::mutex g_mu;
uint32_t* g_host = nullptr;
uint32_t* g_dev = nullptr;
bool g_tried = false;
uint32_t g_serial = 0;
LaunchRec g_rec[kTraceLaunches];
// tags of many more launches than the device ring holds (the SM histories name old serials)
constexpr uint32_t kTagRing = 1u << 17;
struct TagRec { uint32_t serial; char tag[92]; };
TagRec* g_tags = nullptr;
std::string g_overwritten;   // launches whose ring slot was reused before all of their CTAs had finished

std::string describe_locked(uint32_t serial, bool* open_out) {
  const LaunchRec& r = g_rec[serial % kTraceLaunches];
  char line[512];
  int open = 0, never = 0;
  std::string detail;
  const uint8_t flag = trace_flag(serial);
  for (int c = 0; c < r.grid && c < kTraceCtas; ++c) {
    const volatile uint32_t* e = g_host + (static_cast<size_t>(serial % kTraceLaunches) * kTraceCtas + c) * kTraceEntry;
    const uint32_t es = e[0], st = e[1];
    if (es != serial) {
      ++never;
      std::snprintf(line, sizeof(line), "    cta %3d: not started\n", c);
      detail += line;
      continue;
    }
    const volatile uint8_t* b = reinterpret_cast<const volatile uint8_t*>(e + 2);
    if (b[kTraceFreeByte] == flag) continue;
    ++open;
    char roles[4 * (kTraceEntry - 2) + 1];
    int nw = (r.threads + 31) / 32;
    if (nw > kTraceFreeByte) nw = kTraceFreeByte;
    for (int w = 0; w < nw; ++w) roles[w] = b[w] == flag ? '1' : '0';
    roles[nw] = 0;
    std::snprintf(line, sizeof(line), "    cta %3d: sm %3u state %u alloc-returned %d roles-done %s\n", c, st >> 8, st & 0xff,
                  static_cast<int>(b[kTraceAllocByte] == flag), roles);
    detail += line;
  }
  *open_out = open != 0 || never != 0;
  if (!*open_out) return std::string();
  std::snprintf(line, sizeof(line), "  launch %u [%s] grid %d threads %d smem %zu stream %p: %d CTAs open, %d not started\n", serial,
                r.tag, r.grid, r.threads, r.smem, r.stream, open, never);
  return line + detail;
}

cudaStream_t g_copy_stream = nullptr;

// The buffer lives in DEVICE memory (marks and atomics cost an L2 access, not a PCIe round trip: the mapped-host
// version slowed the kernels enough to hide the fault it was built to find); trace_report() snapshots it with an async
// copy on a private stream, which runs even while a kernel of another stream is stuck.
void init_locked() {
  if (g_tried) return;
  g_tried = true;
  const char* v = std::getenv("TT_TRACE");
  if (!v || std::atoi(v) == 0) return;
  void* d = nullptr;
  if (cudaMalloc(&d, kTraceWords * sizeof(uint32_t)) != cudaSuccess) return;
  if (cudaMemset(d, 0, kTraceWords * sizeof(uint32_t)) != cudaSuccess) { cudaFree(d); return; }
  void* h = nullptr;
  if (cudaHostAlloc(&h, kTraceWords * sizeof(uint32_t), cudaHostAllocPortable) != cudaSuccess) { cudaFree(d); return; }
  std::memset(h, 0, kTraceWords * sizeof(uint32_t));
  if (cudaStreamCreateWithFlags(&g_copy_stream, cudaStreamNonBlocking) != cudaSuccess) { cudaFree(d); cudaFreeHost(h); return; }
  cudaDeviceSynchronize();
  g_host = static_cast<uint32_t*>(h);
  g_dev = static_cast<uint32_t*>(d);
}

// device buffer -> g_host; false when the copy did not finish within ~5 s
bool snapshot_locked() {
  if (cudaMemcpyAsync(g_host, g_dev, kTraceWords * sizeof(uint32_t), cudaMemcpyDeviceToHost, g_copy_stream) != cudaSuccess) return false;
  for (int i = 0; i < 5000; ++i) {
    const cudaError_t q = cudaStreamQuery(g_copy_stream);
    if (q == cudaSuccess) return true;
    if (q != cudaErrorNotReady) return false;
    struct timespec ts = {0, 1000000};
    nanosleep(&ts, nullptr);
  }
  return false;
}

}  // namespace

uint32_t* trace_dev() {
  std::lock_guard<std::mutex> lock(g_mu);
  init_locked();
  return g_dev;
}

uint32_t trace_launch(const char* tag, int grid, int threads, size_t smem, cudaStream_t s) {
  std::lock_guard<std::mutex> lock(g_mu);
  init_locked();
  if (!g_dev) return 0;
  const uint32_t serial = ++g_serial;
  LaunchRec& r = g_rec[serial % kTraceLaunches];
  r.serial = serial;
  std::snprintf(r.tag, sizeof(r.tag), "%s", tag ? tag : "");
  if (!g_tags) g_tags = new TagRec[kTagRing]();
  g_tags[serial % kTagRing].serial = serial;
  std::snprintf(g_tags[serial % kTagRing].tag, sizeof(g_tags[0].tag), "%s g%d t%d s%p", tag ? tag : "", grid, threads, static_cast<const void*>(s));
  r.grid = grid; r.threads = threads; r.smem = smem; r.stream = s;
  return serial;
}

std::string trace_report() {
  std::lock_guard<std::mutex> lock(g_mu);
  if (!g_host) return "trace off (TT_TRACE=1 enables it)\n";
  std::string out;
  if (!snapshot_locked()) out += "trace: the device->host snapshot did not complete; showing the previous one\n";
  char line[512];
  std::snprintf(line, sizeof(line), "trace: %u launches recorded\n", g_serial);
  out += line;
  const uint32_t first = g_serial > static_cast<uint32_t>(kTraceLaunches) ? g_serial - kTraceLaunches + 1 : 1;
  for (uint32_t serial = first; serial <= g_serial; ++serial) {
    if (g_rec[serial % kTraceLaunches].serial != serial) continue;
    bool open = false;
    out += describe_locked(serial, &open);
  }
  if (!g_overwritten.empty()) out += "  unfinished when their ring slot was reused:\n" + g_overwritten;
  out += "  resident small CTAs (sm: serial/state ...):\n";
  for (int sm = 0; sm < kTraceSms; ++sm) {
    std::string row;
    for (int k = 0; k < 64; ++k) {
      const uint32_t v = g_host[kTraceGemmWords + static_cast<size_t>(sm) * 64 + k];
      if (v == 0) continue;
      std::snprintf(line, sizeof(line), " %u/%u", v >> 8, v & 0xff);
      row += line;
    }
    if (!row.empty()) {
      std::snprintf(line, sizeof(line), "    sm %3d:", sm);
      out += line;
      out += row + "\n";
    }
  }
  out += "  TMEM event history per SM (oldest first; serial:cta:A = alloc returned, :F = dealloc issued):\n";
  for (int sm = 0; sm < kTraceSms; ++sm) {
    const uint32_t cur = g_host[kTraceHistBase + sm];
    if (cur == 0) continue;
    // balance over the recorded window + the last events
    std::string row;
    const uint32_t n = cur < static_cast<uint32_t>(kTraceHist) ? cur : kTraceHist;
    int show = n > 20 ? 20 : static_cast<int>(n);
    for (uint32_t k = cur - show; k < cur; ++k) {
      const uint32_t* e = g_host + kTraceHistBase + kTraceSms + (static_cast<size_t>(sm) * kTraceHist + k % kTraceHist) * 2;
      std::snprintf(line, sizeof(line), " %u:%u:%c", e[0], (e[1] >> 4) & 0xfff, (e[1] & 0xf) == 1 ? 'A' : 'F');
      row += line;
    }
    std::snprintf(line, sizeof(line), "    sm %3d (%u events):", sm, cur);
    out += line;
    out += row + "\n";
    // an SM whose history is far shorter than the others' stopped taking CTAs: name the launches in its tail
    if (g_tags && cur + 2000 < g_host[kTraceHistBase + ((sm + 8) % 128)]) {
      uint32_t last = 0;
      for (uint32_t k = cur - show; k < cur; ++k) {
        const uint32_t* e = g_host + kTraceHistBase + kTraceSms + (static_cast<size_t>(sm) * kTraceHist + k % kTraceHist) * 2;
        if (e[0] == last) continue;
        last = e[0];
        const TagRec& t = g_tags[e[0] % kTagRing];
        std::snprintf(line, sizeof(line), "        launch %u = [%s]\n", e[0], t.serial == e[0] ? t.tag : "?");
        out += line;
      }
    }
  }
  return out;
}

}  // namespace tt
