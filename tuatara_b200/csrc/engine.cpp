// Page pipeline and the engine half of the C ABI: image_to_data (tuatara.cpp:314-512) for a batch
// of pages, sharded data-parallel over the engine's GPUs (one host worker thread per device, no
// collective: pages are independent), results gathered on the host in page order.
#include "engine.h"

#include <algorithm>
#include <array>
#include <atomic>
#include <map>
#include <tuple>
#include <cstdlib>
#include <cstring>
#include <thread>

#include "common.h"
#include "gemm_tc.cuh"
#include "nn_kernels.cuh"
#include "resize.cuh"
#include "tokenizer.h"

using namespace tt;

namespace {

constexpr int kCompCap = 2048;   // compact per-page result block: components
constexpr int kRowCap = 32768;   //                                row extents
constexpr int kMaxCropsPerPass = 32768;
constexpr int kCraftSubBatch = 8;   // pages per CRAFT / post-processing pass inside a group

#define E_TRY(expr)                                   \
  do {                                                \
    cudaError_t _e = (expr);                          \
    if (_e != cudaSuccess) return 1;                  \
  } while (0)
#define E_CUDA(expr)                                                               \
  do {                                                                             \
    cudaError_t _e = (expr);                                                       \
    if (_e != cudaSuccess) {                                                       \
      set_error(std::string(#expr) + " failed: " + cudaGetErrorString(_e));        \
      return 1;                                                                    \
    }                                                                              \
  } while (0)

struct PageOut {
  std::vector<std::string> text;
  std::vector<std::array<float, 4>> bbox;
};

cudaError_t ensure_post(DeviceCtx& d, int batch, int H, int W, int comp_cap, int row_cap) {
  if (d.post.parent && d.post.batch >= batch && d.post.H == H && d.post.W == W && d.post.comp_cap == comp_cap &&
      d.post.row_cap == row_cap)
    return cudaSuccess;
  post_workspace_free(&d.post);
  return post_workspace_alloc(&d.post, batch, H, W, comp_cap, row_cap, false);
}

// Runs post-processing for `batch` maps on the device and returns the boxes per page (host).
int detect_boxes(DeviceCtx& d, const tt_config& cfg, const float* maps_dev, int batch, int H, int W,
                 std::vector<std::vector<DetBox>>* out) {
  int comp_cap = kCompCap, row_cap = kRowCap;
  for (int attempt = 0; attempt < 2; ++attempt) {
    E_TRY(ensure_post(d, batch, H, W, comp_cap, row_cap));
    PostWorkspace ws = d.post;
    ws.batch = batch;
    PostParams pp;
    pp.low_text = cfg.low_text;
    pp.link_threshold = cfg.link_threshold;
    const size_t bytes = ws.result_stride * batch;
    E_TRY(d.ensure_pinned(bytes));
    stage_begin(d.stream);
    E_TRY(post_run(ws, maps_dev, pp, d.stream));
    E_CUDA(cudaMemcpyAsync(d.pinned, ws.result, bytes, cudaMemcpyDeviceToHost, d.stream));
    stage_end(d.stream, "postprocess", 0.0, 24.0 * batch * H * W);  // 24 B per map pixel; kernels + the result D2H
    E_TRY(stream_sync(d.stream));
    g_d2h_bytes += bytes;
    out->assign(batch, {});
    bool ok = true;
    for (int b = 0; b < batch && ok; ++b)
      ok = collect_boxes(d.pinned + b * ws.result_stride, comp_cap, row_cap, H, W, cfg, &(*out)[b]);
    if (ok) return 0;
    comp_cap = H * W / 2 + 2;  // pathological map: worst-case capacities
    row_cap = H * W;
  }
  set_error("post-process: capacity overflow");
  return 1;
}

// ---------------------------------------------------------------------------------------------------------------
// Scheduling.  A request is cut into DETECTION UNITS: up to kCraftSubBatch pages of one size (pages of equal
// size are bucketed wherever they sit in the request).  Units go into work queues that the execution slots of all GPUs
// drain with an atomic counter: words per page vary, so static page -> GPU assignment would leave GPUs idle.  A slot
//   1. runs a unit: upload, resize, CRAFT, post-processing, one small D2H, host geometry, then the crop kernel, which
//      appends the unit's crops (bf16 patch rows) to the slot's RECOGNITION BATCH -- pages of any size mix there;
//   2. runs PARSeq over the batch once it holds `max_batch_pages` pages' worth of crops (or the queues are empty):
//      the decoder's 27 passes are latency bound, their cost per crop falls with the batch.
// Two slots per GPU by default: one slot's host phases (geometry, result assembly, D2H waits) are covered by the other
// slot's kernels.
struct Unit {
  std::vector<int> pages;   // indices into the request, all of one size
};

struct CropOwner { int page, box; };

// crops waiting for recognition in one slot (device patch rows live in DeviceCtx::patch_buf)
struct Pending {
  std::vector<CropOwner> owner;
  int pages = 0;
};

cudaError_t ensure_patch_buf(DeviceCtx& d, size_t crops, size_t keep_crops) {
  if (crops <= d.patch_cap) return cudaSuccess;
  size_t cap = std::max<size_t>(crops + crops / 2, 4096);
  __nv_bfloat16* nb = nullptr;
  TT_CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&nb), cap * 128 * 96 * sizeof(__nv_bfloat16)));
  if (d.patch_buf && keep_crops)
    TT_CUDA_TRY(cudaMemcpyAsync(nb, d.patch_buf, keep_crops * 128 * 96 * sizeof(__nv_bfloat16), cudaMemcpyDeviceToDevice, d.stream));
  if (d.patch_buf) {
    TT_CUDA_TRY(stream_sync(d.stream));
    cudaFree(d.patch_buf);
  }
  d.patch_buf = nb;
  d.patch_cap = cap;
  return cudaSuccess;
}

// Detection + cropping of one unit (tuatara.cpp:349-418, :437-441).  Fills bbox of every page of the unit, sizes its
// text vector, appends the crops to the slot's recognition batch.
int detect_unit(DeviceCtx& d, const tt_config& cfg, const tt_image* pages, const Unit& u, const tt_ocr_options& opt,
                std::vector<PageOut>* results, Pending* pend) {
  const cudaMemcpyKind page_kind = opt.pages_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
  const int nb = static_cast<int>(u.pages.size());
  const tt_image& first = pages[u.pages[0]];
  int th, tw, h32, w32;
  float ratio;
  resize_plan(first.rows, first.cols, cfg.canvas_size, cfg.mag_ratio, &th, &tw, &h32, &w32, &ratio);
  const size_t page_bytes = static_cast<size_t>(first.rows) * first.cols * 3;
  const size_t in_bytes = static_cast<size_t>(h32) * w32 * 3;
  const size_t page_stride = (page_bytes + 255) & ~size_t(255);
  const size_t need = d.craft_bytes(nb, h32, w32) + nb * (page_stride + 4096 + in_bytes) + (8u << 20);
  if (need > d.arena.capacity()) E_TRY(stream_sync(d.stream));   // growing frees the old block: nothing may still use it
  E_TRY(d.arena.reserve(need));
  d.arena.reset();
  uint8_t* pages_dev = d.arena.get<uint8_t>(nb * page_stride);
  uint8_t* craft_in = d.arena.get<uint8_t>(nb * in_bytes);
  if (!pages_dev || !craft_in) { set_error("arena exhausted (pages)"); return 1; }
  std::vector<PageRef> refs(nb);
  stage_begin(d.stream);
  for (int b = 0; b < nb; ++b) {
    const tt_image& im = pages[u.pages[b]];
    uint8_t* dst = pages_dev + b * page_stride;
    if (opt.pages_on_device) {
      refs[b] = PageRef{im.data, im.rows, im.cols, im.step};  // already resident: read in place
    } else {
      E_CUDA(cudaMemcpy2DAsync(dst, static_cast<size_t>(im.cols) * 3, im.data, im.step, static_cast<size_t>(im.cols) * 3, im.rows,
                               page_kind, d.stream));
      g_h2d_bytes += page_bytes;
      refs[b] = PageRef{dst, im.rows, im.cols, static_cast<size_t>(im.cols) * 3};
    }
    E_TRY(page_resize_pad(refs[b].data, im.rows, im.cols, refs[b].step, craft_in + b * in_bytes, th, tw, h32, w32, d.stream));
  }
  stage_end(d.stream, "preprocess", 0.0, static_cast<double>(nb) * (page_bytes + in_bytes));
  float* maps = nullptr;
  stage_begin(d.stream);
  E_TRY(d.craft_forward(craft_in, nb, h32, w32, &maps));
  // 27 convolutions of CRAFT: 711.4 FLOP per input pixel (= 746.0 GFLOP at 1024 x 1024)
  stage_end(d.stream, "craft", 746.0e9 / (1024.0 * 1024.0) * nb * h32 * w32, 0.0);
  if (opt.score_override) {
    const size_t map_elems = static_cast<size_t>(h32 / 2) * (w32 / 2) * 2;
    for (int b = 0; b < nb; ++b)
      if (opt.score_override[u.pages[b]]) {
        E_CUDA(cudaMemcpyAsync(maps + b * map_elems, opt.score_override[u.pages[b]], map_elems * sizeof(float),
                               opt.override_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, d.stream));
        if (!opt.override_on_device) g_h2d_bytes += map_elems * sizeof(float);
      }
  }
  std::vector<std::vector<DetBox>> det;
  if (detect_boxes(d, cfg, maps, nb, h32 / 2, w32 / 2, &det)) return 1;

  // host: rescale boxes, bounding rects, output bboxes (tuatara.cpp:406-418, :256-274)
  const float inv = 1.f / ratio;  // ratio_w == ratio_h (tuatara.cpp:360-361)
  std::vector<CropBox> crops;
  std::vector<WarpBox> warps;   // cfg.rectify: one perspective warp per box instead of the axis-aligned crop
  const bool rectify = cfg.rectify != 0;
  for (int b = 0; b < nb; ++b) {
    PageOut& po = (*results)[u.pages[b]];
    const tt_image& im = pages[u.pages[b]];
    int k = 0;
    for (const DetBox& db : det[b]) {
      const RotatedRect adj = adjust_rect(db.rect, inv, inv, 2.f);
      std::array<float, 4> bb;
      rect_to_bbox(adj, bb.data());
      po.bbox.push_back(bb);
      const RectI r = rect_bounding(adj);  // the reference throws when this leaves the image; we clamp
      const int x0 = std::max(r.x, 0), y0 = std::max(r.y, 0);
      const int x1 = std::min(r.x + r.w, im.cols), y1 = std::min(r.y + r.h, im.rows);
      crops.push_back(CropBox{b, x0, y0, std::max(x1 - x0, 0), std::max(y1 - y0, 0)});
      if (rectify) {
        Pt2f quad[4];
        rect_to_quad(adj, quad);
        WarpBox wb{};
        wb.page = quad_to_warp(quad, wb.m) ? b : -1;
        warps.push_back(wb);
      }
      pend->owner.push_back(CropOwner{u.pages[b], k++});
    }
    po.text.resize(po.bbox.size());   // zero boxes: the reference crashes in torch::cat({}) (tuatara.cpp:485); we return no items
  }
  const int n = static_cast<int>(crops.size());
  if (opt.detect_only) {   // boxes only: nothing joins the recognition batch
    pend->owner.resize(pend->owner.size() - n);
    return 0;
  }
  pend->pages += nb;
  if (n == 0) return 0;
  const size_t have = pend->owner.size() - n;
  E_TRY(ensure_patch_buf(d, have + n, have));
  PageRef* refs_dev = d.arena.get<PageRef>(nb);
  CropBox* boxes_dev = d.arena.get<CropBox>(n);
  if (!refs_dev || !boxes_dev) { set_error("arena exhausted (crops)"); return 1; }
  E_CUDA(cudaMemcpyAsync(refs_dev, refs.data(), sizeof(PageRef) * nb, cudaMemcpyHostToDevice, d.stream));
  E_CUDA(cudaMemcpyAsync(boxes_dev, crops.data(), sizeof(CropBox) * n, cudaMemcpyHostToDevice, d.stream));
  g_h2d_bytes += sizeof(PageRef) * nb + sizeof(CropBox) * n;
  stage_begin(d.stream);
  if (rectify) {
    WarpBox* warps_dev = d.arena.get<WarpBox>(n);
    if (!warps_dev) { set_error("arena exhausted (warps)"); return 1; }
    E_CUDA(cudaMemcpyAsync(warps_dev, warps.data(), sizeof(WarpBox) * n, cudaMemcpyHostToDevice, d.stream));
    g_h2d_bytes += sizeof(WarpBox) * n;
    E_TRY(crop_warp(refs_dev, warps_dev, n, nullptr, d.patch_buf + have * 128 * 96, d.stream));
  } else {
    E_TRY(crop_resize(refs_dev, boxes_dev, n, nullptr, d.patch_buf + have * 128 * 96, d.stream));
  }
  double src = 0;
  for (const CropBox& c : crops) src += 3.0 * c.w * c.h;
  stage_end(d.stream, "crop_resize", 0.0, src + static_cast<double>(n) * 128 * 96 * 2);  // source rect + bf16 patch rows
  return 0;
}

// PARSeq + tokenizer over the slot's recognition batch (tuatara.cpp:450-505).
int recognise(DeviceCtx& d, std::vector<PageOut>* results, Pending* pend) {
  const int n = static_cast<int>(pend->owner.size());
  const int L = d.w->pd.L;
  for (int c0 = 0; c0 < n; c0 += kMaxCropsPerPass) {
    const int nc = std::min(kMaxCropsPerPass, n - c0);
    const size_t need = d.parseq_bytes(nc) + (4u << 20);
    if (need > d.arena.capacity()) E_TRY(stream_sync(d.stream));
    E_TRY(d.arena.reserve(need));
    d.arena.reset();
    float* logits = nullptr;
    int* ids = nullptr;
    E_TRY(d.parseq_forward(d.patch_buf + static_cast<size_t>(c0) * 128 * 96, nc, nullptr, &logits, &ids));
    const size_t bytes = sizeof(int) * static_cast<size_t>(nc) * L;
    E_TRY(d.ensure_pinned(bytes));
    E_CUDA(cudaMemcpyAsync(d.pinned, ids, bytes, cudaMemcpyDeviceToHost, d.stream));
    g_d2h_bytes += bytes;
    E_TRY(stream_sync(d.stream));
    const int* host_ids = reinterpret_cast<const int*>(d.pinned);
    for (int c = 0; c < nc; ++c) {   // tokenizer (tuatara.cpp:492-505)
      const CropOwner& o = pend->owner[c0 + c];
      (*results)[o.page].text[o.box] = decode_ids(host_ids + static_cast<size_t>(c) * L, L);
    }
  }
  pend->owner.clear();
  pend->pages = 0;
  return 0;
}

struct WorkQueues {
  std::vector<Unit> units;
  std::vector<std::vector<int>> per_dev;   // unit indices bound to a device (pages resident there)
  std::vector<int> shared;                 // unit indices any device may take (host pages)
  std::vector<std::atomic<int>> next_dev;
  std::atomic<int> next_shared{0};
  explicit WorkQueues(int G) : per_dev(G), next_dev(G) { for (auto& a : next_dev) a.store(0); }
  // next unit for a slot of device g, or -1
  int pop(int g) {
    const int i = next_dev[g].fetch_add(1);
    if (i < static_cast<int>(per_dev[g].size())) return per_dev[g][i];
    const int j = next_shared.fetch_add(1);
    if (j < static_cast<int>(shared.size())) return shared[j];
    return -1;
  }
};

// One execution slot draining the queues.
void slot_worker(tt_engine& e, int g, int sidx, int want, const tt_image* pages, const tt_ocr_options& opt, WorkQueues* q,
                 std::vector<PageOut>* results, int* rc, std::string* err) {
  // take an idle slot of this GPU if there is one (concurrent callers spread over the slots), else queue on our own
  DeviceCtx* dp = nullptr;
  for (int k = 0; k < want && dp == nullptr; ++k) {
    DeviceCtx& c = *e.devs[g * kSlotsPerDevice + (sidx + k) % want];
    if (c.mu.try_lock()) dp = &c;
  }
  if (dp == nullptr) {
    dp = e.devs[g * kSlotsPerDevice + sidx].get();
    dp->mu.lock();
  }
  DeviceCtx& d = *dp;
  std::lock_guard<std::mutex> lock(d.mu, std::adopt_lock);
  if (cudaSetDevice(d.device) != cudaSuccess) { *err = "cudaSetDevice failed"; *rc = 1; return; }
  const tt_config& cfg = e.cfg;
  const int max_b = cfg.max_batch_pages > 0 ? cfg.max_batch_pages : 32;
  Pending pend;
  for (int ui = q->pop(g); ui >= 0; ui = q->pop(g)) {
    if (detect_unit(d, cfg, pages, q->units[ui], opt, results, &pend)) { *err = last_error(); *rc = 1; return; }
    if (pend.pages >= max_b && recognise(d, results, &pend)) { *err = last_error(); *rc = 1; return; }
  }
  if (recognise(d, results, &pend)) { *err = last_error(); *rc = 1; return; }
}

}  // namespace

extern "C" {

int tt_engine_create(const char* weights_dir, const int* devices, int n_devices, const tt_config* cfg,
                     tt_engine** out) {
  try {
    if (!weights_dir || !*weights_dir) { set_error("Please provide a value for weights_dir"); return 1; }
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
      set_error("no CUDA device: this library has no CPU fallback");
      return 1;
    }
    std::unique_ptr<tt_engine> e(new tt_engine);
    if (cfg) e->cfg = *cfg; else tt_config_default(&e->cfg);
    e->slots.store(e->cfg.slots_per_gpu);
    std::vector<int> devs;
    if (devices && n_devices > 0) {
      devs.assign(devices, devices + n_devices);
    } else if (const char* ev = std::getenv("TT_DEVICES")) {   // "all" or a comma-separated list
      if (std::strcmp(ev, "all") == 0) {
        for (int i = 0; i < count; ++i) devs.push_back(i);
      } else {
        for (const char* p = ev; *p;) {
          char* end = nullptr;
          const long v = std::strtol(p, &end, 10);
          if (end == p) { set_error(std::string("TT_DEVICES: cannot parse '") + ev + "'"); return 1; }
          devs.push_back(static_cast<int>(v));
          p = (*end == ',') ? end + 1 : end;
        }
      }
    }
    if (devs.empty()) devs.push_back(0);
    for (int dv : devs) {
      if (dv < 0 || dv >= count) { set_error("invalid device index " + std::to_string(dv)); return 1; }
      std::shared_ptr<DeviceWeights> shared;
      for (int sidx = 0; sidx < kSlotsPerDevice; ++sidx) {
        std::unique_ptr<DeviceCtx> d(new DeviceCtx);
        d->device = dv;
        if (d->init(weights_dir, shared) != cudaSuccess) return 1;
        shared = d->w;
        e->devs.push_back(std::move(d));
      }
    }
    e->n_devices = static_cast<int>(devs.size());
    *out = e.release();
    return 0;
  } catch (const std::exception& ex) {
    set_error(std::string("tt_engine_create: ") + ex.what());
    return 1;
  }
}

void tt_engine_destroy(tt_engine* e) { delete e; }

int tt_device_count(void) {
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess) { cudaGetLastError(); return 0; }
  return count;
}

void tt_engine_set_slots(tt_engine* e, int slots) {
  if (e) e->slots.store(slots);
}

void* tt_engine_stream(tt_engine* e, int idx) {
  if (!e || idx < 0 || idx >= e->n_devices) return nullptr;
  return e->devs[idx * kSlotsPerDevice]->stream;
}
void tt_io_bytes(unsigned long long* h2d, unsigned long long* d2h) {
  if (h2d) *h2d = g_h2d_bytes.load();
  if (d2h) *d2h = g_d2h_bytes.load();
}
void tt_profile_enable(int on) { prof_enable(on != 0); }
void tt_profile_collect(double* total_ms, double* total_flops, unsigned long long* launches) {
  prof_collect(total_ms, total_flops, nullptr, launches);
}
int tt_profile_stages(char* buf, int cap) {
  const std::string csv = stage_collect();
  if (!buf || cap <= 0) return static_cast<int>(csv.size());
  const size_t n = std::min(csv.size(), static_cast<size_t>(cap - 1));
  std::memcpy(buf, csv.data(), n);
  buf[n] = 0;
  return static_cast<int>(csv.size());
}
void tt_profile_dump(const char* path, double* total_ms, double* total_flops, unsigned long long* launches) {
  prof_collect(total_ms, total_flops, nullptr, launches, path);
}

int tt_ocr_pages(tt_engine* e, const tt_image* pages, int n_pages, tt_result** out) {
  return tt_ocr_pages_ex(e, pages, n_pages, nullptr, out);
}

int tt_ocr_pages_ex(tt_engine* e, const tt_image* pages, int n_pages, const tt_ocr_options* opt_in, tt_result** out) {
  try {
    tt_ocr_options opt{0, 0, nullptr, 0};
    if (opt_in) opt = *opt_in;
    if (!e || !out) { set_error("tt_ocr_pages: null argument"); return 1; }
    if (n_pages < 0 || (n_pages > 0 && !pages)) { set_error("tt_ocr_pages: bad page array"); return 1; }
    for (int i = 0; i < n_pages; ++i) {
      if (!pages[i].data || pages[i].rows <= 0 || pages[i].cols <= 0 || pages[i].channels != 3) {
        set_error("Error reading image from file");  // tuatara.cpp:344-347
        return 1;
      }
      if (pages[i].step < static_cast<size_t>(pages[i].cols) * 3) { set_error("tt_ocr_pages: row step smaller than cols * 3"); return 1; }
    }
    std::vector<PageOut> results(n_pages);
    const int G = e->n_devices;
    // detection units: pages of equal size, wherever they sit in the request; device-resident pages stay on their GPU
    WorkQueues q(G);
    {
      std::map<std::tuple<int, int, int>, int> open;   // (device or -1, rows, cols) -> unit being filled
      for (int i = 0; i < n_pages; ++i) {
        const int dev = opt.pages_on_device ? i % G : -1;
        const auto key = std::make_tuple(dev, pages[i].rows, pages[i].cols);
        auto it = open.find(key);
        if (it == open.end() || static_cast<int>(q.units[it->second].pages.size()) >= kCraftSubBatch) {
          q.units.emplace_back();
          const int ui = static_cast<int>(q.units.size()) - 1;
          open[key] = ui;
          if (dev >= 0) q.per_dev[dev].push_back(ui); else q.shared.push_back(ui);
          it = open.find(key);
        }
        q.units[it->second].pages.push_back(i);
      }
    }
    static const int env_slots = std::getenv("TT_SLOTS") ? std::atoi(std::getenv("TT_SLOTS")) : 0;  // development override
    const int cfg_slots = e->slots.load();
    const int want = std::max(1, std::min(cfg_slots > 0 ? cfg_slots : env_slots > 0 ? env_slots : 3, kSlotsPerDevice));
    // workers: device-major round robin, so that a short request touches every GPU before any second slot
    struct W { int g, sidx; };
    std::vector<W> ws;
    size_t shared_claimed = 0;
    for (int sidx = 0; sidx < want; ++sidx)
      for (int g = 0; g < G; ++g) {
        if (static_cast<size_t>(sidx) < q.per_dev[g].size()) ws.push_back(W{g, sidx});
        else if (shared_claimed < q.shared.size()) { ++shared_claimed; ws.push_back(W{g, sidx}); }
      }
    if (ws.empty() && n_pages > 0) ws.push_back(W{0, 0});
    std::vector<std::string> errs(ws.size());
    std::vector<int> rcs(ws.size(), 0);
    std::vector<std::thread> workers;
    for (size_t w = 1; w < ws.size(); ++w)
      workers.emplace_back([&, w] { slot_worker(*e, ws[w].g, ws[w].sidx, want, pages, opt, &q, &results, &rcs[w], &errs[w]); });
    if (!ws.empty()) slot_worker(*e, ws[0].g, ws[0].sidx, want, pages, opt, &q, &results, &rcs[0], &errs[0]);
    for (auto& w : workers) w.join();
    for (size_t w = 0; w < ws.size(); ++w)
      if (rcs[w]) { set_error("device " + std::to_string(e->devs[ws[w].g * kSlotsPerDevice]->device) + ": " + errs[w]); return 1; }
    // host-side gather into the C result
    tt_result* r = new tt_result;
    r->n_pages = n_pages;
    r->pages = new tt_page_result[n_pages];
    for (int i = 0; i < n_pages; ++i) {
      const PageOut& po = results[i];
      tt_page_result& pr = r->pages[i];
      pr.n_items = static_cast<int>(po.text.size());
      pr.items = pr.n_items ? new tt_item[pr.n_items] : nullptr;
      for (int k = 0; k < pr.n_items; ++k) {
        pr.items[k].text = new char[po.text[k].size() + 1];
        std::memcpy(pr.items[k].text, po.text[k].c_str(), po.text[k].size() + 1);
        std::memcpy(pr.items[k].bbox, po.bbox[k].data(), sizeof(float) * 4);
      }
    }
    *out = r;
    return 0;
  } catch (const std::exception& ex) {
    set_error(std::string("tt_ocr_pages: ") + ex.what());
    return 1;
  }
}

void tt_result_free(tt_result* r) {
  if (!r) return;
  for (int i = 0; i < r->n_pages; ++i) {
    for (int k = 0; k < r->pages[i].n_items; ++k) delete[] r->pages[i].items[k].text;
    delete[] r->pages[i].items;
  }
  delete[] r->pages;
  delete r;
}

int tt_craft_forward(tt_engine* e, const uint8_t* craft_input, int h32, int w32, float* maps_out) {
  try {
    if (!e) { set_error("tt_craft_forward: null engine"); return 1; }
    DeviceCtx& d = *e->devs[0];
    std::lock_guard<std::mutex> lock(d.mu);
    E_CUDA(cudaSetDevice(d.device));
    const size_t in_bytes = static_cast<size_t>(h32) * w32 * 3;
    E_TRY(d.arena.reserve(d.craft_bytes(1, h32, w32) + in_bytes + (8u << 20)));
    d.arena.reset();
    uint8_t* in = d.arena.get<uint8_t>(in_bytes);
    E_CUDA(cudaMemcpyAsync(in, craft_input, in_bytes, cudaMemcpyHostToDevice, d.stream));
    float* maps = nullptr;
    E_TRY(d.craft_forward(in, 1, h32, w32, &maps));
    E_CUDA(cudaMemcpyAsync(maps_out, maps, sizeof(float) * (h32 / 2) * (w32 / 2) * 2, cudaMemcpyDeviceToHost, d.stream));
    E_TRY(stream_sync(d.stream));
    return 0;
  } catch (const std::exception& ex) {
    set_error(std::string("tt_craft_forward: ") + ex.what());
    return 1;
  }
}

int tt_craft_tap(tt_engine* e, const char* name, float* out, long long capacity, int dims_out[3]) {
  try {
    if (!e || !name) { set_error("tt_craft_tap: null argument"); return 1; }
    DeviceCtx& d = *e->devs[0];
    std::lock_guard<std::mutex> lock(d.mu);
    E_CUDA(cudaSetDevice(d.device));
    for (const CraftTap& t : d.craft_taps) {
      if (std::strcmp(t.name, name) != 0) continue;
      const long long n = static_cast<long long>(t.H) * t.W * t.C;
      if (dims_out) { dims_out[0] = t.H; dims_out[1] = t.W; dims_out[2] = t.C; }
      if (!out) return 0;
      if (capacity < n) { set_error("tt_craft_tap: output buffer too small"); return 1; }
      std::vector<__nv_bfloat16> h(static_cast<size_t>(n));
      E_CUDA(cudaMemcpy(h.data(), t.ptr, sizeof(__nv_bfloat16) * n, cudaMemcpyDeviceToHost));
      for (long long i = 0; i < n; ++i) out[i] = __bfloat162float(h[static_cast<size_t>(i)]);
      return 0;
    }
    set_error(std::string("tt_craft_tap: no activation named '") + name + "' (run tt_craft_forward first)");
    return 1;
  } catch (const std::exception& ex) {
    set_error(std::string("tt_craft_tap: ") + ex.what());
    return 1;
  }
}

int tt_parseq_forward(tt_engine* e, const uint8_t* crops, int n, const int32_t* forced_tokens, float* logits_out,
                      int32_t* ids_out) {
  try {
    if (!e) { set_error("tt_parseq_forward: null engine"); return 1; }
    if (n <= 0) return 0;
    DeviceCtx& d = *e->devs[0];
    std::lock_guard<std::mutex> lock(d.mu);
    E_CUDA(cudaSetDevice(d.device));
    const int L = d.w->pd.L, NC = d.w->pd.n_cls_pad;
    for (int c0 = 0; c0 < n; c0 += kMaxCropsPerPass) {
      const int nc = std::min(kMaxCropsPerPass, n - c0);
      const size_t crop_bytes = static_cast<size_t>(nc) * 32 * 128 * 3;
      E_TRY(d.arena.reserve(d.parseq_bytes(nc) + crop_bytes + static_cast<size_t>(nc) * (128 * 96 * 2 + 4 * L) + (8u << 20)));
      d.arena.reset();
      uint8_t* cu8 = d.arena.get<uint8_t>(crop_bytes);
      __nv_bfloat16* patches = d.arena.get<__nv_bfloat16>(static_cast<size_t>(nc) * 128 * 96);
      int* forced = forced_tokens ? d.arena.get<int>(static_cast<size_t>(nc) * (L - 1)) : nullptr;
      E_CUDA(cudaMemcpyAsync(cu8, crops + static_cast<size_t>(c0) * 32 * 128 * 3, crop_bytes, cudaMemcpyHostToDevice, d.stream));
      if (forced)
        E_CUDA(cudaMemcpyAsync(forced, forced_tokens + static_cast<size_t>(c0) * (L - 1), sizeof(int) * nc * (L - 1),
                               cudaMemcpyHostToDevice, d.stream));
      E_TRY(patchify_u8(cu8, nc, patches, d.stream));
      float* logits = nullptr;
      int* ids = nullptr;
      E_TRY(d.parseq_forward(patches, nc, forced, &logits, &ids));
      if (logits_out)
        E_CUDA(cudaMemcpy2DAsync(logits_out + static_cast<size_t>(c0) * L * d.w->pd.n_cls, sizeof(float) * d.w->pd.n_cls, logits,
                                 sizeof(float) * NC, sizeof(float) * d.w->pd.n_cls, static_cast<size_t>(nc) * L,
                                 cudaMemcpyDeviceToHost, d.stream));
      if (ids_out)
        E_CUDA(cudaMemcpyAsync(ids_out + static_cast<size_t>(c0) * L, ids, sizeof(int) * nc * L, cudaMemcpyDeviceToHost, d.stream));
      E_TRY(stream_sync(d.stream));
    }
    return 0;
  } catch (const std::exception& ex) {
    set_error(std::string("tt_parseq_forward: ") + ex.what());
    return 1;
  }
}

int tt_postprocess_dev(tt_engine* e, const float* maps_dev, int batch, int H, int W, int* n_rects_total, void*) {
  try {
    if (!e) { set_error("tt_postprocess_dev: null engine"); return 1; }
    DeviceCtx& d = *e->devs[0];
    std::lock_guard<std::mutex> lock(d.mu);
    E_CUDA(cudaSetDevice(d.device));
    std::vector<std::vector<DetBox>> det;
    if (detect_boxes(d, e->cfg, maps_dev, batch, H, W, &det)) return 1;
    int total = 0;
    for (auto& v : det) total += static_cast<int>(v.size());
    if (n_rects_total) *n_rects_total = total;
    return 0;
  } catch (const std::exception& ex) {
    set_error(std::string("tt_postprocess_dev: ") + ex.what());
    return 1;
  }
}

}  // extern "C"
