// Page pipeline and the engine half of the C ABI: image_to_data (tuatara.cpp:314-512) for a batch
// of pages, sharded data-parallel over the engine's GPUs (one host worker thread per device, no
// collective: pages are independent), results gathered on the host in page order.
#include "engine.h"

#include <algorithm>
#include <array>
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
constexpr int kMaxCropsPerPass = 16384;
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
    E_CUDA(cudaStreamSynchronize(d.stream));
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

// One group of equally sized pages on one device.  CRAFT + post-processing run in sub-batches of kCraftSubBatch
// pages (its activations stay L2-friendlier: bigger batches ran 10 % slower per page), PARSeq runs once over the
// crops of the whole group (its 27 decoder passes are launch-latency bound: their cost per page falls with the batch).
int run_group(DeviceCtx& d, const tt_config& cfg, const tt_image* pages, const std::vector<int>& idx,
              const tt_ocr_options& opt, std::vector<PageOut>* results) {
  const cudaMemcpyKind page_kind = opt.pages_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
  const int B = static_cast<int>(idx.size());
  const int SB = std::min(B, kCraftSubBatch);
  const tt_image& first = pages[idx[0]];
  int th, tw, h32, w32;
  float ratio;
  resize_plan(first.rows, first.cols, cfg.canvas_size, cfg.mag_ratio, &th, &tw, &h32, &w32, &ratio);
  const size_t page_bytes = static_cast<size_t>(first.rows) * first.cols * 3;
  const size_t in_bytes = static_cast<size_t>(h32) * w32 * 3;
  const size_t page_stride = (page_bytes + 255) & ~size_t(255);
  const size_t need = d.craft_bytes(SB, h32, w32) + B * (page_stride + 4096) + SB * in_bytes + (8u << 20);
  E_TRY(d.arena.reserve(need));
  d.arena.reset();
  uint8_t* pages_dev = d.arena.get<uint8_t>(B * page_stride);
  if (!pages_dev) { set_error("arena exhausted (pages)"); return 1; }
  const size_t mark = d.arena.offset();
  std::vector<PageRef> refs(B);
  std::vector<std::vector<DetBox>> det(B);
  for (int s0 = 0; s0 < B; s0 += SB) {
    const int nb = std::min(SB, B - s0);
    d.arena.reset_to(mark);
    uint8_t* craft_in = d.arena.get<uint8_t>(nb * in_bytes);
    if (!craft_in) { set_error("arena exhausted (CRAFT input)"); return 1; }
    stage_begin(d.stream);
    for (int b = 0; b < nb; ++b) {
      const tt_image& im = pages[idx[s0 + b]];
      uint8_t* dst = pages_dev + (s0 + b) * page_stride;
      if (opt.pages_on_device) {
        refs[s0 + b] = PageRef{im.data, im.rows, im.cols, im.step};  // already resident: read in place
      } else {
        E_CUDA(cudaMemcpy2DAsync(dst, static_cast<size_t>(im.cols) * 3, im.data, im.step,
                                 static_cast<size_t>(im.cols) * 3, im.rows, page_kind, d.stream));
        g_h2d_bytes += page_bytes;
        refs[s0 + b] = PageRef{dst, im.rows, im.cols, static_cast<size_t>(im.cols) * 3};
      }
      E_TRY(page_resize_pad(refs[s0 + b].data, im.rows, im.cols, refs[s0 + b].step, craft_in + b * in_bytes, th, tw, h32, w32,
                            d.stream));
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
        if (opt.score_override[idx[s0 + b]]) {
          E_CUDA(cudaMemcpyAsync(maps + b * map_elems, opt.score_override[idx[s0 + b]], map_elems * sizeof(float),
                                 opt.override_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, d.stream));
          if (!opt.override_on_device) g_h2d_bytes += map_elems * sizeof(float);
        }
    }
    std::vector<std::vector<DetBox>> sub;
    if (detect_boxes(d, cfg, maps, nb, h32 / 2, w32 / 2, &sub)) return 1;
    for (int b = 0; b < nb; ++b) det[s0 + b] = std::move(sub[b]);
  }

  // host: rescale boxes, bounding rects, output bboxes (tuatara.cpp:406-418, :256-274)
  const float inv = 1.f / ratio;  // ratio_w == ratio_h (tuatara.cpp:360-361)
  std::vector<CropBox> crops;
  for (int b = 0; b < B; ++b) {
    PageOut& po = (*results)[idx[b]];
    const tt_image& im = pages[idx[b]];
    for (const DetBox& db : det[b]) {
      const RotatedRect adj = adjust_rect(db.rect, inv, inv, 2.f);
      std::array<float, 4> bb;
      rect_to_bbox(adj, bb.data());
      po.bbox.push_back(bb);
      const RectI r = rect_bounding(adj);  // the reference throws when this leaves the image; we clamp
      const int x0 = std::max(r.x, 0), y0 = std::max(r.y, 0);
      const int x1 = std::min(r.x + r.w, im.cols), y1 = std::min(r.y + r.h, im.rows);
      crops.push_back(CropBox{b, x0, y0, std::max(x1 - x0, 0), std::max(y1 - y0, 0)});
    }
  }
  const int n = static_cast<int>(crops.size());
  if (n == 0) return 0;  // the reference crashes in torch::cat({}) (tuatara.cpp:485); we return no items

  // crops -> PARSeq, in passes of bounded size; the page buffers stay where they are in the arena,
  // everything CRAFT allocated after them is recycled
  std::vector<int> all_ids(static_cast<size_t>(n) * d.w->pd.L);
  for (int c0 = 0; c0 < n; c0 += kMaxCropsPerPass) {
    const int nc = std::min(kMaxCropsPerPass, n - c0);
    // recycle: keep the page buffers by re-reserving on top of them
    const size_t keep = B * page_stride + 8192;
    const size_t need2 = keep + d.parseq_bytes(nc) + static_cast<size_t>(nc) * (128 * 96 * 2 + sizeof(CropBox)) +
                         B * sizeof(PageRef) + (4u << 20);
    if (need2 > d.arena.capacity()) {
      // growing would move the page buffers: re-upload is simpler than copying device to device
      E_TRY(d.arena.reserve(need2));
      d.arena.reset();
      pages_dev = d.arena.get<uint8_t>(B * page_stride);
      for (int b = 0; b < B; ++b) {
        const tt_image& im = pages[idx[b]];
        uint8_t* dst = pages_dev + b * page_stride;
        if (opt.pages_on_device) continue;
        E_CUDA(cudaMemcpy2DAsync(dst, static_cast<size_t>(im.cols) * 3, im.data, im.step,
                                 static_cast<size_t>(im.cols) * 3, im.rows, page_kind, d.stream));
        refs[b].data = dst;
      }
    } else {
      d.arena.reset_to(mark);
    }
    PageRef* refs_dev = d.arena.get<PageRef>(B);
    CropBox* boxes_dev = d.arena.get<CropBox>(nc);
    __nv_bfloat16* patches = d.arena.get<__nv_bfloat16>(static_cast<size_t>(nc) * 128 * 96);
    if (!refs_dev || !boxes_dev || !patches) { set_error("arena exhausted (crops)"); return 1; }
    E_CUDA(cudaMemcpyAsync(refs_dev, refs.data(), sizeof(PageRef) * B, cudaMemcpyHostToDevice, d.stream));
    E_CUDA(cudaMemcpyAsync(boxes_dev, crops.data() + c0, sizeof(CropBox) * nc, cudaMemcpyHostToDevice, d.stream));
    g_h2d_bytes += sizeof(PageRef) * B + sizeof(CropBox) * nc + sizeof(int) * nc * d.w->pd.L /*token init*/;
    g_d2h_bytes += sizeof(int) * nc * d.w->pd.L;
    stage_begin(d.stream);
    E_TRY(crop_resize(refs_dev, boxes_dev, nc, nullptr, patches, d.stream));
    {
      double src = 0;
      for (int c = c0; c < c0 + nc; ++c) src += 3.0 * crops[c].w * crops[c].h;
      stage_end(d.stream, "crop_resize", 0.0, src + static_cast<double>(nc) * 128 * 96 * 2);  // source rect + bf16 patch rows
    }
    float* logits = nullptr;
    int* ids = nullptr;
    E_TRY(d.parseq_forward(patches, nc, nullptr, &logits, &ids));
    E_CUDA(cudaMemcpyAsync(all_ids.data() + static_cast<size_t>(c0) * d.w->pd.L, ids, sizeof(int) * nc * d.w->pd.L,
                           cudaMemcpyDeviceToHost, d.stream));
    E_CUDA(cudaStreamSynchronize(d.stream));
  }
  // tokenizer (tuatara.cpp:492-505)
  int c = 0;
  for (int b = 0; b < B; ++b) {
    PageOut& po = (*results)[idx[b]];
    for (size_t k = 0; k < det[b].size(); ++k, ++c)
      po.text.push_back(decode_ids(all_ids.data() + static_cast<size_t>(c) * d.w->pd.L, d.w->pd.L));
  }
  return 0;
}

// All pages assigned to one GPU: consecutive pages of identical size form groups of <= max_batch pages;
// group k runs on slot k % kSlotsPerDevice, the slots run concurrently from their own host threads.
int run_device(tt_engine& e, int g, const tt_image* pages, const std::vector<int>& mine, const tt_ocr_options& opt,
               std::vector<PageOut>* results, std::string* err) {
  const tt_config& cfg = e.cfg;
  const int max_b = cfg.max_batch_pages > 0 ? cfg.max_batch_pages : 32;
  std::vector<std::vector<int>> groups;
  size_t i = 0;
  while (i < mine.size()) {
    std::vector<int> grp{mine[i]};
    size_t j = i + 1;
    while (j < mine.size() && static_cast<int>(grp.size()) < max_b && pages[mine[j]].rows == pages[mine[i]].rows &&
           pages[mine[j]].cols == pages[mine[i]].cols) {
      grp.push_back(mine[j]);
      ++j;
    }
    groups.push_back(std::move(grp));
    i = j;
  }
  static const int env_slots = std::getenv("TT_SLOTS") ? std::atoi(std::getenv("TT_SLOTS")) : 0;  // development override
  const int dflt = env_slots > 0 ? env_slots : 1;  // 2 concurrent slots are opt-in (bench.py): see DESIGN "Known issue"
  const int want = std::min(cfg.slots_per_gpu > 0 ? cfg.slots_per_gpu : dflt, kSlotsPerDevice);
  const int S = std::min<int>(want, static_cast<int>(groups.size()));
  std::vector<int> rcs(S, 0);
  std::vector<std::string> errs(S);
  auto slot_main = [&](int sidx) {
    // TT_SLOT_STEAL=1 (development, see DESIGN "Known issue"): take any idle slot instead of queueing on slot sidx
    static const bool steal = std::getenv("TT_SLOT_STEAL") && std::atoi(std::getenv("TT_SLOT_STEAL")) != 0;
    DeviceCtx* dp = nullptr;
    for (int k = 0; steal && k < want && dp == nullptr; ++k) {
      DeviceCtx& c = *e.devs[g * kSlotsPerDevice + (sidx + k) % want];
      if (c.mu.try_lock()) dp = &c;
    }
    if (dp == nullptr) {
      dp = e.devs[g * kSlotsPerDevice + sidx].get();
      dp->mu.lock();
    }
    DeviceCtx& d = *dp;
    std::lock_guard<std::mutex> lock(d.mu, std::adopt_lock);
    if (cudaSetDevice(d.device) != cudaSuccess) { errs[sidx] = "cudaSetDevice failed"; rcs[sidx] = 1; return; }
    for (size_t k = sidx; k < groups.size(); k += S)
      if (run_group(d, cfg, pages, groups[k], opt, results)) { errs[sidx] = last_error(); rcs[sidx] = 1; return; }
  };
  std::vector<std::thread> th;
  for (int sidx = 1; sidx < S; ++sidx) th.emplace_back(slot_main, sidx);
  if (S > 0) slot_main(0);
  for (auto& t : th) t.join();
  for (int sidx = 0; sidx < S; ++sidx)
    if (rcs[sidx]) { *err = errs[sidx]; return 1; }
  return 0;
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
    std::vector<int> devs;
    if (devices && n_devices > 0) devs.assign(devices, devices + n_devices);
    else devs.push_back(0);
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

void tt_engine_set_slots(tt_engine* e, int slots) {
  if (e) e->cfg.slots_per_gpu = slots;
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
    tt_ocr_options opt{0, 0, nullptr};
    if (opt_in) opt = *opt_in;
    if (!e || !out) { set_error("tt_ocr_pages: null argument"); return 1; }
    for (int i = 0; i < n_pages; ++i)
      if (!pages[i].data || pages[i].rows <= 0 || pages[i].cols <= 0 || pages[i].channels != 3) {
        set_error("Error reading image from file");  // tuatara.cpp:344-347
        return 1;
      }
    std::vector<PageOut> results(n_pages);
    const int G = e->n_devices;
    std::vector<std::vector<int>> shard(G);
    for (int i = 0; i < n_pages; ++i) shard[i % G].push_back(i);  // page i -> GPU i mod G
    std::vector<std::string> errs(G);
    std::vector<int> rcs(G, 0);
    std::vector<std::thread> workers;
    for (int g = 0; g < G; ++g) {
      if (shard[g].empty()) continue;
      workers.emplace_back([&, g] { rcs[g] = run_device(*e, g, pages, shard[g], opt, &results, &errs[g]); });
    }
    for (auto& w : workers) w.join();
    for (int g = 0; g < G; ++g)
      if (rcs[g]) { set_error("device " + std::to_string(e->devs[g * kSlotsPerDevice]->device) + ": " + errs[g]); return 1; }
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
    E_CUDA(cudaStreamSynchronize(d.stream));
    return 0;
  } catch (const std::exception& ex) {
    set_error(std::string("tt_craft_forward: ") + ex.what());
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
      E_CUDA(cudaStreamSynchronize(d.stream));
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
