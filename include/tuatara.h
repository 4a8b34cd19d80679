// Drop-in replacement for the reference's public header (/root/reference/tuatara.h:1-15).
//
// Same names, same return type: `OutputItem {text, bbox}` and
// `std::vector<OutputItem> image_to_data(image, weights_dir, outputs_dir)`.  The reference's
// signature takes a cv::Mat, which forces OpenCV on every caller; here the real parameter type is
// the OpenCV-free `tuatara::ImageView`, and when <opencv2/core.hpp> is available an inline
// cv::Mat overload forwards to it, so reference callers (examples/resume.cpp:9-11) compile
// unchanged.  Header-only on top of the C ABI (tuatara_c.h): link with -ltuatara_b200.
//
// Behavioural contract kept from tuatara.cpp:314-512: synchronous; one item per detected box in
// CCL label order; empty weights_dir / outputs_dir print to stderr and return {} (:315-323);
// engine errors print to stderr and return {} (the reference's soft-failure convention, :337-347).
// Differences, all deliberate: models are loaded once per (weights_dir) and cached instead of on
// every call (:333, :423); the caller's pixels are not modified (:349 swaps them in place);
// zero detections return {} instead of crashing in torch::cat (:485); a box whose boundingRect
// leaves the image is clamped instead of throwing (:416).
#ifndef TUATARA_H
#define TUATARA_H

#include <cstdint>
#include <cstdlib>
#include <iostream>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "tuatara_c.h"

#if defined(__has_include)
#if __has_include(<opencv2/core.hpp>)
#include <opencv2/core.hpp>
#define TUATARA_HAVE_OPENCV 1
#endif
#endif

struct OutputItem {
  std::string text;
  std::vector<float> bbox;  // [min_x, min_y, max_x, max_y]
};

namespace tuatara {

struct ImageView {  // 8-bit, 3 channels, row-major (what cv::imread / np.uint8[H,W,3] hold)
  const uint8_t* data = nullptr;
  int rows = 0, cols = 0, channels = 3;
  size_t step = 0;  // bytes per row; 0 = cols * channels
};

namespace detail {
inline tt_engine* engine_for(const std::string& weights_dir) {
  static std::mutex mu;
  static std::map<std::string, tt_engine*> cache;  // process-wide, keyed by weights_dir
  std::lock_guard<std::mutex> lock(mu);
  auto it = cache.find(weights_dir);
  if (it != cache.end()) return it->second;
  tt_engine* e = nullptr;
  // devices: the TT_DEVICES environment variable when set ("0,2" / "all"), else every visible GPU -- a batch is cut
  // into detection units that all of them pull from one queue; a single page simply runs on the first
  std::vector<int> devs;
  if (std::getenv("TT_DEVICES") == nullptr)
    for (int i = 0; i < tt_device_count(); ++i) devs.push_back(i);
  if (tt_engine_create(weights_dir.c_str(), devs.empty() ? nullptr : devs.data(), static_cast<int>(devs.size()), nullptr, &e) != 0)
    return nullptr;
  cache[weights_dir] = e;
  return e;
}
}  // namespace detail

// Additive batch API: many pages in one call, spread over the engine's GPUs (all visible ones, or TT_DEVICES).
inline std::vector<std::vector<OutputItem>> image_to_data_batch(const std::vector<ImageView>& images,
                                                                const std::string& weights_dir,
                                                                const std::string& outputs_dir) {
  std::vector<std::vector<OutputItem>> out;
  if (weights_dir.empty()) { std::cerr << "Please provide a value for weights_dir" << std::endl; return out; }
  if (outputs_dir.empty()) { std::cerr << "Please provide a value for outputs_dir" << std::endl; return out; }
  std::vector<tt_image> pages(images.size());
  for (size_t i = 0; i < images.size(); ++i) {
    const ImageView& v = images[i];
    if (!v.data || v.rows <= 0 || v.cols <= 0) { std::cerr << "Error reading image from file"; return out; }
    pages[i] = tt_image{v.data, v.rows, v.cols, v.channels, v.step ? v.step : static_cast<size_t>(v.cols) * v.channels};
  }
  tt_engine* e = detail::engine_for(weights_dir);
  if (!e) { std::cerr << "error loading the models: " << tt_last_error() << std::endl; return out; }
  tt_result* r = nullptr;
  if (tt_ocr_pages(e, pages.data(), static_cast<int>(pages.size()), &r) != 0) {
    std::cerr << "tuatara: " << tt_last_error() << std::endl;
    return out;
  }
  out.resize(r->n_pages);
  for (int p = 0; p < r->n_pages; ++p)
    for (int k = 0; k < r->pages[p].n_items; ++k) {
      const tt_item& it = r->pages[p].items[k];
      out[p].push_back(OutputItem{it.text, std::vector<float>(it.bbox, it.bbox + 4)});
    }
  tt_result_free(r);
  return out;
}

}  // namespace tuatara

inline std::vector<OutputItem> image_to_data(const tuatara::ImageView& image, std::string weights_dir,
                                             std::string outputs_dir) {
  auto pages = tuatara::image_to_data_batch({image}, weights_dir, outputs_dir);
  return pages.empty() ? std::vector<OutputItem>{} : pages[0];
}

#ifdef TUATARA_HAVE_OPENCV
// The reference's exact signature (tuatara.h:13).
inline std::vector<OutputItem> image_to_data(cv::Mat image, std::string weights_dir, std::string outputs_dir) {
  if (image.empty()) { std::cerr << "Error reading image from file"; return {}; }  // tuatara.cpp:344-347
  tuatara::ImageView v;
  v.data = image.data; v.rows = image.rows; v.cols = image.cols; v.channels = image.channels(); v.step = image.step;
  return image_to_data(v, weights_dir, outputs_dir);
}
#endif

#endif  // TUATARA_H
