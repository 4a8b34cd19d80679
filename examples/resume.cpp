// The reference's examples/resume.cpp on the B200 path: same call, same argument order
//   resume <image> <weights_dir> <outputs_dir>
// (reference: examples/resume.cpp:7-13 -- imread, image_to_data(image, argv[2], argv[3])).  With OpenCV headers the
// image is read exactly as there (BGR from cv::imread); without, a binary PPM is read and handed over as BGR too, so
// both builds feed the path the channel order the reference's C++ callers do.  The reference
// discards the result; this prints one line per item (text, then the Tesseract-style box) and, like
// bindings/run_ocr.py:10-82, writes the page with the boxes drawn to <outputs_dir>/annotated.ppm.
#include <algorithm>
#include <cstdio>
#include <string>

#include "pnm.h"
#include "tuatara.h"

#ifdef TUATARA_HAVE_OPENCV
#include <opencv2/imgcodecs.hpp>
#endif

static void draw_box(PnmImage* im, const std::vector<float>& b) {
  const int x0 = std::max(0, static_cast<int>(b[0])), y0 = std::max(0, static_cast<int>(b[1]));
  const int x1 = std::min(im->cols - 1, static_cast<int>(b[2])), y1 = std::min(im->rows - 1, static_cast<int>(b[3]));
  auto put = [&](int x, int y) {
    uint8_t* p = &im->rgb[(static_cast<size_t>(y) * im->cols + x) * 3];
    p[0] = 255; p[1] = 0; p[2] = 0;
  };
  for (int x = x0; x <= x1; ++x) { put(x, y0); put(x, y1); }
  for (int y = y0; y <= y1; ++y) { put(x0, y); put(x1, y); }
}

int main(int argc, const char** argv) {
  if (argc < 4) {
    std::fprintf(stderr, "usage: %s <image.ppm|png> <weights_dir> <outputs_dir>\n", argv[0]);
    return 2;
  }
  const std::string image_path = argv[1];
  std::vector<OutputItem> items;
  PnmImage rgb;
#ifdef TUATARA_HAVE_OPENCV
  cv::Mat image = cv::imread(image_path, cv::IMREAD_COLOR);
  items = image_to_data(image, argv[2], argv[3]);
  rgb.rows = image.rows; rgb.cols = image.cols;
  rgb.rgb.resize(static_cast<size_t>(image.rows) * image.cols * 3);
  for (int y = 0; y < image.rows; ++y)
    for (int x = 0; x < image.cols; ++x)
      for (int c = 0; c < 3; ++c) rgb.rgb[(static_cast<size_t>(y) * image.cols + x) * 3 + c] = image.at<cv::Vec3b>(y, x)[2 - c];
#else
  if (!read_ppm(image_path, &rgb)) {
    std::fprintf(stderr, "Error reading image from file\n");  // tuatara.cpp:344-347
    return 1;
  }
  std::vector<uint8_t> bgr(rgb.rgb.size());
  for (size_t i = 0; i < rgb.rgb.size(); i += 3) { bgr[i] = rgb.rgb[i + 2]; bgr[i + 1] = rgb.rgb[i + 1]; bgr[i + 2] = rgb.rgb[i]; }
  tuatara::ImageView view;
  view.data = bgr.data(); view.rows = rgb.rows; view.cols = rgb.cols; view.channels = 3;
  view.step = static_cast<size_t>(rgb.cols) * 3;
  items = image_to_data(view, argv[2], argv[3]);
#endif
  for (const OutputItem& it : items)
    std::printf("%s\t%.0f\t%.0f\t%.0f\t%.0f\n", it.text.c_str(), it.bbox[0], it.bbox[1], it.bbox[2], it.bbox[3]);
  for (const OutputItem& it : items) draw_box(&rgb, it.bbox);
  write_ppm(std::string(argv[3]) + "/annotated.ppm", rgb);
  return 0;
}
