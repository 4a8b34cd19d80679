// The reference's examples/table.cpp: as resume.cpp with the weight / output directories fixed to the repository
// layout (reference: examples/table.cpp:10 -- image_to_data(image, "../../weights", "../../outputs")).
//   table <image>
#include <algorithm>
#include <cstdio>
#include <string>

#include "pnm.h"
#include "tuatara.h"

#ifdef TUATARA_HAVE_OPENCV
#include <opencv2/imgcodecs.hpp>
#endif

int main(int argc, const char** argv) {
  if (argc < 2) {
    std::fprintf(stderr, "usage: %s <image.ppm|png>\n", argv[0]);
    return 2;
  }
  std::vector<OutputItem> items;
#ifdef TUATARA_HAVE_OPENCV
  cv::Mat image = cv::imread(argv[1], cv::IMREAD_COLOR);
  items = image_to_data(image, "../../weights", "../../outputs");
#else
  PnmImage rgb;
  if (!read_ppm(argv[1], &rgb)) {
    std::fprintf(stderr, "Error reading image from file\n");
    return 1;
  }
  for (size_t i = 0; i < rgb.rgb.size(); i += 3) std::swap(rgb.rgb[i], rgb.rgb[i + 2]);  // BGR like cv::imread
  tuatara::ImageView view;
  view.data = rgb.rgb.data(); view.rows = rgb.rows; view.cols = rgb.cols; view.channels = 3;
  view.step = static_cast<size_t>(rgb.cols) * 3;
  items = image_to_data(view, "../../weights", "../../outputs");
#endif
  for (const OutputItem& it : items)
    std::printf("%s\t%.0f\t%.0f\t%.0f\t%.0f\n", it.text.c_str(), it.bbox[0], it.bbox[1], it.bbox[2], it.bbox[3]);
  return 0;
}
