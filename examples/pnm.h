// Minimal binary PPM (P6, maxval 255) reader / writer for the examples: the reference's examples read PNGs with
// cv::imread (examples/resume.cpp:8-9); without OpenCV a decoded page comes from `convert page.png page.ppm`.
#pragma once
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

struct PnmImage {
  int rows = 0, cols = 0;
  std::vector<uint8_t> rgb;  // rows x cols x 3, R G B
};

inline bool pnm_skip_ws_and_comments(FILE* f) {
  int c;
  while ((c = std::fgetc(f)) != EOF) {
    if (c == '#') { while ((c = std::fgetc(f)) != EOF && c != '\n') {} continue; }
    if (c != ' ' && c != '\t' && c != '\n' && c != '\r') { std::ungetc(c, f); return true; }
  }
  return false;
}

inline bool read_ppm(const std::string& path, PnmImage* out) {
  FILE* f = std::fopen(path.c_str(), "rb");
  if (!f) return false;
  char magic[3] = {0};
  int maxval = 0;
  bool ok = std::fread(magic, 1, 2, f) == 2 && magic[0] == 'P' && magic[1] == '6' && pnm_skip_ws_and_comments(f) &&
            std::fscanf(f, "%d", &out->cols) == 1 && pnm_skip_ws_and_comments(f) && std::fscanf(f, "%d", &out->rows) == 1 &&
            pnm_skip_ws_and_comments(f) && std::fscanf(f, "%d", &maxval) == 1 && maxval == 255 && out->rows > 0 && out->cols > 0;
  if (ok) {
    std::fgetc(f);  // the single whitespace byte after maxval
    out->rgb.resize(static_cast<size_t>(out->rows) * out->cols * 3);
    ok = std::fread(out->rgb.data(), 1, out->rgb.size(), f) == out->rgb.size();
  }
  std::fclose(f);
  return ok;
}

inline bool write_ppm(const std::string& path, const PnmImage& im) {
  FILE* f = std::fopen(path.c_str(), "wb");
  if (!f) return false;
  std::fprintf(f, "P6\n%d %d\n255\n", im.cols, im.rows);
  const bool ok = std::fwrite(im.rgb.data(), 1, im.rgb.size(), f) == im.rgb.size();
  std::fclose(f);
  return ok;
}
