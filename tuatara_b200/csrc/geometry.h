// Host geometry of the OCR path: the OpenCV calls of tuatara.cpp:177-179 (findNonZero +
// minAreaRect after the rectangular dilation of :166-174), :236-253 (adjust_result_coordinates),
// :256-274 (bbox formatting) and :416 (boundingRect), restated so that the floats are identical to
// OpenCV 4.13's.  Compile with -ffp-contract=off.
#pragma once
#include <stdint.h>

#include <vector>

namespace tt {

struct RotatedRect { float cx, cy, w, h, angle; };
struct Pt2f { float x, y; };
struct Pt2i { int x, y; };
struct RectI { int x, y, w, h; };

// cv::minAreaRect for integer pixel coordinates / float points (input order matters, see App. D1).
RotatedRect min_area_rect_i(const Pt2i* pts, int n);
RotatedRect min_area_rect_f(const Pt2f* pts, int n);
// cv::convexHull(points, clockwise=false, returnPoints=false): indices into pts.
std::vector<int> convex_hull_i(const Pt2i* pts, int n);
std::vector<int> convex_hull_f(const Pt2f* pts, int n);
// cv::RotatedRect::points / boundingRect
void rect_points(const RotatedRect& r, Pt2f out[4]);
RectI rect_bounding(const RotatedRect& r);

// One connected component as the GPU reports it (postprocess.cuh) -> the point list
// cv::findNonZero would return after link-only removal and the (1+niter)^2 rectangular dilation
// inside the clamped ROI, reduced to <= 2 points per row in raster order (App. D6), then
// minAreaRect.  rows[i] = {xmin, xmax} of source row top+i (xmin > xmax: empty).
// Returns false when the component is filtered out (area < min_area or max_text < text_threshold).
struct CompIn {
  int left, top, right, bottom, area;
  float max_text;
};
bool component_rect(const CompIn& c, const int* row_xmin_xmax, int img_w, int img_h, int min_area,
                    float text_threshold, RotatedRect* out, std::vector<Pt2i>* scratch);

// tuatara.cpp:236-253 for one box: corners *= (ratio * ratio_net) in fp32, minAreaRect of the 4 corners.
RotatedRect adjust_rect(const RotatedRect& r, float ratio_w, float ratio_h, float ratio_net);
// Opt-in rectification (the TODO at tuatara.cpp:411-415; tt_config.rectify): the box's 4 vertices ordered
// top-left, top-right, bottom-right, bottom-left (smallest / largest x + y, smallest / largest y - x, first index wins)
void rect_to_quad(const RotatedRect& r, Pt2f quad[4]);
// cv::getPerspectiveTransform(quad -> {(0,0),(127,0),(127,31),(0,31)}) followed by the inversion cv::warpPerspective
// does: m_inv maps an output pixel of the 128 x 32 crop to source coordinates.  false: degenerate quad.
bool quad_to_warp(const Pt2f quad[4], double m_inv[9]);
// tuatara.cpp:256-274: [min_x, min_y, max_x, max_y] of the 4 vertices, std::round-ed.
void rect_to_bbox(const RotatedRect& r, float out[4]);

}  // namespace tt
