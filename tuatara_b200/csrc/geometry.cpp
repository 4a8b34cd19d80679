// Float-exact restatement of the OpenCV 4.13 geometry the reference calls.
// Build with -ffp-contract=off: every float product/sum below must round separately.
#include "geometry.h"

#include <float.h>
#include <math.h>

#include <algorithm>
#include <numeric>

namespace tt {

namespace {

constexpr double kPi = 3.1415926535897932384626433832795;  // CV_PI

// ---- App. D1: cv::convexHull(points, clockwise=false, returnPoints=false) -----------------
// sign of the turn a -> b -> p, computed like Sklansky_<T, DotT>: coordinate differences of
// consecutive edges in T, products in DotT (int64 for int points, double for float points).
inline int turn(const Pt2i& a, const Pt2i& b, const Pt2i& p) {
  const int ax = b.x - a.x, ay = b.y - a.y, bx = p.x - b.x, by = p.y - b.y;
  const long long c = static_cast<long long>(ax) * by - static_cast<long long>(ay) * bx;
  return c > 0 ? 1 : (c < 0 ? -1 : 0);
}
inline int turn(const Pt2f& a, const Pt2f& b, const Pt2f& p) {
  const float ax = b.x - a.x, ay = b.y - a.y, bx = p.x - b.x, by = p.y - b.y;
  const double c = static_cast<double>(ax) * by - static_cast<double>(ay) * bx;
  return c > 0 ? 1 : (c < 0 ? -1 : 0);
}

template <class P>
std::vector<int> convex_hull_impl(const P* pts, int n) {
  std::vector<int> hull;
  if (n <= 0) return hull;
  std::vector<int> order(n);
  std::iota(order.begin(), order.end(), 0);
  std::sort(order.begin(), order.end(), [&](int a, int b) {
    return pts[a].x < pts[b].x || (pts[a].x == pts[b].x && pts[a].y < pts[b].y);
  });
  const P& first = pts[order[0]];
  const P& last = pts[order[n - 1]];
  if (first.x == last.x && first.y == last.y) {  // all points coincide
    hull.push_back(order[0]);
    return hull;
  }
  // strict monotone chain: collinear points are dropped
  std::vector<int> lower, upper;
  for (int i = 0; i < n; ++i) {
    const int p = order[i];
    while (lower.size() >= 2 && turn(pts[lower[lower.size() - 2]], pts[lower.back()], pts[p]) <= 0) lower.pop_back();
    lower.push_back(p);
  }
  for (int i = n - 1; i >= 0; --i) {
    const int p = order[i];
    while (upper.size() >= 2 && turn(pts[upper[upper.size() - 2]], pts[upper.back()], pts[p]) <= 0) upper.pop_back();
    upper.push_back(p);
  }
  // start at the last point of the sorted order (max x, then max y): upper chain first
  for (size_t i = 0; i + 1 < upper.size(); ++i) hull.push_back(upper[i]);
  for (size_t i = 0; i + 1 < lower.size(); ++i) hull.push_back(lower[i]);
  // cyclic-shift quirk on the *input indices* of the hull vertices
  const int nout = static_cast<int>(hull.size());
  if (nout >= 3) {
    int min_idx = 0, max_idx = 0, lt = 0;
    for (int i = 1; i < nout; ++i) {
      const int idx = hull[i];
      lt += hull[i - 1] < idx;
      if (lt > 1 && lt <= i - 2) break;
      if (idx < hull[min_idx]) min_idx = i;
      if (idx > hull[max_idx]) max_idx = i;
    }
    const int mmdist = std::abs(max_idx - min_idx);
    if ((mmdist == 1 || mmdist == nout - 1) && (lt <= 1 || lt >= nout - 2)) {
      const bool ascending = (max_idx + 1) % nout == min_idx;
      const int i0 = ascending ? min_idx : max_idx;
      if (i0 > 0) {
        std::vector<int> shifted(nout);
        int j = i0, i = 0;
        for (; i < nout; ++i) {
          const int cur = shifted[i] = hull[j];
          const int next_j = j + 1 < nout ? j + 1 : 0;
          const int nxt = hull[next_j];
          if (i < nout - 1 && (ascending != (cur < nxt))) break;
          j = next_j;
        }
        if (i == nout) hull.swap(shifted);
      }
    }
  }
  return hull;
}

// ---- App. D2: rotatingCalipers, CALIPERS_MINAREARECT ----------------------------------------
inline bool first_vec_is_right(float v1x, float v1y, float v2x, float v2y) {
  const float t = v1y * v2x + (-v1x) * v2y;
  return t < 0.f;
}

void rotating_calipers(const Pt2f* P, int n, float out[6]) {
  std::vector<float> inv(n);
  std::vector<Pt2f> vect(n);
  int left = 0, bottom = 0, right = 0, top = 0;
  float left_x = P[0].x, right_x = P[0].x, top_y = P[0].y, bottom_y = P[0].y;
  for (int i = 0; i < n; ++i) {
    const Pt2f& p0 = P[i];
    const Pt2f& p1 = P[i + 1 < n ? i + 1 : 0];
    if (p0.x < left_x) { left_x = p0.x; left = i; }
    if (p0.x > right_x) { right_x = p0.x; right = i; }
    if (p0.y > top_y) { top_y = p0.y; top = i; }
    if (p0.y < bottom_y) { bottom_y = p0.y; bottom = i; }
    const float dx = p1.x - p0.x, dy = p1.y - p0.y;
    vect[i].x = dx; vect[i].y = dy;
    inv[i] = static_cast<float>(1.0 / sqrt(static_cast<double>(dx) * dx + static_cast<double>(dy) * dy));
  }
  float orientation = 0.f;
  {
    double ax = vect[n - 1].x, ay = vect[n - 1].y;
    for (int i = 0; i < n; ++i) {
      const double bx = vect[i].x, by = vect[i].y;
      const double convexity = ax * by - ay * bx;
      if (convexity != 0) { orientation = convexity > 0 ? 1.f : -1.f; break; }
      ax = bx; ay = by;
    }
  }
  float base_a = orientation, base_b = 0.f;
  int seq[4] = {bottom, right, top, left};
  float minarea = FLT_MAX;
  int rec_left = 0, rec_bottom = 0;
  float rec_a = 0.f, rec_b = 0.f, rec_w = 0.f, rec_h = 0.f;
  for (int k = 0; k < n; ++k) {
    const float r0x = vect[seq[0]].x, r0y = vect[seq[0]].y;
    const float r1x = vect[seq[1]].y, r1y = -vect[seq[1]].x;   // rotate90cw
    const float r2x = -vect[seq[2]].x, r2y = -vect[seq[2]].y;  // rotate180
    const float r3x = -vect[seq[3]].y, r3y = vect[seq[3]].x;   // rotate90ccw
    const float rx[4] = {r0x, r1x, r2x, r3x}, ry[4] = {r0y, r1y, r2y, r3y};
    int me = 0;
    for (int i = 1; i < 4; ++i)
      if (first_vec_is_right(rx[i], ry[i], rx[me], ry[me])) me = i;
    const int pindex = seq[me];
    const float lead_x = vect[pindex].x * inv[pindex];
    const float lead_y = vect[pindex].y * inv[pindex];
    switch (me) {
      case 0: base_a = lead_x; base_b = lead_y; break;
      case 1: base_a = lead_y; base_b = -lead_x; break;
      case 2: base_a = -lead_x; base_b = -lead_y; break;
      default: base_a = -lead_y; base_b = lead_x; break;
    }
    seq[me] += 1;
    if (seq[me] == n) seq[me] = 0;
    float dx = P[seq[1]].x - P[seq[3]].x, dy = P[seq[1]].y - P[seq[3]].y;
    const float width = dx * base_a + dy * base_b;
    dx = P[seq[2]].x - P[seq[0]].x; dy = P[seq[2]].y - P[seq[0]].y;
    const float height = -dx * base_b + dy * base_a;
    const float area = width * height;
    if (area <= minarea) {
      minarea = area;
      rec_left = seq[3]; rec_a = base_a; rec_w = width; rec_b = base_b; rec_h = height; rec_bottom = seq[0];
    }
  }
  const float A1 = rec_a, B1 = rec_b, A2 = -rec_b, B2 = rec_a;
  const float C1 = A1 * P[rec_left].x + P[rec_left].y * B1;
  const float C2 = A2 * P[rec_bottom].x + P[rec_bottom].y * B2;
  const float idet = 1.f / (A1 * B2 - A2 * B1);
  const float px = (C1 * B2 - C2 * B1) * idet;
  const float py = (A1 * C2 - A2 * C1) * idet;
  out[0] = px; out[1] = py;
  out[2] = A1 * rec_w; out[3] = B1 * rec_w;
  out[4] = A2 * rec_h; out[5] = B2 * rec_h;
}

// ---- App. D3: minAreaRect wrap-up on the hull (already float) ---------------------------------
RotatedRect min_area_rect_hull(const std::vector<Pt2f>& h) {
  RotatedRect box{0.f, 0.f, 0.f, 0.f, 0.f};
  const int n = static_cast<int>(h.size());
  double vx = 1.0, vy = 0.0;  // direction of the "width" edge
  if (n > 2) {
    float out[6];
    rotating_calipers(h.data(), n, out);
    box.cx = out[0] + (out[2] + out[4]) * 0.5f;
    box.cy = out[1] + (out[3] + out[5]) * 0.5f;
    box.w = static_cast<float>(sqrt(static_cast<double>(out[2]) * out[2] + static_cast<double>(out[3]) * out[3]));
    box.h = static_cast<float>(sqrt(static_cast<double>(out[4]) * out[4] + static_cast<double>(out[5]) * out[5]));
    vx = out[2]; vy = out[3];
  } else if (n == 2) {
    box.cx = (h[0].x + h[1].x) * 0.5f;
    box.cy = (h[0].y + h[1].y) * 0.5f;
    const double dx = h[1].x - h[0].x, dy = h[1].y - h[0].y;
    box.w = static_cast<float>(sqrt(dx * dx + dy * dy));
    box.h = 0.f;
    vx = dx; vy = dy;
  } else if (n == 1) {
    box.cx = h[0].x; box.cy = h[0].y;
  }
  // Angle normalised into [-90, 0) by quarter turns, swapping width/height at each turn.  The
  // turns are applied to the edge vector *before* atan2 (exact), not subtracted from the angle:
  // the two differ in the last float bit when the result is within ~1e-5 degrees of 0 (fuzzed
  // against cv2 4.13: 0 mismatches in 20000 float-corner cases, all degenerate cases match).
  double deg = atan2(vy, vx) * 180.0 / kPi;
  int turns = 0;
  while (deg >= 0.0) { deg -= 90.0; const double t = vx; vx = vy; vy = -t; ++turns; }
  while (deg < -90.0) { deg += 90.0; const double t = vx; vx = -vy; vy = t; ++turns; }
  if (turns & 1) std::swap(box.w, box.h);
  if (turns) deg = atan2(vy, vx) * 180.0 / kPi;
  box.angle = static_cast<float>(deg);
  return box;
}

}  // namespace

std::vector<int> convex_hull_i(const Pt2i* pts, int n) { return convex_hull_impl(pts, n); }
std::vector<int> convex_hull_f(const Pt2f* pts, int n) { return convex_hull_impl(pts, n); }

RotatedRect min_area_rect_i(const Pt2i* pts, int n) {
  const std::vector<int> hi = convex_hull_impl(pts, n);
  std::vector<Pt2f> h(hi.size());
  for (size_t i = 0; i < hi.size(); ++i) h[i] = Pt2f{static_cast<float>(pts[hi[i]].x), static_cast<float>(pts[hi[i]].y)};
  return min_area_rect_hull(h);
}

RotatedRect min_area_rect_f(const Pt2f* pts, int n) {
  const std::vector<int> hi = convex_hull_impl(pts, n);
  std::vector<Pt2f> h(hi.size());
  for (size_t i = 0; i < hi.size(); ++i) h[i] = pts[hi[i]];
  return min_area_rect_hull(h);
}

// ---- App. D4 ---------------------------------------------------------------------------------
void rect_points(const RotatedRect& r, Pt2f pt[4]) {
  const double ang = static_cast<double>(r.angle) * kPi / 180.0;
  const float b = static_cast<float>(cos(ang)) * 0.5f;
  const float a = static_cast<float>(sin(ang)) * 0.5f;
  pt[0].x = r.cx - a * r.h - b * r.w;
  pt[0].y = r.cy + b * r.h - a * r.w;
  pt[1].x = r.cx + a * r.h - b * r.w;
  pt[1].y = r.cy - b * r.h - a * r.w;
  pt[2].x = 2 * r.cx - pt[0].x;
  pt[2].y = 2 * r.cy - pt[0].y;
  pt[3].x = 2 * r.cx - pt[1].x;
  pt[3].y = 2 * r.cy - pt[1].y;
}

RectI rect_bounding(const RotatedRect& r) {
  Pt2f pt[4];
  rect_points(r, pt);
  const float minx = std::min(std::min(std::min(pt[0].x, pt[1].x), pt[2].x), pt[3].x);
  const float miny = std::min(std::min(std::min(pt[0].y, pt[1].y), pt[2].y), pt[3].y);
  const float maxx = std::max(std::max(std::max(pt[0].x, pt[1].x), pt[2].x), pt[3].x);
  const float maxy = std::max(std::max(std::max(pt[0].y, pt[1].y), pt[2].y), pt[3].y);
  RectI o;
  o.x = static_cast<int>(floor(static_cast<double>(minx)));
  o.y = static_cast<int>(floor(static_cast<double>(miny)));
  o.w = static_cast<int>(ceil(static_cast<double>(maxx))) - o.x + 1;
  o.h = static_cast<int>(ceil(static_cast<double>(maxy))) - o.y + 1;
  return o;
}

// ---- App. D6: component -> reduced point list -> minAreaRect ---------------------------------
bool component_rect(const CompIn& c, const int* rows, int img_w, int img_h, int min_area, float text_threshold,
                    RotatedRect* out, std::vector<Pt2i>* scratch) {
  if (c.area < min_area) return false;               // tuatara.cpp:147-148
  if (c.max_text < text_threshold) return false;     // tuatara.cpp:154
  const int x = c.left, y = c.top, w = c.right - c.left + 1, h = c.bottom - c.top + 1;
  // tuatara.cpp:166 -- integer division, x2 inside the sqrt
  const int niter = static_cast<int>(sqrt(static_cast<double>((c.area * std::min(w, h)) / (w * h) * 2)));
  const int sx = std::max(0, x - niter), sy = std::max(0, y - niter);                          // :168-169
  const int ex = std::min(img_w, x + w + niter + 1), ey = std::min(img_h, y + h + niter + 1);  // :170-171
  // (1+niter)^2 rect kernel, default anchor: grows floor(niter/2) toward -x/-y, ceil(niter/2) toward +x/+y
  const int l = niter / 2, r = niter - l;
  std::vector<Pt2i>& pts = *scratch;
  pts.clear();
  const int yo_lo = std::max(sy, y - l), yo_hi = std::min(ey, y + h + r);
  for (int yo = yo_lo; yo < yo_hi; ++yo) {
    const int s_lo = std::max(y, yo - r), s_hi = std::min(y + h - 1, yo + l);
    int mn = INT32_MAX, mx = -1;
    for (int s = s_lo; s <= s_hi; ++s) {
      const int a = rows[2 * (s - y)], b = rows[2 * (s - y) + 1];
      if (a <= b) { mn = std::min(mn, a); mx = std::max(mx, b); }
    }
    if (mx < 0) continue;
    const int px0 = std::max(sx, mn - l), px1 = std::min(ex - 1, mx + r);
    pts.push_back(Pt2i{px0, yo});
    if (px1 != px0) pts.push_back(Pt2i{px1, yo});
  }
  *out = min_area_rect_i(pts.data(), static_cast<int>(pts.size()));  // tuatara.cpp:177-179
  return true;
}

RotatedRect adjust_rect(const RotatedRect& rr, float ratio_w, float ratio_h, float ratio_net) {
  Pt2f c[4];
  rect_points(rr, c);                        // tuatara.cpp:240-241
  const float sw = ratio_w * ratio_net, sh = ratio_h * ratio_net;
  for (int i = 0; i < 4; ++i) {              // :243-246
    c[i].x *= sw;
    c[i].y *= sh;
  }
  return min_area_rect_f(c, 4);              // :248
}

void rect_to_quad(const RotatedRect& r, Pt2f quad[4]) {
  Pt2f v[4];
  rect_points(r, v);
  int tl = 0, br = 0, tr = 0, bl = 0;
  for (int i = 1; i < 4; ++i) {   // float32 sums / differences, first extremum wins (numpy argmin / argmax)
    const float s = v[i].x + v[i].y, d = v[i].y - v[i].x;
    if (s < v[tl].x + v[tl].y) tl = i;
    if (s > v[br].x + v[br].y) br = i;
    if (d < v[tr].y - v[tr].x) tr = i;
    if (d > v[bl].y - v[bl].x) bl = i;
  }
  quad[0] = v[tl]; quad[1] = v[tr]; quad[2] = v[br]; quad[3] = v[bl];
}

bool quad_to_warp(const Pt2f q[4], double m_inv[9]) {
  // cv::getPerspectiveTransform: 8 x 8 system in double, LU with partial pivoting
  static const double dx[4] = {0.0, 127.0, 127.0, 0.0}, dy[4] = {0.0, 0.0, 31.0, 31.0};
  double a[8][9];
  for (int i = 0; i < 4; ++i) {
    const double sx = q[i].x, sy = q[i].y;
    double* r0 = a[i];
    double* r1 = a[i + 4];
    r0[0] = sx; r0[1] = sy; r0[2] = 1; r0[3] = 0; r0[4] = 0; r0[5] = 0; r0[6] = -sx * dx[i]; r0[7] = -sy * dx[i]; r0[8] = dx[i];
    r1[0] = 0; r1[1] = 0; r1[2] = 0; r1[3] = sx; r1[4] = sy; r1[5] = 1; r1[6] = -sx * dy[i]; r1[7] = -sy * dy[i]; r1[8] = dy[i];
  }
  for (int c = 0; c < 8; ++c) {
    int piv = c;
    for (int r = c + 1; r < 8; ++r)
      if (fabs(a[r][c]) > fabs(a[piv][c])) piv = r;
    if (fabs(a[piv][c]) < 1e-12) return false;
    if (piv != c)
      for (int k = 0; k < 9; ++k) std::swap(a[piv][k], a[c][k]);
    for (int r = c + 1; r < 8; ++r) {
      const double f = a[r][c] / a[c][c];
      for (int k = c; k < 9; ++k) a[r][k] -= f * a[c][k];
    }
  }
  double m[9];
  for (int r = 7; r >= 0; --r) {
    double v = a[r][8];
    for (int k = r + 1; k < 8; ++k) v -= a[r][k] * m[k];
    m[r] = v / a[r][r];
  }
  m[8] = 1.0;
  // cv::invert of the 3 x 3 (cofactors / determinant), as warpPerspective does before sampling
  const double det = m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) + m[2] * (m[3] * m[7] - m[4] * m[6]);
  if (det == 0.0) return false;
  const double id = 1.0 / det;
  m_inv[0] = (m[4] * m[8] - m[5] * m[7]) * id;
  m_inv[1] = (m[2] * m[7] - m[1] * m[8]) * id;
  m_inv[2] = (m[1] * m[5] - m[2] * m[4]) * id;
  m_inv[3] = (m[5] * m[6] - m[3] * m[8]) * id;
  m_inv[4] = (m[0] * m[8] - m[2] * m[6]) * id;
  m_inv[5] = (m[2] * m[3] - m[0] * m[5]) * id;
  m_inv[6] = (m[3] * m[7] - m[4] * m[6]) * id;
  m_inv[7] = (m[1] * m[6] - m[0] * m[7]) * id;
  m_inv[8] = (m[0] * m[4] - m[1] * m[3]) * id;
  return true;
}

void rect_to_bbox(const RotatedRect& rr, float out[4]) {
  Pt2f v[4];
  rect_points(rr, v);                        // tuatara.cpp:258
  const float min_x = std::min(std::min(v[0].x, v[1].x), std::min(v[2].x, v[3].x));
  const float min_y = std::min(std::min(v[0].y, v[1].y), std::min(v[2].y, v[3].y));
  const float max_x = std::max(std::max(v[0].x, v[1].x), std::max(v[2].x, v[3].x));
  const float max_y = std::max(std::max(v[0].y, v[1].y), std::max(v[2].y, v[3].y));
  out[0] = roundf(min_x); out[1] = roundf(min_y); out[2] = roundf(max_x); out[3] = roundf(max_y);  // :266-270
}

}  // namespace tt
