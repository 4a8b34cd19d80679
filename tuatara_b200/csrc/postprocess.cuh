// Score-map post-processing on the GPU: the dense part of get_detected_boxes
// (tuatara.cpp:119-179): min-max normalise, two thresholds, 4-connected component labelling
// with OpenCV's label numbering, per-component stats, per-(component,row) extents of the
// pixels that survive the link-only removal.  Everything after that (rectangular dilation of
// the row extents, convex hull, rotating calipers) is O(rows) per component and runs on the
// host (geometry.cpp).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace tt {

// Per-page header, first in the page's result block (device and host mirror).
struct PostHeader {
  float tmin, tmax, lmin, lmax;  // raw min/max of the two maps
  int n_comp;                    // components found (cv2 nLabels - 1)
  int n_rows;                    // total row-extent slots used (sum of heights)
  int overflow;                  // n_comp > comp_cap or n_rows > row_cap: host must re-run with bigger caps
  int pad;
};

struct PostComp {  // cv::connectedComponentsWithStats row + what tuatara.cpp:150-154 needs
  int left, top, right, bottom;  // inclusive bbox
  int area;
  float max_text;                // max normalised text score over the component
  int row_off;                   // first row-extent slot (rows top..bottom)
  int pad;
};

struct PostRow { int xmin, xmax; };  // xmin > xmax: no surviving pixel in this row

struct PostParams {
  float low_text = 0.4f, link_threshold = 0.4f;
};

// Device workspace for a batch of equally sized maps.
struct PostWorkspace {
  int batch = 0, H = 0, W = 0;
  int comp_cap = 0, row_cap = 0;
  int* parent = nullptr;       // [B][H*W] union-find forest / root index
  int* rank = nullptr;         // [B][H*W] label of a root pixel (valid at roots)
  uint8_t* flags = nullptr;    // [B][H*W] bit0 fg, bit1 text, bit2 link
  int* block_counts = nullptr; // [B][nblk+1]
  int* comp_scan = nullptr;    // [B][HW/2+2] heights scan scratch
  // full-size per-component stats (indexed by label; label count is data dependent, <= HW/2+1)
  int* c_area = nullptr; int* c_minx = nullptr; int* c_miny = nullptr; int* c_maxx = nullptr; int* c_maxy = nullptr;
  unsigned* c_maxv = nullptr; int* c_off = nullptr;
  PostRow* rows_full = nullptr;  // [B][H*W]
  // compact result block per page: header | comps[comp_cap] | rows[row_cap]
  uint8_t* result = nullptr;
  size_t result_stride = 0;
  int* labels = nullptr;       // optional [B][H*W] final labels (parity tests)
  size_t bytes = 0;
};

size_t post_result_stride(int comp_cap, int row_cap);
cudaError_t post_workspace_alloc(PostWorkspace* ws, int batch, int H, int W, int comp_cap, int row_cap, bool want_labels);
void post_workspace_free(PostWorkspace* ws);

// maps: device fp32 [B][H][W][2] (channel 0 text/region, 1 link/affinity), CRAFT's output layout.
// Fills ws->result (device); caller copies `batch * result_stride` bytes to the host.
cudaError_t post_run(const PostWorkspace& ws, const float* maps, const PostParams& p, cudaStream_t s);

}  // namespace tt
