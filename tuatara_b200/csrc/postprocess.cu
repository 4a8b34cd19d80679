// GPU restatement of the dense half of get_detected_boxes (tuatara.cpp:119-179).
//
//   cv::threshold x2 + union (tuatara.cpp:131-137)       -> k_label_init (fused with normalise :120-121)
//   cv::connectedComponentsWithStats(...,4) (:142)       -> warp-run union-find: k_label_init / k_merge /
//                                                           k_compress / k_scan_blocks / k_rank / k_stats
//   labels==k + minMaxLoc per component (:150-152)       -> segmented atomicMax in k_stats
//   segmap / setTo(0, link&&!text) / findNonZero (:156-178) -> per-(component,row) extents in k_rows
//
// Label numbering: OpenCV numbers components in raster order of their first pixel.  The forest
// here always links the larger root index under the smaller one, so a component's root *is* its
// first pixel in raster order; label = 1 + (number of roots before it) -- a prefix count.
// All float ops that feed a comparison use explicit IEEE intrinsics (no FMA contraction, true
// division) so thresholds flip exactly where ATen's CPU sub/div + cv::threshold flip.
#include "postprocess.cuh"

#include <limits.h>

#include <algorithm>

#include "common.h"

namespace tt {

namespace {

constexpr int kPix = 1024;  // pixels (threads) per block in the per-pixel kernels
constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ unsigned enc_f32(float f) {
  const unsigned b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float dec_f32(unsigned e) {
  return __uint_as_float((e & 0x80000000u) ? (e & 0x7fffffffu) : ~e);
}
__device__ __forceinline__ float normalise(float v, float mn, float mx) {
  return __fdiv_rn(__fsub_rn(v, mn), __fsub_rn(mx, mn));  // (x - min) / (max - min), tuatara.cpp:120-121
}

struct Scratch {  // per page, lives at the start of block_counts' allocation
  unsigned mm[4];  // encoded tmin, tmax, lmin, lmax
};

__global__ void k_init(PostWorkspace ws, unsigned* mm) {
  const int page = blockIdx.x;
  if (threadIdx.x == 0) {
    mm[page * 4 + 0] = 0xffffffffu; mm[page * 4 + 1] = 0u;
    mm[page * 4 + 2] = 0xffffffffu; mm[page * 4 + 3] = 0u;
    PostHeader* h = reinterpret_cast<PostHeader*>(ws.result + page * ws.result_stride);
    h->n_comp = 0; h->n_rows = 0; h->overflow = 0; h->pad = 0;
  }
}

// min/max of both channels; maps are [H*W] float2 per page, read as float4 (two pixels) when even.
__global__ void k_minmax(const float* __restrict__ maps, int hw, unsigned* mm) {
  const int page = blockIdx.y;
  const float2* src = reinterpret_cast<const float2*>(maps) + static_cast<size_t>(page) * hw;
  float tmin = INFINITY, tmax = -INFINITY, lmin = INFINITY, lmax = -INFINITY;
  const int stride = gridDim.x * blockDim.x;
  const int pairs = hw >> 1;
  const float4* src4 = reinterpret_cast<const float4*>(src);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < pairs; i += stride) {
    const float4 v = __ldg(src4 + i);
    tmin = fminf(tmin, fminf(v.x, v.z)); tmax = fmaxf(tmax, fmaxf(v.x, v.z));
    lmin = fminf(lmin, fminf(v.y, v.w)); lmax = fmaxf(lmax, fmaxf(v.y, v.w));
  }
  if ((hw & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
    const float2 v = src[hw - 1];
    tmin = fminf(tmin, v.x); tmax = fmaxf(tmax, v.x); lmin = fminf(lmin, v.y); lmax = fmaxf(lmax, v.y);
  }
  for (int o = 16; o > 0; o >>= 1) {
    tmin = fminf(tmin, __shfl_xor_sync(FULL, tmin, o)); tmax = fmaxf(tmax, __shfl_xor_sync(FULL, tmax, o));
    lmin = fminf(lmin, __shfl_xor_sync(FULL, lmin, o)); lmax = fmaxf(lmax, __shfl_xor_sync(FULL, lmax, o));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMin(&mm[page * 4 + 0], enc_f32(tmin)); atomicMax(&mm[page * 4 + 1], enc_f32(tmax));
    atomicMin(&mm[page * 4 + 2], enc_f32(lmin)); atomicMax(&mm[page * 4 + 3], enc_f32(lmax));
  }
}

// thresholds + union of the two binary maps + warp-run initial labels
__global__ void k_label_init(PostWorkspace ws, const float* __restrict__ maps, const unsigned* mm, PostParams pp) {
  const int page = blockIdx.y;
  const int hw = ws.H * ws.W;
  const int i = blockIdx.x * kPix + threadIdx.x;
  const float tmin = dec_f32(mm[page * 4 + 0]), tmax = dec_f32(mm[page * 4 + 1]);
  const float lmin = dec_f32(mm[page * 4 + 2]), lmax = dec_f32(mm[page * 4 + 3]);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    PostHeader* h = reinterpret_cast<PostHeader*>(ws.result + page * ws.result_stride);
    h->tmin = tmin; h->tmax = tmax; h->lmin = lmin; h->lmax = lmax;
  }
  bool fg = false;
  unsigned char fl = 0;
  int x = 0;
  if (i < hw) {
    const float2 v = __ldg(reinterpret_cast<const float2*>(maps) + static_cast<size_t>(page) * hw + i);
    const bool text = normalise(v.x, tmin, tmax) > pp.low_text;        // cv::threshold THRESH_BINARY (:131)
    const bool link = normalise(v.y, lmin, lmax) > pp.link_threshold;  // (:132)
    fg = text || link;                                                  // clip(text+link,0,1) (:136)
    fl = (fg ? 1 : 0) | (text ? 2 : 0) | (link ? 4 : 0);
    x = i % ws.W;
  }
  // start of this pixel's horizontal fg run inside the warp's 32-pixel span (runs break at x == 0)
  const unsigned lane = threadIdx.x & 31;
  const unsigned fgbits = __ballot_sync(FULL, fg);
  const unsigned rowstart = __ballot_sync(FULL, x == 0);
  if (i < hw) {
    int par = -1;
    if (fg) {
      // breaks: lanes that are background, or lanes that start a row (they start a new run themselves)
      const unsigned below = (lane == 0) ? 0u : (FULL >> (32 - lane));        // lanes < lane
      const unsigned bg_below = ~fgbits & below;
      const unsigned rs_le = rowstart & (below | (1u << lane));               // row starts at lanes <= lane
      int start = 0;
      if (bg_below) start = 32 - __clz(bg_below);                             // lane after the last bg lane
      if (rs_le) start = max(start, 31 - __clz(rs_le));
      par = i - (static_cast<int>(lane) - start);
    }
    ws.parent[static_cast<size_t>(page) * hw + i] = par;
    ws.flags[static_cast<size_t>(page) * hw + i] = fl;
  }
}

__device__ __forceinline__ int uf_find(const int* par, int a) {
  int p = par[a];
  while (p != a) { a = p; p = par[a]; }
  return a;
}
__device__ __forceinline__ void uf_union(int* par, int a, int b) {
  while (true) {
    a = uf_find(par, a);
    b = uf_find(par, b);
    if (a == b) return;
    if (a < b) { const int t = a; a = b; b = t; }  // a > b: hang the larger root under the smaller
    const int old = atomicMin(&par[a], b);
    if (old == a) return;
    a = old;
  }
}

__global__ void k_merge(PostWorkspace ws) {
  const int page = blockIdx.y;
  const int hw = ws.H * ws.W, W = ws.W;
  const int i = blockIdx.x * kPix + threadIdx.x;
  if (i >= hw) return;
  int* par = ws.parent + static_cast<size_t>(page) * hw;
  const uint8_t* fl = ws.flags + static_cast<size_t>(page) * hw;
  if (!(fl[i] & 1)) return;
  const int x = i % W;
  const bool left = x > 0 && (fl[i - 1] & 1);
  if (left && (i & 31) == 0) uf_union(par, i, i - 1);  // run continues across a warp-span boundary
  if (i >= W && (fl[i - W] & 1)) {
    const bool upleft = x > 0 && (fl[i - W - 1] & 1);
    if (!(left && upleft)) uf_union(par, i, i - W);   // otherwise the left pixel already made this link
  }
}

__global__ void k_compress(PostWorkspace ws) {
  const int page = blockIdx.y;
  const int hw = ws.H * ws.W;
  const int i = blockIdx.x * kPix + threadIdx.x;
  int* par = ws.parent + static_cast<size_t>(page) * hw;
  bool root = false;
  if (i < hw && par[i] >= 0) {
    const int r = uf_find(par, i);
    par[i] = r;
    root = (r == i);
  }
  const int cnt = __syncthreads_count(root);
  if (threadIdx.x == 0) ws.block_counts[page * (gridDim.x + 1) + blockIdx.x] = cnt;
}

// exclusive scan of `n` ints in place, total returned to all threads; one block, any n.
__device__ int block_scan_inplace(int* data, int n) {
  __shared__ int warp_sums[32];
  __shared__ int carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int base = 0; base < n; base += blockDim.x) {
    const int idx = base + threadIdx.x;
    const int v = idx < n ? data[idx] : 0;
    int inc = v;
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(FULL, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) warp_sums[wid] = inc;
    __syncthreads();
    if (wid == 0) {
      int w = lane < nw ? warp_sums[lane] : 0;
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(FULL, w, o);
        if (lane >= o) w += t;
      }
      warp_sums[lane] = w;  // inclusive scan of warp totals
    }
    __syncthreads();
    const int carry = carry_s;
    const int woff = wid > 0 ? warp_sums[wid - 1] : 0;
    if (idx < n) data[idx] = carry + woff + inc - v;
    __syncthreads();
    if (threadIdx.x == 0) carry_s = carry + warp_sums[nw - 1];
    __syncthreads();
  }
  return carry_s;
}

__global__ void k_scan_blocks(PostWorkspace ws, int nblk) {
  const int page = blockIdx.x;
  const int total = block_scan_inplace(ws.block_counts + page * (nblk + 1), nblk);
  if (threadIdx.x == 0) {
    PostHeader* h = reinterpret_cast<PostHeader*>(ws.result + page * ws.result_stride);
    h->n_comp = total;
  }
}

// roots get their final label (1-based raster rank) and reset their component's accumulators
__global__ void k_rank(PostWorkspace ws) {
  __shared__ int warp_cnt[32];
  const int page = blockIdx.y;
  const int hw = ws.H * ws.W;
  const int i = blockIdx.x * kPix + threadIdx.x;
  const int* par = ws.parent + static_cast<size_t>(page) * hw;
  const bool root = i < hw && par[i] == i;
  const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const unsigned bits = __ballot_sync(FULL, root);
  if (lane == 0) warp_cnt[wid] = __popc(bits);
  __syncthreads();
  if (wid == 0) {
    int w = warp_cnt[lane];
    int inc = w;
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(FULL, inc, o);
      if (lane >= o) inc += t;
    }
    warp_cnt[lane] = inc - w;  // exclusive
  }
  __syncthreads();
  if (root) {
    const int k = 1 + ws.block_counts[page * (gridDim.x + 1) + blockIdx.x] + warp_cnt[wid] +
                  __popc(bits & ((1u << lane) - 1));
    ws.rank[static_cast<size_t>(page) * hw + i] = k;
    const size_t c = static_cast<size_t>(page) * (hw / 2 + 2) + k;
    ws.c_area[c] = 0; ws.c_minx[c] = INT_MAX; ws.c_miny[c] = INT_MAX; ws.c_maxx[c] = -1; ws.c_maxy[c] = -1;
    ws.c_maxv[c] = 0u;
  }
}

// area / bbox / max normalised text per component; lanes of a warp holding the same label are
// reduced with redux.sync first, one atomic set per (warp, label)
__global__ void k_stats(PostWorkspace ws, const float* __restrict__ maps, const unsigned* mm) {
  const int page = blockIdx.y;
  const int hw = ws.H * ws.W, W = ws.W;
  const int i = blockIdx.x * kPix + threadIdx.x;
  const int* par = ws.parent + static_cast<size_t>(page) * hw;
  int k = 0;
  if (i < hw) {
    const int r = par[i];
    if (r >= 0) k = ws.rank[static_cast<size_t>(page) * hw + r];
    if (ws.labels != nullptr) ws.labels[static_cast<size_t>(page) * hw + i] = k;
  }
  const unsigned active = __ballot_sync(FULL, k > 0);
  if (k > 0) {
    const float tmin = dec_f32(mm[page * 4 + 0]), tmax = dec_f32(mm[page * 4 + 1]);
    const float t = __ldg(maps + (static_cast<size_t>(page) * hw + i) * 2);
    const unsigned tv = __float_as_uint(normalise(t, tmin, tmax));  // >= +0 (or NaN): uint order == float order
    const int x = i % W, y = i / W;
    const unsigned peers = __match_any_sync(active, k);
    const int area = __popc(peers);
    const int mnx = __reduce_min_sync(peers, x), mxx = __reduce_max_sync(peers, x);
    const int mny = __reduce_min_sync(peers, y), mxy = __reduce_max_sync(peers, y);
    const unsigned mv = __reduce_max_sync(peers, tv);
    if ((threadIdx.x & 31) == __ffs(peers) - 1) {
      const size_t c = static_cast<size_t>(page) * (hw / 2 + 2) + k;
      atomicAdd(&ws.c_area[c], area);
      atomicMin(&ws.c_minx[c], mnx); atomicMax(&ws.c_maxx[c], mxx);
      atomicMin(&ws.c_miny[c], mny); atomicMax(&ws.c_maxy[c], mxy);
      atomicMax(&ws.c_maxv[c], mv);
    }
  }
}

// exclusive scan of component heights -> first row slot of each component
__global__ void k_scan_heights(PostWorkspace ws) {
  const int page = blockIdx.x;
  const int hw = ws.H * ws.W;
  PostHeader* h = reinterpret_cast<PostHeader*>(ws.result + page * ws.result_stride);
  const int n = h->n_comp;
  const size_t cbase = static_cast<size_t>(page) * (hw / 2 + 2);
  int* off = ws.c_off + cbase + 1;  // entry for label k at off[k-1]
  for (int j = threadIdx.x; j < n; j += blockDim.x)
    off[j] = ws.c_maxy[cbase + 1 + j] - ws.c_miny[cbase + 1 + j] + 1;
  __syncthreads();
  const int total = block_scan_inplace(off, n);
  if (threadIdx.x == 0) {
    h->n_rows = total;
    h->overflow = (n > ws.comp_cap || total > ws.row_cap) ? 1 : 0;
  }
}

__global__ void k_rows_init(PostWorkspace ws) {
  const int page = blockIdx.y;
  const int hw = ws.H * ws.W;
  const PostHeader* h = reinterpret_cast<const PostHeader*>(ws.result + page * ws.result_stride);
  const int n = h->n_rows;
  PostRow* rows = ws.rows_full + static_cast<size_t>(page) * hw;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) rows[j] = PostRow{INT_MAX, -1};
}

// leftmost / rightmost surviving pixel of every (component, row): segmap after
// setTo(0, link_score==1 && text_score==0) (tuatara.cpp:156-160), before dilation
__global__ void k_rows(PostWorkspace ws) {
  const int page = blockIdx.y;
  const int hw = ws.H * ws.W, W = ws.W;
  const int i = blockIdx.x * kPix + threadIdx.x;
  int slot = -1, x = 0;
  if (i < hw) {
    const int r = ws.parent[static_cast<size_t>(page) * hw + i];
    const uint8_t fl = ws.flags[static_cast<size_t>(page) * hw + i];
    const bool keep = r >= 0 && !((fl & 4) && !(fl & 2));
    if (keep) {
      const int k = ws.rank[static_cast<size_t>(page) * hw + r];
      const size_t c = static_cast<size_t>(page) * (hw / 2 + 2) + k;
      x = i % W;
      slot = ws.c_off[c] + (i / W - ws.c_miny[c]);
    }
  }
  const unsigned active = __ballot_sync(FULL, slot >= 0);
  if (slot >= 0) {
    const unsigned peers = __match_any_sync(active, slot);
    const int mn = __reduce_min_sync(peers, x), mx = __reduce_max_sync(peers, x);
    if ((threadIdx.x & 31) == __ffs(peers) - 1) {
      PostRow* row = ws.rows_full + static_cast<size_t>(page) * hw + slot;
      atomicMin(&row->xmin, mn);
      atomicMax(&row->xmax, mx);
    }
  }
}

// compact result block: header | comps[comp_cap] | rows[row_cap]
__global__ void k_pack(PostWorkspace ws) {
  const int page = blockIdx.y;
  const int hw = ws.H * ws.W;
  uint8_t* blk = ws.result + page * ws.result_stride;
  const PostHeader* h = reinterpret_cast<const PostHeader*>(blk);
  PostComp* comps = reinterpret_cast<PostComp*>(blk + sizeof(PostHeader));
  PostRow* rows = reinterpret_cast<PostRow*>(blk + sizeof(PostHeader) + sizeof(PostComp) * ws.comp_cap);
  const int n = min(h->n_comp, ws.comp_cap);
  const int nr = min(h->n_rows, ws.row_cap);
  const size_t cbase = static_cast<size_t>(page) * (hw / 2 + 2);
  const int stride = gridDim.x * blockDim.x;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  for (int j = t; j < n; j += stride) {
    const size_t c = cbase + 1 + j;
    PostComp pc;
    pc.left = ws.c_minx[c]; pc.top = ws.c_miny[c]; pc.right = ws.c_maxx[c]; pc.bottom = ws.c_maxy[c];
    pc.area = ws.c_area[c]; pc.max_text = __uint_as_float(ws.c_maxv[c]); pc.row_off = ws.c_off[c]; pc.pad = 0;
    comps[j] = pc;
  }
  const PostRow* src = ws.rows_full + static_cast<size_t>(page) * hw;
  for (int j = t; j < nr; j += stride) rows[j] = src[j];
}

unsigned* mm_ptr(const PostWorkspace& ws) {
  // encoded min/max live behind the block_counts array
  const int nblk = (ws.H * ws.W + kPix - 1) / kPix;
  return reinterpret_cast<unsigned*>(ws.block_counts + static_cast<size_t>(ws.batch) * (nblk + 1));
}

}  // namespace

size_t post_result_stride(int comp_cap, int row_cap) {
  const size_t s = sizeof(PostHeader) + sizeof(PostComp) * comp_cap + sizeof(PostRow) * row_cap;
  return (s + 255) & ~static_cast<size_t>(255);
}

cudaError_t post_workspace_alloc(PostWorkspace* ws, int batch, int H, int W, int comp_cap, int row_cap,
                                 bool want_labels) {
  *ws = PostWorkspace{};
  ws->batch = batch; ws->H = H; ws->W = W; ws->comp_cap = comp_cap; ws->row_cap = row_cap;
  const size_t hw = static_cast<size_t>(H) * W;
  const size_t ncomp = hw / 2 + 2;
  const int nblk = static_cast<int>((hw + kPix - 1) / kPix);
  ws->result_stride = post_result_stride(comp_cap, row_cap);
  auto al = [](size_t b) { return (b + 255) & ~static_cast<size_t>(255); };
  size_t off = 0;
  const size_t o_parent = off; off += al(batch * hw * sizeof(int));
  const size_t o_rank = off; off += al(batch * hw * sizeof(int));
  const size_t o_flags = off; off += al(batch * hw);
  const size_t o_blk = off; off += al((static_cast<size_t>(batch) * (nblk + 1) + batch * 4) * sizeof(int));
  const size_t o_comp = off; off += 7 * al(batch * ncomp * sizeof(int));
  const size_t o_rows = off; off += al(batch * hw * sizeof(PostRow));
  const size_t o_res = off; off += al(batch * ws->result_stride);
  const size_t o_lab = off; if (want_labels) off += al(batch * hw * sizeof(int));
  uint8_t* base = nullptr;
  TT_CUDA_TRY(cudaMalloc(&base, off));
  ws->bytes = off;
  ws->parent = reinterpret_cast<int*>(base + o_parent);
  ws->rank = reinterpret_cast<int*>(base + o_rank);
  ws->flags = base + o_flags;
  ws->block_counts = reinterpret_cast<int*>(base + o_blk);
  const size_t cs = al(batch * ncomp * sizeof(int));
  ws->c_area = reinterpret_cast<int*>(base + o_comp + 0 * cs);
  ws->c_minx = reinterpret_cast<int*>(base + o_comp + 1 * cs);
  ws->c_miny = reinterpret_cast<int*>(base + o_comp + 2 * cs);
  ws->c_maxx = reinterpret_cast<int*>(base + o_comp + 3 * cs);
  ws->c_maxy = reinterpret_cast<int*>(base + o_comp + 4 * cs);
  ws->c_maxv = reinterpret_cast<unsigned*>(base + o_comp + 5 * cs);
  ws->c_off = reinterpret_cast<int*>(base + o_comp + 6 * cs);
  ws->rows_full = reinterpret_cast<PostRow*>(base + o_rows);
  ws->result = base + o_res;
  ws->labels = want_labels ? reinterpret_cast<int*>(base + o_lab) : nullptr;
  return cudaSuccess;
}

void post_workspace_free(PostWorkspace* ws) {
  if (ws->parent) cudaFree(ws->parent);  // base of the single allocation
  *ws = PostWorkspace{};
}

cudaError_t post_run(const PostWorkspace& ws, const float* maps, const PostParams& p, cudaStream_t s) {
  const int hw = ws.H * ws.W;
  const int nblk = (hw + kPix - 1) / kPix;
  const dim3 gpix(nblk, ws.batch);
  unsigned* mm = mm_ptr(ws);
  k_init<<<ws.batch, 32, 0, s>>>(ws, mm);
  TT_LAUNCH_CHECK();
  k_minmax<<<dim3(std::min(nblk, 64), ws.batch), 256, 0, s>>>(maps, hw, mm);
  TT_LAUNCH_CHECK();
  k_label_init<<<gpix, kPix, 0, s>>>(ws, maps, mm, p);
  TT_LAUNCH_CHECK();
  k_merge<<<gpix, kPix, 0, s>>>(ws);
  TT_LAUNCH_CHECK();
  k_compress<<<gpix, kPix, 0, s>>>(ws);
  TT_LAUNCH_CHECK();
  k_scan_blocks<<<ws.batch, 1024, 0, s>>>(ws, nblk);
  TT_LAUNCH_CHECK();
  k_rank<<<gpix, kPix, 0, s>>>(ws);
  TT_LAUNCH_CHECK();
  k_stats<<<gpix, kPix, 0, s>>>(ws, maps, mm);
  TT_LAUNCH_CHECK();
  k_scan_heights<<<ws.batch, 1024, 0, s>>>(ws);
  TT_LAUNCH_CHECK();
  k_rows_init<<<dim3(32, ws.batch), 256, 0, s>>>(ws);
  TT_LAUNCH_CHECK();
  k_rows<<<gpix, kPix, 0, s>>>(ws);
  TT_LAUNCH_CHECK();
  k_pack<<<dim3(16, ws.batch), 256, 0, s>>>(ws);
  TT_LAUNCH_CHECK();
  return cudaSuccess;
}

}  // namespace tt
