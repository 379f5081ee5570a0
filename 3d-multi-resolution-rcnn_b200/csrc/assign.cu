// Training-time glue between the proposal path and RoIAlign: dense 3D IoU, the
// MaxIoU assigner fused over it, and the bbox2delta3d target encoder, for B200 (sm_100a).
//
// Replaces (reference, /root/reference):
//   bbox_overlaps (6-column branch)      mmdet/core/bbox/geometry.py:49-60
//   MaxIoUAssigner.assign_wrt_overlaps   mmdet/core/bbox/assigners/max_iou_assigner.py:128-171
//   bbox2delta3d                         mmdet/core/bbox/transforms.py:33-63
//
// The reference materialises the [gts x boxes] IoU matrix with ~25 torch elementwise launches (for the RPN assigner
// that is gts x 1.6 M anchors), reduces it twice and then loops over the gts in Python.  Here the matrix is never
// stored: one pass computes each box's row maximum while reducing every gt's column maximum (warp shuffle -> shared
// atomicMax -> one global atomicMax per gt and CTA), a second pass applies the four assignment rules, recomputing the
// few IoUs it needs.  HBM traffic: 24 B read + 20..28 B written per box instead of 4*k bytes per box several times.
//
// Arithmetic: torch evaluates the reference's expression as separate fp32 elementwise kernels, i.e. every operation
// rounded, nothing contracted; the same sequence is pinned here with _rn intrinsics:
//   w = max(min(ax2,bx2) - max(ax1,bx1) + 1, 0) ...; inter = (w*h)*d; area = ((x2-x1+1)*(y2-y1+1))*(z2-z1+1);
//   iou = inter / ((areaA + areaB) - inter).
// Ties (parity unpinned by the reference: torch 1.0.1's max(dim) index on ties): the LOWEST index wins.
#include <float.h>

#include <cmath>

#include "common.cuh"

namespace roi3d {

struct GtBox {
  float x1, y1, x2, y2, z1, z2, area, pad;
};

__device__ __forceinline__ float box_area6(const float *b) {
  return __fmul_rn(__fmul_rn(__fadd_rn(__fsub_rn(b[2], b[0]), 1.0f), __fadd_rn(__fsub_rn(b[3], b[1]), 1.0f)),
                   __fadd_rn(__fsub_rn(b[5], b[4]), 1.0f));
}

// a = gt (bboxes1 of the reference call bbox_overlaps(gt_bboxes, bboxes)), b = candidate box
__device__ __forceinline__ float iou3d_torch(const GtBox &a, const float *b, float area_b) {
  const float w = fmaxf(__fadd_rn(__fsub_rn(fminf(a.x2, b[2]), fmaxf(a.x1, b[0])), 1.0f), 0.0f);
  const float h = fmaxf(__fadd_rn(__fsub_rn(fminf(a.y2, b[3]), fmaxf(a.y1, b[1])), 1.0f), 0.0f);
  const float d = fmaxf(__fadd_rn(__fsub_rn(fminf(a.z2, b[5]), fmaxf(a.z1, b[4])), 1.0f), 0.0f);
  const float inter = __fmul_rn(__fmul_rn(w, h), d);
  const float uni = __fsub_rn(__fadd_rn(a.area, area_b), inter);
  // disjoint boxes (the vast majority of pairs): 0 / uni is +0 for any positive union -- skip the IEEE division;
  // degenerate unions (<= 0, NaN) still go through it so that -0 / NaN come out as torch computes them
  if (inter == 0.0f && uni > 0.0f) return 0.0f;
  return __fdiv_rn(inter, uni);
}

// order-preserving unsigned key of a float for atomicMax, and its inverse.  Degenerate boxes (x2 < x1 - 1) have
// negative "areas", so an IoU can be negative or NaN; a (positive) NaN maps above +inf like torch.max propagates it.
__device__ __forceinline__ unsigned iou_key(float v) {
  const unsigned u = __float_as_uint(v + 0.0f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float iou_from_key(unsigned k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

constexpr int kGtTile = 512;  // gts staged per shared-memory tile (16 KB)

__device__ __forceinline__ void load_gt_tile(GtBox *sg, const float *gt, int k, int t0, int tid, int nthreads) {
  for (int i = tid; i < min(kGtTile, k - t0); i += nthreads) {
    const float *g = gt + (long long)(t0 + i) * 6;
    GtBox b;
    b.x1 = g[0], b.y1 = g[1], b.x2 = g[2], b.y2 = g[3], b.z1 = g[4], b.z2 = g[5];
    b.area = box_area6(g);
    b.pad = 0.0f;
    sg[i] = b;
  }
}

// ------------------------------------------------------------------------------------------------
// Dense IoU matrix [m, n] (bbox_overlaps, 6-column, not aligned).  grid (ceil(n/256), ceil(m/8)).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bbox_overlaps3d_kernel(const float *__restrict__ b1, int m, int s1,
                                                              const float *__restrict__ b2, int n, int s2,
                                                              float *__restrict__ out) {
  __shared__ GtBox rows[8];
  const int r0 = blockIdx.y * 8;
  if (threadIdx.x < 8 && r0 + threadIdx.x < m) {
    const float *g = b1 + (long long)(r0 + threadIdx.x) * s1;
    GtBox b;
    b.x1 = g[0], b.y1 = g[1], b.x2 = g[2], b.y2 = g[3], b.z1 = g[4], b.z2 = g[5];
    b.area = box_area6(g);
    b.pad = 0.0f;
    rows[threadIdx.x] = b;
  }
  __syncthreads();
  const int j = blockIdx.x * 256 + threadIdx.x;
  if (j >= n) return;
  float c[6];
#pragma unroll
  for (int q = 0; q < 6; ++q) c[q] = __ldg(b2 + (long long)j * s2 + q);
  const float area = box_area6(c);
  for (int r = 0; r < min(8, m - r0); ++r) out[(long long)(r0 + r) * n + j] = iou3d_torch(rows[r], c, area);
}

// ------------------------------------------------------------------------------------------------
// Assigner pass 1: per box max / argmax over the gts; per gt max over the boxes (global atomicMax).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) assign_pass1_kernel(const float *__restrict__ boxes, int n, int stride,
                                                           const float *__restrict__ gt, int k,
                                                           float *__restrict__ max_overlaps, int32_t *__restrict__ argmax,
                                                           unsigned *__restrict__ gt_max_key,
                                                           const uint8_t *__restrict__ ignore) {
  __shared__ GtBox sg[kGtTile];
  __shared__ unsigned smax[kGtTile];
  const int j = blockIdx.x * 256 + threadIdx.x;
  const bool live = j < n;
  // a box inside an ignore region: its whole column of the overlap matrix reads -1 (max_iou_assigner.py:111)
  const bool ign = ignore != nullptr && live && __ldg(ignore + j) != 0;
  float c[6] = {0, 0, 0, 0, 0, 0};
  if (live) {
#pragma unroll
    for (int q = 0; q < 6; ++q) c[q] = __ldg(boxes + (long long)j * stride + q);
  }
  const float area = box_area6(c);
  float best = -FLT_MAX;
  int besti = 0;
  bool best_nan = false;
  for (int t0 = 0; t0 < k; t0 += kGtTile) {
    const int tk = min(kGtTile, k - t0);
    __syncthreads();
    load_gt_tile(sg, gt, k, t0, threadIdx.x, 256);
    for (int i = threadIdx.x; i < tk; i += 256) smax[i] = 0u;
    __syncthreads();
    for (int i = 0; i < tk; ++i) {
      const float v = live ? (ign ? -1.0f : iou3d_torch(sg[i], c, area)) : 0.0f;
      // torch.max propagates NaN: the first NaN wins and stays
      if (!best_nan) {
        if (v != v) best = v, besti = t0 + i, best_nan = true;
        else if (v > best) best = v, besti = t0 + i;
      }
      unsigned key = live ? iou_key(v) : 0u;
      key = __reduce_max_sync(0xffffffffu, key);
      if ((threadIdx.x & 31) == 0 && key > smax[i]) atomicMax(&smax[i], key);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < tk; i += 256)
      if (smax[i] != 0u) atomicMax(gt_max_key + t0 + i, smax[i]);
  }
  if (live) {
    max_overlaps[j] = best;
    argmax[j] = besti;
  }
}

// ------------------------------------------------------------------------------------------------
// Assigner pass 2: rules 1-3 per box, and rule 4 either completely (gt_max_assign_all: every box whose IoU with gt
// i equals that gt's maximum gets i + 1, later gts override earlier ones) or its first half (the lowest box index
// attaining each gt's maximum, applied in gt order by assign_pass3_kernel).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) assign_pass2_kernel(const float *__restrict__ boxes, int n, int stride,
                                                           const float *__restrict__ gt, int k,
                                                           const float *__restrict__ max_overlaps,
                                                           const int32_t *__restrict__ argmax,
                                                           const unsigned *__restrict__ gt_max_key, float pos_thr,
                                                           float neg_lo, float neg_hi, float min_pos_iou, int assign_all,
                                                           int64_t *__restrict__ assigned, int32_t *__restrict__ gt_argmax,
                                                           const uint8_t *__restrict__ ignore) {
  __shared__ GtBox sg[kGtTile];
  __shared__ float sgmax[kGtTile];
  const int j = blockIdx.x * 256 + threadIdx.x;
  const bool live = j < n;
  const bool ign = ignore != nullptr && live && __ldg(ignore + j) != 0;
  float c[6] = {0, 0, 0, 0, 0, 0};
  if (live) {
#pragma unroll
    for (int q = 0; q < 6; ++q) c[q] = __ldg(boxes + (long long)j * stride + q);
  }
  const float area = box_area6(c);
  long long a = -1;
  if (live) {
    const float mo = max_overlaps[j];
    if (mo >= neg_lo && mo < neg_hi) a = 0;
    if (mo >= pos_thr) a = (long long)argmax[j] + 1;
  }
  for (int t0 = 0; t0 < k; t0 += kGtTile) {
    const int tk = min(kGtTile, k - t0);
    __syncthreads();
    load_gt_tile(sg, gt, k, t0, threadIdx.x, 256);
    for (int i = threadIdx.x; i < tk; i += 256) sgmax[i] = iou_from_key(gt_max_key[t0 + i]);
    __syncthreads();
    if (!live) continue;
    const float my_max = max_overlaps[j];
    for (int i = 0; i < tk; ++i) {
      const float gm = sgmax[i];
      if (!(gm >= min_pos_iou)) continue;
      // this box's IoU with gt i cannot exceed its own row maximum: nothing to recompute when that is below gm
      // (a NaN row maximum compares false and falls through to the exact test)
      if (my_max < gm) continue;
      if ((ign ? -1.0f : iou3d_torch(sg[i], c, area)) == gm) {
        if (assign_all) a = t0 + i + 1;
        else atomicMin(gt_argmax + t0 + i, j);
      }
    }
  }
  if (live) assigned[j] = a;
}

// gt_max_assign_all == False: assigned[gt_argmax[i]] = i + 1 in gt order (one thread; k is small)
__global__ void assign_pass3_kernel(const int32_t *__restrict__ gt_argmax, int k, int n, int64_t *__restrict__ assigned) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  for (int i = 0; i < k; ++i) {
    const int j = gt_argmax[i];
    if (j >= 0 && j < n) assigned[j] = i + 1;
  }
}

__global__ void __launch_bounds__(256) assign_labels_kernel(const int64_t *__restrict__ assigned, int n,
                                                            const int64_t *__restrict__ gt_labels,
                                                            int64_t *__restrict__ labels) {
  const int j = blockIdx.x * 256 + threadIdx.x;
  if (j >= n) return;
  const long long a = assigned[j];
  labels[j] = a > 0 ? gt_labels[a - 1] : 0;
}

// ------------------------------------------------------------------------------------------------
// bbox2delta3d (transforms.py:33-63): centre / size encoding, (x - mean) / std.
// ------------------------------------------------------------------------------------------------
struct Float6 {
  float v[6];
};

__global__ void __launch_bounds__(256) bbox2delta3d_kernel(const float *__restrict__ prop, int sp,
                                                           const float *__restrict__ gt, int sg, int n, Float6 means,
                                                           Float6 stds, float *__restrict__ out) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  float p[6], g[6];
#pragma unroll
  for (int q = 0; q < 6; ++q) p[q] = __ldg(prop + (long long)i * sp + q), g[q] = __ldg(gt + (long long)i * sg + q);
  float d[6];
  // pairs (0,2) (1,3) (4,5): centre = (lo + hi) * 0.5, size = hi - lo + 1
  const int lo[3] = {0, 1, 4}, hi[3] = {2, 3, 5}, dc[3] = {0, 1, 4}, ds[3] = {2, 3, 5};
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const float pc = __fmul_rn(__fadd_rn(p[lo[a]], p[hi[a]]), 0.5f), ps = __fadd_rn(__fsub_rn(p[hi[a]], p[lo[a]]), 1.0f);
    const float gc = __fmul_rn(__fadd_rn(g[lo[a]], g[hi[a]]), 0.5f), gs = __fadd_rn(__fsub_rn(g[hi[a]], g[lo[a]]), 1.0f);
    d[dc[a]] = __fdiv_rn(__fsub_rn(gc, pc), ps);
    d[ds[a]] = logf(__fdiv_rn(gs, ps));
  }
#pragma unroll
  for (int q = 0; q < 6; ++q) out[(long long)i * 6 + q] = __fdiv_rn(__fsub_rn(d[q], means.v[q]), stds.v[q]);
}

// Inverse of the above for the bbox head: [n, 6k] class-wise deltas applied to [n] boxes (delta2bbox3D,
// mmdet/core/bbox/transforms.py:105-160), one thread per (box, class).  All four size / depth terms are clamped with
// |log(wh_ratio_clip)| as the reference does (:122-128).  Same operation order as the fused proposal decode (proposal.cu).
__global__ void __launch_bounds__(256) delta2bbox3d_kernel(const float *__restrict__ rois, int sr,
                                                           const float *__restrict__ deltas, int n, int k, Float6 means,
                                                           Float6 stds, float max_ratio, float img_h, float img_w,
                                                           float img_d, float *__restrict__ out) {
  const long long t = (long long)blockIdx.x * 256 + threadIdx.x;
  if (t >= (long long)n * k) return;
  const int i = (int)(t / k), c = (int)(t - (long long)i * k);
  float r[6], d[6];
#pragma unroll
  for (int q = 0; q < 6; ++q) {
    r[q] = __ldg(rois + (long long)i * sr + q);
    d[q] = __fadd_rn(__fmul_rn(__ldg(deltas + ((long long)i * k + c) * 6 + q), stds.v[q]), means.v[q]);
  }
  float *o = out + ((long long)i * k + c) * 6;
  // pairs (0,2) (1,3) (4,5); deltas (0,2) (1,3) (4,5) = (centre shift, log size ratio)
  const int lo[3] = {0, 1, 4}, hi[3] = {2, 3, 5};
  const float lim[3] = {img_w, img_h, img_d};
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const float pc = __fmul_rn(__fadd_rn(r[lo[a]], r[hi[a]]), 0.5f), ps = __fadd_rn(__fsub_rn(r[hi[a]], r[lo[a]]), 1.0f);
    const float ds = fminf(fmaxf(d[hi[a]], -max_ratio), max_ratio);
    const float gs = __fmul_rn(ps, expf(ds)), gc = __fadd_rn(pc, __fmul_rn(ps, d[lo[a]]));
    float v1 = __fadd_rn(__fsub_rn(gc, __fmul_rn(gs, 0.5f)), 0.5f), v2 = __fsub_rn(__fadd_rn(gc, __fmul_rn(gs, 0.5f)), 0.5f);
    if (img_w > 0.0f) v1 = fminf(fmaxf(v1, 0.0f), lim[a] - 1.0f), v2 = fminf(fmaxf(v2, 0.0f), lim[a] - 1.0f);
    o[lo[a]] = v1, o[hi[a]] = v2;
  }
}

// ------------------------------------------------------------------------------------------------
// Anchors of one level in closed form + valid / inside flags.  Flat index = ((y*W + x)*D + z)*A + a
// (np.meshgrid(x, y, z) 'xy' order, anchor_generator_3d.py:59-70); anchor = base[a] + (x*s, y*s, x*s, y*s, z*sd, z*sd).
// flag = cell inside the valid extent (valid_flags, :73-92) AND, if allowed_border >= 0, the anchor inside the image
// grown by the border (anchor_inside_flags, anchor_target.py:203-217).
// ------------------------------------------------------------------------------------------------
struct AnchorParams {
  int A, D, H, W;
  float stride, dstride;
  float base[16][6];
  int valid_d, valid_h, valid_w;
  float img_h, img_w, img_d, border;
  int use_border;
  float *anchors;
  unsigned char *flags;
};

__global__ void __launch_bounds__(256) grid_anchors_kernel(const AnchorParams p) {
  const long long total = (long long)p.A * p.D * p.H * p.W;
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= total) return;
  long long t = i;
  const int a = (int)(t % p.A);
  t /= p.A;
  const int z = (int)(t % p.D);
  t /= p.D;
  const int x = (int)(t % p.W);
  const int y = (int)(t / p.W);
  const float sx = (float)x * p.stride, sy = (float)y * p.stride, sz = (float)z * p.dstride;
  float v[6];
  v[0] = __fadd_rn(p.base[a][0], sx), v[1] = __fadd_rn(p.base[a][1], sy), v[2] = __fadd_rn(p.base[a][2], sx);
  v[3] = __fadd_rn(p.base[a][3], sy), v[4] = __fadd_rn(p.base[a][4], sz), v[5] = __fadd_rn(p.base[a][5], sz);
  if (p.anchors != nullptr) {
#pragma unroll
    for (int q = 0; q < 6; ++q) p.anchors[i * 6 + q] = v[q];
  }
  if (p.flags != nullptr) {
    bool f = x < p.valid_w && y < p.valid_h && z < p.valid_d;
    if (p.use_border)
      f = f && v[0] >= -p.border && v[1] >= -p.border && v[4] >= -p.border && v[2] < p.img_w + p.border &&
          v[3] < p.img_h + p.border && v[5] < p.img_d + p.border;
    p.flags[i] = f ? 1 : 0;
  }
}

}  // namespace roi3d

using namespace roi3d;

extern "C" {

int roi3d_bbox_overlaps3d(const float *boxes1_dev, int m, int stride1, const float *boxes2_dev, int n, int stride2,
                          float *iou_dev, void *stream) {
  ROI3D_CHECK_ARG(m >= 0 && n >= 0 && stride1 >= 6 && stride2 >= 6, "bad sizes m=%d n=%d", m, n);
  if (m == 0 || n == 0) return ROI3D_OK;
  ROI3D_CHECK_ARG(boxes1_dev && boxes2_dev && iou_dev, "NULL pointer");
  ROI3D_CHECK_ARG(ceil_div(m, 8) <= 65535, "too many rows (m=%d)", m);
  bbox_overlaps3d_kernel<<<dim3(ceil_div(n, 256), ceil_div(m, 8)), 256, 0, (cudaStream_t)stream>>>(
      boxes1_dev, m, stride1, boxes2_dev, n, stride2, iou_dev);
  ROI3D_LAUNCH_CHECK();
  return ROI3D_OK;
}

size_t roi3d_assign_workspace_bytes(int n, int k) {
  auto up = [](size_t x) { return (x + 255) / 256 * 256; };
  return up(sizeof(int32_t) * (size_t)(n > 0 ? n : 0)) + 2 * up(sizeof(int32_t) * (size_t)(k > 0 ? k : 0));
}

int roi3d_assign_max_iou(const float *bboxes_dev, int n, int stride, const float *gt_dev, int k,
                         const int64_t *gt_labels_dev, float pos_iou_thr, float neg_iou_lo, float neg_iou_hi,
                         float min_pos_iou, int gt_max_assign_all, int64_t *assigned_gt_inds_dev,
                         float *max_overlaps_dev, int64_t *assigned_labels_dev, void *workspace_dev,
                         size_t workspace_bytes, void *stream) {
  return roi3d_assign_max_iou_ignore(bboxes_dev, n, stride, gt_dev, k, gt_labels_dev, nullptr, pos_iou_thr, neg_iou_lo,
                                     neg_iou_hi, min_pos_iou, gt_max_assign_all, assigned_gt_inds_dev, max_overlaps_dev,
                                     assigned_labels_dev, workspace_dev, workspace_bytes, stream);
}

int roi3d_assign_max_iou_ignore(const float *bboxes_dev, int n, int stride, const float *gt_dev, int k,
                                const int64_t *gt_labels_dev, const uint8_t *ignore_flags_dev, float pos_iou_thr,
                                float neg_iou_lo, float neg_iou_hi, float min_pos_iou, int gt_max_assign_all,
                                int64_t *assigned_gt_inds_dev, float *max_overlaps_dev, int64_t *assigned_labels_dev,
                                void *workspace_dev, size_t workspace_bytes, void *stream) {
  ROI3D_CHECK_ARG(n > 0 && k > 0, "No gt or bboxes (n=%d, k=%d)", n, k);  // the reference raises ValueError
  ROI3D_CHECK_ARG(stride >= 6, "bboxes need at least 6 columns");
  ROI3D_CHECK_ARG(bboxes_dev && gt_dev && assigned_gt_inds_dev && max_overlaps_dev && workspace_dev, "NULL pointer");
  ROI3D_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace_dev) & 255) == 0, "workspace must be 256-byte aligned");
  if (workspace_bytes < roi3d_assign_workspace_bytes(n, k)) {
    set_error("assign workspace too small: %zu < %zu", workspace_bytes, roi3d_assign_workspace_bytes(n, k));
    return ROI3D_ENOMEM;
  }
  auto up = [](size_t x) { return (x + 255) / 256 * 256; };
  char *w = static_cast<char *>(workspace_dev);
  int32_t *argmax = reinterpret_cast<int32_t *>(w);
  unsigned *gt_max_key = reinterpret_cast<unsigned *>(w + up(sizeof(int32_t) * (size_t)n));
  int32_t *gt_argmax = reinterpret_cast<int32_t *>(w + up(sizeof(int32_t) * (size_t)n) + up(sizeof(int32_t) * (size_t)k));
  cudaStream_t st = (cudaStream_t)stream;
  ROI3D_CUDA(cudaMemsetAsync(gt_max_key, 0, sizeof(unsigned) * (size_t)k, st));
  ROI3D_CUDA(cudaMemsetAsync(gt_argmax, 0x7f, sizeof(int32_t) * (size_t)k, st));  // 0x7f7f7f7f > any index
  const int blocks = ceil_div(n, 256);
  assign_pass1_kernel<<<blocks, 256, 0, st>>>(bboxes_dev, n, stride, gt_dev, k, max_overlaps_dev, argmax, gt_max_key,
                                              ignore_flags_dev);
  ROI3D_LAUNCH_CHECK();
  assign_pass2_kernel<<<blocks, 256, 0, st>>>(bboxes_dev, n, stride, gt_dev, k, max_overlaps_dev, argmax, gt_max_key,
                                              pos_iou_thr, neg_iou_lo, neg_iou_hi, min_pos_iou, gt_max_assign_all,
                                              assigned_gt_inds_dev, gt_argmax, ignore_flags_dev);
  ROI3D_LAUNCH_CHECK();
  if (!gt_max_assign_all) {
    assign_pass3_kernel<<<1, 32, 0, st>>>(gt_argmax, k, n, assigned_gt_inds_dev);
    ROI3D_LAUNCH_CHECK();
  }
  if (assigned_labels_dev != nullptr) {
    ROI3D_CHECK_ARG(gt_labels_dev != nullptr, "gt_labels is NULL");
    assign_labels_kernel<<<blocks, 256, 0, st>>>(assigned_gt_inds_dev, n, gt_labels_dev, assigned_labels_dev);
    ROI3D_LAUNCH_CHECK();
  }
  return ROI3D_OK;
}

int roi3d_grid_anchors(int A, int D, int H, int W, float stride, float depth_stride, const float *base_anchors_host,
                       int valid_d, int valid_h, int valid_w, float img_h, float img_w, float img_d, int allowed_border,
                       float *anchors_dev, uint8_t *flags_dev, void *stream) {
  ROI3D_CHECK_ARG(A >= 1 && A <= 16, "A=%d out of [1,16]", A);
  ROI3D_CHECK_ARG(D > 0 && H > 0 && W > 0 && base_anchors_host, "bad arguments");
  const long long total = (long long)A * D * H * W;
  ROI3D_CHECK_ARG(ceil_div_ll(total, 256) < 2147483647LL, "too many anchors");
  if (anchors_dev == nullptr && flags_dev == nullptr) return ROI3D_OK;
  AnchorParams p;
  p.A = A, p.D = D, p.H = H, p.W = W, p.stride = stride, p.dstride = depth_stride;
  for (int a = 0; a < A; ++a)
    for (int j = 0; j < 6; ++j) p.base[a][j] = base_anchors_host[a * 6 + j];
  p.valid_d = valid_d, p.valid_h = valid_h, p.valid_w = valid_w;
  p.img_h = img_h, p.img_w = img_w, p.img_d = img_d;
  p.use_border = allowed_border >= 0, p.border = (float)allowed_border;
  p.anchors = anchors_dev, p.flags = flags_dev;
  grid_anchors_kernel<<<(unsigned)ceil_div_ll(total, 256), 256, 0, (cudaStream_t)stream>>>(p);
  ROI3D_LAUNCH_CHECK();
  return ROI3D_OK;
}

int roi3d_bbox2delta3d(const float *proposals_dev, int stride_p, const float *gt_dev, int stride_g, int n,
                       const float *means6, const float *stds6, float *deltas_dev, void *stream) {
  ROI3D_CHECK_ARG(n >= 0 && stride_p >= 6 && stride_g >= 6, "bad sizes");
  if (n == 0) return ROI3D_OK;
  ROI3D_CHECK_ARG(proposals_dev && gt_dev && deltas_dev, "NULL pointer");
  Float6 m, s;
  for (int q = 0; q < 6; ++q) m.v[q] = means6 ? means6[q] : 0.0f, s.v[q] = stds6 ? stds6[q] : 1.0f;
  bbox2delta3d_kernel<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(proposals_dev, stride_p, gt_dev, stride_g, n, m,
                                                                          s, deltas_dev);
  ROI3D_LAUNCH_CHECK();
  return ROI3D_OK;
}

int roi3d_delta2bbox3d(const float *rois_dev, int stride_r, const float *deltas_dev, int n, int num_classes,
                       const float *means6, const float *stds6, float wh_ratio_clip, float img_h, float img_w, float img_d,
                       float *out_dev, void *stream) {
  ROI3D_CHECK_ARG(n >= 0 && num_classes >= 1 && stride_r >= 6, "bad sizes");
  ROI3D_CHECK_ARG(wh_ratio_clip > 0.0f, "wh_ratio_clip must be > 0");
  if (n == 0) return ROI3D_OK;
  ROI3D_CHECK_ARG(rois_dev && deltas_dev && out_dev, "NULL pointer");
  Float6 m, s;
  for (int q = 0; q < 6; ++q) m.v[q] = means6 ? means6[q] : 0.0f, s.v[q] = stds6 ? stds6[q] : 1.0f;
  // max_ratio = np.abs(np.log(wh_ratio_clip)) evaluated in float64, then used as a python float by clamp
  const float max_ratio = (float)std::fabs(std::log((double)wh_ratio_clip));
  const long long total = (long long)n * num_classes;
  delta2bbox3d_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      rois_dev, stride_r, deltas_dev, n, num_classes, m, s, max_ratio, img_h, img_w, img_d, out_dev);
  ROI3D_LAUNCH_CHECK();
  return ROI3D_OK;
}

}  // extern "C"
