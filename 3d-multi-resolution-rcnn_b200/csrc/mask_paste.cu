// Mask paste: the per-detection resize + threshold of FCNMaskHead3D.get_seg_masks on
// the device, for B200 (sm_100a).
//
// Replaces (reference, /root/reference): mmdet/models/mask_heads/fcn_mask_head_3d.py:144-187 -- per detection on the
// host: sigmoid -> numpy, `skimage.transform.resize(mask, (d, h, w))`, `> mask_thr_binary`, paste into a full-volume
// uint8 array.  `resize` is third-party (scikit-image==0.18.0 over scipy==1.5.4, requirements.txt:24-25; neither is
// under /root/reference).  Its n-dimensional branch is restated here from the published algorithm:
//   factors = in / out (float64);  anti-aliasing (default for float input): sigma = max(0, (factors - 1) / 2),
//   scipy.ndimage.gaussian_filter(mode='mirror', truncate 4.0) = per axis (z, y, x) a symmetric correlate1d in float64
//   accumulators with the result rounded to the image dtype (float32) after every axis;
//   coordinates c = factors * (o + 0.5) - 0.5;  scipy.ndimage.map_coordinates(order=1, mode='mirror'): the coordinate
//   is folded into [0, n-1] by reflection, weights (1 - t, t), the 8 corners accumulated z-outer / x-inner in float64,
//   rounded to float32;  clip to the filtered image's [min, max];  `> thr` in float32.
// Parity status: unpinned by the reference (no test exercises it) -- checked against the same restatement running on
// the scipy.ndimage that is installed in this image.
//
// One CTA per detection; the 14x14x10-sized mask lives in shared memory, arithmetic in float64 like scipy.
#include <algorithm>
#include <vector>

#include "common.cuh"

namespace roi3d {

constexpr int kMaskMaxVol = 4096;   // floats of one mask (14*14*10 = 1960)
constexpr int kMaskMaxTaps = 257;   // gaussian half width up to 128

__device__ __forceinline__ int mirror_index(int i, int n) {
  if (n == 1) return 0;
  const int p = 2 * n - 2;
  i %= p;
  if (i < 0) i += p;
  return i < n ? i : p - i;
}

// scipy's coordinate fold for mode 'mirror' (ni_interpolation.c, map_coordinate)
__device__ __forceinline__ double mirror_coord(double c, int n) {
  if (n <= 1) return 0.0;
  const double len = (double)n - 1.0;
  if (c < 0.0) {
    const double sz2 = 2.0 * len;
    c = sz2 * (double)(long long)(-c / sz2) + c;
    c = c <= -len ? c + sz2 : -c;
  } else if (c > len) {
    const double sz2 = 2.0 * len;
    c -= sz2 * (double)(long long)(c / sz2);
    if (c > len) c = sz2 - c;
  }
  return c;
}

__global__ void __launch_bounds__(256) mask_paste_kernel(const float *__restrict__ logits, int Dm, int Hm, int Wm,
                                                         const int32_t *__restrict__ boxes,
                                                         const int64_t *__restrict__ offsets, float thr,
                                                         unsigned char *__restrict__ out) {
  __shared__ float A[kMaskMaxVol];
  __shared__ float B[kMaskMaxVol];
  __shared__ double wts[kMaskMaxTaps];
  __shared__ float red_min[256], red_max[256];
  const int det = blockIdx.x, tid = threadIdx.x;
  const int vol = Dm * Hm * Wm;
  const int32_t *bx = boxes + (long long)det * 6;
  const int ow = max(bx[2] - bx[0] + 1, 1), oh = max(bx[3] - bx[1] + 1, 1), od = max(bx[5] - bx[4] + 1, 1);
  const float *src = logits + (long long)det * vol;
  for (int i = tid; i < vol; i += 256) A[i] = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-src[i])));  // torch sigmoid
  __syncthreads();
  const int dims[3] = {Dm, Hm, Wm};
  const int outs[3] = {od, oh, ow};
  const int strides[3] = {Hm * Wm, Wm, 1};
  float *cur = A, *nxt = B;
  // ---- anti-aliasing gaussian, axis by axis ----
  for (int ax = 0; ax < 3; ++ax) {
    const double factor = (double)dims[ax] / (double)outs[ax];
    const double sigma = fmax(0.0, (factor - 1.0) / 2.0);
    if (!(sigma > 1e-15)) continue;
    int lw = (int)(4.0 * sigma + 0.5);
    if (lw > (kMaskMaxTaps - 1) / 2) lw = (kMaskMaxTaps - 1) / 2;  // (cannot happen for masks up to 64 voxels per axis)
    for (int i = tid; i <= 2 * lw; i += 256) {
      const double x = (double)(i - lw);
      wts[i] = exp(-0.5 / (sigma * sigma) * (x * x));
    }
    __syncthreads();
    if (tid == 0) {
      double sum = 0.0;
      for (int i = 0; i <= 2 * lw; ++i) sum += wts[i];
      for (int i = 0; i <= 2 * lw; ++i) wts[i] = wts[i] / sum;
    }
    __syncthreads();
    const int n = dims[ax], st = strides[ax];
    for (int e = tid; e < vol; e += 256) {
      const int pos = (e / st) % n;
      const int base = e - pos * st;
      double tmp = (double)cur[e] * wts[lw];
      for (int ll = -lw; ll < 0; ++ll) {
        const double l = (double)cur[base + mirror_index(pos + ll, n) * st];
        const double r = (double)cur[base + mirror_index(pos - ll, n) * st];
        tmp += (l + r) * wts[ll + lw];
      }
      nxt[e] = (float)tmp;
    }
    __syncthreads();
    float *t = cur;
    cur = nxt, nxt = t;
  }
  // ---- min / max of the filtered image (clip=True) ----
  float mn = INFINITY, mx = -INFINITY;
  for (int i = tid; i < vol; i += 256) mn = fminf(mn, cur[i]), mx = fmaxf(mx, cur[i]);
  red_min[tid] = mn, red_max[tid] = mx;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (tid < o) red_min[tid] = fminf(red_min[tid], red_min[tid + o]), red_max[tid] = fmaxf(red_max[tid], red_max[tid + o]);
    __syncthreads();
  }
  mn = red_min[0], mx = red_max[0];
  // ---- trilinear resample at the box resolution, threshold ----
  const double fz = (double)Dm / (double)od, fy = (double)Hm / (double)oh, fx = (double)Wm / (double)ow;
  const long long total = (long long)od * oh * ow;
  unsigned char *dst = out + offsets[det];
  for (long long o = tid; o < total; o += 256) {
    const int x = (int)(o % ow);
    const long long t1 = o / ow;
    const int y = (int)(t1 % oh), z = (int)(t1 / oh);
    const double cz = mirror_coord(fz * ((double)z + 0.5) - 0.5, Dm);
    const double cy = mirror_coord(fy * ((double)y + 0.5) - 0.5, Hm);
    const double cx = mirror_coord(fx * ((double)x + 0.5) - 0.5, Wm);
    const int z0 = (int)floor(cz), y0 = (int)floor(cy), x0 = (int)floor(cx);
    const double tz = cz - (double)z0, ty = cy - (double)y0, tx = cx - (double)x0;
    const double wz[2] = {1.0 - tz, tz}, wy[2] = {1.0 - ty, ty}, wx[2] = {1.0 - tx, tx};
    const int zi[2] = {mirror_index(z0, Dm), mirror_index(z0 + 1, Dm)};
    const int yi[2] = {mirror_index(y0, Hm), mirror_index(y0 + 1, Hm)};
    const int xi[2] = {mirror_index(x0, Wm), mirror_index(x0 + 1, Wm)};
    double acc = 0.0;
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          double coeff = (double)cur[(zi[a] * Hm + yi[b]) * Wm + xi[c]];
          coeff *= wz[a];
          coeff *= wy[b];
          coeff *= wx[c];
          acc += coeff;
        }
    float v = (float)acc;
    v = fminf(fmaxf(v, mn), mx);
    dst[o] = v > thr ? 1 : 0;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Mask targets (training half): mask_target_single, mmdet/core/mask/mask_target.py:17-50.
// Per positive proposal on the host the reference crops the assigned ground-truth mask to the proposal's int32 box,
// runs `255 * skimage.transform.resize(crop, (mask_size_depth, mask_size, mask_size))`, casts to uint8 (truncation)
// and sets every non-zero voxel to 1.  The crop is a uint8 array, so `resize` first maps it to float64 with
// v * (1 / 255) (skimage.util.img_as_float) and everything after that -- gaussian anti-aliasing, order-1
// map_coordinates, clip -- runs in float64.  Same restatement of the two scipy.ndimage primitives as mask_paste_kernel
// (operation order included: with {0,1} masks the reference's result hinges on the last bit of 255 * (sum w_i / 255)).
// One CTA per proposal; the filtered crop lives in a global float64 workspace (two buffers of the crop's volume).
// ---------------------------------------------------------------------------------------------------------------
struct MaskTargetJob {
  long long ws_off;   // doubles from the workspace base: two buffers of vol doubles each
  int gt;             // ground-truth mask index
  int x1, y1, z1;     // crop origin (>= 0)
  int w, h, d;        // crop size after numpy's slice clipping; any <= 0 -> all-zero target
};

__global__ void __launch_bounds__(256) mask_target_kernel(const unsigned char *__restrict__ gt_masks, int D, int H, int W,
                                                          const MaskTargetJob *__restrict__ jobs, int Md, int Mh, int Mw,
                                                          double *__restrict__ ws, float *__restrict__ out) {
  __shared__ double wts[kMaskMaxTaps];
  __shared__ double red_min[256], red_max[256];
  const int job = blockIdx.x, tid = threadIdx.x;
  const MaskTargetJob j = jobs[job];
  float *dst = out + (long long)job * Md * Mh * Mw;
  const int ovol = Md * Mh * Mw;
  if (j.w <= 0 || j.h <= 0 || j.d <= 0) {
    for (int i = tid; i < ovol; i += 256) dst[i] = 0.0f;
    return;
  }
  const long long vol = (long long)j.d * j.h * j.w;
  double *cur = ws + j.ws_off, *nxt = cur + vol;
  const unsigned char *src = gt_masks + (long long)j.gt * D * H * W;
  for (long long e = tid; e < vol; e += 256) {
    const int x = (int)(e % j.w);
    const long long t = e / j.w;
    const int y = (int)(t % j.h), z = (int)(t / j.h);
    cur[e] = (double)src[((long long)(j.z1 + z) * H + (j.y1 + y)) * W + (j.x1 + x)] * (1.0 / 255.0);  // img_as_float
  }
  __syncthreads();
  const int dims[3] = {j.d, j.h, j.w};
  const int outs[3] = {Md, Mh, Mw};
  const long long strides[3] = {(long long)j.h * j.w, (long long)j.w, 1};
  for (int ax = 0; ax < 3; ++ax) {
    const double factor = (double)dims[ax] / (double)outs[ax];
    const double sigma = fmax(0.0, (factor - 1.0) / 2.0);
    if (!(sigma > 1e-15)) continue;   // scipy's gaussian_filter skips axes with sigma <= 1e-15
    int lw = (int)(4.0 * sigma + 0.5);
    if (lw > (kMaskMaxTaps - 1) / 2) lw = (kMaskMaxTaps - 1) / 2;
    for (int i = tid; i <= 2 * lw; i += 256) {
      const double x = (double)(i - lw);
      wts[i] = exp(-0.5 / (sigma * sigma) * (x * x));
    }
    __syncthreads();
    if (tid == 0) {
      double sum = 0.0;
      for (int i = 0; i <= 2 * lw; ++i) sum += wts[i];
      for (int i = 0; i <= 2 * lw; ++i) wts[i] = wts[i] / sum;
    }
    __syncthreads();
    const int n = dims[ax];
    const long long st = strides[ax];
    for (long long e = tid; e < vol; e += 256) {
      const int pos = (int)((e / st) % n);
      const long long base = e - (long long)pos * st;
      double tmp = __dmul_rn(cur[e], wts[lw]);   // scipy's correlate1d: separate multiplies and adds, no contraction
      for (int ll = -lw; ll < 0; ++ll) {
        const double l = cur[base + (long long)mirror_index(pos + ll, n) * st];
        const double r = cur[base + (long long)mirror_index(pos - ll, n) * st];
        tmp = __dadd_rn(tmp, __dmul_rn(__dadd_rn(l, r), wts[ll + lw]));
      }
      nxt[e] = tmp;
    }
    __syncthreads();
    double *t = cur;
    cur = nxt, nxt = t;
  }
  double mn = INFINITY, mx = -INFINITY;
  for (long long i = tid; i < vol; i += 256) mn = fmin(mn, cur[i]), mx = fmax(mx, cur[i]);
  red_min[tid] = mn, red_max[tid] = mx;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (tid < o) red_min[tid] = fmin(red_min[tid], red_min[tid + o]), red_max[tid] = fmax(red_max[tid], red_max[tid + o]);
    __syncthreads();
  }
  mn = red_min[0], mx = red_max[0];
  const double fz = (double)j.d / (double)Md, fy = (double)j.h / (double)Mh, fx = (double)j.w / (double)Mw;
  for (int o = tid; o < ovol; o += 256) {
    const int x = o % Mw;
    const int t1 = o / Mw;
    const int y = t1 % Mh, z = t1 / Mh;
    const double cz = mirror_coord(fz * ((double)z + 0.5) - 0.5, j.d);
    const double cy = mirror_coord(fy * ((double)y + 0.5) - 0.5, j.h);
    const double cx = mirror_coord(fx * ((double)x + 0.5) - 0.5, j.w);
    const int z0 = (int)floor(cz), y0 = (int)floor(cy), x0 = (int)floor(cx);
    const double tz = cz - (double)z0, ty = cy - (double)y0, tx = cx - (double)x0;
    const double wz[2] = {1.0 - tz, tz}, wy[2] = {1.0 - ty, ty}, wx[2] = {1.0 - tx, tx};
    const int zi[2] = {mirror_index(z0, j.d), mirror_index(z0 + 1, j.d)};
    const int yi[2] = {mirror_index(y0, j.h), mirror_index(y0 + 1, j.h)};
    const int xi[2] = {mirror_index(x0, j.w), mirror_index(x0 + 1, j.w)};
    double acc = 0.0;
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          double coeff = cur[((long long)zi[a] * j.h + yi[b]) * j.w + xi[c]];
          coeff = __dmul_rn(coeff, wz[a]);
          coeff = __dmul_rn(coeff, wy[b]);
          coeff = __dmul_rn(coeff, wx[c]);
          acc = __dadd_rn(acc, coeff);
        }
    acc = fmin(fmax(acc, mn), mx);                       // clip=True
    const double scaled = __dmul_rn(255.0, acc);         // `255 * resize(...)`
    const unsigned char u = (unsigned char)(int)scaled;  // .astype(np.uint8): truncation (values are in [0, 255])
    dst[o] = u > 0 ? 1.0f : 0.0f;                        // target[target > 0] = 1, then .float()
  }
}

}  // namespace roi3d

using namespace roi3d;

extern "C" {

int roi3d_mask_paste(const float *mask_logits_dev, int n, int Dm, int Hm, int Wm, const int32_t *boxes_dev,
                     const int64_t *offsets_dev, float thr, uint8_t *out_dev, void *stream) {
  ROI3D_CHECK_ARG(n >= 0 && Dm > 0 && Hm > 0 && Wm > 0, "bad sizes");
  if (n == 0) return ROI3D_OK;
  ROI3D_CHECK_ARG(mask_logits_dev && boxes_dev && offsets_dev && out_dev, "NULL pointer");
  ROI3D_CHECK_ARG((long long)Dm * Hm * Wm <= kMaskMaxVol, "mask of %dx%dx%d exceeds %d voxels", Dm, Hm, Wm, kMaskMaxVol);
  ROI3D_CHECK_ARG(Dm <= 64 && Hm <= 64 && Wm <= 64, "mask axes longer than 64 are not supported");
  mask_paste_kernel<<<n, 256, 0, (cudaStream_t)stream>>>(mask_logits_dev, Dm, Hm, Wm, boxes_dev, offsets_dev, thr, out_dev);
  ROI3D_LAUNCH_CHECK();
  return ROI3D_OK;
}

size_t roi3d_mask_target_workspace_bytes(const int32_t *crop_dhw_host, int n) {
  size_t doubles = 0;
  for (int i = 0; i < n; ++i) {
    const long long d = crop_dhw_host[i * 3], h = crop_dhw_host[i * 3 + 1], w = crop_dhw_host[i * 3 + 2];
    if (d > 0 && h > 0 && w > 0) doubles += 2 * (size_t)(d * h * w);
  }
  return doubles * sizeof(double) + 256 + (size_t)n * sizeof(MaskTargetJob);
}

int roi3d_mask_target(const uint8_t *gt_masks_dev, int G, int D, int H, int W, const int32_t *boxes_host,
                      const int64_t *gt_inds_host, int n, int Md, int Mh, int Mw, float *out_dev, void *workspace_dev,
                      size_t workspace_bytes, void *stream) {
  ROI3D_CHECK_ARG(n >= 0 && G >= 0 && D > 0 && H > 0 && W > 0 && Md > 0 && Mh > 0 && Mw > 0, "bad sizes");
  if (n == 0) return ROI3D_OK;
  ROI3D_CHECK_ARG(gt_masks_dev && boxes_host && gt_inds_host && out_dev && workspace_dev, "NULL pointer");
  ROI3D_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace_dev) & 255) == 0, "workspace must be 256-byte aligned");
  std::vector<MaskTargetJob> jobs((size_t)n);
  std::vector<int32_t> crops((size_t)n * 3);
  long long off = 0;
  for (int i = 0; i < n; ++i) {
    const int32_t *b = boxes_host + (size_t)i * 6;  // x1, y1, x2, y2, z1, z2 after `.astype(np.int32)`
    ROI3D_CHECK_ARG(gt_inds_host[i] >= 0 && gt_inds_host[i] < G, "proposal %d: gt index %lld out of [0,%d)", i,
                    (long long)gt_inds_host[i], G);
    ROI3D_CHECK_ARG(b[0] >= 0 && b[1] >= 0 && b[4] >= 0, "proposal %d starts at a negative coordinate", i);
    const int w = std::max(b[2] - b[0] + 1, 1), h = std::max(b[3] - b[1] + 1, 1), d = std::max(b[5] - b[4] + 1, 1);
    MaskTargetJob &j = jobs[(size_t)i];
    j.gt = (int)gt_inds_host[i], j.x1 = b[0], j.y1 = b[1], j.z1 = b[4];
    j.w = std::min(b[0] + w, W) - b[0], j.h = std::min(b[1] + h, H) - b[1], j.d = std::min(b[4] + d, D) - b[4];  // slice clipping
    j.ws_off = off;
    crops[(size_t)i * 3] = j.d, crops[(size_t)i * 3 + 1] = j.h, crops[(size_t)i * 3 + 2] = j.w;
    if (j.w > 0 && j.h > 0 && j.d > 0) off += 2LL * j.d * j.h * j.w;
  }
  const size_t need = roi3d_mask_target_workspace_bytes(crops.data(), n);
  if (workspace_bytes < need) {
    set_error("mask_target workspace too small: %zu < %zu", workspace_bytes, need);
    return ROI3D_ENOMEM;
  }
  cudaStream_t st = (cudaStream_t)stream;
  double *ws = static_cast<double *>(workspace_dev);
  MaskTargetJob *jobs_dev = reinterpret_cast<MaskTargetJob *>(static_cast<char *>(workspace_dev) + ((size_t)off * sizeof(double) + 255) / 256 * 256);
  ROI3D_CUDA(cudaMemcpyAsync(jobs_dev, jobs.data(), jobs.size() * sizeof(MaskTargetJob), cudaMemcpyHostToDevice, st));
  ROI3D_CUDA(cudaStreamSynchronize(st));  // `jobs` is a stack-owned pageable buffer
  mask_target_kernel<<<n, 256, 0, st>>>(gt_masks_dev, D, H, W, jobs_dev, Md, Mh, Mw, ws, out_dev);
  ROI3D_LAUNCH_CHECK();
  return ROI3D_OK;
}

}  // extern "C"
