// Library plumbing (error text, device info) and the host-buffer entry points: the same calls the
// reference's Python wrappers make with numpy / CPU tensors (mmdet/ops/nms/nms_wrapper.py:29-32),
// with the H2D / D2H copies inside the call.  These are what bench.py's `e2e` leg times.
#include <stdarg.h>
#include <string.h>

#include <algorithm>
#include <cmath>
#include <mutex>
#include <vector>

#include "common.cuh"

namespace roi3d {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// A tiny grow-only device/pinned scratch cache so repeated host-buffer calls do not pay cudaMalloc.
struct Scratch {
  void *dev = nullptr;
  size_t dev_bytes = 0;
  void *pin = nullptr;
  size_t pin_bytes = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t compute = nullptr, d2h = nullptr;  // the pipelined host-buffer forward uses three streams
  std::vector<cudaEvent_t> events;
  std::mutex mu;
};
static Scratch g_scratch_dev[64];  // one grow-only scratch per CUDA device (one process may drive several)

static int current_scratch(Scratch **out) {
  int dev = 0;
  ROI3D_CUDA(cudaGetDevice(&dev));
  ROI3D_CHECK_ARG(dev >= 0 && dev < 64, "device index %d out of range", dev);
  *out = &g_scratch_dev[dev];
  return ROI3D_OK;
}

static int ensure(Scratch &s, size_t dev_bytes, size_t pin_bytes) {
  if (!s.stream) ROI3D_CUDA(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
  if (dev_bytes > s.dev_bytes) {
    if (s.dev) cudaFree(s.dev);
    s.dev = nullptr, s.dev_bytes = 0;
    ROI3D_CUDA(cudaMalloc(&s.dev, dev_bytes));
    s.dev_bytes = dev_bytes;
  }
  if (pin_bytes > s.pin_bytes) {
    if (s.pin) cudaFreeHost(s.pin);
    s.pin = nullptr, s.pin_bytes = 0;
    ROI3D_CUDA(cudaMallocHost(&s.pin, pin_bytes));
    s.pin_bytes = pin_bytes;
  }
  return ROI3D_OK;
}

static size_t up256(size_t x) { return (x + 255) / 256 * 256; }

int g_host_pipeline_kb = 0;  // roi3d_set_tuning key 4: -1 = never pipeline the host forward, 0 = auto (>= 32 MB), N = >= N KB


// Device -> mapped pinned host rows: out_host[rows[i]] = staged[i] for one group of finished RoIs.  One kernel per
// group instead of one DMA descriptor per RoI (a 351 KB copy costs about 4 us of fixed overhead; 512 of them were
// 2 ms of the call).  16-byte coalesced stores, enough CTAs in flight to keep the PCIe write path busy.
__global__ void __launch_bounds__(256) scatter_rows_to_host_kernel(const float4 *__restrict__ staged, float4 *__restrict__ host,
                                                                  const int32_t *__restrict__ rows, int cnt, long long row_vec4) {
  const long long total = (long long)cnt * row_vec4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / row_vec4, c = i - r * row_vec4;
    __stcs(host + (long long)__ldg(rows + r) * row_vec4 + c, __ldcs(staged + i));
  }
}

}  // namespace roi3d

using namespace roi3d;

extern "C" {

int roi3d_abi_version(void) { return ROI3D_ABI_VERSION; }

const char *roi3d_last_error(void) { return g_err; }

int roi3d_device_info(int *sm_count, int *cc_major, int *cc_minor, size_t *l2_bytes) {
  int dev = 0;
  ROI3D_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  ROI3D_CUDA(cudaGetDeviceProperties(&prop, dev));
  if (sm_count) *sm_count = prop.multiProcessorCount;
  if (cc_major) *cc_major = prop.major;
  if (cc_minor) *cc_minor = prop.minor;
  if (l2_bytes) *l2_bytes = (size_t)prop.l2CacheSize;
  return ROI3D_OK;
}

int roi3d_nms3d_host(const float *dets_host, int n, float iou_thr, int64_t *keep_host, int32_t *num_keep_host) {
  ROI3D_CHECK_ARG(n >= 0 && num_keep_host, "bad arguments");
  if (n == 0) {
    *num_keep_host = 0;
    return ROI3D_OK;
  }
  ROI3D_CHECK_ARG(dets_host && keep_host, "NULL pointer");
  Scratch *sc = nullptr;
  {
    int rc0 = current_scratch(&sc);
    if (rc0) return rc0;
  }
  Scratch &g_scratch = *sc;
  std::lock_guard<std::mutex> lock(g_scratch.mu);
  const size_t dets_b = up256(sizeof(float) * 7 * (size_t)n), keep_b = up256(sizeof(int64_t) * (size_t)n);
  const size_t ws_b = roi3d_nms3d_workspace_bytes(1, n);
  int rc = ensure(g_scratch, dets_b + keep_b + 256 + ws_b, 0);
  if (rc) return rc;
  char *d = static_cast<char *>(g_scratch.dev);
  float *dets_dev = reinterpret_cast<float *>(d);
  int64_t *keep_dev = reinterpret_cast<int64_t *>(d + dets_b);
  int32_t *cnt_dev = reinterpret_cast<int32_t *>(d + dets_b + keep_b);
  void *ws = d + dets_b + keep_b + 256;
  cudaStream_t st = g_scratch.stream;
  ROI3D_CUDA(cudaMemcpyAsync(dets_dev, dets_host, sizeof(float) * 7 * (size_t)n, cudaMemcpyHostToDevice, st));
  rc = roi3d_nms3d_batched(dets_dev, nullptr, 1, n, iou_thr, keep_dev, nullptr, cnt_dev, ws, ws_b, st);
  if (rc) return rc;
  ROI3D_CUDA(cudaMemcpyAsync(num_keep_host, cnt_dev, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  // the keep list is at most n entries: copy all of it in the same stream, then one sync
  ROI3D_CUDA(cudaMemcpyAsync(keep_host, keep_dev, sizeof(int64_t) * (size_t)n, cudaMemcpyDeviceToHost, st));
  ROI3D_CUDA(cudaStreamSynchronize(st));
  return ROI3D_OK;
}

// Events for the pipelined forward, created once per device.
static int ensure_pipeline(Scratch &s, size_t n_events) {
  if (!s.compute) ROI3D_CUDA(cudaStreamCreateWithFlags(&s.compute, cudaStreamNonBlocking));
  if (!s.d2h) ROI3D_CUDA(cudaStreamCreateWithFlags(&s.d2h, cudaStreamNonBlocking));
  while (s.events.size() < n_events) {
    cudaEvent_t e;
    ROI3D_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    s.events.push_back(e);
  }
  return ROI3D_OK;
}

// Last feature slice (z index) a RoI can touch: samples lie below max((z2 + 1) * s, z1 * s), the trilinear tap
// reaches one voxel further (csrc/common.cuh axis_tap), plus one voxel of margin for fp32 rounding.
static int roi_last_slice(const float *roi, float scale_d, int D) {
  const float zend = (std::max(roi[5], roi[6]) + 1.0f) * scale_d;
  if (!(zend == zend) || zend >= (float)D) return D - 1;  // NaN or past the volume: wait for everything
  if (zend < 0.0f) return 0;
  return std::min(D - 1, (int)std::floor(zend) + 2);
}

int roi3d_roi_align3d_forward_host(const float *feats_host, int layout, int B, int C, int D, int H, int W,
                                   const float *rois_host, int K, int PD, int PH, int PW, float spatial_scale,
                                   float spatial_scale_depth, int sample_num, float *out_host) {
  ROI3D_CHECK_ARG(B > 0 && C > 0 && D > 0 && H > 0 && W > 0 && K >= 0 && PD > 0 && PH > 0 && PW > 0, "bad sizes");
  if (K == 0) return ROI3D_OK;
  ROI3D_CHECK_ARG(feats_host && rois_host && out_host, "NULL pointer");
  ROI3D_CHECK_ARG(layout == ROI3D_NCDHW || layout == ROI3D_NDHWC, "bad layout");
  Scratch *sc = nullptr;
  {
    int rc0 = current_scratch(&sc);
    if (rc0) return rc0;
  }
  Scratch &g_scratch = *sc;
  std::lock_guard<std::mutex> lock(g_scratch.mu);
  const size_t S = (size_t)D * H * W;
  const size_t feat_b = up256(sizeof(float) * (size_t)B * C * S);
  const size_t rois_b = up256(sizeof(float) * 7 * (size_t)K);
  const size_t roi_out = (size_t)C * PD * PH * PW;  // floats per RoI
  const size_t out_b = up256(sizeof(float) * (size_t)K * roi_out);
  const size_t conv_b = layout == ROI3D_NCDHW ? feat_b : 0;
  const size_t rows_b = up256(sizeof(int32_t) * (size_t)K);
  const size_t perm_b = up256(sizeof(float) * 7 * (size_t)K);
  int rc = ensure(g_scratch, feat_b + conv_b + rois_b + out_b + rows_b, perm_b + rows_b);
  if (rc) return rc;
  char *d = static_cast<char *>(g_scratch.dev);
  float *feat_dev = reinterpret_cast<float *>(d);
  float *conv_dev = reinterpret_cast<float *>(d + feat_b);
  float *rois_dev = reinterpret_cast<float *>(d + feat_b + conv_b);
  float *out_dev = reinterpret_cast<float *>(d + feat_b + conv_b + rois_b);
  cudaStream_t st = g_scratch.stream;

  // ---- pipelined path (one volume, large enough to matter): the volume travels in z slabs; RoIs are grouped by
  //      the last slab they touch, each group runs as soon as its slabs are resident and converted, and its
  //      output rows go back to the host while later slabs are still arriving (PCIe is full duplex).  The
  //      whole call then costs about max(H2D, D2H) instead of their sum.
  const bool pipelined = B == 1 && D >= 8 && g_host_pipeline_kb >= 0 &&
                         (size_t)C * S * sizeof(float) >= (g_host_pipeline_kb > 0 ? (size_t)g_host_pipeline_kb << 10 : (size_t)32 << 20) &&
                         (((size_t)H * W) % 4 == 0 || layout == ROI3D_NDHWC);
  if (pipelined) {
    // eight slabs of D/10 slices, then four of D/20: the rows of the last group are all that is left to return
    // once the upload has finished, so the last slabs are the thin ones
    std::vector<int> slab_end;  // exclusive z bound of slab j
    for (int j = 1; j <= 12; ++j) {
      const int w = j <= 8 ? 2 * j : 16 + (j - 8);
      const int end = j == 12 ? D : (int)((long long)D * w / 20);
      if (end > (slab_end.empty() ? 0 : slab_end.back())) slab_end.push_back(end);
    }
    const int nslab = (int)slab_end.size();
    rc = ensure_pipeline(g_scratch, 2 * (size_t)nslab);
    if (rc) return rc;
    // group of a RoI = first slab whose end lies past the RoI's last slice
    std::vector<int> group(K), order(K);
    for (int k = 0; k < K; ++k) {
      const int last = roi_last_slice(rois_host + (size_t)k * 7, spatial_scale_depth, D);
      group[k] = (int)(std::upper_bound(slab_end.begin(), slab_end.end(), last) - slab_end.begin());
      if (group[k] >= nslab) group[k] = nslab - 1;
      order[k] = k;
    }
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return group[a] < group[b]; });
    // If the caller's output buffer is pinned (mapped) host memory, each group's rows are written into it by one
    // small kernel (scatter_rows_to_host_kernel); otherwise they are copied back one cudaMemcpyAsync per RoI.
    float *out_mapped = nullptr;
    {
      cudaPointerAttributes attr;
      if (cudaPointerGetAttributes(&attr, out_host) == cudaSuccess && attr.type == cudaMemoryTypeHost &&
          attr.devicePointer != nullptr && roi_out % 4 == 0 && (reinterpret_cast<uintptr_t>(attr.devicePointer) % 16) == 0)
        out_mapped = static_cast<float *>(attr.devicePointer);
      else
        (void)cudaGetLastError();
    }
    // permuted RoI list + row map through the pinned scratch
    int32_t *rows_dev = reinterpret_cast<int32_t *>(d + feat_b + conv_b + rois_b + out_b);
    float *rois_perm = static_cast<float *>(g_scratch.pin);
    int32_t *rows_perm = reinterpret_cast<int32_t *>(static_cast<char *>(g_scratch.pin) + perm_b);
    for (int i = 0; i < K; ++i) {
      memcpy(rois_perm + (size_t)i * 7, rois_host + (size_t)order[i] * 7, sizeof(float) * 7);
      rows_perm[i] = order[i];
    }
    ROI3D_CUDA(cudaMemcpyAsync(rois_dev, rois_perm, sizeof(float) * 7 * (size_t)K, cudaMemcpyHostToDevice, st));
    ROI3D_CUDA(cudaMemcpyAsync(rows_dev, rows_perm, sizeof(int32_t) * (size_t)K, cudaMemcpyHostToDevice, st));
    const float *src_dev = layout == ROI3D_NCDHW ? conv_dev : feat_dev;
    int pos = 0;
    for (int j = 0; j < nslab; ++j) {
      const int z0 = j == 0 ? 0 : slab_end[j - 1], z1 = slab_end[j];
      const size_t n0 = (size_t)z0 * H * W, ns = (size_t)(z1 - z0) * H * W;  // voxel range of the slab
      if (layout == ROI3D_NCDHW) {
        // [C][slab voxels] gathered compactly, then converted into its place of the channels-last volume
        float *compact = feat_dev + (size_t)C * n0;
        ROI3D_CUDA(cudaMemcpy2DAsync(compact, ns * sizeof(float), feats_host + n0, S * sizeof(float), ns * sizeof(float),
                                     (size_t)C, cudaMemcpyHostToDevice, st));
        rc = roi3d_ncdhw_to_ndhwc(compact, conv_dev + n0 * C, 1, C, z1 - z0, H, W, st);
        if (rc) return rc;
      } else {
        ROI3D_CUDA(cudaMemcpyAsync(feat_dev + n0 * C, feats_host + n0 * C, ns * C * sizeof(float), cudaMemcpyHostToDevice, st));
      }
      int cnt = 0;
      while (pos + cnt < K && group[order[pos + cnt]] == j) ++cnt;
      if (cnt == 0) continue;
      cudaEvent_t ready = g_scratch.events[2 * j], done = g_scratch.events[2 * j + 1];
      ROI3D_CUDA(cudaEventRecord(ready, st));
      ROI3D_CUDA(cudaStreamWaitEvent(g_scratch.compute, ready, 0));
      rc = roi3d_roi_align3d_forward(src_dev, ROI3D_NDHWC, 1, C, D, H, W, rois_dev + (size_t)pos * 7, cnt, PD, PH, PW,
                                     spatial_scale, spatial_scale_depth, sample_num, out_dev + (size_t)pos * roi_out,
                                     g_scratch.compute);
      if (rc) return rc;
      ROI3D_CUDA(cudaEventRecord(done, g_scratch.compute));
      ROI3D_CUDA(cudaStreamWaitEvent(g_scratch.d2h, done, 0));
      if (out_mapped) {
        scatter_rows_to_host_kernel<<<64, 256, 0, g_scratch.d2h>>>(
            reinterpret_cast<const float4 *>(out_dev + (size_t)pos * roi_out), reinterpret_cast<float4 *>(out_mapped),
            rows_dev + pos, cnt, (long long)(roi_out / 4));
        ROI3D_LAUNCH_CHECK();
        pos += cnt;
        continue;
      }
      // rows back to their original positions; runs of consecutive original indices travel as one copy
      for (int i = 0; i < cnt;) {
        int run = 1;
        while (i + run < cnt && order[pos + i + run] == order[pos + i] + run) ++run;
        ROI3D_CUDA(cudaMemcpyAsync(out_host + (size_t)order[pos + i] * roi_out, out_dev + (size_t)(pos + i) * roi_out,
                                   sizeof(float) * roi_out * run, cudaMemcpyDeviceToHost, g_scratch.d2h));
        i += run;
      }
      pos += cnt;
    }
    ROI3D_CUDA(cudaStreamSynchronize(g_scratch.compute));
    ROI3D_CUDA(cudaStreamSynchronize(g_scratch.d2h));
    ROI3D_CUDA(cudaStreamSynchronize(st));
    return ROI3D_OK;
  }

  ROI3D_CUDA(cudaMemcpyAsync(feat_dev, feats_host, sizeof(float) * (size_t)B * C * S, cudaMemcpyHostToDevice, st));
  ROI3D_CUDA(cudaMemcpyAsync(rois_dev, rois_host, sizeof(float) * 7 * (size_t)K, cudaMemcpyHostToDevice, st));
  const float *src = feat_dev;
  if (layout == ROI3D_NCDHW) {
    rc = roi3d_ncdhw_to_ndhwc(feat_dev, conv_dev, B, C, D, H, W, st);
    if (rc) return rc;
    src = conv_dev;
  }
  rc = roi3d_roi_align3d_forward(src, ROI3D_NDHWC, B, C, D, H, W, rois_dev, K, PD, PH, PW, spatial_scale,
                                 spatial_scale_depth, sample_num, out_dev, st);
  if (rc) return rc;
  ROI3D_CUDA(cudaMemcpyAsync(out_host, out_dev, sizeof(float) * (size_t)K * roi_out, cudaMemcpyDeviceToHost, st));
  ROI3D_CUDA(cudaStreamSynchronize(st));
  return ROI3D_OK;
}

}  // extern "C"
