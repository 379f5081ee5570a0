// Library plumbing (error text, device info) and the host-buffer entry points: the same calls the
// reference's Python wrappers make with numpy / CPU tensors (mmdet/ops/nms/nms_wrapper.py:29-32),
// with the H2D / D2H copies inside the call.  These are what bench.py's `e2e` leg times.
#include <stdarg.h>
#include <string.h>

#include <mutex>

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
  const size_t feat_b = up256(sizeof(float) * (size_t)B * C * D * H * W);
  const size_t rois_b = up256(sizeof(float) * 7 * (size_t)K);
  const size_t out_b = up256(sizeof(float) * (size_t)K * C * PD * PH * PW);
  const size_t conv_b = layout == ROI3D_NCDHW ? feat_b : 0;
  int rc = ensure(g_scratch, feat_b + conv_b + rois_b + out_b, 0);
  if (rc) return rc;
  char *d = static_cast<char *>(g_scratch.dev);
  float *feat_dev = reinterpret_cast<float *>(d);
  float *conv_dev = reinterpret_cast<float *>(d + feat_b);
  float *rois_dev = reinterpret_cast<float *>(d + feat_b + conv_b);
  float *out_dev = reinterpret_cast<float *>(d + feat_b + conv_b + rois_b);
  cudaStream_t st = g_scratch.stream;
  ROI3D_CUDA(cudaMemcpyAsync(feat_dev, feats_host, sizeof(float) * (size_t)B * C * D * H * W, cudaMemcpyHostToDevice, st));
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
  ROI3D_CUDA(cudaMemcpyAsync(out_host, out_dev, sizeof(float) * (size_t)K * C * PD * PH * PW, cudaMemcpyDeviceToHost, st));
  ROI3D_CUDA(cudaStreamSynchronize(st));
  return ROI3D_OK;
}

}  // extern "C"
