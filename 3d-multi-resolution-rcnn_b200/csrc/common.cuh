// Shared helpers for libroi3d_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "../../include/roi3d_b200.h"

namespace roi3d {

void set_error(const char *fmt, ...);

#define ROI3D_CHECK_ARG(cond, ...)   \
  do {                               \
    if (!(cond)) {                   \
      roi3d::set_error(__VA_ARGS__); \
      return ROI3D_EINVAL;           \
    }                                \
  } while (0)

#define ROI3D_CUDA(call)                                                                       \
  do {                                                                                         \
    cudaError_t e__ = (call);                                                                  \
    if (e__ != cudaSuccess) {                                                                  \
      roi3d::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
      return ROI3D_ECUDA;                                                                      \
    }                                                                                          \
  } while (0)

#define ROI3D_LAUNCH_CHECK()                                                                   \
  do {                                                                                         \
    cudaError_t e__ = cudaGetLastError();                                                      \
    if (e__ != cudaSuccess) {                                                                  \
      roi3d::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(e__), __FILE__, __LINE__); \
      return ROI3D_ECUDA;                                                                      \
    }                                                                                          \
  } while (0)

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline long long ceil_div_ll(long long a, long long b) { return (a + b - 1) / b; }

// ---------------------------------------------------------------------------------------------
// RoIAlign3D geometry, bit-for-bit the arithmetic of the COMPILED reference kernels
// (ROIAlignForward3D / ROIAlignBackward3D, roi_align_kernel.cu:214-291, :519-636): explicit _rn
// intrinsics pin every rounding and every contraction, independent of -fmad.  The FFMA sites were
// read from the reference's SASS (see oracle/roi3d_oracle.c header):
//   start = c1*s;  size = max(FFMA(c2+1, s, -start), 0);  bin = size / P (IEEE);
//   S = sample_num > 0 ? sample_num : (int)ceil(bin);
//   coord(p, i) = FFMA((float)p, bin, start) + ((i + .5f) * bin) / (float)S.
// ---------------------------------------------------------------------------------------------
struct Axis {
  float start, bin;
  int S;
};

__device__ __forceinline__ Axis axis_setup(float c1, float c2, float scale, int P, int sample_num) {
  Axis a;
  a.start = __fmul_rn(c1, scale);
  float size = fmaxf(__fmaf_rn(__fadd_rn(c2, 1.0f), scale, -a.start), 0.0f);
  a.bin = __fdiv_rn(size, (float)P);
  a.S = sample_num > 0 ? sample_num : (int)ceilf(a.bin);
  return a;
}

__device__ __forceinline__ float axis_coord(const Axis &a, int p, int i) {
  float base = __fmaf_rn((float)p, a.bin, a.start);
  return __fadd_rn(base, __fdiv_rn(__fmul_rn((float)i + 0.5f, a.bin), (float)a.S));
}

// One axis of bilinear_interpolate_3d (roi_align_kernel.cu:64-110): out of [-1, size] -> invalid;
// clamp <= 0 to 0; low = (int)c; low >= size-1 -> high = low = size-1, c = low; l = c - low; h = 1 - l.
struct Tap {
  int valid, low, high;
  float l, h;
};

__device__ __forceinline__ Tap axis_tap(float c, int size) {
  Tap t;
  t.valid = !(c < -1.0f || c > (float)size);
  if (c <= 0.0f) c = 0.0f;
  t.low = (int)c;
  if (t.low >= size - 1) {
    t.high = t.low = size - 1;
    c = (float)t.low;
  } else {
    t.high = t.low + 1;
  }
  t.l = __fsub_rn(c, (float)t.low);
  t.h = __fsub_rn(1.0f, t.l);
  return t;
}

// FPN level of one RoI: SingleRoIExtractor.map_roi_levels (single_level.py:73-81) as torch's CUDA
// elementwise kernels evaluate it: sqrt (IEEE), `/ finest_scale` as a multiply by the fp32 reciprocal
// (torch's CUDA div-by-python-scalar), + 1e-6, log2f, floor, clamp(0, L-1), .long().
__device__ __forceinline__ int roi_level(const float *roi, int num_levels, float inv_finest) {
  float vol = __fmul_rn(__fmul_rn(__fadd_rn(__fsub_rn(roi[3], roi[1]), 1.0f), __fadd_rn(__fsub_rn(roi[4], roi[2]), 1.0f)),
                        __fadd_rn(__fsub_rn(roi[6], roi[5]), 1.0f));
  float s = __fsqrt_rn(vol);
  float q = __fadd_rn(__fmul_rn(s, inv_finest), 1e-6f);
  float t = floorf(log2f(q));
  if (!(t > 0.0f)) t = 0.0f;
  if (t > (float)(num_levels - 1)) t = (float)(num_levels - 1);
  return (int)t;
}


// cudaFuncSetAttribute (the opt-in to more than 48 KB of dynamic shared memory) is a per-DEVICE property of a kernel:
// a process that drives several GPUs has to set it on each.  `need(bytes)` is true until `mark(bytes)` has recorded
// at least that size for the current device (two threads racing set the attribute twice: harmless).
struct PerDeviceSmemOptIn {
  std::atomic<size_t> set_bytes[64] = {};
  static int dev_index() {
    int dev = 0;
    return (cudaGetDevice(&dev) == cudaSuccess && dev >= 0 && dev < 64) ? dev : -1;
  }
  bool need(size_t bytes) const {
    const int d = dev_index();
    return d < 0 || set_bytes[d].load(std::memory_order_acquire) < bytes;
  }
  void mark(size_t bytes) {
    const int d = dev_index();
    if (d >= 0) set_bytes[d].store(bytes, std::memory_order_release);
  }
};


// SM count of the current device, looked up once per device.
inline int current_sm_count(int *out) {
  static std::atomic<int> cached[64] = {};
  int dev = 0;
  ROI3D_CUDA(cudaGetDevice(&dev));
  int n = (dev >= 0 && dev < 64) ? cached[dev].load(std::memory_order_relaxed) : 0;
  if (n == 0) {
    ROI3D_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
    if (dev >= 0 && dev < 64) cached[dev].store(n, std::memory_order_relaxed);
  }
  *out = n;
  return ROI3D_OK;
}

}  // namespace roi3d
