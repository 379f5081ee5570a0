// 3D IoU NMS for B200 (sm_100a), batched over independent segments, entirely on the device.
//
// Replaces (reference, /root/reference): nms_cuda_3d + nms_kernel_3d + devIoU3d,
//   mmdet/ops/nms/src/nms_kernel.cu:196-257, :81-129, :23-33, reached through
//   mmdet/ops/nms/nms_wrapper.py:42-44 from rpn_head_3d.py:134 and bbox_nms.py:89.
//
// The reference sorts with ATen, computes the FULL N x N/64 bit matrix with 64-thread blocks, copies it
// to the host with a blocking cudaMemcpy, sweeps it on the CPU and copies the result back.  Here:
//   1. rank kernel     -- stable descending order by counting (score desc, index asc; O(n^2) compares,
//                         embarrassingly parallel, n <= a few thousand) and gather of the boxes into
//                         sorted 32-byte records that also carry the per-box volume factors;
//   2. mask kernel     -- upper-triangular 64x64 tiles only, column boxes staged in shared memory,
//                         one uint64 word per (row, column-tile);
//   3. sweep kernel    -- one CTA per segment walks the 64-box diagonal tiles: a single thread resolves
//                         a tile from its 64 diagonal words held in shared memory, then all threads OR the
//                         kept rows into the `removed` words of the later tiles; the kept set is then
//                         compacted twice (ascending original index, and descending score).
// No host synchronisation, no D2H copy; many (volume, level, class) segments share the three launches.
//
// Bit-exactness: IoU uses explicit _rn intrinsics in exactly the operation order of the compiled
// reference (SASS: interS and Sa plain FMULs, union = FFMA(bw*bh, bd, Sa) - interS, IEEE division,
// strict '>' against the threshold).
#include "common.cuh"

namespace roi3d {

struct __align__(16) SortedBox {
  float x1, y1, x2, y2, z1, z2;
  float sxy;  // (x2-x1+1)*(y2-y1+1)
  float sz;   // (z2-z1+1)
};

__device__ __forceinline__ unsigned score_key(float s) {
  s = s + 0.0f;  // -0 -> +0 so that +-0 tie like torch's comparison does
  unsigned u = __float_as_uint(s);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// devIoU3d(a = row box, b = column box) > thr, compiled-reference arithmetic.
__device__ __forceinline__ bool iou3d_gt(const SortedBox &a, float Sa, const SortedBox &b, float thr) {
  const float left = fmaxf(a.x1, b.x1), right = fminf(a.x2, b.x2);
  const float top = fmaxf(a.y1, b.y1), bottom = fminf(a.y2, b.y2);
  const float front = fmaxf(a.z1, b.z1), back = fminf(a.z2, b.z2);
  const float w = fmaxf(__fadd_rn(__fsub_rn(right, left), 1.0f), 0.0f);
  const float h = fmaxf(__fadd_rn(__fsub_rn(bottom, top), 1.0f), 0.0f);
  const float d = fmaxf(__fadd_rn(__fsub_rn(back, front), 1.0f), 0.0f);
  const float inter = __fmul_rn(__fmul_rn(w, h), d);
  const float uni = __fsub_rn(__fmaf_rn(b.sxy, b.sz, Sa), inter);
  return __fdiv_rn(inter, uni) > thr;
}

// ------------------------------------------------------------------------------------------------
// 1. rank + gather.  grid (ceil(n_max/32), nseg), block 256 = 32 boxes x 8 slices of the j range
//    (each warp scans one slice with warp-uniform loads; partial ranks are summed through smem).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) nms3d_rank_kernel(const float *__restrict__ dets, const int32_t *seg_counts,
                                                         int n_max, SortedBox *__restrict__ sorted,
                                                         int32_t *__restrict__ order) {
  const int seg = blockIdx.y;
  const int n = seg_counts ? min(max(seg_counts[seg], 0), n_max) : n_max;
  if ((int)(blockIdx.x * 32) >= n) return;
  const float *d = dets + (long long)seg * n_max * 7;
  const int lane = threadIdx.x & 31, slice = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + lane;
  const unsigned ki = i < n ? score_key(__ldg(d + (long long)i * 7 + 6)) : 0u;
  // keys are staged through shared memory in tiles: one coalesced-ish gather with every load in flight at once
  // (the scores sit 28 bytes apart), then warp-uniform broadcast reads; slice s of each tile is scanned by warp s.
  constexpr int kTile = 2048;
  __shared__ unsigned keys[kTile];
  int rank = 0;
  for (int t0 = 0; t0 < n; t0 += kTile) {
    const int tn = min(kTile, n - t0);
    __syncthreads();
#pragma unroll
    for (int q = 0; q < kTile / 256; ++q) {
      const int j = q * 256 + threadIdx.x;
      if (j < tn) keys[j] = score_key(__ldg(d + (long long)(t0 + j) * 7 + 6));
    }
    __syncthreads();
    const int per = (tn + 7) >> 3;
    const int j0 = slice * per, j1 = min(tn, j0 + per);
#pragma unroll 8
    for (int j = j0; j < j1; ++j) {
      const unsigned kj = keys[j];
      rank += (kj > ki) || (kj == ki && (t0 + j) < i);
    }
  }
  __shared__ int part[8][32];
  part[slice][lane] = rank;
  __syncthreads();
  if (slice == 0 && i < n) {
    rank = 0;
#pragma unroll
    for (int q = 0; q < 8; ++q) rank += part[q][lane];
    const float *b = d + (long long)i * 7;
    SortedBox sb;
    sb.x1 = b[0], sb.y1 = b[1], sb.x2 = b[2], sb.y2 = b[3], sb.z1 = b[4], sb.z2 = b[5];
    sb.sxy = __fmul_rn(__fadd_rn(__fsub_rn(sb.x2, sb.x1), 1.0f), __fadd_rn(__fsub_rn(sb.y2, sb.y1), 1.0f));
    sb.sz = __fadd_rn(__fsub_rn(sb.z2, sb.z1), 1.0f);
    sorted[(long long)seg * n_max + rank] = sb;
    order[(long long)seg * n_max + rank] = i;
  }
}

// ------------------------------------------------------------------------------------------------
// 2. suppression bit matrix, upper triangle.  grid (ceil(cb/4), cb, nseg), block 256 =
//    64 rows x 4 column tiles.  mask[seg][row][cbm] (cbm = ceil(n_max/64)).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) nms3d_mask_kernel(const SortedBox *__restrict__ sorted,
                                                         const int32_t *seg_counts, int n_max, float thr,
                                                         unsigned long long *__restrict__ mask) {
  const int seg = blockIdx.z;
  const int n = seg_counts ? min(max(seg_counts[seg], 0), n_max) : n_max;
  const int cb = (n + 63) >> 6, cbm = (n_max + 63) >> 6;
  const int rb = blockIdx.y;
  const int c0 = blockIdx.x * 4;
  if (rb >= cb || c0 >= cb || c0 + 3 < rb) return;
  const SortedBox *sb = sorted + (long long)seg * n_max;
  __shared__ SortedBox cols[4][64];
  const int rl = threadIdx.x & 63, cq = threadIdx.x >> 6;
  {
    const int j = (c0 + cq) * 64 + rl;
    if (j < n) cols[cq][rl] = sb[j];
  }
  __syncthreads();
  const int cblk = c0 + cq;
  const int row = rb * 64 + rl;
  if (cblk < rb || cblk >= cb || row >= n) return;
  const SortedBox a = sb[row];
  const float Sa = __fmul_rn(a.sxy, a.sz);
  const int csize = min(64, n - cblk * 64);
  const int start = (cblk == rb) ? rl + 1 : 0;
  unsigned long long t = 0;
  for (int j = start; j < csize; ++j) {
    if (iou3d_gt(a, Sa, cols[cq][j], thr)) t |= 1ULL << j;
  }
  mask[((long long)seg * n_max + row) * cbm + cblk] = t;
}

// ------------------------------------------------------------------------------------------------
// 3. greedy sweep + compaction.  One CTA (256 threads) per segment.
// ------------------------------------------------------------------------------------------------
constexpr int kMaxColBlocks = 512;   // n_max <= 32768 (static shared memory budget)
constexpr int kPanelW = 32;           // words of each mask row held in shared memory per diagonal tile

__global__ void __launch_bounds__(256) nms3d_sweep_kernel(const unsigned long long *__restrict__ mask,
                                                          const int32_t *__restrict__ order,
                                                          const int32_t *seg_counts, int n_max,
                                                          unsigned char *__restrict__ flags,
                                                          int64_t *__restrict__ keep,
                                                          int64_t *__restrict__ keep_by_score,
                                                          int32_t *__restrict__ num_keep) {
  const int seg = blockIdx.x;
  const int n = seg_counts ? min(max(seg_counts[seg], 0), n_max) : n_max;
  const int cb = (n + 63) >> 6, cbm = (n_max + 63) >> 6;
  const unsigned long long *m = mask + (long long)seg * n_max * cbm;
  const int32_t *ord = order + (long long)seg * n_max;
  unsigned char *fl = flags + (long long)seg * n_max;
  int64_t *kp = keep + (long long)seg * n_max;
  int64_t *kps = keep_by_score ? keep_by_score + (long long)seg * n_max : nullptr;
  const int tid = threadIdx.x;

  __shared__ unsigned long long remv[kMaxColBlocks];
  __shared__ unsigned long long keptw[kMaxColBlocks];
  __shared__ unsigned long long panel[2][64][kPanelW];
  __shared__ unsigned long long s_kept;
  __shared__ int s_warp[8];
  __shared__ int s_base;

  for (int i = tid; i < cb; i += 256) remv[i] = 0ULL;
  for (int i = tid; i < n; i += 256) fl[i] = 0;

  // Row panels: for diagonal tile `blk` the words [blk, blk+kPanelW) of its 64 mask rows, copied to shared
  // memory with cp.async one tile ahead of the resolve (the rows do not depend on the sweep state).
  auto prefetch = [&](int blk, int buf) {
    if (blk < cb) {
      const int nwp = min(kPanelW, cb - blk);
      for (int idx = tid; idx < 64 * kPanelW; idx += 256) {
        const int row = idx / kPanelW, w = idx - row * kPanelW;
        const int grow = blk * 64 + row;
        unsigned long long *dst = &panel[buf][row][w];
        if (grow < n && w < nwp) {
          const unsigned sd = (unsigned)__cvta_generic_to_shared(dst);
          asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(sd), "l"(m + (long long)grow * cbm + blk + w)
                       : "memory");
        } else {
          *dst = 0ULL;
        }
      }
    }
    asm volatile("cp.async.commit_group;\n" ::: "memory");
  };
  prefetch(0, 0);
  __syncthreads();

  for (int blk = 0; blk < cb; ++blk) {
    const int buf = blk & 1;
    prefetch(blk + 1, buf ^ 1);
    asm volatile("cp.async.wait_group 1;\n" ::: "memory");
    __syncthreads();
    const int bs = min(64, n - blk * 64);
    if (tid == 0) {
      // serial resolve of the diagonal tile with all 64 diagonal words in registers (fully unrolled:
      // the bit tests use compile-time positions, the chain is test -> predicated OR)
      // 32-bit halves: the test of step i is a single LOP3 on a compile-time bit, the update a predicated OR.
      // Row i only carries bits j > i, so rows >= 32 have an empty low word.
      unsigned dlo[32], dhi[64];
#pragma unroll
      for (int i = 0; i < 64; ++i) {
        const unsigned long long d = panel[buf][i][0];
        if (i < 32) dlo[i] = (unsigned)d;
        dhi[i] = (unsigned)(d >> 32);
      }
      unsigned long long r0 = remv[blk];
      if (bs < 64) r0 |= ~0ULL << bs;  // rows past n never count as kept
      unsigned rlo = (unsigned)r0, rhi = (unsigned)(r0 >> 32), klo = 0u, khi = 0u;
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        if (!(rlo & (1u << i))) {
          klo |= 1u << i;
          rlo |= dlo[i];
          rhi |= dhi[i];
        }
      }
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        if (!(rhi & (1u << i))) {
          khi |= 1u << i;
          rhi |= dhi[32 + i];
        }
      }
      const unsigned long long kept = ((unsigned long long)khi << 32) | klo;
      s_kept = kept;
      keptw[blk] = kept;
    }
    __syncthreads();
    const unsigned long long kept = s_kept;
    const int nw = cb - blk - 1;
    if (nw > 0) {
      // words blk+1 .. blk+kPanelW-1 from the panel: thread -> (word w, 8-row group g)
      const int nwp = min(kPanelW - 1, nw);
      {
        const int w = 1 + (tid & (kPanelW - 1)), g = tid / kPanelW;  // kPanelW == 32, 256 threads -> g in [0,8)
        if (w <= nwp) {
          unsigned long long v = 0ULL;
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const int row = g * 8 + q;
            if ((kept >> row) & 1ULL) v |= panel[buf][row][w];
          }
          if (v) atomicOr(&remv[blk + w], v);
        }
      }
      // words beyond the panel (only when n > 64*kPanelW): straight from global memory
      for (int idx = tid; idx < (nw - nwp) * 64; idx += 256) {
        const int i = idx / (nw - nwp), w = blk + 1 + nwp + idx % (nw - nwp);
        if ((kept >> i) & 1ULL) {
          const unsigned long long v = m[((long long)blk * 64 + i) * cbm + w];
          if (v) atomicOr(&remv[w], v);
        }
      }
    }
    __syncthreads();
  }
  asm volatile("cp.async.wait_group 0;\n" ::: "memory");

  // ---- kept set -> (a) descending-score list, (b) flags by original index ----
  if (tid == 0) s_base = 0;
  __syncthreads();
  for (int w0 = 0; w0 < cb; w0 += 256) {
    const int w = w0 + tid;
    const unsigned long long kw = w < cb ? keptw[w] : 0ULL;
    const int cnt = __popcll(kw);
    // block exclusive scan of cnt
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if ((tid & 31) >= o) incl += t;
    }
    if ((tid & 31) == 31) s_warp[tid >> 5] = incl;
    __syncthreads();
    int woff = 0, tot = 0;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      if (q < (tid >> 5)) woff += s_warp[q];
      tot += s_warp[q];
    }
    int pos = s_base + woff + incl - cnt;
    unsigned long long bits = kw;
    while (bits) {
      const int b = __ffsll((long long)bits) - 1;
      bits &= bits - 1;
      const int orig = ord[w * 64 + b];
      if (kps) kps[pos] = orig;
      fl[orig] = 1;
      ++pos;
    }
    __syncthreads();
    if (tid == 0) s_base += tot;
    __syncthreads();
  }
  const int total = s_base;
  __syncthreads();

  // ---- ascending original index: ordered compaction of flags ----
  if (tid == 0) s_base = 0;
  __syncthreads();
  for (int i0 = 0; i0 < n; i0 += 256) {
    const int i = i0 + tid;
    const int f = (i < n) ? fl[i] : 0;
    const unsigned bal = __ballot_sync(0xffffffffu, f);
    const int lane = tid & 31;
    if (lane == 0) s_warp[tid >> 5] = __popc(bal);
    __syncthreads();
    int woff = 0, tot = 0;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      if (q < (tid >> 5)) woff += s_warp[q];
      tot += s_warp[q];
    }
    if (f) kp[s_base + woff + __popc(bal & ((1u << lane) - 1u))] = i;
    __syncthreads();
    if (tid == 0) s_base += tot;
    __syncthreads();
  }
  if (tid == 0) num_keep[seg] = total;
}

struct NmsWorkspace {
  SortedBox *sorted;
  int32_t *order;
  unsigned long long *mask;
  unsigned char *flags;
  size_t bytes;
};

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

static NmsWorkspace carve(void *base, int nseg, int n_max) {
  NmsWorkspace w;
  const size_t cbm = (size_t)(n_max + 63) / 64;
  size_t off = 0;
  char *b = static_cast<char *>(base);
  w.sorted = reinterpret_cast<SortedBox *>(b + off);
  off = align_up(off + sizeof(SortedBox) * (size_t)nseg * n_max, 256);
  w.order = reinterpret_cast<int32_t *>(b + off);
  off = align_up(off + sizeof(int32_t) * (size_t)nseg * n_max, 256);
  w.mask = reinterpret_cast<unsigned long long *>(b + off);
  off = align_up(off + sizeof(unsigned long long) * (size_t)nseg * n_max * cbm, 256);
  w.flags = reinterpret_cast<unsigned char *>(b + off);
  off = align_up(off + (size_t)nseg * n_max, 256);
  w.bytes = off;
  return w;
}

}  // namespace roi3d

using namespace roi3d;

extern "C" {

size_t roi3d_nms3d_workspace_bytes(int nseg, int n_max) {
  if (nseg <= 0 || n_max <= 0) return 256;
  return carve(nullptr, nseg, n_max).bytes;
}

int roi3d_nms3d_batched(const float *dets_dev, const int32_t *seg_counts_dev, int nseg, int n_max, float iou_thr,
                        int64_t *keep_dev, int64_t *keep_by_score_dev, int32_t *num_keep_dev, void *workspace_dev,
                        size_t workspace_bytes, void *stream) {
  ROI3D_CHECK_ARG(nseg >= 0 && n_max >= 0, "bad sizes nseg=%d n_max=%d", nseg, n_max);
  if (nseg == 0) return ROI3D_OK;
  cudaStream_t st = (cudaStream_t)stream;
  ROI3D_CHECK_ARG(num_keep_dev != nullptr, "num_keep is NULL");
  if (n_max == 0) {
    ROI3D_CUDA(cudaMemsetAsync(num_keep_dev, 0, sizeof(int32_t) * nseg, st));
    return ROI3D_OK;
  }
  ROI3D_CHECK_ARG(dets_dev && keep_dev && workspace_dev, "NULL pointer");
  ROI3D_CHECK_ARG(n_max <= kMaxColBlocks * 64, "n_max=%d exceeds %d", n_max, kMaxColBlocks * 64);
  ROI3D_CHECK_ARG(nseg <= 65535, "nseg=%d exceeds 65535", nseg);
  ROI3D_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace_dev) & 255) == 0, "workspace must be 256-byte aligned");
  NmsWorkspace w = carve(workspace_dev, nseg, n_max);
  if (workspace_bytes < w.bytes) {
    set_error("nms3d workspace too small: %zu < %zu", workspace_bytes, w.bytes);
    return ROI3D_ENOMEM;
  }
  const int cbm = (n_max + 63) / 64;
  nms3d_rank_kernel<<<dim3(ceil_div(n_max, 32), nseg), 256, 0, st>>>(dets_dev, seg_counts_dev, n_max, w.sorted,
                                                                      w.order);
  ROI3D_LAUNCH_CHECK();
  nms3d_mask_kernel<<<dim3(ceil_div(cbm, 4), cbm, nseg), 256, 0, st>>>(w.sorted, seg_counts_dev, n_max, iou_thr,
                                                                       w.mask);
  ROI3D_LAUNCH_CHECK();
  nms3d_sweep_kernel<<<nseg, 256, 0, st>>>(w.mask, w.order, seg_counts_dev, n_max, w.flags, keep_dev,
                                           keep_by_score_dev, num_keep_dev);
  ROI3D_LAUNCH_CHECK();
  return ROI3D_OK;
}

}  // extern "C"
