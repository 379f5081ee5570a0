// RoIAlign3D forward, "planar" kernel for B200 (sm_100a): reads the reference's NCDHW layout natively.
//
// Replaces (reference, /root/reference):
//   ROIAlignForward3D / bilinear_interpolate_3d   mmdet/ops/roi_align/src/roi_align_kernel.cu:214-291, :64-149
//   (its NCDHW-contiguous input contract: roi_align_cuda.cpp:35-39, :72-74)
//   SingleRoIExtractor.forward / map_roi_levels    mmdet/models/roi_extractors/single_level.py:58-104
//
// The streamed kernel (roi_align3d_stream.cu) keeps channels in the lanes, which needs channels-last memory.  In
// NCDHW a voxel's channels are D*H*W floats apart, but the x-runs of a RoI footprint are contiguous, so here the
// roles are swapped: a CTA takes (RoI, group of CG channels), copies the footprint of each channel -- (z, y) rows of
// RXB floats, 16-byte cp.async, rows start at a 16-byte boundary -- into one shared-memory PLANE per channel, and the
// lanes run over OUTPUT ELEMENTS of the three separable contractions (same per-axis tap tables as the other kernels,
// built with the compiled reference's rounding sequence, common.cuh):
//     x: T1[c][z,y][pw]  = sum_t wx[pw][t] * in[c][z,y][xo[pw] + t]
//     y: T2[c][z][ph,pw] = sum_t wy[ph][t] * T1[c][z, yo[ph] + t][pw]
//     z: out[c][pd,ph,pw] = 1/count * sum_t wz[pd][t] * T2[c][zo[pd] + t][ph,pw]
// Offsets and weights of an element are looked up once and reused for every channel of the group.  The last stage
// writes the [K, C, PD, PH, PW] output directly: consecutive lanes = consecutive floats, 128-byte stores, no staging
// -- which also makes this the kernel for the 14 x 14 x 14 mask branch, where the output (2.9 GB at C3) is 93 % of
// the traffic.  Channels-last levels are accepted too (4-byte cp.async that transposes into the planes).
// A CTA walks the 64 channels of its item in passes of up to 16 channels; the next pass's planes are copied while
// the current pass is reduced (two input buffers), and two CTAs are resident per SM.
//
// RoIs with a bin of more than four taps, or a footprint that does not fit the planes even one channel at a time,
// are evaluated literally (reference sample loops, bit-exact) by the same CTA.
#include "roi_align3d_shared.cuh"

namespace roi3d {

namespace {

constexpr int PL_THREADS = 256;
constexpr int PL_MAXP = 16;          // bins per axis the tables hold
constexpr int PL_MAXCG = 16;         // channels reduced per pass
constexpr int PL_SMEM_FLOATS = 27648;  // plane storage per CTA (108 KB): two CTAs per SM

struct PlanarTables {
  int xo[PL_MAXP], yo[PL_MAXP], zo[PL_MAXP];       // first tap of a bin, from the box origin (x: from the aligned origin)
  float xw[PL_MAXP][4], yw[PL_MAXP][4], zw[PL_MAXP][4];
  int lo[48], hi[48];                              // per (axis, bin): tap range in level coordinates
  int box[20];                                     // xa, RXB, y0, RY, z0, RZ, flags(1 = empty, 2 = slow), NTX; per-axis min / max / widest bin
};

__device__ __forceinline__ void cp_async16_pl(void *dst, const void *src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src)
               : "memory");
}
__device__ __forceinline__ void cp_async4_pl(void *dst, const void *src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src)
               : "memory");
}

// q / d for the small operands of the copy loops: magic = ceil(2^32 / d), exact while q * d < 2^32
__device__ __forceinline__ unsigned fast_magic(unsigned d) { return d <= 1 ? 0u : 0xFFFFFFFFu / d + 1u; }
__device__ __forceinline__ unsigned fast_div(unsigned q, unsigned d, unsigned magic) { return d <= 1 ? q : __umulhi(q, magic); }

// Literal evaluation of one output bin for one channel, any layout (element strides sc / sz / sy / sx): the
// reference's sample loops with its corner-weight / FFMA-chain arithmetic (roi_align_kernel.cu:134-146 + SASS).
__device__ float literal_bin_strided(const Axis &axw, const Axis &axh, const Axis &axd, int D, int H, int W,
                                     const float *fc, long long sz, long long sy, long long sx, int pd, int ph, int pw) {
  float acc = 0.0f;
  for (int iz = 0; iz < axd.S; ++iz) {
    const Tap tz = axis_tap(axis_coord(axd, pd, iz), D);
    for (int iy = 0; iy < axh.S; ++iy) {
      const Tap ty = axis_tap(axis_coord(axh, ph, iy), H);
      for (int ix = 0; ix < axw.S; ++ix) {
        const Tap tx = axis_tap(axis_coord(axw, pw, ix), W);
        if (!(tz.valid && ty.valid && tx.valid)) continue;  // contributes 0, still counted
        const float hxhy = __fmul_rn(tx.h, ty.h), lxhy = __fmul_rn(tx.l, ty.h);
        const float hxly = __fmul_rn(tx.h, ty.l), lxly = __fmul_rn(tx.l, ty.l);
        const float w1 = __fmul_rn(hxhy, tz.h), w2 = __fmul_rn(lxhy, tz.h), w3 = __fmul_rn(hxly, tz.h),
                    w4 = __fmul_rn(lxly, tz.h), w5 = __fmul_rn(hxhy, tz.l), w6 = __fmul_rn(lxhy, tz.l),
                    w7 = __fmul_rn(hxly, tz.l), w8 = __fmul_rn(lxly, tz.l);
        const long long zl = tz.low * sz, zh = tz.high * sz, yl = ty.low * sy, yh = ty.high * sy;
        const long long xl = tx.low * sx, xh = tx.high * sx;
        float t = __fmul_rn(w2, __ldg(fc + zl + yl + xh));
        t = __fmaf_rn(w1, __ldg(fc + zl + yl + xl), t);
        t = __fmaf_rn(w3, __ldg(fc + zl + yh + xl), t);
        t = __fmaf_rn(w4, __ldg(fc + zl + yh + xh), t);
        t = __fmaf_rn(w5, __ldg(fc + zh + yl + xl), t);
        t = __fmaf_rn(w6, __ldg(fc + zh + yl + xh), t);
        t = __fmaf_rn(w7, __ldg(fc + zh + yh + xl), t);
        t = __fmaf_rn(w8, __ldg(fc + zh + yh + xh), t);
        acc = __fadd_rn(acc, t);
      }
    }
  }
  return __fdiv_rn(acc, (float)(axd.S * axh.S * axw.S));
}

// One contraction stage for one output element over the channels of a pass: NT taps at s[t * TS] (TS compile-time),
// source planes src_plane floats apart, results handed to `put(c, v)`.  Four channels per trip keep twelve to sixteen
// independent shared-memory loads in flight per thread.
template <int NT, int TS, typename Put>
__device__ __forceinline__ void planar_taps(const float *s, int src_plane, int nch, const float (&w)[4], Put put) {
  int c = 0;
  for (; c + 4 <= nch; c += 4) {
    float v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float *q = s + u * src_plane;
      float a = w[0] * q[0];
      if (NT > 1) a = fmaf(w[1], q[TS], a);
      if (NT > 2) a = fmaf(w[2], q[2 * TS], a);
      if (NT > 3) a = fmaf(w[3], q[3 * TS], a);
      v[u] = a;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) put(c + u, v[u]);
    s += 4 * src_plane;
  }
  for (; c < nch; ++c) {
    float a = w[0] * s[0];
    if (NT > 1) a = fmaf(w[1], s[TS], a);
    if (NT > 2) a = fmaf(w[2], s[2 * TS], a);
    if (NT > 3) a = fmaf(w[3], s[3 * TS], a);
    put(c, a);
    s += src_plane;
  }
}

// Tap count as a template argument from a runtime value in 1..4.
template <int TS, typename Put>
__device__ __forceinline__ void planar_taps_n(int nt, const float *s, int src_plane, int nch, const float (&w)[4], Put put) {
  if (nt >= 4) planar_taps<4, TS>(s, src_plane, nch, w, put);
  else if (nt == 3) planar_taps<3, TS>(s, src_plane, nch, w, put);
  else if (nt == 2) planar_taps<2, TS>(s, src_plane, nch, w, put);
  else planar_taps<1, TS>(s, src_plane, nch, w, put);
}

// Visit order of the RoIs: by (level, volume, z, y, x) of their first corner.  In NCDHW a RoI uses 40 to 70 bytes of
// each 512-byte feature row it touches while DRAM is fetched in 128-byte lines; x-neighbours processed close in time
// find the rest of the line in L2.  Rank by counting over keys staged in shared memory (K <= 8192).
__global__ void __launch_bounds__(256) roi_align3d_order_kernel(const RoiParams p, int *order) {
  extern __shared__ unsigned long long okeys[];
  for (int k = threadIdx.x; k < p.K; k += blockDim.x) {
    float r[7];
#pragma unroll
    for (int i = 0; i < 7; ++i) r[i] = __ldg(p.rois + (long long)k * 7 + i);
    const int lvl = p.num_levels > 1 ? roi_level(r, p.num_levels, p.inv_finest) : 0;
    const float s = p.lv[lvl].scale, sd = p.lv[lvl].scale_d;
    const unsigned z = (unsigned)fminf(fmaxf(r[5] * sd, 0.0f), 1023.0f), y = (unsigned)fminf(fmaxf(r[2] * s, 0.0f), 4095.0f);
    const unsigned x = (unsigned)fminf(fmaxf(r[1] * s, 0.0f), 4095.0f);
    const unsigned b = (unsigned)fminf(fmaxf(r[0], 0.0f), 4095.0f);
    // 8 z slices x 16 rows form a cell; inside a cell RoIs are walked along x
    okeys[k] = ((unsigned long long)lvl << 56) | ((unsigned long long)b << 44) | ((unsigned long long)(z >> 3) << 36) |
               ((unsigned long long)(y >> 4) << 26) | ((unsigned long long)x << 14) | (unsigned long long)(k & 0x3fff);
  }
  __syncthreads();
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= p.K) return;
  const unsigned long long mine = okeys[k];
  int rank = 0;
  for (int j = 0; j < p.K; ++j) rank += (okeys[j] < mine) || (okeys[j] == mine && j < k);
  order[rank] = k;
}

template <int P>  // PW == PH == P
__global__ void __launch_bounds__(PL_THREADS) roi_align3d_fwd_planar_kernel(const RoiParams p, int CG, int ngroups, int ndhwc,
                                                                            const int *__restrict__ order) {
  extern __shared__ __align__(16) float planes[];
  __shared__ PlanarTables T;
  const int tid = threadIdx.x;
  const int kslot = blockIdx.x / ngroups, g = blockIdx.x - kslot * ngroups;
  const int k = order != nullptr ? __ldg(order + kslot) : kslot;
  const int c_first = g * CG;
  const int nch_all = min(CG, p.C - c_first);
  const int PD = p.PD;

  // ---- RoI geometry (every thread), tap tables (threads 0..47: axis = tid / 16, bin = tid % 16)
  float r[7];
#pragma unroll
  for (int i = 0; i < 7; ++i) r[i] = __ldg(p.rois + (long long)k * 7 + i);
  const int lvl = p.num_levels > 1 ? roi_level(r, p.num_levels, p.inv_finest) : 0;
  const LevelDev L = p.lv[lvl];
  const int b = (int)r[0];
  const bool ok = b >= 0 && b < p.B;
  const long long krow = p.out_rows != nullptr ? __ldg(p.out_rows + k) : k;
  if (p.lvls_out != nullptr && g == 0 && tid == 0) p.lvls_out[k] = lvl;
  const Axis axw = axis_setup(r[1], r[3], L.scale, P, p.sample_num);
  const Axis axh = axis_setup(r[2], r[4], L.scale, P, p.sample_num);
  const Axis axd = axis_setup(r[5], r[6], L.scale_d, PD, p.sample_num);
  const long long out_elems = (long long)PD * P * P;
  float *out_roi = p.out + (krow * p.C + c_first) * out_elems;
  {
    const int axis = tid >> 4, bin = tid & 15;
    if (tid < 48) {
      const int nb = axis == 2 ? PD : P;
      const Axis ax = axis == 0 ? axw : axis == 1 ? axh : axd;
      const int asize = axis == 0 ? L.W : axis == 1 ? L.H : L.D;
      int lo = INT_MAX, hi = -1;
      if (bin < nb && ok) {
        for (int i = 0; i < ax.S; ++i) {
          const Tap t = axis_tap(axis_coord(ax, bin, i), asize);
          if (t.valid) lo = min(lo, t.low), hi = max(hi, t.high);
        }
      }
      T.lo[tid] = lo, T.hi[tid] = hi;
    }
  }
  if (tid < 64) {
    // box of the footprint: min / max over the 16 bins of an axis (threads 0..15 x, 16..31 y, 32..47 z)
    const int lo = tid < 48 ? T.lo[tid] : INT_MAX, hi = tid < 48 ? T.hi[tid] : -1;   // own values (same thread wrote them)
    const bool has = hi >= lo;
    int mn = has ? lo : INT_MAX, mx = has ? hi : -1, wide = has ? hi - lo + 1 : 0;
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
      mn = min(mn, __shfl_xor_sync(FULL, mn, o)), mx = max(mx, __shfl_xor_sync(FULL, mx, o));
      wide = max(wide, __shfl_xor_sync(FULL, wide, o));
    }
    if ((tid & 15) == 0 && tid < 48) {
      const int axis = tid >> 4;
      T.box[8 + axis * 3 + 0] = mn, T.box[8 + axis * 3 + 1] = mx, T.box[8 + axis * 3 + 2] = wide;
    }
  }
  __syncthreads();
  if (tid == 0) {
    const int *q = T.box + 8;
    const bool empty = !ok || q[1] < q[0] || q[4] < q[3] || q[7] < q[6];
    const bool slow = q[2] > 4 || q[5] > 4 || q[8] > 4;
    const int xa = empty ? 0 : (q[0] & ~3);                          // 16-byte aligned row start
    const int RXB = empty ? 4 : max(4, ((q[1] - xa + 1) + 3) & ~3);  // floats per plane row (whole 16-byte pieces)
    T.box[0] = xa, T.box[1] = RXB;
    T.box[2] = empty ? 0 : q[3], T.box[3] = empty ? 0 : q[4] - q[3] + 1;
    T.box[4] = empty ? 0 : q[6], T.box[5] = empty ? 0 : q[7] - q[6] + 1;
    T.box[6] = (empty ? 1 : 0) | (slow && !empty ? 2 : 0);
    T.box[7] = q[2] <= 3 ? 3 : 4;   // taps per x bin the x stage reads (RXB >= 4 holds either)
  }
  __syncthreads();
  const int xa = T.box[0], RXB = T.box[1], y0 = T.box[2], RY = T.box[3], z0 = T.box[4], RZ = T.box[5];
  const int flags = T.box[6], NTX = T.box[7];
  // planes: A holds the footprint, then (after the x stage) T2; B holds T1.  +1 float per plane against bank conflicts
  const int rows = RZ * RY;
  const int szA = ((max(rows * RXB, RZ * P * P) + 3) & ~3) + 4, szB = ((rows * P + 3) & ~3) + 4;  // 16-byte aligned planes
  // two input buffers (the next pass is copied while this one is reduced) + one T1 buffer
  int CGs = (flags == 0) ? min(min(nch_all, PL_MAXCG), PL_SMEM_FLOATS / (2 * szA + szB)) : 0;
  if (CGs > 4) CGs &= ~3;  // whole groups of four channels (the stage loops walk four at a time)
  const long long vox = (long long)L.D * L.H * L.W;
  // element strides of the level: NCDHW (sc = vox, sx = 1) or channels-last (sc = 1, sx = C)
  const long long sc = ndhwc ? 1 : vox, sx = ndhwc ? p.C : 1, sy = sx * L.W, sz = sy * L.H;
  const float *fb = L.feats + (long long)(ok ? b : 0) * vox * p.C;

  if (flags & 1) {  // no sample inside the level (or batch index out of range): 0 / count, NaN for count == 0
    const float v = __fmul_rn(0.0f, __frcp_rn((float)(axd.S * axh.S * axw.S)));
    for (long long i = tid; i < (long long)nch_all * out_elems; i += PL_THREADS) __stcs(out_roi + i, v);
    return;
  }
  if (CGs == 0) {  // literal path: lanes over output elements, one channel at a time (rare)
    for (int c = 0; c < nch_all; ++c) {
      const float *fc = fb + (long long)(c_first + c) * sc;
      for (int e = tid; e < (int)out_elems; e += PL_THREADS) {
        const int pw = e % P, ph = (e / P) % P, pd = e / (P * P);
        __stcs(out_roi + c * out_elems + e, literal_bin_strided(axw, axh, axd, L.D, L.H, L.W, fc, sz, sy, sx, pd, ph, pw));
      }
    }
    return;
  }

  // ---- tap tables: the taps of a bin are NT consecutive voxels of its axis (NT = 4, or the axis extent if smaller; 3
  //      along x when no bin needs more), weights summed per voxel, first tap relative to the box origin (x: to the
  //      aligned row start) and shifted left where the NT taps would leave the box (the shifted-in weights are 0), so
  //      that every tap reads copied data
  if (tid < 48) {
    const int axis = tid >> 4, bin = tid & 15;
    const int nb = axis == 2 ? PD : P;
    if (bin < nb) {
      const Axis ax = axis == 0 ? axw : axis == 1 ? axh : axd;
      const int asize = axis == 0 ? L.W : axis == 1 ? L.H : L.D;
      const int lo = T.lo[tid], hi = T.hi[tid];
      float w0 = 0.0f, w1 = 0.0f, w2 = 0.0f, w3 = 0.0f;
      int off = 0;
      if (hi >= lo) {
        off = lo - (axis == 0 ? xa : axis == 1 ? y0 : z0);
        const int extent = axis == 0 ? RXB : axis == 1 ? RY : RZ;
        const int nt = axis == 0 ? NTX : min(4, extent);
        const int sh = max(0, off + nt - extent);
        off -= sh;
        for (int i = 0; i < ax.S; ++i) {
          const Tap t = axis_tap(axis_coord(ax, bin, i), asize);
          if (t.valid) {
            const int a0 = t.low - lo + sh, a1 = t.high - lo + sh;
            if (a0 == 0) w0 += t.h; else if (a0 == 1) w1 += t.h; else if (a0 == 2) w2 += t.h; else w3 += t.h;
            if (a1 == 0) w0 += t.l; else if (a1 == 1) w1 += t.l; else if (a1 == 2) w2 += t.l; else w3 += t.l;
          }
        }
      }
      int *po = axis == 0 ? T.xo : axis == 1 ? T.yo : T.zo;
      float(*pwt)[4] = axis == 0 ? T.xw : axis == 1 ? T.yw : T.zw;
      po[bin] = off;
      pwt[bin][0] = w0, pwt[bin][1] = w1, pwt[bin][2] = w2, pwt[bin][3] = w3;
    }
  }
  __syncthreads();
  const float inv = __frcp_rn((float)(axd.S * axh.S * axw.S));  // exact for the power-of-two counts of fixed sample_num
  float *A0 = planes, *A1 = planes + (size_t)CGs * szA, *B = planes + (size_t)2 * CGs * szA;
  const int npass = (nch_all + CGs - 1) / CGs;
  auto issue_copy = [&](int c0, float *Ad) {
    const int nch = min(CGs, nch_all - c0);
    // ---- copy the footprint planes (32-bit element offsets from the volume's first element: checked by the launcher)
    if (!ndhwc) {
      const unsigned npc = (unsigned)RXB >> 2;       // 16-byte pieces per row
      const unsigned per_ch = (unsigned)rows * npc;
      const unsigned m_npc = fast_magic(npc), m_ry = fast_magic((unsigned)RY);
      const unsigned W_ = (unsigned)L.W, HW = (unsigned)(L.H * L.W);
      const unsigned base0 = (unsigned)((z0 * L.H + y0) * L.W + xa);
      // a thread's (row, piece) walk is the same for every channel: decode once per row-piece, loop channels inside
      for (unsigned rr = tid; rr < per_ch; rr += PL_THREADS) {
        const unsigned row = fast_div(rr, npc, m_npc), pc = rr - row * npc;
        const unsigned z = fast_div(row, (unsigned)RY, m_ry), y = row - z * (unsigned)RY;
        const int x = xa + (int)pc * 4;
        const unsigned goff = base0 + z * HW + y * W_ + pc * 4;
        float *dst = Ad + row * RXB + pc * 4;
        const float *src = fb + (size_t)(c_first + c0) * (size_t)vox + goff;
        if (x + 4 <= L.W) {
          for (int c = 0; c < nch; ++c) cp_async16_pl(dst + c * szA, src + (size_t)c * (size_t)vox);
        } else {
          for (int c = 0; c < nch; ++c)
#pragma unroll
            for (int j = 0; j < 4; ++j) dst[c * szA + j] = (x + j < L.W) ? __ldg(src + (size_t)c * (size_t)vox + j) : 0.0f;
        }
      }
    } else {
      // channels-last: the nch channels of a voxel are contiguous; thread -> (voxel, channel), channels fastest
      const unsigned per_vox = (unsigned)(rows * RXB);
      const unsigned m_nch = fast_magic((unsigned)nch), m_rxb = fast_magic((unsigned)RXB), m_ry = fast_magic((unsigned)RY);
      const unsigned C_ = (unsigned)p.C, WC = (unsigned)L.W * C_, HWC = (unsigned)L.H * WC;
      const float *src0 = fb + (size_t)((z0 * L.H + y0) * L.W + xa) * C_ + (c_first + c0);
      for (unsigned q = tid; q < (unsigned)nch * per_vox; q += PL_THREADS) {
        const unsigned v = fast_div(q, (unsigned)nch, m_nch), c = q - v * (unsigned)nch;
        const unsigned row = fast_div(v, (unsigned)RXB, m_rxb), xx = v - row * (unsigned)RXB;
        const unsigned z = fast_div(row, (unsigned)RY, m_ry), y = row - z * (unsigned)RY;
        float *dst = Ad + c * szA + v;
        if (xa + (int)xx < L.W)
          cp_async4_pl(dst, src0 + (z * HWC + y * WC + xx * C_ + c));
        else
          *dst = 0.0f;
      }
    }
    asm volatile("cp.async.commit_group;\n" ::: "memory");
  };
  issue_copy(0, A0);
  for (int ip = 0; ip < npass; ++ip) {
    const int c0 = ip * CGs;
    const int nch = min(CGs, nch_all - c0);
    float *A = (ip & 1) ? A1 : A0;
    if (ip + 1 < npass) {
      issue_copy(c0 + CGs, (ip & 1) ? A0 : A1);
      asm volatile("cp.async.wait_group 1;\n" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;\n" ::: "memory");
    }
    __syncthreads();
    // ---- x stage: A[c][row][RXB] -> B[c][row][P]
    for (int e = tid; e < rows * P; e += PL_THREADS) {
      const int row = e / P, pw = e - row * P;
      const float4 w4 = *reinterpret_cast<const float4 *>(&T.xw[pw][0]);
      const float w[4] = {w4.x, w4.y, w4.z, w4.w};
      float *d = B + e;
      auto put = [&](int c, float v) { d[c * szB] = v; };
      planar_taps_n<1>(NTX, A + row * RXB + T.xo[pw], szA, nch, w, put);
    }
    __syncthreads();
    // ---- y stage: B[c][z][y][P] -> A[c][z][ph][P]
    for (int e = tid; e < RZ * P * P; e += PL_THREADS) {
      const int pw = e % P, t1 = e / P;
      const int ph = t1 % P, z = t1 / P;
      const float4 w4 = *reinterpret_cast<const float4 *>(&T.yw[ph][0]);
      const float w[4] = {w4.x, w4.y, w4.z, w4.w};
      float *d = A + e;
      auto put = [&](int c, float v) { d[c * szA] = v; };
      planar_taps_n<P>(min(RY, 4), B + (z * RY + T.yo[ph]) * P + pw, szB, nch, w, put);
    }
    __syncthreads();
    // ---- z stage: A[c][z][P*P] -> out[c][pd][P*P], scaled by 1 / count, streamed to global
    {
      constexpr int PP = P * P;
      float *oc = out_roi + (long long)c0 * out_elems;
      for (int e = tid; e < PD * PP; e += PL_THREADS) {
        const int pd = e / PP, q = e - pd * PP;
        const float4 w4 = *reinterpret_cast<const float4 *>(&T.zw[pd][0]);
        const float w[4] = {w4.x * inv, w4.y * inv, w4.z * inv, w4.w * inv};
        float *d = oc + e;
        auto put = [&](int c, float v) { __stcs(d + c * out_elems, v); };
        planar_taps_n<PP>(min(RZ, 4), A + T.zo[pd] * PP + q, szA, nch, w, put);
      }
    }
    __syncthreads();
  }
}

}  // namespace

bool fwd_planar_ok(const RoiParams &p, int layout) {
  if (p.PW != p.PH || (p.PW != 7 && p.PW != 14) || p.PD < 1 || p.PD > PL_MAXP) return false;
  if ((long long)p.K * ((p.C + 3) / 4) >= 2147483647LL) return false;
  for (int l = 0; l < p.num_levels; ++l)
    if ((long long)p.lv[l].D * p.lv[l].H * p.lv[l].W * p.C >= 2147483647LL) return false;  // 32-bit offsets inside a volume
  for (int l = 0; l < p.num_levels; ++l) {
    if (layout == ROI3D_NCDHW) {
      // 16-byte row pieces: rows must start on 16-byte boundaries
      if ((reinterpret_cast<uintptr_t>(p.lv[l].feats) & 15) != 0 || p.lv[l].W % 4 != 0) return false;
    }
  }
  return true;
}

int launch_fwd_planar(RoiParams &p, int layout, cudaStream_t st) {
  // a CTA takes 64 channels of a RoI (its prologue -- RoI geometry, tap tables -- is paid once for them) and walks them
  // in passes of up to 16 channels, as many as the planes of the RoI's footprint leave room for
  int CG = 64;
  if (p.C < CG) CG = p.C;
  const int ngroups = ceil_div(p.C, CG);
  const size_t smem = (size_t)PL_SMEM_FLOATS * sizeof(float);
  const long long blocks = (long long)p.K * ngroups;
  ROI3D_CHECK_ARG(blocks < 2147483647LL, "roi_align3d forward: too many work items");
  const int ndhwc = layout == ROI3D_NDHWC ? 1 : 0;
  int *order = nullptr;
  if (p.K > 1 && p.K <= 8192) {
    cudaMemPool_t pool;
    const int rc = stream_pool(&pool);
    if (rc) return rc;
    ROI3D_CUDA(cudaMallocFromPoolAsync(reinterpret_cast<void **>(&order), (size_t)p.K * sizeof(int), pool, st));
    roi_align3d_order_kernel<<<ceil_div(p.K, 256), 256, (size_t)p.K * sizeof(unsigned long long), st>>>(p, order);
    ROI3D_LAUNCH_CHECK();
  }
  if (p.PW == 7) {
    static bool set7 = false;
    if (!set7) {
      ROI3D_CUDA(cudaFuncSetAttribute(roi_align3d_fwd_planar_kernel<7>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      set7 = true;
    }
    roi_align3d_fwd_planar_kernel<7><<<(unsigned)blocks, PL_THREADS, smem, st>>>(p, CG, ngroups, ndhwc, order);
  } else {
    static bool set14 = false;
    if (!set14) {
      ROI3D_CUDA(cudaFuncSetAttribute(roi_align3d_fwd_planar_kernel<14>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      set14 = true;
    }
    roi_align3d_fwd_planar_kernel<14><<<(unsigned)blocks, PL_THREADS, smem, st>>>(p, CG, ngroups, ndhwc, order);
  }
  ROI3D_LAUNCH_CHECK();
  if (order != nullptr) ROI3D_CUDA(cudaFreeAsync(order, st));
  return ROI3D_OK;
}

}  // namespace roi3d
