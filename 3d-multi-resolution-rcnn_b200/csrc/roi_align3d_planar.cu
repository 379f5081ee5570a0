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
// A CTA walks the 64 channels of its item in passes of up to 16 channels (packed four by four: LDS.128 + FFMA2 in the
// y / z stages) and, where the footprint does not fit, in chunks of z slices; the next step's planes are copied while
// the current one is reduced (two input buffers); three CTAs are resident per SM (64 registers, 72 KB).
//
// RoIs with a bin of more than four taps, or a footprint that does not fit the planes even one channel at a time,
// are evaluated literally (reference sample loops, bit-exact) by the same CTA.
#include "roi_align3d_shared.cuh"

namespace roi3d {

namespace {

constexpr int PL_THREADS = 256;
constexpr int PL_MAXP = 16;          // bins per axis the tables hold
constexpr int PL_MAXCG = 16;         // channels reduced per pass
constexpr int PL_SMEM_FLOATS = 18432;  // plane storage per CTA (72 KB): three CTAs per SM (measured: 2 x 108 KB 782 us, 3 x 72 KB 733 us, 4 x 54 KB 731 us on C3)

struct PlanarTables {
  int xo[PL_MAXP], yo[PL_MAXP], zo[PL_MAXP];       // first tap of a bin, from the box origin (x: from the aligned origin)
  int nz[PL_MAXP];                                 // slices of z bin pd
  alignas(16) float xw[PL_MAXP][4];
  alignas(16) float yw[PL_MAXP][4];
  alignas(16) float zw[PL_MAXP][4];                // pre-multiplied by 1 / count
  int lo[48], hi[48];                              // per (axis, bin): tap range in level coordinates
  int box[20];                                     // xa, RXB, y0, RY, z0, RZ, flags(1 = empty, 2 = slow), NTX; per-axis min / max / widest bin; CGs, ZC, NTY
};

__device__ __forceinline__ void cp_async16_pl(void *dst, const void *src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src)
               : "memory");
}

// q / d for the small operands of the copy loops: magic = ceil(2^32 / d), exact while q * d < 2^32
__device__ __forceinline__ unsigned fast_magic(unsigned d) { return d <= 1 ? 0u : 0xFFFFFFFFu / d + 1u; }
__device__ __forceinline__ unsigned fast_div(unsigned q, unsigned d, unsigned magic) { return d <= 1 ? q : __umulhi(q, magic); }

// acc += w * v for the four channels of a packed entry: two packed-fp32 FFMA2, each half rounded like fmaf
__device__ __forceinline__ void fma4(float4 &a, float w, const float4 v) {
  const float2 w2 = make_float2(w, w);
  const float2 lo = __ffma2_rn(w2, make_float2(v.x, v.y), make_float2(a.x, a.y));
  const float2 hi = __ffma2_rn(w2, make_float2(v.z, v.w), make_float2(a.z, a.w));
  a = make_float4(lo.x, lo.y, hi.x, hi.y);
}
__device__ __forceinline__ float4 mul4(float w, const float4 v) {
  const float2 w2 = make_float2(w, w);
  const float2 lo = __fmul2_rn(w2, make_float2(v.x, v.y)), hi = __fmul2_rn(w2, make_float2(v.z, v.w));
  return make_float4(lo.x, lo.y, hi.x, hi.y);
}
// NT taps, STRIDE floats apart, of packed (four-channel) entries: one LDS.128 + two FFMA2 per tap
template <int NT, int STRIDE>
__device__ __forceinline__ float4 taps4(const float *s, const float4 w) {
  float4 a = mul4(w.x, *reinterpret_cast<const float4 *>(s));
  if (NT > 1) fma4(a, w.y, *reinterpret_cast<const float4 *>(s + STRIDE));
  if (NT > 2) fma4(a, w.z, *reinterpret_cast<const float4 *>(s + 2 * STRIDE));
  if (NT > 3) fma4(a, w.w, *reinterpret_cast<const float4 *>(s + 3 * STRIDE));
  return a;
}
// the z stage: bins of one RoI differ in how many slices they touch.  The tables widen every bin to two slices where
// the footprint has two (TWO: uniform over the CTA, a template argument so that the common case has no predicated
// merges); a third / fourth tap is rare.  Accumulators are two packed pairs updated in place.
struct Acc4 {
  float2 lo, hi;
};
template <int STRIDE, bool TWO>
__device__ __forceinline__ Acc4 taps4_n(int n, const float *s, const float4 w) {
  Acc4 a;
  {
    const float4 v = *reinterpret_cast<const float4 *>(s);
    a.lo = __fmul2_rn(make_float2(w.x, w.x), make_float2(v.x, v.y));
    a.hi = __fmul2_rn(make_float2(w.x, w.x), make_float2(v.z, v.w));
  }
  if (TWO) {
    const float4 v = *reinterpret_cast<const float4 *>(s + STRIDE);
    a.lo = __ffma2_rn(make_float2(w.y, w.y), make_float2(v.x, v.y), a.lo);
    a.hi = __ffma2_rn(make_float2(w.y, w.y), make_float2(v.z, v.w), a.hi);
  }
  if (n > 2) {
    const float4 v = *reinterpret_cast<const float4 *>(s + 2 * STRIDE);
    a.lo = __ffma2_rn(make_float2(w.z, w.z), make_float2(v.x, v.y), a.lo);
    a.hi = __ffma2_rn(make_float2(w.z, w.z), make_float2(v.z, v.w), a.hi);
    if (n > 3) {
      const float4 u = *reinterpret_cast<const float4 *>(s + 3 * STRIDE);
      a.lo = __ffma2_rn(make_float2(w.w, w.w), make_float2(u.x, u.y), a.lo);
      a.hi = __ffma2_rn(make_float2(w.w, w.w), make_float2(u.z, u.w), a.hi);
    }
  }
  return a;
}
// four channel planes of one output element, OE floats apart (compile-time when the output depth is)
template <bool FULL4>
__device__ __forceinline__ void store4(float *d, long long oe, const Acc4 a, int left) {
  __stcs(d, a.lo.x);
  if (FULL4 || left > 1) __stcs(d + oe, a.lo.y);
  if (FULL4 || left > 2) __stcs(d + 2 * oe, a.hi.x);
  if (FULL4 || left > 3) __stcs(d + 3 * oe, a.hi.y);
}

// ---- x stage: footprint rows -> T1[group][row][pw][4 channels].  NCDHW: the footprint lies in one plane per channel
//      (rows of RXB floats), four scalar tap rows per group; channels-last: packed entries, one LDS.128 per tap.
// Wide outputs (P = 14) have enough (row, pw) tasks per step to fill the CTA: a thread decodes its task once and walks
// the channel groups; narrow ones spread (group, row, pw) over the threads.
template <int P, bool CL, int NT>
__device__ __forceinline__ float4 planar_x_one(const float *raw, const float4 w, int cg, int row, int xo, int rowsMax, int RXB,
                                               int szR) {
  if constexpr (CL) {
    return taps4<NT, 4>(raw + ((cg * rowsMax + row) * RXB + xo) * 4, w);
  } else {
    const float *s = raw + (4 * cg) * szR + row * RXB + xo;
    float v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float *q = s + u * szR;
      float x = __fmul_rn(w.x, q[0]);
      if (NT > 1) x = __fmaf_rn(w.y, q[1], x);
      if (NT > 2) x = __fmaf_rn(w.z, q[2], x);
      if (NT > 3) x = __fmaf_rn(w.w, q[3], x);
      v[u] = x;
    }
    return make_float4(v[0], v[1], v[2], v[3]);
  }
}

template <int P, bool CL, int NT>
__device__ __forceinline__ void planar_x_stage(const PlanarTables &T, const float *raw, float *T1, int rowsC, int rowsMax,
                                               int RXB, int szR, int NG, int tid) {
  const unsigned ntask = (unsigned)(rowsC * P);
  if constexpr (P >= 14) {
    for (unsigned t = tid; t < ntask; t += PL_THREADS) {
      const unsigned row = t / P, pw = t - row * P;
      const float4 w = *reinterpret_cast<const float4 *>(&T.xw[pw][0]);
      const int xo = T.xo[pw];
      float *d = T1 + (row * P + pw) * 4;
      for (int cg = 0; cg < NG; ++cg) {
        *reinterpret_cast<float4 *>(d) = planar_x_one<P, CL, NT>(raw, w, cg, row, xo, rowsMax, RXB, szR);
        d += rowsMax * P * 4;
      }
    }
  } else {
    const unsigned total = ntask * (unsigned)NG, m_nt = fast_magic(ntask);
    for (unsigned e = tid; e < total; e += PL_THREADS) {
      const unsigned cg = fast_div(e, ntask, m_nt), t = e - cg * ntask;
      const unsigned row = t / P, pw = t - row * P;
      const float4 w = *reinterpret_cast<const float4 *>(&T.xw[pw][0]);
      *reinterpret_cast<float4 *>(T1 + ((cg * rowsMax + row) * P + pw) * 4) =
          planar_x_one<P, CL, NT>(raw, w, cg, row, T.xo[pw], rowsMax, RXB, szR);
    }
  }
}

// ---- y stage: T1[group][zr * RY + y][pw][4] -> T2[group][z][ph * P + pw][4]
template <int P, int NT>
__device__ __forceinline__ void planar_y_stage(const PlanarTables &T, const float *T1, float *T2, int zb, int zc, int rowsMax,
                                               int RY, int RZ, int NG, int tid) {
  constexpr int PP = P * P;
  const unsigned ntask = (unsigned)(zc * PP);
  if constexpr (P >= 14) {
    for (unsigned t = tid; t < ntask; t += PL_THREADS) {
      const unsigned zr = t / PP, q = t - zr * PP;
      const unsigned ph = q / P, pw = q - ph * P;
      const float4 w = *reinterpret_cast<const float4 *>(&T.yw[ph][0]);
      const float *sp = T1 + ((zr * RY + T.yo[ph]) * P + pw) * 4;
      float *d = T2 + ((zb + zr) * PP + q) * 4;
      for (int cg = 0; cg < NG; ++cg) {
        *reinterpret_cast<float4 *>(d) = taps4<NT, P * 4>(sp, w);
        sp += rowsMax * P * 4, d += RZ * PP * 4;
      }
    }
  } else {
    const unsigned total = ntask * (unsigned)NG, m_nt = fast_magic(ntask);
    for (unsigned e = tid; e < total; e += PL_THREADS) {
      const unsigned cg = fast_div(e, ntask, m_nt), t = e - cg * ntask;
      const unsigned zr = t / PP, q = t - zr * PP;
      const unsigned ph = q / P, pw = q - ph * P;
      const float4 w = *reinterpret_cast<const float4 *>(&T.yw[ph][0]);
      const float4 a = taps4<NT, P * 4>(T1 + ((cg * rowsMax + zr * RY + T.yo[ph]) * P + pw) * 4, w);
      *reinterpret_cast<float4 *>(T2 + ((cg * RZ + zb + zr) * PP + q) * 4) = a;
    }
  }
}

// ---- z stage: T2[group][z][q][4] -> out[c][pd][q]
template <int P, bool TWO>
__device__ __forceinline__ void planar_z_stage(const PlanarTables &T, const float *T2, float *oc, long long out_elems, int PD,
                                               int RZ, int NG, int nch, int tid) {
  constexpr int PP = P * P;
  const int gT2 = RZ * PP * 4;
  const bool full4 = (nch & 3) == 0;
  if constexpr (PP >= PL_THREADS / 2) {
    // wide outputs: a thread keeps its element and walks the channel groups (offsets / weights looked up once)
    for (int e = tid; e < PD * PP; e += PL_THREADS) {
      const int pd = e / PP, q = e - pd * PP;
      const int n = T.nz[pd];
      const float4 w = *reinterpret_cast<const float4 *>(&T.zw[pd][0]);
      const float *sp = T2 + (T.zo[pd] * PP + q) * 4;
      float *d = oc + e;
      if (full4) {
        for (int cg = 0; cg < NG; ++cg) {
          store4<true>(d, out_elems, taps4_n<PP * 4, TWO>(n, sp, w), 4);
          sp += gT2, d += 4 * out_elems;
        }
      } else {
        for (int cg = 0; cg < NG; ++cg) {
          store4<false>(d, out_elems, taps4_n<PP * 4, TWO>(n, sp, w), nch - cg * 4);
          sp += gT2, d += 4 * out_elems;
        }
      }
    }
  } else {
    const unsigned ntask = (unsigned)(PD * PP), total = ntask * (unsigned)NG, m_nt = fast_magic(ntask);
    for (unsigned e = tid; e < total; e += PL_THREADS) {
      const unsigned cg = fast_div(e, ntask, m_nt), t = e - cg * ntask;
      const unsigned pd = t / PP, q = t - pd * PP;
      const float4 w = *reinterpret_cast<const float4 *>(&T.zw[pd][0]);
      const Acc4 a = taps4_n<PP * 4, TWO>(T.nz[pd], T2 + ((cg * RZ + T.zo[pd]) * PP + q) * 4, w);
      float *d = oc + (long long)(cg * 4) * out_elems + t;
      if (full4) store4<true>(d, out_elems, a, 4);
      else store4<false>(d, out_elems, a, nch - (int)cg * 4);
    }
  }
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

template <int P, bool CL, int PDT>  // PW == PH == P; CL: channels-last levels; PDT: output depth (0 = p.PD at run time)
__global__ void __launch_bounds__(PL_THREADS, 4)
    roi_align3d_fwd_planar_kernel(const RoiParams p, int CG, int ngroups, int smem_floats, const int *__restrict__ order) {
  extern __shared__ __align__(16) float planes[];
  __shared__ PlanarTables T;
  constexpr int PP = P * P;
  const int tid = threadIdx.x;
  // channel-group-major: the RoIs of one 64-channel group run together, so the part of the level they share (a quarter
  // of a 256-channel level: L2-sized) is fetched from HBM once
  const int g = blockIdx.x / p.K, kslot = blockIdx.x - g * p.K;
  const int k = order != nullptr ? __ldg(order + kslot) : kslot;
  const int c_first = g * CG;
  const int nch_all = min(CG, p.C - c_first);
  const int PD = PDT > 0 ? PDT : p.PD;

  // ---- RoI geometry (every thread), tap ranges (threads 0..47: axis = tid / 16, bin = tid % 16)
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
  const long long out_elems = PDT > 0 ? (long long)PDT * PP : (long long)PD * PP;   // a constant for PDT > 0: immediates
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
    // NCDHW rows are copied in whole 16-byte pieces from a 16-byte aligned start; channels-last rows voxel by voxel
    const int xa = empty ? 0 : (CL ? q[0] : (q[0] & ~3));
    const int RXB = empty ? 4 : (CL ? q[1] - q[0] + 1 : max(4, ((q[1] - xa + 1) + 3) & ~3));
    const int RY = empty ? 0 : q[4] - q[3] + 1, RZ = empty ? 0 : q[7] - q[6] + 1;
    // Shared-memory plan: CGs channels per pass (whole groups of four) and ZC slices per step.  T2 holds every slice of
    // the pass; the footprint (two buffers: the next step is copied while this one is reduced) and T1 hold one step.
    // Fewest steps wins, ties go to more channels per pass.
    int bestCG = 0, bestZC = 0, bestSteps = INT_MAX;
    if (!empty && !slow) {
      const int ncap = min((nch_all + 3) & ~3, PL_MAXCG);
      for (int cg = ncap; cg >= 4; cg -= 4) {
        const int rem = smem_floats - cg * RZ * PP;
        const int per = cg * (2 * RY * RXB + RY * P);
        const int zmax = rem > 0 ? min(RZ, rem / per) : 0;
        if (zmax < 1) continue;
        const int nchunk = (RZ + zmax - 1) / zmax;
        const int steps = ((nch_all + cg - 1) / cg) * nchunk;
        if (steps < bestSteps) bestSteps = steps, bestCG = cg, bestZC = (RZ + nchunk - 1) / nchunk;
      }
    }
    T.box[0] = xa, T.box[1] = RXB;
    T.box[2] = empty ? 0 : q[3], T.box[3] = RY;
    T.box[4] = empty ? 0 : q[6], T.box[5] = RZ;
    T.box[6] = (empty ? 1 : 0) | (slow && !empty ? 2 : 0);
    T.box[7] = max(1, q[2]);    // taps per x bin (widest bin)
    T.box[17] = bestCG, T.box[18] = bestZC;
    T.box[19] = max(1, q[5]);   // taps per y bin
  }
  __syncthreads();
  const int xa = T.box[0], RXB = T.box[1], y0 = T.box[2], RY = T.box[3], z0 = T.box[4], RZ = T.box[5];
  const int flags = T.box[6], NTX = T.box[7], CGs = T.box[17], ZC = T.box[18], NTY = T.box[19];
  const long long vox = (long long)L.D * L.H * L.W;
  const float *fb = L.feats + (long long)(ok ? b : 0) * vox * p.C;

  if (flags & 1) {  // no sample inside the level (or batch index out of range): 0 / count, NaN for count == 0
    const float v = __fmul_rn(0.0f, __frcp_rn((float)(axd.S * axh.S * axw.S)));
    for (long long i = tid; i < (long long)nch_all * out_elems; i += PL_THREADS) __stcs(out_roi + i, v);
    return;
  }
  if (CGs == 0) {  // literal path: lanes over output elements, one channel at a time (rare)
    const long long sc = CL ? 1 : vox, sx = CL ? p.C : 1, sy = sx * L.W, sz = sy * L.H;
    for (int c = 0; c < nch_all; ++c) {
      const float *fc = fb + (long long)(c_first + c) * sc;
      for (int e = tid; e < (int)out_elems; e += PL_THREADS) {
        const int pw = e % P, ph = (e / P) % P, pd = e / PP;
        __stcs(out_roi + c * out_elems + e, literal_bin_strided(axw, axh, axd, L.D, L.H, L.W, fc, sz, sy, sx, pd, ph, pw));
      }
    }
    return;
  }

  // ---- tap tables.  x / y: every bin reads NTX / NTY consecutive voxels (the widest bin of the axis), first tap
  //      relative to the box origin and shifted left where the taps would leave the box (the shifted-in weights are 0),
  //      so that every tap reads copied data.  z: exactly the slices of the bin, weights pre-multiplied by 1 / count
  //      (exact for the power-of-two counts of fixed sample_num).
  if (tid < 48) {
    const int axis = tid >> 4, bin = tid & 15;
    const int nb = axis == 2 ? PD : P;
    if (bin < nb) {
      const Axis ax = axis == 0 ? axw : axis == 1 ? axh : axd;
      const int asize = axis == 0 ? L.W : axis == 1 ? L.H : L.D;
      const int lo = T.lo[tid], hi = T.hi[tid];
      float w0 = 0.0f, w1 = 0.0f, w2 = 0.0f, w3 = 0.0f;
      int off = 0, n = 0;
      if (hi >= lo) {
        n = hi - lo + 1;
        off = lo - (axis == 0 ? xa : axis == 1 ? y0 : z0);
        int sh = 0;
        if (axis < 2) {
          const int extent = axis == 0 ? RXB : RY, nt = axis == 0 ? NTX : NTY;
          sh = max(0, off + nt - extent);
          off -= sh;
        }
        for (int i = 0; i < ax.S; ++i) {
          const Tap t = axis_tap(axis_coord(ax, bin, i), asize);
          if (t.valid) {
            const int a0 = t.low - lo + sh, a1 = t.high - lo + sh;
            if (a0 == 0) w0 += t.h; else if (a0 == 1) w1 += t.h; else if (a0 == 2) w2 += t.h; else w3 += t.h;
            if (a1 == 0) w0 += t.l; else if (a1 == 1) w1 += t.l; else if (a1 == 2) w2 += t.l; else w3 += t.l;
          }
        }
      }
      if (axis == 2) {
        const float inv = __frcp_rn((float)(axd.S * axh.S * axw.S));
        w0 *= inv, w1 *= inv, w2 *= inv, w3 *= inv;
        if (n < 2 && RZ >= 2) {  // at least two slices per bin (weight 0 on the added one): see taps4_n
          if (off + 1 >= RZ) off -= 1, w1 = w0, w0 = 0.0f;
          n = 2;
        }
        T.nz[bin] = max(n, 1);   // a bin without a valid sample reads slice 0 with weight 0
      }
      int *po = axis == 0 ? T.xo : axis == 1 ? T.yo : T.zo;
      float(*pwt)[4] = axis == 0 ? T.xw : axis == 1 ? T.yw : T.zw;
      po[bin] = off;
      pwt[bin][0] = w0, pwt[bin][1] = w1, pwt[bin][2] = w2, pwt[bin][3] = w3;
    }
  }
  // (visible to all threads after the first barrier of the step loop)

  const int rowsMax = ZC * RY;             // rows of a step's footprint / T1 planes
  const int szR = rowsMax * RXB;           // NCDHW: floats per channel plane of a step
  const int rawsz = CGs * szR;             // one footprint buffer (either layout)
  float *raw0 = planes, *raw1 = planes + rawsz, *T1 = planes + 2 * rawsz, *T2 = T1 + CGs * rowsMax * P;
  const int npass = (nch_all + CGs - 1) / CGs;
  const int nchunk = (RZ + ZC - 1) / ZC;
  const int nsteps = npass * nchunk;

  auto issue_copy = [&](int s, float *dstbuf) {
    const int pass = s / nchunk, chunk = s - pass * nchunk;
    const int c0 = pass * CGs, nch = min(CGs, nch_all - c0);
    const int zb = chunk * ZC, zc = min(ZC, RZ - zb);
    const unsigned rows = (unsigned)(zc * RY);
    const unsigned m_ry = fast_magic((unsigned)RY);
    // 32-bit element offsets from the volume's first element: checked by the launcher
    if constexpr (!CL) {
      const unsigned npc = (unsigned)RXB >> 2;       // 16-byte pieces per row; every piece lies inside the row (W % 4 == 0)
      const unsigned per_ch = rows * npc;
      const unsigned m_npc = fast_magic(npc);
      const unsigned W_ = (unsigned)L.W, HW = (unsigned)(L.H * L.W);
      const unsigned base0 = (unsigned)(((z0 + zb) * L.H + y0) * L.W + xa);
      const float *src0 = fb + (size_t)(c_first + c0) * (size_t)vox;
      // a thread's (row, piece) walk is the same for every channel: decode once per row-piece, loop channels inside
      for (unsigned rr = tid; rr < per_ch; rr += PL_THREADS) {
        const unsigned row = fast_div(rr, npc, m_npc), pc = rr - row * npc;
        const unsigned z = fast_div(row, (unsigned)RY, m_ry), y = row - z * (unsigned)RY;
        unsigned dst = (unsigned)__cvta_generic_to_shared(dstbuf + row * RXB + pc * 4);
        const float *src = src0 + (base0 + z * HW + y * W_ + pc * 4);
        for (int c = 0; c < nch; ++c) {
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(src) : "memory");
          dst += (unsigned)szR * 4u, src += vox;
        }
      }
    } else {
      // channels-last: four channels of a voxel per 16-byte copy, straight into the packed layout [group][row][x][4]
      const unsigned NG = (unsigned)(nch >> 2);
      const unsigned total = rows * (unsigned)RXB * NG;
      const unsigned m_ng = fast_magic(NG), m_rxb = fast_magic((unsigned)RXB);
      const unsigned C_ = (unsigned)p.C, WC = (unsigned)L.W * C_, HWC = (unsigned)L.H * WC;
      const float *src0 = fb + (size_t)(((z0 + zb) * L.H + y0) * L.W + xa) * C_ + (c_first + c0);
      for (unsigned q = tid; q < total; q += PL_THREADS) {
        const unsigned v = fast_div(q, NG, m_ng), cg = q - v * NG;
        const unsigned row = fast_div(v, (unsigned)RXB, m_rxb), xx = v - row * (unsigned)RXB;
        const unsigned z = fast_div(row, (unsigned)RY, m_ry), y = row - z * (unsigned)RY;
        cp_async16_pl(dstbuf + ((cg * rowsMax + row) * RXB + xx) * 4, src0 + (z * HWC + y * WC + xx * C_ + cg * 4));
      }
    }
    asm volatile("cp.async.commit_group;\n" ::: "memory");
  };

  issue_copy(0, raw0);
  for (int s = 0; s < nsteps; ++s) {
    const int pass = s / nchunk, chunk = s - pass * nchunk;
    const int c0 = pass * CGs, nch = min(CGs, nch_all - c0), NG = (nch + 3) >> 2;
    const int zb = chunk * ZC, zc = min(ZC, RZ - zb);
    const float *raw = (s & 1) ? raw1 : raw0;
    if (s + 1 < nsteps) {
      issue_copy(s + 1, (s & 1) ? raw0 : raw1);
      asm volatile("cp.async.wait_group 1;\n" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;\n" ::: "memory");
    }
    __syncthreads();
    const int rowsC = zc * RY;
    if (NTX >= 4) planar_x_stage<P, CL, 4>(T, raw, T1, rowsC, rowsMax, RXB, szR, NG, tid);
    else if (NTX == 3) planar_x_stage<P, CL, 3>(T, raw, T1, rowsC, rowsMax, RXB, szR, NG, tid);
    else if (NTX == 2) planar_x_stage<P, CL, 2>(T, raw, T1, rowsC, rowsMax, RXB, szR, NG, tid);
    else planar_x_stage<P, CL, 1>(T, raw, T1, rowsC, rowsMax, RXB, szR, NG, tid);
    __syncthreads();
    if (NTY >= 4) planar_y_stage<P, 4>(T, T1, T2, zb, zc, rowsMax, RY, RZ, NG, tid);
    else if (NTY == 3) planar_y_stage<P, 3>(T, T1, T2, zb, zc, rowsMax, RY, RZ, NG, tid);
    else if (NTY == 2) planar_y_stage<P, 2>(T, T1, T2, zb, zc, rowsMax, RY, RZ, NG, tid);
    else planar_y_stage<P, 1>(T, T1, T2, zb, zc, rowsMax, RY, RZ, NG, tid);
    if (chunk + 1 < nchunk) continue;
    __syncthreads();
    // ---- z stage: T2[group][z][q][4] -> out[c][pd][q], streamed to global: consecutive lanes = consecutive floats
    if (RZ >= 2) planar_z_stage<P, true>(T, T2, out_roi + (long long)c0 * out_elems, out_elems, PD, RZ, NG, nch, tid);
    else planar_z_stage<P, false>(T, T2, out_roi + (long long)c0 * out_elems, out_elems, PD, RZ, NG, nch, tid);
  }
}

// =================================================================================================================
// Backward of the planar formulation (mask branch: 14-wide outputs; channels-last gradients).
//
// Replaces (reference): ROIAlignBackward3D / bilinear_interpolate_gradient_3d, roi_align_kernel.cu:519-636, :383-442 -- 64
// scalar atomics per output element.  Here the three contractions run transposed, lanes over elements, four channels per
// 16-byte entry like the forward:
//     z^T: T2[z][ph,pw]  = sum_pd wz[pd][z] * g[pd][ph,pw]   scatter form, but a thread owns its (ph, pw, group) column for
//                          all pd, so the slices live in a 4-deep register window (bins are monotone in z): grad_out is
//                          read from global exactly once, coalesced, and every T2 entry is written once;
//     y^T: T1[z,y][pw]   = sum_ph wy[ph][y] * T2[z][ph,pw]   gather over the bins whose support holds row y;
//     x^T: d[z,y,x]      = sum_pw wx[pw][x] * T1[z,y][pw]    gather, then ONE 16-byte vector red per (voxel, 4 channels):
//                          each voxel of the footprint receives exactly one red per RoI and channel group (the per-warp
//                          kernels issue one per (voxel, z slice of every pd bin), ~2.6x as many).
// =================================================================================================================
constexpr int PLB_MAXR = 32;   // footprint rows / voxels per row the dense gather tables hold

struct PlanarBwdTables {
  int xo[PL_MAXP], yo[PL_MAXP], zo[PL_MAXP];
  int nx[PL_MAXP], ny[PL_MAXP], nz[PL_MAXP];       // taps per bin (0: the bin has no valid sample)
  alignas(16) float xw[PL_MAXP][4];
  alignas(16) float yw[PL_MAXP][4];
  alignas(16) float zw[PL_MAXP][4];                // pre-multiplied by 1 / count
  int lo[48], hi[48];
  int box[20];
  int ylo[PLB_MAXR], yhi[PLB_MAXR], xlo[PLB_MAXR], xhi[PLB_MAXR];   // bins whose support holds row y / voxel x
  alignas(16) float ywd[PLB_MAXR][PL_MAXP];        // dense: weight of row y in bin ph
  alignas(16) float xwd[PLB_MAXR][PL_MAXP];
  alignas(16) float zwd[PLB_MAXR][PL_MAXP];        // dense: weight of slice z in bin pd, pre-multiplied by 1 / count
};

__device__ __forceinline__ void red4(float *dst, const float4 v) { atomicAdd(reinterpret_cast<float4 *>(dst), v); }

// Literal gradient of one output bin for one channel (reference sample loops, roi_align_kernel.cu:590-632)
__device__ void literal_bin_bwd_strided(const Axis &axw, const Axis &axh, const Axis &axd, int D, int H, int W, float *gc,
                                        long long sz, long long sy, long long sx, int pd, int ph, int pw, float top) {
  const float count = (float)(axd.S * axh.S * axw.S);
  for (int iz = 0; iz < axd.S; ++iz) {
    const Tap tz = axis_tap(axis_coord(axd, pd, iz), D);
    for (int iy = 0; iy < axh.S; ++iy) {
      const Tap ty = axis_tap(axis_coord(axh, ph, iy), H);
      for (int ix = 0; ix < axw.S; ++ix) {
        const Tap tx = axis_tap(axis_coord(axw, pw, ix), W);
        if (!(tz.valid && ty.valid && tx.valid)) continue;
        const float hxhy = __fmul_rn(tx.h, ty.h), lxhy = __fmul_rn(tx.l, ty.h);
        const float hxly = __fmul_rn(tx.h, ty.l), lxly = __fmul_rn(tx.l, ty.l);
        const float w[8] = {__fmul_rn(hxhy, tz.h), __fmul_rn(lxhy, tz.h), __fmul_rn(hxly, tz.h), __fmul_rn(lxly, tz.h),
                            __fmul_rn(hxhy, tz.l), __fmul_rn(lxhy, tz.l), __fmul_rn(hxly, tz.l), __fmul_rn(lxly, tz.l)};
        const long long zl = tz.low * sz, zh = tz.high * sz, yl = ty.low * sy, yh = ty.high * sy;
        const long long xl = tx.low * sx, xh = tx.high * sx;
        const long long off[8] = {zl + yl + xl, zl + yl + xh, zl + yh + xl, zl + yh + xh,
                                  zh + yl + xl, zh + yl + xh, zh + yh + xl, zh + yh + xh};
#pragma unroll
        for (int q = 0; q < 8; ++q) atomicAdd(gc + off[q], __fdiv_rn(__fmul_rn(top, w[q]), count));
      }
    }
  }
}

template <int P, int PDT>
__global__ void __launch_bounds__(PL_THREADS, 2)
    roi_align3d_bwd_planar_kernel(const RoiParams p, int CG, int smem_floats, const int *__restrict__ order) {
  extern __shared__ __align__(16) float planes[];
  __shared__ PlanarBwdTables T;
  constexpr int PP = P * P;
  const int tid = threadIdx.x;
  const int g = blockIdx.x / p.K, kslot = blockIdx.x - g * p.K;
  const int k = order != nullptr ? __ldg(order + kslot) : kslot;
  const int c_first = g * CG;
  const int nch_all = min(CG, p.C - c_first);
  const int PD = PDT > 0 ? PDT : p.PD;

  float r[7];
#pragma unroll
  for (int i = 0; i < 7; ++i) r[i] = __ldg(p.rois + (long long)k * 7 + i);
  const int lvl = p.num_levels > 1 ? roi_level(r, p.num_levels, p.inv_finest) : 0;
  const LevelDev L = p.lv[lvl];
  const int b = (int)r[0];
  const bool ok = b >= 0 && b < p.B;
  const Axis axw = axis_setup(r[1], r[3], L.scale, P, p.sample_num);
  const Axis axh = axis_setup(r[2], r[4], L.scale, P, p.sample_num);
  const Axis axd = axis_setup(r[5], r[6], L.scale_d, PD, p.sample_num);
  const long long out_elems = PDT > 0 ? (long long)PDT * PP : (long long)PD * PP;
  const float *g_roi = p.grad_out + ((long long)k * p.C + c_first) * out_elems;
  if (!ok) return;   // the forward wrote zeros for this RoI: no gradient
  {
    const int axis = tid >> 4, bin = tid & 15;
    if (tid < 48) {
      const int nb = axis == 2 ? PD : P;
      const Axis ax = axis == 0 ? axw : axis == 1 ? axh : axd;
      const int asize = axis == 0 ? L.W : axis == 1 ? L.H : L.D;
      int lo = INT_MAX, hi = -1;
      if (bin < nb) {
        for (int i = 0; i < ax.S; ++i) {
          const Tap t = axis_tap(axis_coord(ax, bin, i), asize);
          if (t.valid) lo = min(lo, t.low), hi = max(hi, t.high);
        }
      }
      T.lo[tid] = lo, T.hi[tid] = hi;
    }
    for (int i = tid; i < PLB_MAXR * PL_MAXP; i += PL_THREADS)
      (&T.ywd[0][0])[i] = 0.0f, (&T.xwd[0][0])[i] = 0.0f, (&T.zwd[0][0])[i] = 0.0f;
    if (tid < PLB_MAXR) T.ylo[tid] = INT_MAX, T.yhi[tid] = -1, T.xlo[tid] = INT_MAX, T.xhi[tid] = -1;
  }
  if (tid < 64) {
    const int lo = tid < 48 ? T.lo[tid] : INT_MAX, hi = tid < 48 ? T.hi[tid] : -1;
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
    const bool empty = q[1] < q[0] || q[4] < q[3] || q[7] < q[6];
    const int RX = empty ? 0 : q[1] - q[0] + 1, RY = empty ? 0 : q[4] - q[3] + 1, RZ = empty ? 0 : q[7] - q[6] + 1;
    const bool slow = q[2] > 4 || q[5] > 4 || q[8] > 4 || RX > PLB_MAXR || RY > PLB_MAXR || RZ > PLB_MAXR;
    int CGs = 0;
    if (!empty && !slow) {
      const int per = RZ * PP + RZ * RY * P;   // T2 + T1 floats per channel
      for (int cg = min((nch_all + 3) & ~3, PL_MAXCG); cg >= 4; cg -= 4)
        if (cg * per <= smem_floats) {
          CGs = cg;
          break;
        }
    }
    T.box[0] = empty ? 0 : q[0], T.box[1] = RX;
    T.box[2] = empty ? 0 : q[3], T.box[3] = RY;
    T.box[4] = empty ? 0 : q[6], T.box[5] = RZ;
    T.box[6] = empty ? 1 : 0;
    T.box[17] = CGs;
  }
  __syncthreads();
  const int x0 = T.box[0], RX = T.box[1], y0 = T.box[2], RY = T.box[3], z0 = T.box[4], RZ = T.box[5];
  const int CGs = T.box[17];
  if (T.box[6] & 1) return;   // no sample inside the level: no gradient
  const long long vox = (long long)L.D * L.H * L.W;
  float *gb = L.grad + (long long)b * vox * p.C;
  if (CGs == 0) {  // literal path (rare): one (channel, element) per thread and trip, scalar atomics
    const long long sx = p.C, sy = sx * L.W, sz = sy * L.H;
    for (long long i = tid; i < (long long)nch_all * out_elems; i += PL_THREADS) {
      const int c = (int)(i / out_elems), e = (int)(i - (long long)c * out_elems);
      const int pw = e % P, ph = (e / P) % P, pd = e / PP;
      literal_bin_bwd_strided(axw, axh, axd, L.D, L.H, L.W, gb + c_first + c, sz, sy, sx, pd, ph, pw, __ldg(g_roi + i));
    }
    return;
  }

  // ---- per-bin taps (exact ranges), then the dense gather tables of the y and x axes
  if (tid < 48) {
    const int axis = tid >> 4, bin = tid & 15;
    const int nb = axis == 2 ? PD : P;
    if (bin < nb) {
      const Axis ax = axis == 0 ? axw : axis == 1 ? axh : axd;
      const int asize = axis == 0 ? L.W : axis == 1 ? L.H : L.D;
      const int lo = T.lo[tid], hi = T.hi[tid];
      float w[4] = {0.0f, 0.0f, 0.0f, 0.0f};
      int off = 0, n = 0;
      if (hi >= lo) {
        n = hi - lo + 1;
        off = lo - (axis == 0 ? x0 : axis == 1 ? y0 : z0);
        for (int i = 0; i < ax.S; ++i) {
          const Tap t = axis_tap(axis_coord(ax, bin, i), asize);
          if (t.valid) {
            const int a0 = t.low - lo, a1 = t.high - lo;
            if (a0 == 0) w[0] += t.h; else if (a0 == 1) w[1] += t.h; else if (a0 == 2) w[2] += t.h; else w[3] += t.h;
            if (a1 == 0) w[0] += t.l; else if (a1 == 1) w[1] += t.l; else if (a1 == 2) w[2] += t.l; else w[3] += t.l;
          }
        }
      }
      if (axis == 2) {
        const float inv = __frcp_rn((float)(axd.S * axh.S * axw.S));
#pragma unroll
        for (int t = 0; t < 4; ++t) w[t] *= inv;
        for (int t = 0; t < n; ++t) T.zwd[off + t][bin] = t == 0 ? w[0] : t == 1 ? w[1] : t == 2 ? w[2] : w[3];
      } else {
        float(*dense)[PL_MAXP] = axis == 0 ? T.xwd : T.ywd;
        int *blo = axis == 0 ? T.xlo : T.ylo, *bhi = axis == 0 ? T.xhi : T.yhi;
        for (int t = 0; t < n; ++t) {
          dense[off + t][bin] = t == 0 ? w[0] : t == 1 ? w[1] : t == 2 ? w[2] : w[3];
          atomicMin(&blo[off + t], bin);
          atomicMax(&bhi[off + t], bin);
        }
      }
    }
  }
  __syncthreads();

  const int rows = RZ * RY;
  float *T2 = planes, *T1 = planes + (size_t)CGs * RZ * PP;   // T2: [g][z][q][4] = CGs * RZ * PP floats; then T1
  const int npass = (nch_all + CGs - 1) / CGs;
  for (int ip = 0; ip < npass; ++ip) {
    const int c0 = ip * CGs, nch = min(CGs, nch_all - c0), NG = (nch + 3) >> 2;
    // ---- z^T: a thread owns (q, group): the grad_out values of all PD bins (4 channels each) are requested together --
    //      one exposed global latency per task, read exactly once, coalesced -- and stay in registers; every slice is
    //      then one pass over the bins with the slice's dense weights (0 outside a bin's support; warp-uniform).
    {
      constexpr int NPD = PDT > 0 ? PDT : PL_MAXP;
      for (int t = tid; t < PP * NG; t += PL_THREADS) {
        const int cg = t / PP, q = t - cg * PP;
        const float *gp = g_roi + (long long)(c0 + cg * 4) * out_elems + q;
        const int left = nch - cg * 4;
        float4 gv[NPD];
#pragma unroll
        for (int pd = 0; pd < NPD; ++pd) {
          gv[pd] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (PDT > 0 || pd < PD) {
            gv[pd].x = __ldcs(gp + pd * PP);
            if (left > 1) gv[pd].y = __ldcs(gp + out_elems + pd * PP);
            if (left > 2) gv[pd].z = __ldcs(gp + 2 * out_elems + pd * PP);
            if (left > 3) gv[pd].w = __ldcs(gp + 3 * out_elems + pd * PP);
          }
        }
        float *dst = T2 + ((cg * RZ) * PP + q) * 4;
        for (int z = 0; z < RZ; ++z) {
          float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
          const float *wz = &T.zwd[z][0];
#pragma unroll
          for (int pd = 0; pd < NPD; ++pd) {
            const float w = wz[pd];
            if (w != 0.0f) fma4(a, w, gv[pd]);
          }
          *reinterpret_cast<float4 *>(dst + z * PP * 4) = a;
        }
      }
    }
    __syncthreads();
    // ---- y^T: T1[g][z * RY + y][pw] = sum over the bins ph that hold row y
    {
      const unsigned ntask = (unsigned)(rows * P), total = ntask * (unsigned)NG;
      const unsigned m_nt = fast_magic(ntask), m_ry = fast_magic((unsigned)RY);
      for (unsigned t = tid; t < total; t += PL_THREADS) {
        const unsigned cg = fast_div(t, ntask, m_nt), u = t - cg * ntask;
        const unsigned row = u / P, pw = u - row * P;
        const unsigned z = fast_div(row, (unsigned)RY, m_ry), y = row - z * (unsigned)RY;
        const int plo = T.ylo[y], phi = T.yhi[y];
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        const float *sp = T2 + ((cg * RZ + z) * PP + pw) * 4;
        const float *wp = &T.ywd[y][0];
        for (int ph = plo; ph <= phi; ++ph) fma4(a, wp[ph], *reinterpret_cast<const float4 *>(sp + ph * P * 4));
        *reinterpret_cast<float4 *>(T1 + ((cg * rows + row) * P + pw) * 4) = a;
      }
    }
    __syncthreads();
    // ---- x^T + scatter: one vector red per (voxel, four channels); groups fastest: 4 lanes = 64 contiguous bytes
    {
      const unsigned total = (unsigned)(rows * RX * NG);
      const unsigned m_ng = fast_magic((unsigned)NG), m_rx = fast_magic((unsigned)RX), m_ry = fast_magic((unsigned)RY);
      for (unsigned t = tid; t < total; t += PL_THREADS) {
        const unsigned u = fast_div(t, (unsigned)NG, m_ng), cg = t - u * (unsigned)NG;
        const unsigned row = fast_div(u, (unsigned)RX, m_rx), x = u - row * (unsigned)RX;
        const unsigned z = fast_div(row, (unsigned)RY, m_ry), y = row - z * (unsigned)RY;
        const int plo = T.xlo[x], phi = T.xhi[x];
        if (phi < plo) continue;   // no bin samples this voxel column
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        const float *sp = T1 + ((cg * rows + row) * P) * 4;
        const float *wp = &T.xwd[x][0];
        for (int pw = plo; pw <= phi; ++pw) fma4(a, wp[pw], *reinterpret_cast<const float4 *>(sp + pw * 4));
        float *dst = gb + ((long long)((z0 + (int)z) * L.H + (y0 + (int)y)) * L.W + (x0 + (int)x)) * p.C + c_first + c0 + cg * 4;
        const int left = nch - (int)cg * 4;
        if (left >= 4) {
          red4(dst, a);
        } else {
          atomicAdd(dst, a.x);
          if (left > 1) atomicAdd(dst + 1, a.y);
          if (left > 2) atomicAdd(dst + 2, a.z);
        }
      }
    }
    // (the next pass's z^T writes T2 only; its barrier separates this pass's T1 reads from the next y^T's writes)
  }
}

}  // namespace

bool fwd_planar_ok(const RoiParams &p, int layout) {
  if (p.PW != p.PH || (p.PW != 7 && p.PW != 14) || p.PD < 1 || p.PD > PL_MAXP) return false;
  if ((long long)p.K * ((p.C + 3) / 4) >= 2147483647LL) return false;
  for (int l = 0; l < p.num_levels; ++l)
    if ((long long)p.lv[l].D * p.lv[l].H * p.lv[l].W * p.C >= 2147483647LL) return false;  // 32-bit offsets inside a volume
  for (int l = 0; l < p.num_levels; ++l) {
    if ((reinterpret_cast<uintptr_t>(p.lv[l].feats) & 15) != 0) return false;
    if (layout == ROI3D_NCDHW) {
      if (p.lv[l].W % 4 != 0) return false;   // 16-byte row pieces: rows must start on 16-byte boundaries
    } else {
      if (p.C % 4 != 0) return false;         // 16-byte copies of four channels of a voxel
    }
  }
  return true;
}

int g_planar_smem_floats = 0;  // roi3d_set_tuning key 10: plane storage per CTA in floats (0 = default, three CTAs per SM)

namespace {
template <int P, bool CL, int PDT>
int launch_planar_cfg(const RoiParams &p, int CG, int ngroups, long long blocks, const int *order, cudaStream_t st) {
  const int smem_floats = g_planar_smem_floats > 0 ? g_planar_smem_floats : PL_SMEM_FLOATS;
  const size_t smem = (size_t)smem_floats * sizeof(float);
  static PerDeviceSmemOptIn opt_in;
  if (opt_in.need(smem)) {
    ROI3D_CUDA(cudaFuncSetAttribute(roi_align3d_fwd_planar_kernel<P, CL, PDT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    opt_in.mark(smem);
  }
  roi_align3d_fwd_planar_kernel<P, CL, PDT><<<(unsigned)blocks, PL_THREADS, smem, st>>>(p, CG, ngroups, smem_floats, order);
  ROI3D_LAUNCH_CHECK();
  return ROI3D_OK;
}
}  // namespace

int launch_fwd_planar(RoiParams &p, int layout, cudaStream_t st) {
  // a CTA takes 64 channels of a RoI (its prologue -- RoI geometry, tap tables -- is paid once for them) and walks them
  // in passes of up to 16 channels, as many as the planes of the RoI's footprint leave room for
  int CG = 64;
  if (p.C < CG) CG = p.C;
  const int ngroups = ceil_div(p.C, CG);
  const long long blocks = (long long)p.K * ngroups;
  ROI3D_CHECK_ARG(blocks < 2147483647LL, "roi_align3d forward: too many work items");
  const bool cl = layout == ROI3D_NDHWC;
  int *order = nullptr;
  if (p.K > 1 && p.K <= 8192) {
    cudaMemPool_t pool;
    const int rc = stream_pool(&pool);
    if (rc) return rc;
    ROI3D_CUDA(cudaMallocFromPoolAsync(reinterpret_cast<void **>(&order), (size_t)p.K * sizeof(int), pool, st));
    roi_align3d_order_kernel<<<ceil_div(p.K, 256), 256, (size_t)p.K * sizeof(unsigned long long), st>>>(p, order);
    ROI3D_LAUNCH_CHECK();
  }
  // output depths with their own instantiation (compile-time plane strides): the cubic shapes and the real config's
  // 7 x 7 x 3 / 14 x 14 x 10 (configs/3d-multi-resolution-rcnn.py); any other depth runs with PD as a run-time value
  int rc;
#define ROI3D_PLANAR_GO(P_, PDT_) \
  (cl ? launch_planar_cfg<P_, true, PDT_>(p, CG, ngroups, blocks, order, st) : launch_planar_cfg<P_, false, PDT_>(p, CG, ngroups, blocks, order, st))
  if (p.PW == 7) rc = p.PD == 7 ? ROI3D_PLANAR_GO(7, 7) : p.PD == 3 ? ROI3D_PLANAR_GO(7, 3) : ROI3D_PLANAR_GO(7, 0);
  else rc = p.PD == 14 ? ROI3D_PLANAR_GO(14, 14) : p.PD == 10 ? ROI3D_PLANAR_GO(14, 10) : ROI3D_PLANAR_GO(14, 0);
#undef ROI3D_PLANAR_GO
  if (order != nullptr) ROI3D_CUDA(cudaFreeAsync(order, st));
  return rc;
}

bool bwd_planar_ok(const RoiParams &p) {
  if (p.PW != p.PH || (p.PW != 14 && p.PW != 7) || p.PD < 1 || p.PD > PL_MAXP) return false;
  if (p.bug_compat || p.C % 4 != 0) return false;
  if ((long long)p.K * ((p.C + 63) / 64) >= 2147483647LL) return false;
  for (int l = 0; l < p.num_levels; ++l)
    if ((reinterpret_cast<uintptr_t>(p.lv[l].grad) & 15) != 0) return false;
  return (reinterpret_cast<uintptr_t>(p.grad_out) & 3) == 0;
}

namespace {
template <int P, int PDT>
int launch_bwd_planar_cfg(const RoiParams &p, int CG, long long blocks, int smem_floats, const int *order, cudaStream_t st) {
  const size_t smem = (size_t)smem_floats * sizeof(float);
  static PerDeviceSmemOptIn opt_in;
  if (opt_in.need(smem)) {
    ROI3D_CUDA(cudaFuncSetAttribute(roi_align3d_bwd_planar_kernel<P, PDT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    opt_in.mark(smem);
  }
  roi_align3d_bwd_planar_kernel<P, PDT><<<(unsigned)blocks, PL_THREADS, smem, st>>>(p, CG, smem_floats, order);
  ROI3D_LAUNCH_CHECK();
  return ROI3D_OK;
}
}  // namespace

int launch_bwd_planar(RoiParams &p, cudaStream_t st) {
  int CG = 64;
  if (p.C < CG) CG = p.C;
  const int ngroups = ceil_div(p.C, CG);
  const long long blocks = (long long)p.K * ngroups;
  const int smem_floats = g_planar_smem_floats > 0 ? g_planar_smem_floats : PL_SMEM_FLOATS;
  int *order = nullptr;
  if (p.K > 1 && p.K <= 8192) {
    cudaMemPool_t pool;
    const int rc = stream_pool(&pool);
    if (rc) return rc;
    ROI3D_CUDA(cudaMallocFromPoolAsync(reinterpret_cast<void **>(&order), (size_t)p.K * sizeof(int), pool, st));
    roi_align3d_order_kernel<<<ceil_div(p.K, 256), 256, (size_t)p.K * sizeof(unsigned long long), st>>>(p, order);
    ROI3D_LAUNCH_CHECK();
  }
  int rc;
  if (p.PW == 14) rc = p.PD == 14 ? launch_bwd_planar_cfg<14, 14>(p, CG, blocks, smem_floats, order, st)
                                  : launch_bwd_planar_cfg<14, 0>(p, CG, blocks, smem_floats, order, st);
  else rc = p.PD == 7 ? launch_bwd_planar_cfg<7, 7>(p, CG, blocks, smem_floats, order, st)
                      : launch_bwd_planar_cfg<7, 0>(p, CG, blocks, smem_floats, order, st);
  if (rc) return rc;
  if (order != nullptr) ROI3D_CUDA(cudaFreeAsync(order, st));
  return ROI3D_OK;
}

}  // namespace roi3d
