// RPN proposal path pieces for B200 (sm_100a): segmented radix-select top-k and the fused
// anchor-generation + delta2bbox3D decode.
//
// Replaces (reference, /root/reference):
//   scores.topk(nms_pre) / scores.topk(max_num)      mmdet/models/anchor_heads/rpn_head_3d.py:108-112, :147
//   sigmoid over every anchor                         rpn_head_3d.py:87-90
//   AnchorGenerator3D.grid_anchors (numpy + H2D)      mmdet/core/anchor/anchor_generator_3d.py:56-71
//   delta2bbox3D (~25 elementwise launches)           mmdet/core/bbox/transforms.py:105-160
//
// Top-k.  Keys are 64-bit: (order-preserving score bits) << 32 | ~logical_index, so keys are unique and
// "descending key" is exactly "descending score, ties -> lower index" (the build's tie rule; torch's own
// tie order is unspecified).  The logical index of an RPN score is its position after
// permute(2,3,1,0).reshape(-1) of the [A,D,H,W] map -- the order the reference's anchors are generated
// in -- while memory is read in its native, coalesced order.  Six MSB-first histogram passes
// (11/11/10 bits over the score word, 11/11/10 over the index word) find the k-th largest key exactly;
// one collect pass appends the k keys >= it; a rank-by-counting pass sorts them.  Every (volume, level)
// segment shares each launch; nothing syncs the host.
#include "common.cuh"

namespace roi3d {

constexpr int kMaxSeg = 64;
constexpr int kBins = 2048;
constexpr int kItemsPerThread = 16;
constexpr int kTopkThreads = 256;
constexpr int kItemsPerCta = kItemsPerThread * kTopkThreads;

struct SegDesc {
  long long off;
  unsigned len;
  int A, D, H, W;  // A == 0: logical index == memory index
  const unsigned char *mask;  // optional, indexed by LOGICAL index: 0 = the element does not take part (rpn_head_3d.py:97-106)
  long long koff;             // this segment's first entry in the key buffer (see topk_hist_kernel)
};

struct SegTable {
  SegDesc s[kMaxSeg];
};

struct SegState {           // device, per segment
  unsigned long long prefix;  // key bits decided so far (left-aligned value of the decided digits)
  int k_rem;                // how many keys still to take inside the current prefix
  int k_take;               // min(k, len)
  int cand_count;
  int bnd_count;            // keys inside the boundary bin after the second digit pass (topk_split_kernel)
  unsigned eff_len;         // elements taking part: len, or the number of set mask bytes (known after digit pass 0)
};

__device__ __forceinline__ unsigned okey(float s) {
  s = s + 0.0f;
  unsigned u = __float_as_uint(s);
  const unsigned k = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
  return k != 0u ? k : 1u;  // 0 is reserved for "masked out" in the key buffer (only -NaN with a full payload maps there)
}

__device__ __forceinline__ float okey_inv(unsigned k) {
  return __uint_as_float((k & 0x80000000u) ? (k ^ 0x80000000u) : ~k);
}

// torch's fp32 sigmoid: 1 / (1 + exp(-x))   (rpn_head_3d.py:90)
__device__ __forceinline__ float sigmoid_ref(float x) { return __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-x))); }

__device__ __forceinline__ unsigned logical_index(const SegDesc &d, unsigned m) {
  if (d.A == 0) return m;
  // memory m = ((a*D + z)*H + y)*W + x  ->  logical ((y*W + x)*D + z)*A + a
  const unsigned x = m % d.W;
  unsigned t = m / d.W;
  const unsigned y = t % d.H;
  t /= d.H;
  const unsigned z = t % d.D;
  const unsigned a = t / d.D;
  return ((y * d.W + x) * d.D + z) * d.A + a;
}

// pass p: digit position and width.  Score word: bits 63..53, 52..42, 41..32; index word: 31..21, 20..10, 9..0
__device__ __forceinline__ void pass_geometry(int pass, int &shift, int &bits) {
  const int sh[6] = {53, 42, 32, 21, 10, 0};
  const int bw[6] = {11, 11, 10, 11, 11, 10};
  shift = sh[pass], bits = bw[pass];
}

// Pick the digit where the descending cumulative count crosses k_rem (256 threads; run by the LAST CTA of a
// segment's histogram pass, see topk_hist_kernel).
__device__ __forceinline__ void scan_segment(unsigned *__restrict__ hist, SegState *__restrict__ state, int seg, int pass) {
  SegState st = state[seg];
  unsigned *g = hist + (long long)seg * kBins;
  __shared__ unsigned part[256];
  __shared__ int s_digit, s_krem;
  // thread t owns bins [8t, 8t+8) ; descending order means high bins first
  unsigned loc[8], sum = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    loc[i] = __ldcg(g + threadIdx.x * 8 + i);  // written by other CTAs' atomics: read through L2
    sum += loc[i];
    g[threadIdx.x * 8 + i] = 0;  // leave the histogram clean for the next pass
  }
  part[threadIdx.x] = sum;
  __syncthreads();
  // suffix sum over threads above me (256 entries; simple serial-in-smem log scan)
  for (int o = 1; o < 256; o <<= 1) {
    unsigned v = (threadIdx.x + o < 256) ? part[threadIdx.x + o] : 0u;
    __syncthreads();
    part[threadIdx.x] += v;
    __syncthreads();
  }
  const unsigned above = part[threadIdx.x] - sum;  // keys in bins strictly above my 8 bins
  if (pass == 0) {
    // the first histogram counts every element that takes part: a masked segment learns its effective length here
    const unsigned total = part[0];
    st.eff_len = total;
    if (total < (unsigned)st.k_take) st.k_take = (int)total, st.k_rem = (int)total;
  }
  const unsigned k = (unsigned)st.k_rem;
  if (k == 0) {  // nothing to select (all masked out)
    __syncthreads();
    if (threadIdx.x == 0) state[seg] = st;
    return;
  }
  if (above < k && above + sum >= k) {
    unsigned cum = above;
    for (int i = 7; i >= 0; --i) {
      if (cum + loc[i] >= k) {
        s_digit = threadIdx.x * 8 + i;
        s_krem = (int)(k - cum);
        break;
      }
      cum += loc[i];
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int shift, bits;
    pass_geometry(pass, shift, bits);
    st.prefix |= (unsigned long long)(unsigned)s_digit << shift;
    st.k_rem = s_krem;
    state[seg] = st;
  }
}

constexpr int kBndCap = 4096;  // boundary keys kept per segment after two digit passes; more (mass ties) -> full passes

// bnd != nullptr (passes 2..5): if the segment's boundary bin fitted kBndCap keys, the pass runs over those keys only
// instead of re-reading and re-scoring the whole segment.
// keys (optional u32 buffer, one entry per score): digit pass 0 stores every element's order-preserving key there (0 =
// the element does not take part: masked out), so that the sigmoid, the mask lookup and the HBM read happen ONCE; the
// later full passes (digit pass 1, the split, and the rare overflow passes) read the keys back from L2.
// Bin counts are aggregated per warp (__match_any_sync) before they touch shared memory: sigmoid scores crowd into a
// handful of exponent bins, which would serialise the shared-memory atomics 32-fold.
template <bool SIGMOID>
__global__ void __launch_bounds__(kTopkThreads) topk_hist_kernel(const float *__restrict__ scores, const SegTable tab,
                                                                 const SegState *__restrict__ state, int pass,
                                                                 unsigned *__restrict__ hist /*[nseg][kBins]*/,
                                                                 const unsigned long long *__restrict__ bnd,
                                                                 int *__restrict__ tickets, unsigned *__restrict__ keys) {
  const int seg = blockIdx.y;
  const SegDesc d = tab.s[seg];
  const SegState st = state[seg];
  const bool use_b = bnd != nullptr && st.bnd_count <= kBndCap;
  const long long len = use_b ? (long long)st.bnd_count : (long long)d.len;
  if ((long long)blockIdx.x * kItemsPerCta >= len) return;
  if (st.k_take <= 0) return;
  __shared__ unsigned h[kBins];
  for (int i = threadIdx.x; i < kBins; i += kTopkThreads) h[i] = 0;
  __syncthreads();
  int shift, bits;
  pass_geometry(pass, shift, bits);
  const unsigned dmask = (1u << bits) - 1u;
  const bool low_word = pass >= 3;
  const unsigned pre_hi = (unsigned)(st.prefix >> 32);
  const float *src = scores + d.off;
  const unsigned long long *bsrc = bnd + (long long)seg * kBndCap;
  unsigned *kbuf = keys != nullptr ? keys + d.koff : nullptr;
  const unsigned lane = threadIdx.x & 31u;
  // grid-stride over the segment: the boundary passes are launched with a few CTAs per segment (the usual case
  // needs one); a segment that overflowed the boundary buffer is then walked by those few CTAs
  for (long long base = (long long)blockIdx.x * kItemsPerCta; base < len; base += (long long)gridDim.x * kItemsPerCta)
#pragma unroll 4
  for (int it = 0; it < kItemsPerThread; ++it) {
    const long long m = base + (long long)it * kTopkThreads + threadIdx.x;
    unsigned bin = 0xFFFFFFFFu;  // no contribution
    if (m < len) {
      unsigned kh, kl_b = 0;
      bool take = true;
      if (use_b) {
        const unsigned long long key = bsrc[m];
        kh = (unsigned)(key >> 32), kl_b = (unsigned)key;
      } else if (kbuf != nullptr && pass > 0) {
        kh = __ldg(kbuf + m);
        take = kh != 0u;
      } else {
        take = d.mask == nullptr || __ldg(d.mask + logical_index(d, (unsigned)m)) != 0;
        float v = __ldg(src + m);
        if (SIGMOID) v = sigmoid_ref(v);
        kh = take ? okey(v) : 0u;
        if (kbuf != nullptr) kbuf[m] = kh;
      }
      if (take) {
        if (!low_word) {
          // participates iff the already-decided high digits match
          const int decided = 32 - (shift - 32) - bits;  // number of decided bits of the score word
          const bool match = decided == 0 || (kh >> (32 - decided)) == (pre_hi >> (32 - decided));
          if (match) bin = (kh >> (shift - 32)) & dmask;
        } else if (kh == pre_hi) {
          const unsigned kl = use_b ? kl_b : ~logical_index(d, (unsigned)m);
          const unsigned pre_lo = (unsigned)st.prefix;
          const int decided = 32 - shift - bits;
          const bool match = decided == 0 || (kl >> (32 - decided)) == (pre_lo >> (32 - decided));
          if (match) bin = (kl >> shift) & dmask;
        }
      }
    }
    // one shared-memory atomic per distinct bin of the warp
    const unsigned peers = __match_any_sync(0xffffffffu, bin);
    if (bin != 0xFFFFFFFFu && lane == (unsigned)(__ffs(peers) - 1)) atomicAdd(&h[bin], (unsigned)__popc(peers));
  }
  __syncthreads();
  unsigned *g = hist + (long long)seg * kBins;
  for (int i = threadIdx.x; i < kBins; i += kTopkThreads) {
    const unsigned c = h[i];
    if (c) atomicAdd(&g[i], c);
  }
  // the last CTA of this segment to get here scans the finished histogram (no separate scan launch)
  __threadfence();
  __syncthreads();
  __shared__ int s_last;
  if (threadIdx.x == 0) {
    const long long want = (len + kItemsPerCta - 1) / kItemsPerCta;
    const int nct = (int)(want < (long long)gridDim.x ? want : (long long)gridDim.x);
    s_last = atomicAdd(&tickets[seg], 1) == nct - 1;
    if (s_last) tickets[seg] = 0;
  }
  __syncthreads();
  if (s_last) {
    __threadfence();
    scan_segment(hist, const_cast<SegState *>(state), seg, pass);
  }
}

template <bool SIGMOID>
__global__ void __launch_bounds__(kTopkThreads) topk_collect_kernel(const float *__restrict__ scores,
                                                                    const SegTable tab, SegState *__restrict__ state,
                                                                    int k, unsigned long long *__restrict__ cand,
                                                                    const unsigned long long *__restrict__ bnd,
                                                                    const unsigned *__restrict__ keys) {
  const int seg = blockIdx.y;
  const SegDesc d = tab.s[seg];
  const long long base = (long long)blockIdx.x * kItemsPerCta;
  const SegState st = state[seg];
  const bool use_b = bnd != nullptr && st.bnd_count <= kBndCap;  // keys above the boundary bin are in cand already
  const long long len = use_b ? (long long)st.bnd_count : (long long)d.len;
  if (base >= len) return;
  if (st.k_take <= 0) return;
  const unsigned thr_hi = (unsigned)(st.prefix >> 32);
  const float *src = scores + d.off;
  const unsigned long long *bsrc = bnd + (long long)seg * kBndCap;
  for (long long base2 = base; base2 < len; base2 += (long long)gridDim.x * kItemsPerCta)
#pragma unroll 4
  for (int it = 0; it < kItemsPerThread; ++it) {
    const long long m = base2 + (long long)it * kTopkThreads + threadIdx.x;
    if (m < len) {
      unsigned long long key;
      if (use_b) {
        key = bsrc[m];
      } else if (keys != nullptr) {
        const unsigned kh = __ldg(keys + d.koff + m);   // 0 (masked out) is below every threshold that selects anything
        if (kh == 0u || kh < thr_hi) continue;
        key = ((unsigned long long)kh << 32) | (unsigned)~logical_index(d, (unsigned)m);
      } else {
        float v = __ldg(src + m);
        if (SIGMOID) v = sigmoid_ref(v);
        const unsigned kh = okey(v);
        if (kh < thr_hi) continue;
        const unsigned li = logical_index(d, (unsigned)m);
        if (d.mask != nullptr && !__ldg(d.mask + li)) continue;
        key = ((unsigned long long)kh << 32) | (unsigned)~li;
      }
      if (key >= st.prefix) {
        const int pos = atomicAdd(&state[seg].cand_count, 1);
        if (pos < k) cand[(long long)seg * k + pos] = key;
      }
    }
  }
}

// After the first two digit passes (22 bits of the score word decided): one pass over the whole segment that
// (a) appends every key ABOVE the boundary bin to the result candidates -- they are certainly selected -- and
// (b) copies the keys INSIDE the boundary bin to a small buffer, on which the remaining four digit passes and the
// final collect run.  Five of the seven full passes over the scores disappear.
template <bool SIGMOID>
__global__ void __launch_bounds__(kTopkThreads) topk_split_kernel(const float *__restrict__ scores, const SegTable tab,
                                                                  SegState *__restrict__ state, int k,
                                                                  unsigned long long *__restrict__ cand,
                                                                  unsigned long long *__restrict__ bnd,
                                                                  const unsigned *__restrict__ keys) {
  const int seg = blockIdx.y;
  const SegDesc d = tab.s[seg];
  const long long base = (long long)blockIdx.x * kItemsPerCta;
  if (base >= (long long)d.len) return;
  const SegState st = state[seg];
  if (st.k_take <= 0) return;
  const unsigned p22 = (unsigned)(st.prefix >> 42);  // the 22 decided bits
  const float *src = scores + d.off;
#pragma unroll 4
  for (int it = 0; it < kItemsPerThread; ++it) {
    const long long m = base + (long long)it * kTopkThreads + threadIdx.x;
    if (m < (long long)d.len) {
      unsigned kh;
      if (keys != nullptr) {
        kh = __ldg(keys + d.koff + m);
        if (kh == 0u) continue;
      } else {
        float v = __ldg(src + m);
        if (SIGMOID) v = sigmoid_ref(v);
        kh = okey(v);
      }
      const unsigned h22 = kh >> 10;
      if (h22 >= p22) {
        const unsigned li = logical_index(d, (unsigned)m);
        if (keys == nullptr && d.mask != nullptr && !__ldg(d.mask + li)) continue;
        const unsigned long long key = ((unsigned long long)kh << 32) | (unsigned)~li;
        if (h22 > p22) {
          const int pos = atomicAdd(&state[seg].cand_count, 1);
          if (pos < k) cand[(long long)seg * k + pos] = key;
        } else {
          const int pos = atomicAdd(&state[seg].bnd_count, 1);
          if (pos < kBndCap) bnd[(long long)seg * kBndCap + pos] = key;
        }
      }
    }
  }
}

// Boundary bin larger than the buffer (mass ties): forget the split, the full-data passes and collect take over.
__global__ void topk_after_split_kernel(SegState *state, int nseg) {
  const int s = threadIdx.x;
  if (s < nseg && state[s].bnd_count > kBndCap) state[s].cand_count = 0;
}

// rank-by-counting sort of the (unique) candidate keys, descending.  grid (ceil(k/256), nseg).
// by_index (a segment no longer than k, when the caller asks for it): the whole segment is selected and is returned
// in ascending logical index instead -- the reference only sorts a level that has more than nms_pre anchors
// (rpn_head_3d.py:96,108-112).  The low key word is ~index, so "descending low word" is ascending index.
__global__ void __launch_bounds__(256) topk_sort_kernel(const unsigned long long *__restrict__ cand,
                                                        const SegState *__restrict__ state, const SegTable tab,
                                                        int small_by_index, int k, int64_t *__restrict__ out_idx,
                                                        float *__restrict__ out_val) {
  const int seg = blockIdx.y;
  const int n = state[seg].k_take;
  const bool by_index = small_by_index && state[seg].eff_len <= (unsigned)k;
  if ((int)(blockIdx.x * 256) >= n) return;
  const unsigned long long *c = cand + (long long)seg * k;
  const int i = blockIdx.x * 256 + threadIdx.x;
  const unsigned long long ki = i < n ? c[i] : 0ULL;
  __shared__ unsigned long long keys[256];
  int rank = 0;
  for (int j0 = 0; j0 < n; j0 += 256) {
    __syncthreads();
    keys[threadIdx.x] = (j0 + threadIdx.x) < n ? c[j0 + threadIdx.x] : 0ULL;
    __syncthreads();
    const int lim = min(256, n - j0);
    if (i < n) {
#pragma unroll 8
      for (int t = 0; t < lim; ++t) rank += by_index ? ((unsigned)keys[t] > (unsigned)ki) : (keys[t] > ki);
    }
  }
  if (i < n) {
    out_idx[(long long)seg * k + rank] = (int64_t)(unsigned)~(unsigned)ki;
    out_val[(long long)seg * k + rank] = okey_inv((unsigned)(ki >> 32));
  }
}

// Same ordering by a bitonic sort in shared memory, one CTA per segment, for k <= kBitonicMax (rank-by-counting is
// O(k^2): 4 M key compares per segment at k = 2000).  Unused slots hold key 0, which sorts last in either order.
constexpr int kBitonicMax = 4096;
__global__ void __launch_bounds__(1024) topk_bitonic_kernel(const unsigned long long *__restrict__ cand,
                                                            const SegState *__restrict__ state, const SegTable tab,
                                                            int small_by_index, int k, int npow2,
                                                            int64_t *__restrict__ out_idx, float *__restrict__ out_val) {
  extern __shared__ __align__(16) unsigned long long sk[];
  const int seg = blockIdx.x, tid = threadIdx.x;
  const int n = state[seg].k_take;
  if (n <= 0) return;
  const bool by_index = small_by_index && state[seg].eff_len <= (unsigned)k;
  const unsigned long long *c = cand + (long long)seg * k;
  // sort key: the 64-bit key itself (descending), or only its low word = ~index (descending = ascending index)
  for (int i = tid; i < npow2; i += 1024) {
    unsigned long long v = i < n ? c[i] : 0ULL;
    if (by_index && i < n) v = ((unsigned long long)(unsigned)v << 32) | (v >> 32);  // swap words: index word leads
    sk[i] = v;
  }
  __syncthreads();
  for (int size = 2; size <= npow2; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = tid; t < (npow2 >> 1); t += 1024) {
        const int lo = 2 * t - (t & (stride - 1));
        const int hi = lo + stride;
        const bool desc = (lo & size) == 0;  // descending blocks first -> whole array descending at the end
        const unsigned long long a = sk[lo], b = sk[hi];
        if (desc ? (a < b) : (a > b)) sk[lo] = b, sk[hi] = a;
      }
      __syncthreads();
    }
  }
  for (int i = tid; i < n; i += 1024) {
    unsigned long long v = sk[i];
    if (by_index) v = ((unsigned long long)(unsigned)v << 32) | (v >> 32);
    out_idx[(long long)seg * k + i] = (int64_t)(unsigned)~(unsigned)v;
    out_val[(long long)seg * k + i] = okey_inv((unsigned)(v >> 32));
  }
}

__global__ void topk_init_kernel(SegState *state, int *tickets, const SegTable tab, int nseg, int k) {
  const int s = threadIdx.x;
  if (s >= nseg) return;
  tickets[s] = 0;
  SegState st;
  st.prefix = 0ULL;
  st.k_take = (int)min((unsigned)k, tab.s[s].len);
  st.k_rem = st.k_take;
  st.cand_count = 0;
  st.bnd_count = 0;
  st.eff_len = tab.s[s].len;
  state[s] = st;
}

// What the callers downstream need to know about each segment without a host read: how many rows came back and
// whether they are in score order (1) or, for a small segment returned whole, in ascending logical index (0).
__global__ void topk_report_kernel(const SegState *state, int nseg, int k, int small_by_index, int32_t *out_count,
                                   unsigned char *out_sorted) {
  const int s = threadIdx.x;
  if (s >= nseg) return;
  if (out_count != nullptr) out_count[s] = state[s].k_take;
  if (out_sorted != nullptr) out_sorted[s] = !(small_by_index && state[s].eff_len <= (unsigned)k);
}

// ------------------------------------------------------------------------------------------------
// Fused anchors + delta2bbox3D + score append for selected anchors of one level.
// ------------------------------------------------------------------------------------------------
struct DecodeParams {
  const float *bbox_pred;  // [6A, D, H, W]
  int A, D, H, W;
  float stride, dstride;
  float base[16][6];
  float means[6], stds[6];
  float img_h, img_w, img_d;
  float max_ratio;
  const int64_t *idx;
  const float *scores;
  int n;
  float *out;
};

__global__ void __launch_bounds__(256) decode_proposals_kernel(const DecodeParams p) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.n) return;
  float *o = p.out + (long long)i * 7;
  const long long li = p.idx[i];
  if (li < 0) {
#pragma unroll
    for (int j = 0; j < 7; ++j) o[j] = 0.0f;
    return;
  }
  // logical index ((y*W + x)*D + z)*A + a   (anchor_generator_3d.py:59-70, np.meshgrid 'xy' order)
  long long t = li;
  const int a = (int)(t % p.A);
  t /= p.A;
  const int z = (int)(t % p.D);
  t /= p.D;
  const int x = (int)(t % p.W);
  const int y = (int)(t / p.W);
  const float sx = (float)x * p.stride, sy = (float)y * p.stride, sz = (float)z * p.dstride;
  const float ax1 = __fadd_rn(p.base[a][0], sx), ay1 = __fadd_rn(p.base[a][1], sy);
  const float ax2 = __fadd_rn(p.base[a][2], sx), ay2 = __fadd_rn(p.base[a][3], sy);
  const float az1 = __fadd_rn(p.base[a][4], sz), az2 = __fadd_rn(p.base[a][5], sz);
  const long long plane = (long long)p.D * p.H * p.W;
  const long long sp = ((long long)z * p.H + y) * p.W + x;
  float d[6];
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    const float raw = __ldg(p.bbox_pred + (long long)(a * 6 + j) * plane + sp);
    d[j] = __fadd_rn(__fmul_rn(raw, p.stds[j]), p.means[j]);  // deltas * stds + means (transforms.py:114)
  }
  const float mr = p.max_ratio;
  const float dx = d[0], dy = d[1];
  const float dw = fminf(fmaxf(d[2], -mr), mr), dh = fminf(fmaxf(d[3], -mr), mr);
  const float dz = fminf(fmaxf(d[4], -mr), mr), dd = fminf(fmaxf(d[5], -mr), mr);
  const float px = __fmul_rn(__fadd_rn(ax1, ax2), 0.5f), py = __fmul_rn(__fadd_rn(ay1, ay2), 0.5f);
  const float pz = __fmul_rn(__fadd_rn(az1, az2), 0.5f);
  const float pw = __fadd_rn(__fsub_rn(ax2, ax1), 1.0f), ph = __fadd_rn(__fsub_rn(ay2, ay1), 1.0f);
  const float pdz = __fadd_rn(__fsub_rn(az2, az1), 1.0f);
  const float gw = __fmul_rn(pw, expf(dw)), gh = __fmul_rn(ph, expf(dh)), gd = __fmul_rn(pdz, expf(dd));
  const float gx = __fadd_rn(px, __fmul_rn(pw, dx)), gy = __fadd_rn(py, __fmul_rn(ph, dy));
  const float gz = __fadd_rn(pz, __fmul_rn(pdz, dz));
  float x1 = __fadd_rn(__fsub_rn(gx, __fmul_rn(gw, 0.5f)), 0.5f), y1 = __fadd_rn(__fsub_rn(gy, __fmul_rn(gh, 0.5f)), 0.5f);
  float x2 = __fsub_rn(__fadd_rn(gx, __fmul_rn(gw, 0.5f)), 0.5f), y2 = __fsub_rn(__fadd_rn(gy, __fmul_rn(gh, 0.5f)), 0.5f);
  float z1 = __fadd_rn(__fsub_rn(gz, __fmul_rn(gd, 0.5f)), 0.5f), z2 = __fsub_rn(__fadd_rn(gz, __fmul_rn(gd, 0.5f)), 0.5f);
  if (p.img_w > 0.0f) {
    x1 = fminf(fmaxf(x1, 0.0f), p.img_w - 1.0f), x2 = fminf(fmaxf(x2, 0.0f), p.img_w - 1.0f);
    y1 = fminf(fmaxf(y1, 0.0f), p.img_h - 1.0f), y2 = fminf(fmaxf(y2, 0.0f), p.img_h - 1.0f);
    z1 = fminf(fmaxf(z1, 0.0f), p.img_d - 1.0f), z2 = fminf(fmaxf(z2, 0.0f), p.img_d - 1.0f);
  }
  o[0] = x1, o[1] = y1, o[2] = x2, o[3] = y2, o[4] = z1, o[5] = z2;
  o[6] = p.scores ? p.scores[i] : 0.0f;
}

// Batched form: one launch decodes the selected anchors of every (image, level) segment.
struct DecodeSeg {
  const float *bbox_pred;  // [6A, D, H, W] of this (image, level)
  int A, D, H, W;
  int level;               // index into DecodeBatch::base / stride tables
  float img_h, img_w, img_d;
};

struct DecodeBatch {
  DecodeSeg seg[kMaxSeg];
  float base[ROI3D_MAX_LEVELS][4][6];
  float stride[ROI3D_MAX_LEVELS], dstride[ROI3D_MAX_LEVELS];
  float means[6], stds[6];
  float max_ratio;
  const int64_t *idx;   // [nseg, k]
  const float *scores;  // [nseg, k]
  int k;
  float *out;           // [nseg, k, 7]
};

__global__ void __launch_bounds__(256) decode_proposals_batched_kernel(const DecodeBatch b) {
  const int s = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= b.k) return;
  DecodeParams p;
  const DecodeSeg &sg = b.seg[s];
  p.bbox_pred = sg.bbox_pred, p.A = sg.A, p.D = sg.D, p.H = sg.H, p.W = sg.W;
  p.stride = b.stride[sg.level], p.dstride = b.dstride[sg.level];
  float *o = b.out + ((long long)s * b.k + i) * 7;
  const long long li = b.idx[(long long)s * b.k + i];
  if (li < 0) {
#pragma unroll
    for (int j = 0; j < 7; ++j) o[j] = 0.0f;
    return;
  }
  long long t = li;
  const int a = (int)(t % p.A);
  t /= p.A;
  const int z = (int)(t % p.D);
  t /= p.D;
  const int x = (int)(t % p.W);
  const int y = (int)(t / p.W);
  const float *bs = b.base[sg.level][a];
  const float sx = (float)x * p.stride, sy = (float)y * p.stride, sz = (float)z * p.dstride;
  const float ax1 = __fadd_rn(bs[0], sx), ay1 = __fadd_rn(bs[1], sy);
  const float ax2 = __fadd_rn(bs[2], sx), ay2 = __fadd_rn(bs[3], sy);
  const float az1 = __fadd_rn(bs[4], sz), az2 = __fadd_rn(bs[5], sz);
  const long long plane = (long long)p.D * p.H * p.W;
  const long long sp = ((long long)z * p.H + y) * p.W + x;
  float d[6];
#pragma unroll
  for (int j = 0; j < 6; ++j)
    d[j] = __fadd_rn(__fmul_rn(__ldg(p.bbox_pred + (long long)(a * 6 + j) * plane + sp), b.stds[j]), b.means[j]);
  const float mr = b.max_ratio;
  const float dw = fminf(fmaxf(d[2], -mr), mr), dh = fminf(fmaxf(d[3], -mr), mr);
  const float dz = fminf(fmaxf(d[4], -mr), mr), dd = fminf(fmaxf(d[5], -mr), mr);
  const float px = __fmul_rn(__fadd_rn(ax1, ax2), 0.5f), py = __fmul_rn(__fadd_rn(ay1, ay2), 0.5f);
  const float pz = __fmul_rn(__fadd_rn(az1, az2), 0.5f);
  const float pw = __fadd_rn(__fsub_rn(ax2, ax1), 1.0f), ph = __fadd_rn(__fsub_rn(ay2, ay1), 1.0f);
  const float pdz = __fadd_rn(__fsub_rn(az2, az1), 1.0f);
  const float gw = __fmul_rn(pw, expf(dw)), gh = __fmul_rn(ph, expf(dh)), gd = __fmul_rn(pdz, expf(dd));
  const float gx = __fadd_rn(px, __fmul_rn(pw, d[0])), gy = __fadd_rn(py, __fmul_rn(ph, d[1]));
  const float gz = __fadd_rn(pz, __fmul_rn(pdz, dz));
  float x1 = __fadd_rn(__fsub_rn(gx, __fmul_rn(gw, 0.5f)), 0.5f), y1 = __fadd_rn(__fsub_rn(gy, __fmul_rn(gh, 0.5f)), 0.5f);
  float x2 = __fsub_rn(__fadd_rn(gx, __fmul_rn(gw, 0.5f)), 0.5f), y2 = __fsub_rn(__fadd_rn(gy, __fmul_rn(gh, 0.5f)), 0.5f);
  float z1 = __fadd_rn(__fsub_rn(gz, __fmul_rn(gd, 0.5f)), 0.5f), z2 = __fsub_rn(__fadd_rn(gz, __fmul_rn(gd, 0.5f)), 0.5f);
  if (sg.img_w > 0.0f) {
    x1 = fminf(fmaxf(x1, 0.0f), sg.img_w - 1.0f), x2 = fminf(fmaxf(x2, 0.0f), sg.img_w - 1.0f);
    y1 = fminf(fmaxf(y1, 0.0f), sg.img_h - 1.0f), y2 = fminf(fmaxf(y2, 0.0f), sg.img_h - 1.0f);
    z1 = fminf(fmaxf(z1, 0.0f), sg.img_d - 1.0f), z2 = fminf(fmaxf(z2, 0.0f), sg.img_d - 1.0f);
  }
  o[0] = x1, o[1] = y1, o[2] = x2, o[3] = y2, o[4] = z1, o[5] = z2;
  o[6] = b.scores ? b.scores[(long long)s * b.k + i] : 0.0f;
}

// ------------------------------------------------------------------------------------------------
// Tail of RPNHead3D.get_bboxes_single for every (image, level) segment at once (rpn_head_3d.py:135-148):
// proposals[:nms_post] of each level in NMS return order, concatenated per image in level order; scores of the
// unused tail are -inf so that the following top-k(max_num) ignores them.
// grid (ceil(P/256), B*L).  keep_sel[s] picks, per segment, the score-ordered or the index-ordered keep list.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) rpn_collect_kernel(const float *__restrict__ dets, int k,
                                                          const int64_t *__restrict__ keep_by_score,
                                                          const int64_t *__restrict__ keep_by_index,
                                                          const int32_t *__restrict__ num_keep,
                                                          const unsigned char *__restrict__ use_index_order, int L, int P,
                                                          float *__restrict__ cat_props, float *__restrict__ cat_scores,
                                                          int32_t *__restrict__ n_valid) {
  const int seg = blockIdx.y, b = seg / L, l = seg - b * L;
  const int i = blockIdx.x * 256 + threadIdx.x;
  int before = 0, total = 0, mine = 0;
  for (int q = 0; q < L; ++q) {
    const int c = min(max(num_keep[b * L + q], 0), P);
    if (q < l) before += c;
    if (q == l) mine = c;
    total += c;
  }
  if (l == 0 && i == 0) n_valid[b] = total;
  if (i >= P) return;
  const long long row0 = (long long)b * L * P;
  if (i < mine) {
    const int64_t *keep = (use_index_order != nullptr && use_index_order[seg]) ? keep_by_index : keep_by_score;
    long long r = keep[(long long)seg * k + i];
    r = r < 0 ? 0 : (r >= k ? k - 1 : r);
    const float *src = dets + ((long long)seg * k + r) * 7;
    float *dst = cat_props + (row0 + before + i) * 7;
#pragma unroll
    for (int q = 0; q < 7; ++q) dst[q] = src[q];
    cat_scores[row0 + before + i] = src[6];
  } else {
    // this segment's unused slots fill the image's tail: total + (slack of earlier levels) + own slack position
    const int slack_before = l * P - before;
    const long long pos = row0 + total + slack_before + (i - mine);
    cat_scores[pos] = -INFINITY;
  }
}

// out[b][j] = rows[b][idx[b][j]] (7 floats); idx < 0 -> zeros.  grid (ceil(n/256), B)
__global__ void __launch_bounds__(256) gather_rows7_kernel(const float *__restrict__ rows, int rows_per_seg,
                                                           const int64_t *__restrict__ idx, int n,
                                                           float *__restrict__ out) {
  const int b = blockIdx.y, j = blockIdx.x * 256 + threadIdx.x;
  if (j >= n) return;
  const long long r = idx[(long long)b * n + j];
  float *dst = out + ((long long)b * n + j) * 7;
  if (r < 0 || r >= rows_per_seg) {
#pragma unroll
    for (int q = 0; q < 7; ++q) dst[q] = 0.0f;
    return;
  }
  const float *src = rows + ((long long)b * rows_per_seg + r) * 7;
#pragma unroll
  for (int q = 0; q < 7; ++q) dst[q] = src[q];
}

}  // namespace roi3d

using namespace roi3d;

extern "C" {

static const size_t kStateBytes = 4096;                                     // kMaxSeg * sizeof(SegState) + tickets, rounded up
static const size_t kHistBytes = (size_t)kMaxSeg * kBins * sizeof(unsigned);  // 512 KiB

size_t roi3d_topk_workspace_bytes(int nseg, int k) {
  if (nseg <= 0 || k <= 0) return 256;
  const size_t ns = (size_t)(nseg < kMaxSeg ? nseg : kMaxSeg);
  const size_t b = kStateBytes + kHistBytes + ns * (size_t)k * sizeof(unsigned long long) +
                   ns * (size_t)kBndCap * sizeof(unsigned long long);
  return (b + 255) / 256 * 256;
}

size_t roi3d_topk_workspace_bytes_keys(int nseg, int k, int64_t total_len) {
  const size_t base = roi3d_topk_workspace_bytes(nseg, k);
  if (nseg <= 0 || k <= 0 || total_len <= 0) return base;
  return base + ((size_t)total_len * sizeof(unsigned) + 255) / 256 * 256;
}

int roi3d_topk_segmented(const float *scores_dev, const int64_t *seg_off, const int64_t *seg_len,
                         const int32_t *seg_adhw, int nseg, int k, int apply_sigmoid, int64_t *out_idx_dev,
                         float *out_val_dev, void *workspace_dev, size_t workspace_bytes, void *stream) {
  return roi3d_topk_segmented_ex(scores_dev, seg_off, seg_len, seg_adhw, nseg, k, apply_sigmoid, 0, out_idx_dev,
                                 out_val_dev, workspace_dev, workspace_bytes, stream);
}

int roi3d_topk_segmented_ex(const float *scores_dev, const int64_t *seg_off, const int64_t *seg_len,
                            const int32_t *seg_adhw, int nseg, int k, int apply_sigmoid, int small_segments_in_index_order,
                            int64_t *out_idx_dev, float *out_val_dev, void *workspace_dev, size_t workspace_bytes,
                            void *stream) {
  return roi3d_topk_segmented_masked(scores_dev, seg_off, seg_len, seg_adhw, nullptr, nseg, k, apply_sigmoid,
                                     small_segments_in_index_order, out_idx_dev, out_val_dev, nullptr, nullptr,
                                     workspace_dev, workspace_bytes, stream);
}

int roi3d_topk_segmented_masked(const float *scores_dev, const int64_t *seg_off, const int64_t *seg_len,
                                const int32_t *seg_adhw, const uint8_t *const *seg_mask_dev_ptrs, int nseg, int k,
                                int apply_sigmoid, int small_segments_in_index_order, int64_t *out_idx_dev,
                                float *out_val_dev, int32_t *out_count_dev, uint8_t *out_sorted_dev, void *workspace_dev,
                                size_t workspace_bytes, void *stream) {
  ROI3D_CHECK_ARG(nseg >= 0 && k >= 0, "bad sizes");
  if (nseg == 0 || k == 0) return ROI3D_OK;
  ROI3D_CHECK_ARG(scores_dev && seg_off && seg_len && out_idx_dev && out_val_dev && workspace_dev, "NULL pointer");
  ROI3D_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace_dev) & 255) == 0, "workspace must be 256-byte aligned");
  if (workspace_bytes < roi3d_topk_workspace_bytes(nseg, k)) {
    set_error("topk workspace too small: %zu < %zu", workspace_bytes, roi3d_topk_workspace_bytes(nseg, k));
    return ROI3D_ENOMEM;
  }
  cudaStream_t st = (cudaStream_t)stream;
  ROI3D_CUDA(cudaMemsetAsync(out_idx_dev, 0xFF, sizeof(int64_t) * (size_t)nseg * k, st));
  ROI3D_CUDA(cudaMemsetAsync(out_val_dev, 0, sizeof(float) * (size_t)nseg * k, st));
  // optional key buffer behind the base workspace (roi3d_topk_workspace_bytes_keys): one u32 per score of a batch of
  // kMaxSeg segments, written by the first digit pass and read back (from L2) by the later full passes
  long long total_len = 0;
  for (int s = 0; s < nseg; ++s) total_len += seg_len[s] > 0 ? seg_len[s] : 0;
  const size_t base_bytes = roi3d_topk_workspace_bytes(nseg, k);
  unsigned *keys = workspace_bytes >= roi3d_topk_workspace_bytes_keys(nseg, k, total_len) && total_len > 0
                       ? reinterpret_cast<unsigned *>(static_cast<char *>(workspace_dev) + base_bytes)
                       : nullptr;
  for (int s0 = 0; s0 < nseg; s0 += kMaxSeg) {
    const int ns = nseg - s0 < kMaxSeg ? nseg - s0 : kMaxSeg;
    SegTable tab;
    long long maxlen = 0, koff = 0;
    for (int s = 0; s < ns; ++s) {
      ROI3D_CHECK_ARG(seg_len[s0 + s] >= 0 && seg_len[s0 + s] < 4294967295LL, "segment %d too long", s0 + s);
      tab.s[s].off = seg_off[s0 + s];
      tab.s[s].len = (unsigned)seg_len[s0 + s];
      if (seg_adhw) {
        tab.s[s].A = seg_adhw[(s0 + s) * 4 + 0], tab.s[s].D = seg_adhw[(s0 + s) * 4 + 1];
        tab.s[s].H = seg_adhw[(s0 + s) * 4 + 2], tab.s[s].W = seg_adhw[(s0 + s) * 4 + 3];
        ROI3D_CHECK_ARG(tab.s[s].A == 0 || (long long)tab.s[s].A * tab.s[s].D * tab.s[s].H * tab.s[s].W == seg_len[s0 + s],
                        "segment %d: A*D*H*W != len", s0 + s);
      } else {
        tab.s[s].A = tab.s[s].D = tab.s[s].H = tab.s[s].W = 0;
      }
      tab.s[s].mask = seg_mask_dev_ptrs != nullptr ? seg_mask_dev_ptrs[s0 + s] : nullptr;
      tab.s[s].koff = koff;
      koff += seg_len[s0 + s];
      if (seg_len[s0 + s] > maxlen) maxlen = seg_len[s0 + s];
    }
    static_assert(sizeof(SegState) * kMaxSeg + sizeof(int) * kMaxSeg <= 4096, "state block");
    char *b = static_cast<char *>(workspace_dev);
    SegState *state = reinterpret_cast<SegState *>(b);
    unsigned *hist = reinterpret_cast<unsigned *>(b + kStateBytes);
    unsigned long long *cand = reinterpret_cast<unsigned long long *>(b + kStateBytes + kHistBytes);
    unsigned long long *bnd = cand + (size_t)(nseg < kMaxSeg ? nseg : kMaxSeg) * k;
    int *tickets = reinterpret_cast<int *>(b + sizeof(SegState) * kMaxSeg);  // inside the 4 KB state block
    topk_init_kernel<<<1, kMaxSeg, 0, st>>>(state, tickets, tab, ns, k);
    ROI3D_LAUNCH_CHECK();
    if (maxlen == 0) {
      if (out_count_dev != nullptr || out_sorted_dev != nullptr) {
        topk_report_kernel<<<1, kMaxSeg, 0, st>>>(state, ns, k, small_segments_in_index_order,
                                                  out_count_dev ? out_count_dev + s0 : nullptr,
                                                  out_sorted_dev ? out_sorted_dev + s0 : nullptr);
        ROI3D_LAUNCH_CHECK();
      }
      continue;
    }
    ROI3D_CUDA(cudaMemsetAsync(hist, 0, (size_t)ns * kBins * sizeof(unsigned), st));
    const dim3 grid((unsigned)ceil_div_ll(maxlen, kItemsPerCta), ns);
    const dim3 grid2((unsigned)(grid.x < 8 ? grid.x : 8), ns);  // boundary passes: see topk_hist_kernel
    for (int pass = 0; pass < 6; ++pass) {
      const unsigned long long *b2 = pass >= 2 ? bnd : nullptr;
      const dim3 g = pass >= 2 ? grid2 : grid;
      if (apply_sigmoid)
        topk_hist_kernel<true><<<g, kTopkThreads, 0, st>>>(scores_dev, tab, state, pass, hist, b2, tickets, keys);
      else
        topk_hist_kernel<false><<<g, kTopkThreads, 0, st>>>(scores_dev, tab, state, pass, hist, b2, tickets, keys);
      ROI3D_LAUNCH_CHECK();
      if (pass == 1) {  // 22 bits decided: split off the certain keys and the boundary bin
        if (apply_sigmoid)
          topk_split_kernel<true><<<grid, kTopkThreads, 0, st>>>(scores_dev, tab, state, k, cand, bnd, keys);
        else
          topk_split_kernel<false><<<grid, kTopkThreads, 0, st>>>(scores_dev, tab, state, k, cand, bnd, keys);
        ROI3D_LAUNCH_CHECK();
        topk_after_split_kernel<<<1, kMaxSeg, 0, st>>>(state, ns);
        ROI3D_LAUNCH_CHECK();
      }
    }
    if (apply_sigmoid)
      topk_collect_kernel<true><<<grid2, kTopkThreads, 0, st>>>(scores_dev, tab, state, k, cand, bnd, keys);
    else
      topk_collect_kernel<false><<<grid2, kTopkThreads, 0, st>>>(scores_dev, tab, state, k, cand, bnd, keys);
    ROI3D_LAUNCH_CHECK();
    if (k <= kBitonicMax) {
      int npow2 = 2;
      while (npow2 < k) npow2 <<= 1;
      topk_bitonic_kernel<<<ns, 1024, (size_t)npow2 * sizeof(unsigned long long), st>>>(
          cand, state, tab, small_segments_in_index_order, k, npow2, out_idx_dev + (size_t)s0 * k,
          out_val_dev + (size_t)s0 * k);
    } else {
      topk_sort_kernel<<<dim3(ceil_div(k, 256), ns), 256, 0, st>>>(cand, state, tab, small_segments_in_index_order, k,
                                                                  out_idx_dev + (size_t)s0 * k,
                                                                  out_val_dev + (size_t)s0 * k);
    }
    ROI3D_LAUNCH_CHECK();
    if (out_count_dev != nullptr || out_sorted_dev != nullptr) {
      topk_report_kernel<<<1, kMaxSeg, 0, st>>>(state, ns, k, small_segments_in_index_order,
                                                out_count_dev ? out_count_dev + s0 : nullptr,
                                                out_sorted_dev ? out_sorted_dev + s0 : nullptr);
      ROI3D_LAUNCH_CHECK();
    }
  }
  return ROI3D_OK;
}

int roi3d_decode_proposals(const float *bbox_pred_dev, int A, int D, int H, int W, float stride, float depth_stride,
                           const float *base_anchors_host, const int64_t *idx_dev, const float *scores_dev, int n,
                           const float *means6_host, const float *stds6_host, float img_h, float img_w, float img_d,
                           float *out_dev, void *stream) {
  ROI3D_CHECK_ARG(n >= 0, "bad n");
  if (n == 0) return ROI3D_OK;
  ROI3D_CHECK_ARG(bbox_pred_dev && base_anchors_host && idx_dev && out_dev, "NULL pointer");
  ROI3D_CHECK_ARG(A >= 1 && A <= 16, "A=%d out of [1,16]", A);
  ROI3D_CHECK_ARG(D > 0 && H > 0 && W > 0, "bad dims");
  DecodeParams p;
  p.bbox_pred = bbox_pred_dev, p.A = A, p.D = D, p.H = H, p.W = W, p.stride = stride, p.dstride = depth_stride;
  for (int a = 0; a < A; ++a)
    for (int j = 0; j < 6; ++j) p.base[a][j] = base_anchors_host[a * 6 + j];
  for (int j = 0; j < 6; ++j) {
    p.means[j] = means6_host ? means6_host[j] : 0.0f;
    p.stds[j] = stds6_host ? stds6_host[j] : 1.0f;
  }
  p.img_h = img_h, p.img_w = img_w, p.img_d = img_d;
  // max_ratio = np.abs(np.log(16/1000)) evaluated in float64 then used as a python float by clamp
  p.max_ratio = (float)4.135166556742356;
  p.idx = idx_dev, p.scores = scores_dev, p.n = n, p.out = out_dev;
  decode_proposals_kernel<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(p);
  ROI3D_LAUNCH_CHECK();
  return ROI3D_OK;
}

int roi3d_decode_proposals_batched(const float *const *bbox_pred_dev_ptrs, const int32_t *seg_adhw,
                                   const int32_t *seg_level, const float *seg_img_hwd, int nseg, int num_levels,
                                   int A, const float *base_anchors_host, const float *strides_host,
                                   const float *depth_strides_host, const int64_t *idx_dev, const float *scores_dev,
                                   int k, const float *means6_host, const float *stds6_host, float *out_dev,
                                   void *stream) {
  ROI3D_CHECK_ARG(nseg >= 0 && k >= 0, "bad sizes");
  if (nseg == 0 || k == 0) return ROI3D_OK;
  ROI3D_CHECK_ARG(bbox_pred_dev_ptrs && seg_adhw && seg_level && base_anchors_host && strides_host &&
                      depth_strides_host && idx_dev && out_dev,
                  "NULL pointer");
  ROI3D_CHECK_ARG(A >= 1 && A <= 4, "batched decode supports 1..4 base anchors per level, got %d", A);
  ROI3D_CHECK_ARG(num_levels >= 1 && num_levels <= ROI3D_MAX_LEVELS, "num_levels out of range");
  cudaStream_t st = (cudaStream_t)stream;
  for (int s0 = 0; s0 < nseg; s0 += kMaxSeg) {
    const int ns = nseg - s0 < kMaxSeg ? nseg - s0 : kMaxSeg;
    DecodeBatch b;
    for (int s = 0; s < ns; ++s) {
      DecodeSeg &sg = b.seg[s];
      sg.bbox_pred = bbox_pred_dev_ptrs[s0 + s];
      sg.A = seg_adhw[(s0 + s) * 4 + 0], sg.D = seg_adhw[(s0 + s) * 4 + 1];
      sg.H = seg_adhw[(s0 + s) * 4 + 2], sg.W = seg_adhw[(s0 + s) * 4 + 3];
      sg.level = seg_level[s0 + s];
      ROI3D_CHECK_ARG(sg.A == A && sg.level >= 0 && sg.level < num_levels && sg.bbox_pred, "segment %d: bad descriptor", s0 + s);
      sg.img_h = seg_img_hwd ? seg_img_hwd[(s0 + s) * 3 + 0] : 0.0f;
      sg.img_w = seg_img_hwd ? seg_img_hwd[(s0 + s) * 3 + 1] : 0.0f;
      sg.img_d = seg_img_hwd ? seg_img_hwd[(s0 + s) * 3 + 2] : 0.0f;
    }
    for (int l = 0; l < num_levels; ++l) {
      for (int a = 0; a < A; ++a)
        for (int j = 0; j < 6; ++j) b.base[l][a][j] = base_anchors_host[(l * A + a) * 6 + j];
      b.stride[l] = strides_host[l], b.dstride[l] = depth_strides_host[l];
    }
    for (int j = 0; j < 6; ++j) {
      b.means[j] = means6_host ? means6_host[j] : 0.0f;
      b.stds[j] = stds6_host ? stds6_host[j] : 1.0f;
    }
    b.max_ratio = (float)4.135166556742356;
    b.idx = idx_dev + (size_t)s0 * k, b.scores = scores_dev ? scores_dev + (size_t)s0 * k : nullptr;
    b.k = k, b.out = out_dev + (size_t)s0 * k * 7;
    decode_proposals_batched_kernel<<<dim3(ceil_div(k, 256), ns), 256, 0, st>>>(b);
    ROI3D_LAUNCH_CHECK();
  }
  return ROI3D_OK;
}

int roi3d_rpn_collect(const float *dets_dev, int num_images, int num_levels, int k, const int64_t *keep_by_score_dev,
                      const int64_t *keep_by_index_dev, const int32_t *num_keep_dev,
                      const uint8_t *use_index_order_dev, int nms_post, float *cat_props_dev, float *cat_scores_dev,
                      int32_t *n_valid_dev, void *stream) {
  ROI3D_CHECK_ARG(num_images >= 0 && num_levels >= 1 && k >= 1 && nms_post >= 1, "bad sizes");
  if (num_images == 0) return ROI3D_OK;
  ROI3D_CHECK_ARG(dets_dev && keep_by_score_dev && num_keep_dev && cat_props_dev && cat_scores_dev && n_valid_dev,
                  "NULL pointer");
  ROI3D_CHECK_ARG(use_index_order_dev == nullptr || keep_by_index_dev != nullptr, "keep_by_index is NULL");
  ROI3D_CHECK_ARG((long long)num_images * num_levels <= 65535, "too many segments");
  const int P = nms_post < k ? nms_post : k;
  rpn_collect_kernel<<<dim3(ceil_div(P, 256), num_images * num_levels), 256, 0, (cudaStream_t)stream>>>(
      dets_dev, k, keep_by_score_dev, keep_by_index_dev, num_keep_dev, use_index_order_dev, num_levels, P, cat_props_dev,
      cat_scores_dev, n_valid_dev);
  ROI3D_LAUNCH_CHECK();
  return ROI3D_OK;
}

int roi3d_gather_rows7(const float *rows_dev, int nseg, int rows_per_seg, const int64_t *idx_dev, int n, float *out_dev,
                       void *stream) {
  ROI3D_CHECK_ARG(nseg >= 0 && rows_per_seg >= 0 && n >= 0 && nseg <= 65535, "bad sizes");
  if (nseg == 0 || n == 0) return ROI3D_OK;
  ROI3D_CHECK_ARG(rows_dev && idx_dev && out_dev, "NULL pointer");
  gather_rows7_kernel<<<dim3(ceil_div(n, 256), nseg), 256, 0, (cudaStream_t)stream>>>(rows_dev, rows_per_seg, idx_dev, n,
                                                                                      out_dev);
  ROI3D_LAUNCH_CHECK();
  return ROI3D_OK;
}

}  // extern "C"
