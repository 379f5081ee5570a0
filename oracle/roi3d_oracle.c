/*
 * roi3d_oracle.c -- TEST INFRASTRUCTURE ONLY.  CPU restatement of the reference's 3D RoI hot path.
 *
 * Nothing in the product path (3d-multi-resolution-rcnn_b200/) may link, import or call this file.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it,
 * and only as the checker / the CPU baseline.
 *
 * Every function restates one reference function (file:line under /root/reference, cited at each
 * definition).  The reference's kernels are CUDA; nvcc's default -fmad=true contracts some of their
 * mul+add pairs into FFMA.  Which pairs were contracted was read from the SASS of the reference's
 * own .cu files compiled for sm_100a with the reference's (absent) math flags
 * (oracle/build_ref.sh keeps that SASS in the .sass files under oracle/_ref/).  Each function therefore takes
 * `contract`:
 *     contract = 1   the arithmetic the COMPILED reference kernel executes (explicit fmaf where the
 *                    SASS shows FFMA) -- this is the parity target for the CUDA product path;
 *     contract = 0   the source-literal arithmetic with no contraction (what a numpy restatement
 *                    of the .cu text would compute).
 * Build with -ffp-contract=off so gcc adds no contraction of its own (oracle/Makefile does).
 *
 * Pinning: see oracle/README.md -- the four IoU known answers of
 * mmdet/core/bbox/geometry.py:81-102 are checked in tests/test_oracle_golden.py, and the GPU tests
 * compare this file against the reference's own kernels (oracle/_ref) on the B200 box; fixtures
 * generated there are committed under tests/golden/.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define API __attribute__((visibility("default")))

API int oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

API void oracle_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* ------------------------------------------------------------------------------------------ */
/* 3D IoU: mmdet/ops/nms/src/nms_kernel.cu:23-33 (devIoU3d).  a = row box, b = column box.      */
/* SASS (oracle/_ref/ref_nms_kernel.sass, nms_kernel_3d): Sa and interS are plain FMULs, the    */
/* union is FFMA(bw*bh, bd, Sa) followed by FADD(-interS); the division is IEEE (div.rn).       */
/* ------------------------------------------------------------------------------------------ */
API float oracle_iou3d(const float *a, const float *b, int contract) {
  float left = fmaxf(a[0], b[0]), right = fminf(a[2], b[2]);
  float top = fmaxf(a[1], b[1]), bottom = fminf(a[3], b[3]);
  float front = fmaxf(a[4], b[4]), back = fminf(a[5], b[5]);
  float width = fmaxf(right - left + 1.0f, 0.0f);
  float height = fmaxf(bottom - top + 1.0f, 0.0f);
  float depth = fmaxf(back - front + 1.0f, 0.0f);
  float interS = width * height * depth;
  float Sa = (a[2] - a[0] + 1.0f) * (a[3] - a[1] + 1.0f) * (a[5] - a[4] + 1.0f);
  float sbxy = (b[2] - b[0] + 1.0f) * (b[3] - b[1] + 1.0f);
  float sbz = b[5] - b[4] + 1.0f;
  float uni;
  if (contract) {
    uni = fmaf(sbxy, sbz, Sa) - interS;
  } else {
    float Sb = sbxy * sbz;
    uni = Sa + Sb - interS;
  }
  return interS / uni;
}

/* Stable descending order: score desc, ties by lower original index.  The reference sorts with  */
/* scores.sort(0, descending=true) (nms_kernel.cu:199-200), which is unstable on ties; the build */
/* rule is the stable order below.                                                    */
typedef struct {
  float s;
  int64_t i;
} score_idx_t;

static int cmp_desc_stable(const void *pa, const void *pb) {
  const score_idx_t *a = (const score_idx_t *)pa, *b = (const score_idx_t *)pb;
  if (a->s > b->s) return -1;
  if (a->s < b->s) return 1;
  return (a->i > b->i) - (a->i < b->i);
}

static int cmp_i64(const void *pa, const void *pb) {
  int64_t a = *(const int64_t *)pa, b = *(const int64_t *)pb;
  return (a > b) - (a < b);
}

API void oracle_argsort_desc_stable(const float *scores, int64_t n, int64_t stride, int64_t *order) {
  score_idx_t *v = (score_idx_t *)malloc(sizeof(score_idx_t) * (size_t)(n > 0 ? n : 1));
  for (int64_t i = 0; i < n; i++) {
    v[i].s = scores[i * stride];
    v[i].i = i;
  }
  qsort(v, (size_t)n, sizeof(score_idx_t), cmp_desc_stable);
  for (int64_t i = 0; i < n; i++) order[i] = v[i].i;
  free(v);
}

/* ------------------------------------------------------------------------------------------ */
/* 3D NMS: nms_cuda_3d, mmdet/ops/nms/src/nms_kernel.cu:196-257 + nms_kernel_3d :81-129.         */
/*   sort by column 6 descending (:199-201); bit (i,j), j>i, set iff IoU(box_i, box_j) > thr     */
/*   (:112-128, row box is `a`); greedy sweep in sorted order (:238-249); kept sorted positions  */
/*   mapped back through the order and returned ASCENDING by original index (:253-256).          */
/* dets: [n,7] = x1,y1,x2,y2,z1,z2,score.  keep: capacity n.  Returns the number kept.          */
/* If order_out != NULL it receives the kept ORIGINAL indices in score order as well.            */
/* ------------------------------------------------------------------------------------------ */
API int64_t oracle_nms3d(const float *dets, int64_t n, float thr, int64_t *keep, int64_t *keep_score_order,
                         int contract) {
  if (n <= 0) return 0;
  int64_t *order = (int64_t *)malloc(sizeof(int64_t) * (size_t)n);
  unsigned char *removed = (unsigned char *)calloc((size_t)n, 1);
  oracle_argsort_desc_stable(dets + 6, n, 7, order);
  int64_t m = 0;
  for (int64_t i = 0; i < n; i++) {
    if (removed[i]) continue;
    const float *a = dets + order[i] * 7;
    keep[m] = order[i];
    if (keep_score_order) keep_score_order[m] = order[i];
    m++;
    for (int64_t j = i + 1; j < n; j++) {
      if (removed[j]) continue; /* does not change the result: a removed box never suppresses */
      if (oracle_iou3d(a, dets + order[j] * 7, contract) > thr) removed[j] = 1;
    }
  }
  qsort(keep, (size_t)m, sizeof(int64_t), cmp_i64);
  free(order);
  free(removed);
  return m;
}

/* Full pairwise suppression matrix in SORTED order, as nms_kernel_3d writes it: word            */
/* mask[i*col_blocks + c] has bit b set iff j = 64c+b > i (within the diagonal tile; every j in  */
/* off-diagonal tiles, including j < i -- the reference computes the full matrix, :86) and       */
/* IoU(sorted_i, sorted_j) > thr.  Used to check the product's mask kernel word for word on the  */
/* upper triangle.                                                                                */
API void oracle_nms3d_mask(const float *sorted_boxes, int64_t n, float thr, uint64_t *mask, int contract) {
  int64_t cb = (n + 63) / 64;
  for (int64_t i = 0; i < n; i++) {
    for (int64_t c = 0; c < cb; c++) {
      uint64_t t = 0;
      int64_t start = (c == i / 64) ? (i % 64) + 1 : 0;
      int64_t cs = n - c * 64 < 64 ? n - c * 64 : 64;
      for (int64_t b = start; b < cs; b++) {
        if (oracle_iou3d(sorted_boxes + i * 7, sorted_boxes + (c * 64 + b) * 7, contract) > thr) t |= 1ULL << b;
      }
      mask[i * cb + c] = t;
    }
  }
}

/* ------------------------------------------------------------------------------------------ */
/* RoIAlign 3D geometry shared by forward and backward.                                         */
/* Reference: ROIAlignForward3D, mmdet/ops/roi_align/src/roi_align_kernel.cu:214-291 and         */
/* ROIAlignBackward3D :519-636.  Per axis (w,h use spatial_scale; d uses spatial_scale_depth):   */
/*   start = c1*s (:233-238); end = (c2+1)*s; size = max(end-start, 0) (:241-243; no min-1       */
/*   clamp); bin = size/P (:245-247); S = sample_num>0 ? sample_num : ceil(size/P) (:252-259);   */
/*   coord(p,i) = start + p*bin + (i+.5)*bin/S (:271-281).                                        */
/* SASS (ROIAlignForward3D<float> and Backward3D<float>): size = max(FFMA(c2+1, s, -start), 0);   */
/* start + p*bin = FFMA((float)p, bin, start); (i+.5)*bin is an FMUL, the division by S is IEEE,  */
/* the final sum is an FADD.                                                                      */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
  float start, bin;
  int S;
} axis_t;

static axis_t axis_setup(float c1, float c2, float scale, int P, int sample_num, int contract) {
  axis_t a;
  a.start = c1 * scale;
  float size;
  if (contract) {
    size = fmaf(c2 + 1.0f, scale, -a.start);
  } else {
    float end = (c2 + 1.0f) * scale;
    size = end - a.start;
  }
  size = fmaxf(size, 0.0f);
  a.bin = size / (float)P;
  a.S = sample_num > 0 ? sample_num : (int)ceilf(a.bin);
  return a;
}

static inline float axis_coord(const axis_t *a, int p, int i, int contract) {
  float base;
  if (contract) {
    base = fmaf((float)p, a->bin, a->start);
  } else {
    base = a->start + (float)p * a->bin;
  }
  return base + ((float)i + 0.5f) * a->bin / (float)a->S;
}

/* One axis of bilinear_interpolate_3d / bilinear_interpolate_gradient_3d                        */
/* (roi_align_kernel.cu:64-110 and :383-438): returns 0 when the coordinate is outside [-1,size]; */
/* clamps <=0 to 0; low=(int)c; if low >= size-1: high=low=size-1 and c=low; l=c-low; h=1-l.      */
typedef struct {
  int valid, low, high;
  float l, h;
} tap_t;

static inline tap_t axis_tap(float c, int size) {
  tap_t t;
  t.valid = !(c < -1.0f || c > (float)size);
  if (c <= 0.0f) c = 0.0f;
  t.low = (int)c;
  if (t.low >= size - 1) {
    t.high = t.low = size - 1;
    c = (float)t.low;
  } else {
    t.high = t.low + 1;
  }
  t.l = c - (float)t.low;
  t.h = 1.0f - t.l;
  return t;
}

/* The eight corner weights, in the reference's order w1..w8 (roi_align_kernel.cu:134-137,        */
/* :431-434): w1=hx*hy*hz w2=lx*hy*hz w3=hx*ly*hz w4=lx*ly*hz w5=hx*hy*lz w6=lx*hy*lz            */
/* w7=hx*ly*lz w8=lx*ly*lz, each evaluated (x*y)*z.                                               */
static inline void corner_weights(const tap_t *tz, const tap_t *ty, const tap_t *tx, float w[8]) {
  w[0] = tx->h * ty->h * tz->h;
  w[1] = tx->l * ty->h * tz->h;
  w[2] = tx->h * ty->l * tz->h;
  w[3] = tx->l * ty->l * tz->h;
  w[4] = tx->h * ty->h * tz->l;
  w[5] = tx->l * ty->h * tz->l;
  w[6] = tx->h * ty->l * tz->l;
  w[7] = tx->l * ty->l * tz->l;
}

/* ------------------------------------------------------------------------------------------ */
/* Forward.  feats: NCDHW contiguous [B,C,D,H,W]; rois [K,7] = (batch, x1,y1,x2,y2,z1,z2);       */
/* out [K,C,PD,PH,PW].  roi_align_kernel.cu:214-291 + :64-149.                                   */
/* SASS of the 8-term sum (:139-146): t = w2*f2; then FFMA(w1,f1,t), w3, w4, w5, w6, w7, w8.     */
/* ------------------------------------------------------------------------------------------ */
API void oracle_roi_align3d_forward(const float *feats, int B, int C, int D, int H, int W, const float *rois,
                                    int K, int PD, int PH, int PW, float scale, float scale_d, int sample_num,
                                    float *out, int contract) {
  (void)B;
  const int64_t plane = (int64_t)D * H * W;
  const int64_t total = (int64_t)K * C;
#pragma omp parallel for schedule(dynamic, 8)
  for (int64_t kc = 0; kc < total; kc++) {
    int k = (int)(kc / C), c = (int)(kc % C);
    const float *r = rois + (int64_t)k * 7;
    int b = (int)r[0];
    axis_t aw = axis_setup(r[1], r[3], scale, PW, sample_num, contract);
    axis_t ah = axis_setup(r[2], r[4], scale, PH, sample_num, contract);
    axis_t ad = axis_setup(r[5], r[6], scale_d, PD, sample_num, contract);
    const float *f = feats + ((int64_t)b * C + c) * plane;
    float *o = out + kc * ((int64_t)PD * PH * PW);
    const float count = (float)(ad.S * ah.S * aw.S);
    for (int pd = 0; pd < PD; pd++)
      for (int ph = 0; ph < PH; ph++)
        for (int pw = 0; pw < PW; pw++) {
          float acc = 0.0f;
          for (int iz = 0; iz < ad.S; iz++) {
            tap_t tz = axis_tap(axis_coord(&ad, pd, iz, contract), D);
            for (int iy = 0; iy < ah.S; iy++) {
              tap_t ty = axis_tap(axis_coord(&ah, ph, iy, contract), H);
              for (int ix = 0; ix < aw.S; ix++) {
                tap_t tx = axis_tap(axis_coord(&aw, pw, ix, contract), W);
                float val = 0.0f;
                if (tz.valid && ty.valid && tx.valid) {
                  float w[8];
                  corner_weights(&tz, &ty, &tx, w);
                  const float f1 = f[tx.low + (int64_t)W * (ty.low + (int64_t)H * tz.low)];
                  const float f2 = f[tx.high + (int64_t)W * (ty.low + (int64_t)H * tz.low)];
                  const float f3 = f[tx.low + (int64_t)W * (ty.high + (int64_t)H * tz.low)];
                  const float f4 = f[tx.high + (int64_t)W * (ty.high + (int64_t)H * tz.low)];
                  const float f5 = f[tx.low + (int64_t)W * (ty.low + (int64_t)H * tz.high)];
                  const float f6 = f[tx.high + (int64_t)W * (ty.low + (int64_t)H * tz.high)];
                  const float f7 = f[tx.low + (int64_t)W * (ty.high + (int64_t)H * tz.high)];
                  const float f8 = f[tx.high + (int64_t)W * (ty.high + (int64_t)H * tz.high)];
                  if (contract) {
                    float t = w[1] * f2;
                    t = fmaf(w[0], f1, t);
                    t = fmaf(w[2], f3, t);
                    t = fmaf(w[3], f4, t);
                    t = fmaf(w[4], f5, t);
                    t = fmaf(w[5], f6, t);
                    t = fmaf(w[6], f7, t);
                    val = fmaf(w[7], f8, t);
                  } else {
                    val = w[0] * f1 + w[1] * f2 + w[2] * f3 + w[3] * f4 + w[4] * f5 + w[5] * f6 + w[6] * f7 +
                          w[7] * f8;
                  }
                }
                acc += val;
              }
            }
          }
          o[((int64_t)pd * PH + ph) * PW + pw] = acc / count;
        }
  }
}

/* ------------------------------------------------------------------------------------------ */
/* Backward.  roi_align_kernel.cu:519-636 + :383-442.  g_k = top_diff*w_k/count (:600-608),       */
/* scattered to the 8 corners when the sample is in range (:610-619).                             */
/* bug_compat = 1 reproduces the reference's top_diff index                                       */
/*   pd*pooled_depth*pooled_width + ph*pooled_width + pw      (:554-555)              */
/* which equals the correct pd*PH*PW + ph*PW + pw only when PD == PH.                             */
/* The reference accumulates with fp32 atomicAdd in nondeterministic order; the oracle            */
/* accumulates the fp32 products g_k in float64 and rounds once, i.e. it is the limit every       */
/* fp32 ordering is within a few ulp of.  grad_in: NCDHW [B,C,D,H,W], overwritten.                */
/* ------------------------------------------------------------------------------------------ */
API void oracle_roi_align3d_backward(const float *grad_out, const float *rois, int K, int PD, int PH, int PW,
                                     float scale, float scale_d, int sample_num, float *grad_in, int B, int C,
                                     int D, int H, int W, int bug_compat, int contract) {
  const int64_t plane = (int64_t)D * H * W;
  const int64_t bins = (int64_t)PD * PH * PW;
  /* parallel over channels: every (b,c) plane is owned by one thread -> no races, fixed order */
#pragma omp parallel
  {
    double *acc = (double *)malloc(sizeof(double) * (size_t)plane * (size_t)(B > 0 ? B : 1));
#pragma omp for schedule(dynamic, 1)
    for (int c = 0; c < C; c++) {
      memset(acc, 0, sizeof(double) * (size_t)plane * (size_t)B);
      for (int k = 0; k < K; k++) {
        const float *r = rois + (int64_t)k * 7;
        int b = (int)r[0];
        axis_t aw = axis_setup(r[1], r[3], scale, PW, sample_num, contract);
        axis_t ah = axis_setup(r[2], r[4], scale, PH, sample_num, contract);
        axis_t ad = axis_setup(r[5], r[6], scale_d, PD, sample_num, contract);
        const float count = (float)(ad.S * ah.S * aw.S);
        const float *go = grad_out + ((int64_t)k * C + c) * bins;
        double *g = acc + (int64_t)b * plane;
        for (int pd = 0; pd < PD; pd++)
          for (int ph = 0; ph < PH; ph++)
            for (int pw = 0; pw < PW; pw++) {
              int64_t off = bug_compat ? ((int64_t)pd * PD * PW + (int64_t)ph * PW + pw)
                                       : (((int64_t)pd * PH + ph) * PW + pw);
              /* bug_compat can index past this (k,c) block exactly as the reference does; stay   */
              /* inside the whole tensor so the oracle never reads out of bounds.                  */
              int64_t abs_off = ((int64_t)k * C + c) * bins + off;
              if (abs_off >= (int64_t)K * C * bins) continue;
              const float top = go[off];
              for (int iz = 0; iz < ad.S; iz++) {
                tap_t tz = axis_tap(axis_coord(&ad, pd, iz, contract), D);
                for (int iy = 0; iy < ah.S; iy++) {
                  tap_t ty = axis_tap(axis_coord(&ah, ph, iy, contract), H);
                  for (int ix = 0; ix < aw.S; ix++) {
                    tap_t tx = axis_tap(axis_coord(&aw, pw, ix, contract), W);
                    if (!(tz.valid && ty.valid && tx.valid)) continue;
                    float w[8];
                    corner_weights(&tz, &ty, &tx, w);
                    g[((int64_t)tz.low * H + ty.low) * W + tx.low] += (double)(top * w[0] / count);
                    g[((int64_t)tz.low * H + ty.low) * W + tx.high] += (double)(top * w[1] / count);
                    g[((int64_t)tz.low * H + ty.high) * W + tx.low] += (double)(top * w[2] / count);
                    g[((int64_t)tz.low * H + ty.high) * W + tx.high] += (double)(top * w[3] / count);
                    g[((int64_t)tz.high * H + ty.low) * W + tx.low] += (double)(top * w[4] / count);
                    g[((int64_t)tz.high * H + ty.low) * W + tx.high] += (double)(top * w[5] / count);
                    g[((int64_t)tz.high * H + ty.high) * W + tx.low] += (double)(top * w[6] / count);
                    g[((int64_t)tz.high * H + ty.high) * W + tx.high] += (double)(top * w[7] / count);
                  }
                }
              }
            }
      }
      for (int b = 0; b < B; b++) {
        float *dst = grad_in + ((int64_t)b * C + c) * plane;
        const double *src = acc + (int64_t)b * plane;
        for (int64_t i = 0; i < plane; i++) dst[i] = (float)src[i];
      }
    }
    free(acc);
  }
}

/* Number of distinct feature voxels (b,z,y,x) the RoI set touches with a non-skipped sample --   */
/* the `U` of SURVEY 8(d)'s algorithmic-bytes formula.  touched: optional [B*D*H*W] byte map.     */
API int64_t oracle_roi_align3d_unique_voxels(const float *rois, int K, int B, int D, int H, int W, int PD,
                                             int PH, int PW, float scale, float scale_d, int sample_num,
                                             unsigned char *touched_out) {
  const int64_t plane = (int64_t)D * H * W;
  unsigned char *touched = touched_out ? touched_out : (unsigned char *)calloc((size_t)(plane * B), 1);
  if (touched_out) memset(touched, 0, (size_t)(plane * B));
  for (int k = 0; k < K; k++) {
    const float *r = rois + (int64_t)k * 7;
    int b = (int)r[0];
    axis_t aw = axis_setup(r[1], r[3], scale, PW, sample_num, 1);
    axis_t ah = axis_setup(r[2], r[4], scale, PH, sample_num, 1);
    axis_t ad = axis_setup(r[5], r[6], scale_d, PD, sample_num, 1);
    /* per-axis touched sets are a product set: mark per axis, then take the outer product */
    unsigned char *mz = (unsigned char *)calloc((size_t)D, 1), *my = (unsigned char *)calloc((size_t)H, 1),
                  *mx = (unsigned char *)calloc((size_t)W, 1);
    for (int p = 0; p < PD; p++)
      for (int i = 0; i < ad.S; i++) {
        tap_t t = axis_tap(axis_coord(&ad, p, i, 1), D);
        if (t.valid) mz[t.low] = mz[t.high] = 1;
      }
    for (int p = 0; p < PH; p++)
      for (int i = 0; i < ah.S; i++) {
        tap_t t = axis_tap(axis_coord(&ah, p, i, 1), H);
        if (t.valid) my[t.low] = my[t.high] = 1;
      }
    for (int p = 0; p < PW; p++)
      for (int i = 0; i < aw.S; i++) {
        tap_t t = axis_tap(axis_coord(&aw, p, i, 1), W);
        if (t.valid) mx[t.low] = mx[t.high] = 1;
      }
    for (int z = 0; z < D; z++)
      if (mz[z])
        for (int y = 0; y < H; y++)
          if (my[y])
            for (int x = 0; x < W; x++)
              if (mx[x]) touched[(int64_t)b * plane + ((int64_t)z * H + y) * W + x] = 1;
    free(mz);
    free(my);
    free(mx);
  }
  int64_t u = 0;
  for (int64_t i = 0; i < plane * B; i++) u += touched[i];
  if (!touched_out) free(touched);
  return u;
}

/* ------------------------------------------------------------------------------------------ */
/* FPN level mapping: SingleRoIExtractor.map_roi_levels,                                         */
/* mmdet/models/roi_extractors/single_level.py:58-82:                                            */
/*   scale = sqrt((x2-x1+1)*(y2-y1+1)*(z2-z1+1)); lvl = floor(log2(scale/finest + 1e-6))         */
/*   clamped to [0, L-1].  The reference evaluates this with torch CUDA elementwise ops; torch's  */
/*   CUDA `tensor / python_scalar` multiplies by the fp32 reciprocal (THC TensorDivConstantOp     */
/*   <float>, and BinaryDivTrueKernel.cu today), its CPU path divides.  recip_div selects which.  */
/*   Parity of sqrt/log2 last-bit behaviour is UNPINNED by the reference; levels can  */
/*   only differ when scale/finest+1e-6 is within an ulp of a power of two.                       */
/* ------------------------------------------------------------------------------------------ */
API void oracle_map_roi_levels(const float *rois, int64_t K, int num_levels, float finest_scale, int recip_div,
                               int64_t *lvls) {
  const float inv = 1.0f / finest_scale;
  for (int64_t k = 0; k < K; k++) {
    const float *r = rois + k * 7;
    float vol = (r[3] - r[1] + 1.0f) * (r[4] - r[2] + 1.0f) * (r[6] - r[5] + 1.0f);
    float s = sqrtf(vol);
    float q = recip_div ? s * inv : s / finest_scale;
    float t = floorf(log2f(q + 1e-6f));
    /* clamp(min=0, max=L-1) on the float, then .long(); NaN (negative volume) -> torch gives an */
    /* implementation-defined integer; the oracle maps NaN to level 0.                            */
    if (!(t > 0.0f)) t = 0.0f;
    if (t > (float)(num_levels - 1)) t = (float)(num_levels - 1);
    lvls[k] = (int64_t)t;
  }
}

/* ------------------------------------------------------------------------------------------ */
/* delta2bbox3D: mmdet/core/bbox/transforms.py:105-160, one box per row (deltas [n,6]).          */
/*   denorm d*std+mean (:112-114); order dx,dy,dw,dh,dz,dd (:115-120); clamp dw,dh,dz,dd to       */
/*   +-|log(16/1000)| (:122-128; note the centre delta dz is clamped too); p* from the anchor     */
/*   with +1 sizes (:130-135); g = p*exp(d), gx = px + pw*dx (addcmul, :137-142); corners         */
/*   g -+ .5*size +- .5 (:144-149); clamp to max_shape = img_shape (H, W, 3, D): x in [0,W-1],    */
/*   y in [0,H-1], z in [0, max_shape[3]-1] (:151-157).                                           */
/* torch evaluates each line as a separate fp32 kernel, so there is no contraction except         */
/* addcmul, whose CUDA kernel computes  px + 1*(pw*dx)  in fp32 (no fma guarantee) -- the oracle   */
/* uses plain mul+add.  exp is expf.                                                              */
/* ------------------------------------------------------------------------------------------ */
API void oracle_delta2bbox3d(const float *anchors, const float *deltas, int64_t n, const float *means,
                             const float *stds, int has_max_shape, float max_h, float max_w, float max_d,
                             float *out) {
  const float max_ratio = (float)fabs(log(16.0 / 1000.0));
  for (int64_t i = 0; i < n; i++) {
    const float *a = anchors + i * 6, *d = deltas + i * 6;
    float dx = d[0] * stds[0] + means[0], dy = d[1] * stds[1] + means[1];
    float dw = d[2] * stds[2] + means[2], dh = d[3] * stds[3] + means[3];
    float dz = d[4] * stds[4] + means[4], dd = d[5] * stds[5] + means[5];
    dw = fminf(fmaxf(dw, -max_ratio), max_ratio);
    dh = fminf(fmaxf(dh, -max_ratio), max_ratio);
    dz = fminf(fmaxf(dz, -max_ratio), max_ratio);
    dd = fminf(fmaxf(dd, -max_ratio), max_ratio);
    float px = (a[0] + a[2]) * 0.5f, py = (a[1] + a[3]) * 0.5f, pz = (a[4] + a[5]) * 0.5f;
    float pw = a[2] - a[0] + 1.0f, ph = a[3] - a[1] + 1.0f, pd = a[5] - a[4] + 1.0f;
    float gw = pw * expf(dw), gh = ph * expf(dh), gd = pd * expf(dd);
    float gx = px + pw * dx, gy = py + ph * dy, gz = pz + pd * dz;
    float x1 = gx - gw * 0.5f + 0.5f, y1 = gy - gh * 0.5f + 0.5f;
    float x2 = gx + gw * 0.5f - 0.5f, y2 = gy + gh * 0.5f - 0.5f;
    float z1 = gz - gd * 0.5f + 0.5f, z2 = gz + gd * 0.5f - 0.5f;
    if (has_max_shape) {
      x1 = fminf(fmaxf(x1, 0.0f), max_w - 1.0f);
      y1 = fminf(fmaxf(y1, 0.0f), max_h - 1.0f);
      x2 = fminf(fmaxf(x2, 0.0f), max_w - 1.0f);
      y2 = fminf(fmaxf(y2, 0.0f), max_h - 1.0f);
      z1 = fminf(fmaxf(z1, 0.0f), max_d - 1.0f);
      z2 = fminf(fmaxf(z2, 0.0f), max_d - 1.0f);
    }
    float *o = out + i * 6;
    o[0] = x1, o[1] = y1, o[2] = x2, o[3] = y2, o[4] = z1, o[5] = z2;
  }
}

/* ------------------------------------------------------------------------------------------ */
/* top-k: torch.topk(scores, k) as used at mmdet/models/anchor_heads/rpn_head_3d.py:109,147.      */
/* Tie order is unspecified by torch (parity unpinned); the build rule is: the k       */
/* largest, descending, ties broken by LOWER index.                                               */
/* ------------------------------------------------------------------------------------------ */
API void oracle_topk_desc_stable(const float *scores, int64_t n, int64_t k, int64_t *idx_out) {
  int64_t *order = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n > 0 ? n : 1));
  oracle_argsort_desc_stable(scores, n, 1, order);
  for (int64_t i = 0; i < k && i < n; i++) idx_out[i] = order[i];
  free(order);
}

/* sigmoid as torch computes it in fp32: 1/(1+exp(-x)) (rpn_head_3d.py:90) */
API void oracle_sigmoid(const float *x, int64_t n, float *y) {
  for (int64_t i = 0; i < n; i++) y[i] = 1.0f / (1.0f + expf(-x[i]));
}
