/*
 * tlod_oracle.c -- CPU restatement of the reference's RoI / proposal hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product package may import, link
 * or call this file; it exists so that tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline leg have an independent checker.
 *
 * Every function restates the arithmetic of one reference routine (cited as
 * /root/reference file:line).  The loops are reorganised (per-RoI geometry is
 * hoisted, images are independent) but each output element sees exactly the
 * same sequence of IEEE operations as the reference source, with FMA
 * contraction disabled (-ffp-contract=off): that is the "source-level fp32"
 * contract SURVEY.md section 7 pins for NMS / IoU / decode.
 *
 * Parity pinning: oracle/validate_against_reference.py runs the reference's own
 * C (roi_align.c via oracle/_ref/libref_cpu.so) and Python layers
 * (lib/model/rpn/*.py) in this container and asserts bit-equality with the
 * functions below; tests/golden holds the vectors it generated.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------ */
/* RoIAlign geometry shared by forward and backward.                   */
/* lib/model/roi_align/src/roi_align_kernel.cu:31-49 (same text in     */
/* roi_align.c:97-116).                                                */
/* ------------------------------------------------------------------ */
typedef struct {
  int start;   /* min(floor(p), size-2) */
  int valid;   /* !(p < 0 || p >= size) */
  float ratio; /* p - start             */
} orc_axis_t;

static void orc_align_axis(float roi_lo_px, float roi_hi_px, float scale, int aligned, int size,
                           orc_axis_t* out) {
  float lo = roi_lo_px * scale;
  float hi = roi_hi_px * scale;
  /* `hi - lo + 1.` : float subtract, then double add, rounded back by fmaxf() */
  float extent = fmaxf((float)((double)(hi - lo) + 1.), 0.f);
  /* `extent / (aligned - 1.)` : double divide stored to float */
  float bin = (float)((double)extent / ((double)aligned - 1.));
  for (int p = 0; p < aligned; ++p) {
    float pos = (float)p * bin + lo; /* two roundings: contraction is off */
    int start = (int)fminf((float)floor((double)pos), (float)(size - 2));
    out[p].start = start;
    out[p].valid = !(pos < 0 || pos >= size);
    out[p].ratio = pos - (float)start;
  }
}

/* roi_align_kernel.cu:15-70  (ROIAlignForward) */
void orc_roi_align_fwd(const float* bottom, float scale, int num_rois, int H, int W, int C, int AH,
                       int AW, const float* rois, float* top) {
#pragma omp parallel for schedule(dynamic, 1)
  for (int n = 0; n < num_rois; ++n) {
    orc_axis_t* ay = (orc_axis_t*)malloc(sizeof(orc_axis_t) * (size_t)AH);
    orc_axis_t* ax = (orc_axis_t*)malloc(sizeof(orc_axis_t) * (size_t)AW);
    const float* r = rois + (size_t)n * 5;
    orc_align_axis(r[1], r[3], scale, AW, W, ax);
    orc_align_axis(r[2], r[4], scale, AH, H, ay);
    /* `int img_start = roi_batch_ind * channels * height * width` is a float product */
    int img_start = (int)(r[0] * C * H * W);
    for (int c = 0; c < C; ++c) {
      const float* plane = bottom + img_start + (size_t)c * H * W;
      float* o = top + (((size_t)n * C + c) * AH) * AW;
      for (int ph = 0; ph < AH; ++ph) {
        for (int pw = 0; pw < AW; ++pw) {
          if (!(ay[ph].valid && ax[pw].valid)) {
            o[ph * AW + pw] = 0.f;
            continue;
          }
          /* C typing of :62-65: `d * (1. - h_ratio)` is double, but `d * h_ratio` is a
           * float product (both operands float) that is only then widened. */
          float hrf = ay[ph].ratio, wrf = ax[pw].ratio;
          double hr = hrf, wr = wrf;
          const float* ul = plane + ay[ph].start * W + ax[pw].start;
          double v = (double)ul[0] * (1. - hr) * (1. - wr) + (double)ul[1] * (1. - hr) * wr +
                     (double)(ul[W] * hrf) * (1. - wr) + (double)(ul[W + 1] * hrf * wrf);
          o[ph * AW + pw] = (float)v;
        }
      }
    }
    free(ay);
    free(ax);
  }
}

/* roi_align_kernel.cu:94-143 (ROIAlignBackward).  The reference's CPU backward
 * (roi_align.c:138-190) has an inverted bounds test and is NOT followed.
 * Contributions are added in output-index order, each rounded to float before
 * the add exactly as atomicAdd(float*, float) receives it.  With
 * accumulate_double != 0 the sums are carried in double (used to bound the
 * order-dependent fp32 error in the tolerance tests). */
void orc_roi_align_bwd(const float* top_diff, float scale, int batch, int num_rois, int H, int W,
                       int C, int AH, int AW, const float* rois, float* bottom_diff,
                       int accumulate_double) {
  size_t total = (size_t)batch * C * H * W;
  double* acc = NULL;
  if (accumulate_double) acc = (double*)calloc(total, sizeof(double));
  memset(bottom_diff, 0, total * sizeof(float));
  orc_axis_t* ay = (orc_axis_t*)malloc(sizeof(orc_axis_t) * (size_t)AH);
  orc_axis_t* ax = (orc_axis_t*)malloc(sizeof(orc_axis_t) * (size_t)AW);
  for (int n = 0; n < num_rois; ++n) {
    const float* r = rois + (size_t)n * 5;
    orc_align_axis(r[1], r[3], scale, AW, W, ax);
    orc_align_axis(r[2], r[4], scale, AH, H, ay);
    int img_start = (int)(r[0] * C * H * W);
    for (int c = 0; c < C; ++c) {
      size_t pbase = (size_t)img_start + (size_t)c * H * W;
      const float* g = top_diff + (((size_t)n * C + c) * AH) * AW;
      for (int ph = 0; ph < AH; ++ph) {
        for (int pw = 0; pw < AW; ++pw) {
          if (!(ay[ph].valid && ax[pw].valid)) continue;
          float hrf = ay[ph].ratio, wrf = ax[pw].ratio;
          double hr = hrf;
          double t = g[ph * AW + pw];
          size_t ul = pbase + (size_t)ay[ph].start * W + ax[pw].start;
          /* `(1 - w_ratio)` is an int minus float, i.e. float; `(1. - h_ratio)` is double */
          float one_minus_wr = 1 - wrf;
          float v00 = (float)(t * (1. - hr) * (double)one_minus_wr);
          float v01 = (float)(t * (1. - hr) * (double)wrf);
          float tf = g[ph * AW + pw];
          float v10 = tf * hrf * one_minus_wr; /* all-float products, :133-134 */
          float v11 = tf * hrf * wrf;
          if (acc) {
            acc[ul] += v00;
            acc[ul + 1] += v01;
            acc[ul + W] += v10;
            acc[ul + W + 1] += v11;
          } else {
            bottom_diff[ul] += v00;
            bottom_diff[ul + 1] += v01;
            bottom_diff[ul + W] += v10;
            bottom_diff[ul + W + 1] += v11;
          }
        }
      }
    }
  }
  if (acc) {
    for (size_t i = 0; i < total; ++i) bottom_diff[i] = (float)acc[i];
    free(acc);
  }
  free(ay);
  free(ax);
}

/* ------------------------------------------------------------------ */
/* RoIPool.  lib/model/roi_pooling/src/roi_pooling_kernel.cu            */
/* ------------------------------------------------------------------ */
typedef struct {
  int batch, sw, sh, ew, eh; /* rounded RoI corners in feature cells */
  float bin_h, bin_w;
} orc_pool_roi_t;

/* roi_pooling_kernel.cu:44-55 */
static orc_pool_roi_t orc_pool_roi(const float* r, float scale, int PH, int PW) {
  orc_pool_roi_t g;
  g.batch = (int)r[0];
  g.sw = (int)round((double)(r[1] * scale));
  g.sh = (int)round((double)(r[2] * scale));
  g.ew = (int)round((double)(r[3] * scale));
  g.eh = (int)round((double)(r[4] * scale));
  int rw = (int)fmaxf((float)(g.ew - g.sw + 1), 1.f);
  int rh = (int)fmaxf((float)(g.eh - g.sh + 1), 1.f);
  g.bin_h = (float)rh / (float)PH;
  g.bin_w = (float)rw / (float)PW;
  return g;
}

static int orc_clampi(int v, int lo, int hi) { return (int)fminf(fmaxf((float)v, (float)lo), (float)hi); }

/* roi_pooling_kernel.cu:24-93 (ROIPoolForward) */
void orc_roi_pool_fwd(const float* bottom, float scale, int num_rois, int H, int W, int C, int PH,
                      int PW, const float* rois, float* top, int* argmax) {
#pragma omp parallel for schedule(dynamic, 1)
  for (int n = 0; n < num_rois; ++n) {
    orc_pool_roi_t g = orc_pool_roi(rois + (size_t)n * 5, scale, PH, PW);
    for (int c = 0; c < C; ++c) {
      int plane_off = g.batch * C * H * W + c * H * W;
      for (int ph = 0; ph < PH; ++ph) {
        int hs = (int)floor((double)((float)ph * g.bin_h));
        int he = (int)ceil((double)((float)(ph + 1) * g.bin_h));
        hs = orc_clampi(hs + g.sh, 0, H);
        he = orc_clampi(he + g.sh, 0, H);
        for (int pw = 0; pw < PW; ++pw) {
          int ws = (int)floor((double)((float)pw * g.bin_w));
          int we = (int)ceil((double)((float)(pw + 1) * g.bin_w));
          ws = orc_clampi(ws + g.sw, 0, W);
          we = orc_clampi(we + g.sw, 0, W);
          int empty = (he <= hs) || (we <= ws);
          float best = empty ? 0.f : -FLT_MAX;
          int besti = -1;
          for (int h = hs; h < he; ++h)
            for (int w = ws; w < we; ++w) {
              float v = bottom[plane_off + h * W + w];
              if (v > best) {
                best = v;
                besti = plane_off + h * W + w;
              }
            }
          size_t o = (((size_t)n * C + c) * PH + ph) * PW + pw;
          top[o] = best;
          if (argmax) argmax[o] = besti;
        }
      }
    }
  }
}

/* roi_pooling_kernel.cu:128-203 (ROIPoolBackward).  The reference is a gather:
 * one thread per input element sums, over RoIs in ascending order and
 * candidate bins in (ph, pw) order, the top_diff whose argmax equals the
 * element.  Here the RoI loop is outermost, which leaves every element's
 * summation sequence (0 + a + b + ...) unchanged. */
void orc_roi_pool_bwd(const float* top_diff, const int* argmax, float scale, int batch,
                      int num_rois, int H, int W, int C, int PH, int PW, const float* rois,
                      float* bottom_diff) {
  memset(bottom_diff, 0, (size_t)batch * C * H * W * sizeof(float));
  for (int n = 0; n < num_rois; ++n) {
    orc_pool_roi_t g = orc_pool_roi(rois + (size_t)n * 5, scale, PH, PW);
    if (g.batch < 0 || g.batch >= batch) continue;
    /* in_roi test, :157-159 */
    int h0 = g.sh < 0 ? 0 : g.sh, h1 = g.eh >= H ? H - 1 : g.eh;
    int w0 = g.sw < 0 ? 0 : g.sw, w1 = g.ew >= W ? W - 1 : g.ew;
    for (int c = 0; c < C; ++c) {
      const float* td = top_diff + ((size_t)n * C + c) * PH * PW;
      const int* am = argmax + ((size_t)n * C + c) * PH * PW;
      for (int h = h0; h <= h1; ++h) {
        /* :178-186 feasible pooled rows */
        int phs = (int)floor((double)((float)(h - g.sh) / g.bin_h));
        int phe = (int)ceil((double)((float)(h - g.sh + 1) / g.bin_h));
        phs = orc_clampi(phs, 0, PH);
        phe = orc_clampi(phe, 0, PH);
        for (int w = w0; w <= w1; ++w) {
          int pws = (int)floor((double)((float)(w - g.sw) / g.bin_w));
          int pwe = (int)ceil((double)((float)(w - g.sw + 1) / g.bin_w));
          pws = orc_clampi(pws, 0, PW);
          pwe = orc_clampi(pwe, 0, PW);
          int index = ((g.batch * C + c) * H + h) * W + w;
          float grad = bottom_diff[index];
          for (int ph = phs; ph < phe; ++ph)
            for (int pw = pws; pw < pwe; ++pw)
              if (am[ph * PW + pw] == index) grad += td[ph * PW + pw];
          bottom_diff[index] = grad;
        }
      }
    }
  }
}

/* ------------------------------------------------------------------ */
/* NMS.  lib/model/nms/src/nms_cuda_kernel.cu                           */
/* ------------------------------------------------------------------ */
/* :31-39 devIoU, every operation rounded to fp32 (no contraction) */
static inline float orc_dev_iou(const float* a, const float* b) {
  float left = fmaxf(a[0], b[0]), right = fminf(a[2], b[2]);
  float top = fmaxf(a[1], b[1]), bottom = fminf(a[3], b[3]);
  float width = fmaxf(right - left + 1, 0.f), height = fmaxf(bottom - top + 1, 0.f);
  float interS = width * height;
  float Sa = (a[2] - a[0] + 1) * (a[3] - a[1] + 1);
  float Sb = (b[2] - b[0] + 1) * (b[3] - b[1] + 1);
  return interS / (Sa + Sb - interS);
}

/* Result of nms_kernel (:41-85) followed by the host greedy scan (:132-144):
 * box i is kept iff no kept j < i has IoU(j, i) > thresh.  `boxes` has
 * `stride` floats per row (5 in the reference: x1 y1 x2 y2 score), already
 * sorted by score.  Returns the number kept; keep[0..ret) ascending.  Stops
 * after max_keep survivors when max_keep > 0 (the caller's `[:post_nms_topN]`
 * slice, proposal_layer.py:151-152). */
int orc_nms(const float* boxes, int n, int stride, float thresh, int max_keep, int* keep) {
  unsigned char* dead = (unsigned char*)calloc((size_t)(n > 0 ? n : 1), 1);
  int k = 0;
  for (int i = 0; i < n; ++i) {
    if (dead[i]) continue;
    keep[k++] = i;
    if (max_keep > 0 && k >= max_keep) break;
    const float* a = boxes + (size_t)i * stride;
    for (int j = i + 1; j < n; ++j) {
      if (!dead[j] && orc_dev_iou(a, boxes + (size_t)j * stride) > thresh) dead[j] = 1;
    }
  }
  free(dead);
  return k;
}

/* ------------------------------------------------------------------ */
/* Proposal layer.  lib/model/rpn/proposal_layer.py:49-163,             */
/* lib/model/rpn/bbox_transform.py:77-133                               */
/* ------------------------------------------------------------------ */
typedef struct {
  float score;
  int idx;
} orc_key_t;

/* torch.sort(scores, 1, True): descending, ties by lower index (stable) */
static int orc_key_cmp(const void* pa, const void* pb) {
  const orc_key_t* a = (const orc_key_t*)pa;
  const orc_key_t* b = (const orc_key_t*)pb;
  if (a->score > b->score) return -1;
  if (a->score < b->score) return 1;
  return (a->idx > b->idx) - (a->idx < b->idx);
}

/* Decode one anchor: bbox_transform.py:77-103 then clip_boxes :125-133.
 * exp_dw / exp_dh are supplied by the caller so that the test can inject the
 * values of the platform's exp (torch CPU, torch CUDA) -- exp is the one
 * transcendental on the path and differs by an ulp between libraries. */
static void orc_decode(const float* anchor, float sx, float sy, float dx, float dy, float exp_dw,
                       float exp_dh, float im_h, float im_w, float* out) {
  float ax1 = anchor[0] + sx, ay1 = anchor[1] + sy, ax2 = anchor[2] + sx, ay2 = anchor[3] + sy;
  float w = ax2 - ax1 + 1.0f;
  float h = ay2 - ay1 + 1.0f;
  float cx = ax1 + 0.5f * w;
  float cy = ay1 + 0.5f * h;
  float pcx = dx * w + cx;
  float pcy = dy * h + cy;
  float pw = exp_dw * w;
  float ph = exp_dh * h;
  float x1 = pcx - 0.5f * pw, y1 = pcy - 0.5f * ph, x2 = pcx + 0.5f * pw, y2 = pcy + 0.5f * ph;
  float mx = im_w - 1, my = im_h - 1;
  out[0] = fminf(fmaxf(x1, 0.f), mx);
  out[1] = fminf(fmaxf(y1, 0.f), my);
  out[2] = fminf(fmaxf(x2, 0.f), mx);
  out[3] = fminf(fmaxf(y2, 0.f), my);
}

/* Whole layer for a batch.
 *   scores   (B, 2A, H, W)  -- fg probabilities are channels [A, 2A)
 *   deltas   (B, 4A, H, W)
 *   exp_d    (B, 4A, H, W) or NULL: element-wise exp of `deltas` computed by the
 *            caller's library (only the dw/dh channels are read); NULL -> expf
 *   im_info  (B, 3)  [h, w, scale]
 *   anchors  (A, 4)
 *   out      (B, post_nms_topN, 5) zero-filled here, col 0 = image index
 *   order_out (B, n_sorted) or NULL: sorted flat anchor index per rank
 *   boxes_out (B, n_sorted, 4) or NULL: decoded+clipped boxes per rank
 *   num_out  (B) or NULL: survivors written per image
 * n_sorted = pre_nms_topN if 0 < pre_nms_topN < B*K*A else K*A   (:138)
 */
void orc_proposals(const float* scores, const float* deltas, const float* exp_d,
                   const float* im_info, const float* anchors, int B, int A, int H, int W,
                   int feat_stride, int pre_nms_topN, int post_nms_topN, float nms_thresh,
                   float* out, int* order_out, float* boxes_out, int* num_out) {
  int K = H * W, N = K * A;
  long numel = (long)B * N;
  int n_sorted = (pre_nms_topN > 0 && pre_nms_topN < numel) ? pre_nms_topN : N;
  if (n_sorted > N) n_sorted = N;
  memset(out, 0, (size_t)B * post_nms_topN * 5 * sizeof(float));
#pragma omp parallel for schedule(dynamic, 1)
  for (int b = 0; b < B; ++b) {
    orc_key_t* keys = (orc_key_t*)malloc(sizeof(orc_key_t) * (size_t)N);
    float* boxes = (float*)malloc(sizeof(float) * 4 * (size_t)n_sorted);
    int* keep = (int*)malloc(sizeof(int) * (size_t)n_sorted);
    const float* sc = scores + ((size_t)b * 2 * A + A) * K;
    const float* dl = deltas + (size_t)b * 4 * A * K;
    const float* ex = exp_d ? exp_d + (size_t)b * 4 * A * K : NULL;
    /* permute(0,2,3,1): flat index i = (y*W + x)*A + a   (:98-103) */
    for (int k = 0; k < K; ++k)
      for (int a = 0; a < A; ++a) {
        keys[k * A + a].score = sc[(size_t)a * K + k];
        keys[k * A + a].idx = k * A + a;
      }
    qsort(keys, (size_t)N, sizeof(orc_key_t), orc_key_cmp);
    for (int r = 0; r < n_sorted; ++r) {
      int i = keys[r].idx, a = i % A, k = i / A;
      float sx = (float)((k % W) * feat_stride), sy = (float)((k / W) * feat_stride);
      float dw = dl[(size_t)(a * 4 + 2) * K + k], dh = dl[(size_t)(a * 4 + 3) * K + k];
      float edw = ex ? ex[(size_t)(a * 4 + 2) * K + k] : expf(dw);
      float edh = ex ? ex[(size_t)(a * 4 + 3) * K + k] : expf(dh);
      orc_decode(anchors + a * 4, sx, sy, dl[(size_t)(a * 4 + 0) * K + k],
                 dl[(size_t)(a * 4 + 1) * K + k], edw, edh, im_info[b * 3 + 0],
                 im_info[b * 3 + 1], boxes + (size_t)r * 4);
      if (order_out) order_out[(size_t)b * n_sorted + r] = i;
    }
    if (boxes_out)
      memcpy(boxes_out + (size_t)b * n_sorted * 4, boxes, sizeof(float) * 4 * (size_t)n_sorted);
    int nk = orc_nms(boxes, n_sorted, 4, nms_thresh, post_nms_topN, keep);
    if (post_nms_topN > 0 && nk > post_nms_topN) nk = post_nms_topN;
    float* o = out + (size_t)b * post_nms_topN * 5;
    for (int r = 0; r < post_nms_topN; ++r) o[r * 5] = (float)b;
    for (int r = 0; r < nk; ++r) memcpy(o + r * 5 + 1, boxes + (size_t)keep[r] * 4, 4 * sizeof(float));
    if (num_out) num_out[b] = nk;
    free(keys);
    free(boxes);
    free(keep);
  }
}

/* ------------------------------------------------------------------ */
/* Batched IoU.  lib/model/rpn/bbox_transform.py:168-257                */
/* anchors (Bn, N, 4) with Bn in {1, B}; gt (B, K, gt_stride>=4).       */
/* ------------------------------------------------------------------ */
void orc_bbox_overlaps_batch(const float* anchors, int anchors_batched, const float* gt,
                             int gt_stride, int B, int N, int K, float* overlaps) {
#pragma omp parallel for schedule(static)
  for (int b = 0; b < B; ++b) {
    const float* an = anchors + (anchors_batched ? (size_t)b * N * 4 : 0);
    for (int i = 0; i < N; ++i) {
      const float* a = an + (size_t)i * 4;
      float aw = a[2] - a[0] + 1, ah = a[3] - a[1] + 1;
      float aarea = aw * ah;
      int azero = (aw == 1) && (ah == 1);
      for (int k = 0; k < K; ++k) {
        const float* g = gt + ((size_t)b * K + k) * gt_stride;
        float gw = g[2] - g[0] + 1, gh = g[3] - g[1] + 1;
        float garea = gw * gh;
        int gzero = (gw == 1) && (gh == 1);
        float iw = fminf(a[2], g[2]) - fmaxf(a[0], g[0]) + 1;
        if (iw < 0) iw = 0;
        float ih = fminf(a[3], g[3]) - fmaxf(a[1], g[1]) + 1;
        if (ih < 0) ih = 0;
        float ua = aarea + garea - (iw * ih);
        float ov = iw * ih / ua;
        if (gzero) ov = 0.f;
        if (azero) ov = -1.f;
        overlaps[((size_t)b * N + i) * K + k] = ov;
      }
    }
  }
}

/* Anchor-target labels before the host-side random subsampling:
 * lib/model/rpn/anchor_target_layer.py:98-116.
 *   overlaps (B, N, K) -> labels (B, N) in {-1,0,1}, argmax (B, N) int32,
 *   max_overlaps (B, N).  torch.max ties -> lowest index. */
void orc_anchor_labels(const float* overlaps, int B, int N, int K, float neg_thresh,
                       float pos_thresh, int clobber_positives, float* labels, int* argmax,
                       float* max_overlaps) {
  for (int b = 0; b < B; ++b) {
    const float* ov = overlaps + (size_t)b * N * K;
    float* gt_max = (float*)malloc(sizeof(float) * (size_t)K);
    for (int k = 0; k < K; ++k) {
      float m = -INFINITY;
      for (int i = 0; i < N; ++i)
        if (ov[(size_t)i * K + k] > m) m = ov[(size_t)i * K + k];
      gt_max[k] = (m == 0.f) ? 1e-5f : m; /* :106 */
    }
    for (int i = 0; i < N; ++i) {
      float m = -INFINITY;
      int am = 0;
      for (int k = 0; k < K; ++k)
        if (ov[(size_t)i * K + k] > m) {
          m = ov[(size_t)i * K + k];
          am = k;
        }
      float lab = -1.f;
      if (!clobber_positives && m < neg_thresh) lab = 0.f;
      int hit = 0;
      for (int k = 0; k < K; ++k) hit += (ov[(size_t)i * K + k] == gt_max[k]);
      if (hit > 0) lab = 1.f;
      if (m >= pos_thresh) lab = 1.f;
      if (clobber_positives && m < neg_thresh) lab = 0.f;
      labels[(size_t)b * N + i] = lab;
      argmax[(size_t)b * N + i] = am;
      max_overlaps[(size_t)b * N + i] = m;
    }
    free(gt_max);
  }
}

/* ------------------------------------------------------------------ */
/* RoICrop = bilinear sampler over a (y, x) grid in [-1, 1]            */
/* lib/model/roi_crop/src/roi_crop_cuda_kernel.cu:12-23 (getTopLeft),  */
/* :45-108 (forward), :111-190 (backward: gradient w.r.t. the images   */
/* only -- the kernel never writes gradGrids).                         */
/* input (ib, C, H, W) NCHW, grid (ob, GH, GW, 2) = (y, x), output     */
/* (ob, C, GH, GW); output batch b samples image b / (ob / ib).        */
/* ------------------------------------------------------------------ */
static void orc_top_left(float x, int size, int* point, float* weight) {
  float coord = (x + 1.0f) * (float)(size - 1) / 2.0f;
  float fl = floorf(coord);
  *point = (int)fl;
  *weight = 1.0f - (coord - fl);
}

void orc_roi_crop_fwd(const float* input, const float* grid, float* output, int ib, int C, int H,
                      int W, int ob, int GH, int GW) {
  const int per = ob / ib;
#pragma omp parallel for schedule(static)
  for (int b = 0; b < ob; ++b) {
    const int bi = b / per;
    for (int y = 0; y < GH; ++y)
      for (int x = 0; x < GW; ++x) {
        const float yf = grid[(((size_t)b * GH + y) * GW + x) * 2 + 0];
        const float xf = grid[(((size_t)b * GH + y) * GW + x) * 2 + 1];
        int xi, yi;
        float xw, yw;
        orc_top_left(xf, W, &xi, &xw);
        orc_top_left(yf, H, &yi, &yw);
        const int tl = xi >= 0 && xi <= W - 1 && yi >= 0 && yi <= H - 1;
        const int tr = xi + 1 >= 0 && xi + 1 <= W - 1 && yi >= 0 && yi <= H - 1;
        const int bl = xi >= 0 && xi <= W - 1 && yi + 1 >= 0 && yi + 1 <= H - 1;
        const int br = xi + 1 >= 0 && xi + 1 <= W - 1 && yi + 1 >= 0 && yi + 1 <= H - 1;
        for (int c = 0; c < C; ++c) {
          const float* p = input + ((size_t)bi * C + c) * H * W;
          float v = 0.f;
          if (tl || tr || bl || br) {
            const float a = tl ? p[yi * W + xi] : 0.f;
            const float bb = tr ? p[yi * W + xi + 1] : 0.f;
            const float cc = bl ? p[(yi + 1) * W + xi] : 0.f;
            const float d = br ? p[(yi + 1) * W + xi + 1] : 0.f;
            v = xw * yw * a + (1 - xw) * yw * bb + xw * (1 - yw) * cc + (1 - xw) * (1 - yw) * d;
          }
          output[(((size_t)b * C + c) * GH + y) * GW + x] = v;
        }
      }
  }
}

/* grad_input (ib, C, H, W) accumulated in double, then rounded */
void orc_roi_crop_bwd(const float* grad_out, const float* grid, float* grad_input, int ib, int C,
                      int H, int W, int ob, int GH, int GW) {
  const int per = ob / ib;
  const size_t total = (size_t)ib * C * H * W;
  double* acc = (double*)calloc(total, sizeof(double));
  for (int b = 0; b < ob; ++b) {
    const int bi = b / per;
    for (int y = 0; y < GH; ++y)
      for (int x = 0; x < GW; ++x) {
        const float yf = grid[(((size_t)b * GH + y) * GW + x) * 2 + 0];
        const float xf = grid[(((size_t)b * GH + y) * GW + x) * 2 + 1];
        int xi, yi;
        float xw, yw;
        orc_top_left(xf, W, &xi, &xw);
        orc_top_left(yf, H, &yi, &yw);
        const int tl = xi >= 0 && xi <= W - 1 && yi >= 0 && yi <= H - 1;
        const int tr = xi + 1 >= 0 && xi + 1 <= W - 1 && yi >= 0 && yi <= H - 1;
        const int bl = xi >= 0 && xi <= W - 1 && yi + 1 >= 0 && yi + 1 <= H - 1;
        const int br = xi + 1 >= 0 && xi + 1 <= W - 1 && yi + 1 >= 0 && yi + 1 <= H - 1;
        for (int c = 0; c < C; ++c) {
          const float g = grad_out[(((size_t)b * C + c) * GH + y) * GW + x];
          double* p = acc + ((size_t)bi * C + c) * H * W;
          if (tl) p[yi * W + xi] += (double)(xw * yw * g);
          if (tr) p[yi * W + xi + 1] += (double)((1 - xw) * yw * g);
          if (bl) p[(yi + 1) * W + xi] += (double)(xw * (1 - yw) * g);
          if (br) p[(yi + 1) * W + xi + 1] += (double)((1 - xw) * (1 - yw) * g);
        }
      }
  }
  for (size_t i = 0; i < total; ++i) grad_input[i] = (float)acc[i];
  free(acc);
}

int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

void orc_set_num_threads(int n) {
#ifdef _OPENMP
  omp_set_num_threads(n);
#else
  (void)n;
#endif
}
