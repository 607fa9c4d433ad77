// Box arithmetic and anchor-target assignment for sm_100a.
//
// Reference semantics (all fp32, every torch op individually rounded, so no
// FMA contraction here either):
//   bbox_transform_batch   lib/model/rpn/bbox_transform.py:36-75
//   bbox_transform_inv     lib/model/rpn/bbox_transform.py:77-103
//   clip_boxes             lib/model/rpn/bbox_transform.py:125-133
//   bbox_overlaps_batch    lib/model/rpn/bbox_transform.py:168-257
//   _AnchorTargetLayer     lib/model/rpn/anchor_target_layer.py:98-116, :147-191
//   _ProposalTargetLayer   lib/model/rpn/proposal_target_layer_cascade.py:33-57, :113-212
#include <limits.h>

#include "common.cuh"

namespace tlod {

// ---------------------------------------------------------------------------
// IoU of one (anchor, gt) pair, bbox_transform.py:181-214
// ---------------------------------------------------------------------------
struct GtBox {
  float x1, y1, x2, y2, area;
  int zero;  // w == 1 && h == 1 (zero padding)
};

__device__ __forceinline__ GtBox make_gt(const float* __restrict__ g) {
  GtBox b;
  b.x1 = __ldg(g);
  b.y1 = __ldg(g + 1);
  b.x2 = __ldg(g + 2);
  b.y2 = __ldg(g + 3);
  const float w = __fadd_rn(__fsub_rn(b.x2, b.x1), 1.f), h = __fadd_rn(__fsub_rn(b.y2, b.y1), 1.f);
  b.area = __fmul_rn(w, h);
  b.zero = (w == 1.f) && (h == 1.f);
  return b;
}

__device__ __forceinline__ float pair_overlap(const float4 a, float a_area, bool a_zero,
                                              const GtBox& g) {
  float iw = __fadd_rn(__fsub_rn(fminf(a.z, g.x2), fmaxf(a.x, g.x1)), 1.f);
  if (iw < 0.f) iw = 0.f;
  float ih = __fadd_rn(__fsub_rn(fminf(a.w, g.y2), fmaxf(a.y, g.y1)), 1.f);
  if (ih < 0.f) ih = 0.f;
  const float inter = __fmul_rn(iw, ih);
  const float ua = __fsub_rn(__fadd_rn(a_area, g.area), inter);
  float ov = __fdiv_rn(inter, ua);
  if (g.zero) ov = 0.f;
  if (a_zero) ov = -1.f;
  return ov;
}

__device__ __forceinline__ float4 load4(const float* __restrict__ p) {
  return make_float4(__ldg(p), __ldg(p + 1), __ldg(p + 2), __ldg(p + 3));
}

__global__ void __launch_bounds__(256)
    overlaps_kernel(const float* __restrict__ anchors, int anchors_batched, int astride, int aoff,
                    const float* __restrict__ gt, int gstride, float* __restrict__ out, int n, int k) {
  const int b = blockIdx.y;
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (long long)n * k) return;
  const int i = (int)(e / k), j = (int)(e - (long long)i * k);
  const float* ap = anchors + ((size_t)(anchors_batched ? b : 0) * n + i) * astride + aoff;
  const float4 a = load4(ap);
  const float aw = __fadd_rn(__fsub_rn(a.z, a.x), 1.f), ah = __fadd_rn(__fsub_rn(a.w, a.y), 1.f);
  const GtBox g = make_gt(gt + ((size_t)b * k + j) * gstride);
  out[(size_t)b * n * k + e] = pair_overlap(a, __fmul_rn(aw, ah), (aw == 1.f) && (ah == 1.f), g);
}

// Monotone float -> int map (valid for every non-NaN float).
__device__ __forceinline__ int f2ord(float f) {
  const int b = __float_as_int(f);
  return b >= 0 ? b : b ^ 0x7fffffff;
}
__device__ __forceinline__ float ord2f(int o) { return __int_as_float(o >= 0 ? o : o ^ 0x7fffffff); }

constexpr int AT_MAXK = 1024;

// The gt boxes of the CTA's image, staged once in shared memory in chunks of AT_GT_CHUNK (the
// kernels read every gt for every anchor; from global memory that was five loads per pair).
constexpr int AT_GT_CHUNK = 64;

// pass 1: per-gt maximum over anchors -> gtmax[b * k + j] (ordered ints, pre-set to 0x80808080).
// One CTA per (gt box, image, anchor segment): every thread strides over its segment of the anchors
// (one 128-bit load per anchor, four in flight), block reduction, ONE global atomicMax per CTA --
// GTM_SEGS per address.  (The anchor-major version -- a thread per anchor looping over the gt boxes,
// shared-memory atomicMax, then one global atomicMax per CTA and gt -- took 25-30 us for two
// 600x1200 images; a warp REDUX in front of its atomics did not help.)
constexpr int GTM_SEGS = 4;
__global__ void __launch_bounds__(256)
    gt_max_kernel(const float* __restrict__ anchors, const float* __restrict__ gt, int gstride,
                  int* __restrict__ gtmax, int n, int k) {
  __shared__ int wmax[8];
  const int j = blockIdx.x, b = blockIdx.y;
  const GtBox g = make_gt(gt + ((size_t)b * k + j) * gstride);
  const int per = (n + GTM_SEGS - 1) / GTM_SEGS;
  const int lo = blockIdx.z * per, hi = min(n, lo + per);
  const float4* __restrict__ a4 = reinterpret_cast<const float4*>(anchors);
  const bool vec = (((uintptr_t)anchors) & 15) == 0;
  int m = INT_MIN;
#pragma unroll 4
  for (int i = lo + threadIdx.x; i < hi; i += 256) {
    const float4 a = vec ? __ldg(a4 + i) : load4(anchors + (size_t)i * 4);
    const float aw = __fadd_rn(__fsub_rn(a.z, a.x), 1.f), ah = __fadd_rn(__fsub_rn(a.w, a.y), 1.f);
    m = max(m, f2ord(pair_overlap(a, __fmul_rn(aw, ah), (aw == 1.f) && (ah == 1.f), g)));
  }
  m = __reduce_max_sync(0xffffffffu, m);
  if ((threadIdx.x & 31) == 0) wmax[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int w = 1; w < 8; ++w) m = max(m, wmax[w]);
    if (m != INT_MIN) atomicMax(&gtmax[(size_t)b * k + j], m);
  }
}

// pass 2: labels, anchor_target_layer.py:100-116
__global__ void __launch_bounds__(256)
    anchor_labels_kernel(const float* __restrict__ anchors, const float* __restrict__ gt, int gstride,
                         const int* __restrict__ gtmax, float* __restrict__ labels,
                         int* __restrict__ argmax, float* __restrict__ max_overlaps, int n, int k,
                         float neg, float pos, int clobber) {
  __shared__ GtBox sgt[AT_GT_CHUNK];
  __shared__ float sgtmax[AT_GT_CHUNK];
  const int b = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool have = i < n;
  const float4 a = have ? load4(anchors + (size_t)i * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
  const float aw = __fadd_rn(__fsub_rn(a.z, a.x), 1.f), ah = __fadd_rn(__fsub_rn(a.w, a.y), 1.f);
  const float aarea = __fmul_rn(aw, ah);
  const bool azero = (aw == 1.f) && (ah == 1.f);
  float best = -INFINITY;
  int besti = 0, hit = 0;
  for (int j0 = 0; j0 < k; j0 += AT_GT_CHUNK) {
    const int kc = min(AT_GT_CHUNK, k - j0);
    __syncthreads();
    for (int j = threadIdx.x; j < kc; j += blockDim.x) {
      sgt[j] = make_gt(gt + ((size_t)b * k + j0 + j) * gstride);
      const float m = ord2f(gtmax[(size_t)b * k + j0 + j]);
      sgtmax[j] = (m == 0.f) ? 1e-5f : m;  // :106
    }
    __syncthreads();
    for (int j = 0; j < kc; ++j) {
      const float ov = pair_overlap(a, aarea, azero, sgt[j]);
      if (ov > best) {
        best = ov;
        besti = j0 + j;
      }
      hit |= (ov == sgtmax[j]);
    }
  }
  if (!have) return;
  float lab = -1.f;
  if (!clobber && best < neg) lab = 0.f;
  if (hit) lab = 1.f;
  if (best >= pos) lab = 1.f;
  if (clobber && best < neg) lab = 0.f;
  const size_t o = (size_t)b * n + i;
  labels[o] = lab;
  argmax[o] = besti;
  if (max_overlaps) max_overlaps[o] = best;
}

// bbox_transform_batch for one pair, bbox_transform.py:38-53
__device__ __forceinline__ float4 encode_box(const float4 ex, const float4 gt) {
  const float ew = __fadd_rn(__fsub_rn(ex.z, ex.x), 1.0f), eh = __fadd_rn(__fsub_rn(ex.w, ex.y), 1.0f);
  const float ecx = __fadd_rn(ex.x, __fmul_rn(0.5f, ew)), ecy = __fadd_rn(ex.y, __fmul_rn(0.5f, eh));
  const float gw = __fadd_rn(__fsub_rn(gt.z, gt.x), 1.0f), gh = __fadd_rn(__fsub_rn(gt.w, gt.y), 1.0f);
  const float gcx = __fadd_rn(gt.x, __fmul_rn(0.5f, gw)), gcy = __fadd_rn(gt.y, __fmul_rn(0.5f, gh));
  float4 t;
  t.x = __fdiv_rn(__fsub_rn(gcx, ecx), ew);
  t.y = __fdiv_rn(__fsub_rn(gcy, ecy), eh);
  t.z = logf(__fdiv_rn(gw, ew));
  t.w = logf(__fdiv_rn(gh, eh));
  return t;
}

__global__ void __launch_bounds__(256)
    transform_batch_kernel(const float* __restrict__ ex, int ex_batched, const float* __restrict__ gt,
                           float* __restrict__ out, int n) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 e = load4(ex + ((size_t)(ex_batched ? b : 0) * n + i) * 4);
  const float4 g = load4(gt + ((size_t)b * n + i) * 4);
  const float4 t = encode_box(e, g);
  float* o = out + ((size_t)b * n + i) * 4;
  o[0] = t.x; o[1] = t.y; o[2] = t.z; o[3] = t.w;
}

// bbox_transform_inv (+ optional clip), bbox_transform.py:77-103, :125-133
__global__ void __launch_bounds__(256)
    transform_inv_kernel(const float* __restrict__ boxes, int boxes_batched,
                         const float* __restrict__ deltas, const float* __restrict__ im_info,
                         float* __restrict__ out, int n) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 a = load4(boxes + ((size_t)(boxes_batched ? b : 0) * n + i) * 4);
  const float4 d = load4(deltas + ((size_t)b * n + i) * 4);
  const float w = __fadd_rn(__fsub_rn(a.z, a.x), 1.0f), h = __fadd_rn(__fsub_rn(a.w, a.y), 1.0f);
  const float cx = __fadd_rn(a.x, __fmul_rn(0.5f, w)), cy = __fadd_rn(a.y, __fmul_rn(0.5f, h));
  const float pcx = __fadd_rn(__fmul_rn(d.x, w), cx), pcy = __fadd_rn(__fmul_rn(d.y, h), cy);
  const float pw = __fmul_rn(expf(d.z), w), ph = __fmul_rn(expf(d.w), h);
  const float hw = __fmul_rn(0.5f, pw), hh = __fmul_rn(0.5f, ph);
  float x1 = __fsub_rn(pcx, hw), y1 = __fsub_rn(pcy, hh), x2 = __fadd_rn(pcx, hw), y2 = __fadd_rn(pcy, hh);
  if (im_info) {
    const float mx = __fsub_rn(__ldg(im_info + b * 3 + 1), 1.f), my = __fsub_rn(__ldg(im_info + b * 3), 1.f);
    x1 = fminf(fmaxf(x1, 0.f), mx); y1 = fminf(fmaxf(y1, 0.f), my);
    x2 = fminf(fmaxf(x2, 0.f), mx); y2 = fminf(fmaxf(y2, 0.f), my);
  }
  float* o = out + ((size_t)b * n + i) * 4;
  o[0] = x1; o[1] = y1; o[2] = x2; o[3] = y2;
}

__global__ void __launch_bounds__(256)
    clip_kernel(float* __restrict__ boxes, const float* __restrict__ im_info, long long per_image) {
  const int b = blockIdx.y;
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= per_image) return;
  const float lim = __fsub_rn(__ldg(im_info + b * 3 + ((e & 1) ? 0 : 1)), 1.f);  // even cols: x -> width
  float* p = boxes + (size_t)b * per_image + e;
  *p = fminf(fmaxf(*p, 0.f), lim);
}

// Final anchor-target maps, anchor_target_layer.py:147-191
__global__ void __launch_bounds__(256)
    anchor_finalize_kernel(const float* __restrict__ labels, const int* __restrict__ argmax,
                           const float* __restrict__ anchors, const float* __restrict__ gt,
                           int gstride, const int* __restrict__ inv_index,
                           float* __restrict__ labels_out, float* __restrict__ targets_out,
                           float* __restrict__ inside_out, float* __restrict__ outside_out, int n,
                           int k, int A, int H, int W, float inside_w, float pos_w, float neg_w,
                           const float* __restrict__ weights_dev) {
  if (weights_dev) {  // {inside, positive, negative} in device memory (a captured launch cannot bake them in)
    inside_w = __ldg(weights_dev);
    pos_w = __ldg(weights_dev + 1);
    neg_w = __ldg(weights_dev + 2);
  }
  const int b = blockIdx.y;
  const int K = H * W;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;  // e = a * K + cell
  if (e >= A * K) return;
  const int a = e / K, cell = e - a * K;
  const int pos = __ldg(inv_index + (size_t)cell * A + a);
  float lab = -1.f;
  float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
  float iw = 0.f, ow = 0.f;
  if (pos >= 0) {
    lab = __ldg(labels + (size_t)b * n + pos);
    const int am = __ldg(argmax + (size_t)b * n + pos);
    t = encode_box(load4(anchors + (size_t)pos * 4), load4(gt + ((size_t)b * k + am) * gstride));
    iw = (lab == 1.f) ? inside_w : 0.f;
    ow = (lab == 1.f) ? pos_w : ((lab == 0.f) ? neg_w : 0.f);
  }
  labels_out[(size_t)b * A * K + e] = lab;  // (B, 1, A*H, W)
  const size_t base = ((size_t)b * 4 * A + a * 4) * K + cell;
  targets_out[base] = t.x;
  targets_out[base + K] = t.y;
  targets_out[base + 2 * (size_t)K] = t.z;
  targets_out[base + 3 * (size_t)K] = t.w;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    inside_out[base + (size_t)j * K] = iw;
    outside_out[base + (size_t)j * K] = ow;
  }
}

// ---------------------------------------------------------------------------
// _ProposalTargetLayer, device part 1 (proposal_target_layer_cascade.py:118-130):
// per RoI the best-overlapping gt box (ties -> lowest index), that overlap and the gt's class.
// rois (B, n, roi_stride) with the box at column roi_off; gt (B, k, >= 5).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    roi_gt_assign_kernel(const float* __restrict__ rois, int roi_stride, int roi_off,
                         const float* __restrict__ gt, int gstride, float* __restrict__ max_overlaps,
                         int* __restrict__ assignment, float* __restrict__ labels, int n, int k) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 a = load4(rois + ((size_t)b * n + i) * roi_stride + roi_off);
  const float aw = __fadd_rn(__fsub_rn(a.z, a.x), 1.f), ah = __fadd_rn(__fsub_rn(a.w, a.y), 1.f);
  const float aarea = __fmul_rn(aw, ah);
  const bool azero = (aw == 1.f) && (ah == 1.f);
  float best = -INFINITY;
  int besti = 0;
  for (int j = 0; j < k; ++j) {
    const GtBox g = make_gt(gt + ((size_t)b * k + j) * gstride);
    const float ov = pair_overlap(a, aarea, azero, g);
    if (ov > best) {
      best = ov;
      besti = j;
    }
  }
  const size_t o = (size_t)b * n + i;
  max_overlaps[o] = best;
  assignment[o] = besti;
  labels[o] = __ldg(gt + ((size_t)b * k + besti) * gstride + 4);
}

// device part 2 (:183-212): gather the sampled RoIs, clamp background labels, encode and
// normalise the regression targets, inside / outside weights.
// keep (B, P) int32 indices into the n candidates; fg_count (B): positions >= fg_count[b] are
// background.  rois_out (B, P, 5), labels_out (B, P), targets / inside / outside (B, P, 4).
__global__ void __launch_bounds__(256)
    proposal_targets_kernel(const float* __restrict__ rois, int roi_stride, int roi_off,
                            const float* __restrict__ gt, int gstride, const int* __restrict__ assignment,
                            const float* __restrict__ labels, const int* __restrict__ keep,
                            const int* __restrict__ fg_count, float* __restrict__ rois_out,
                            float* __restrict__ labels_out, float* __restrict__ targets_out,
                            float* __restrict__ inside_out, float* __restrict__ outside_out, int n, int k,
                            int P, float4 means, float4 stds, float4 inside_w, int normalize) {
  const int b = blockIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= P) return;
  const int src = __ldg(keep + (size_t)b * P + j);
  const size_t so = (size_t)b * n + src;
  const float4 box = load4(rois + so * roi_stride + roi_off);
  float lab = __ldg(labels + so);
  if (j >= __ldg(fg_count + b)) lab = 0.f;  // :194-195
  const float4 g = load4(gt + ((size_t)b * k + __ldg(assignment + so)) * gstride);
  float4 t = encode_box(box, g);
  if (normalize) {  // :106-109: (targets - means) / stds, each op rounded
    t.x = __fdiv_rn(__fsub_rn(t.x, means.x), stds.x);
    t.y = __fdiv_rn(__fsub_rn(t.y, means.y), stds.y);
    t.z = __fdiv_rn(__fsub_rn(t.z, means.z), stds.z);
    t.w = __fdiv_rn(__fsub_rn(t.w, means.w), stds.w);
  }
  const bool fg = lab > 0.f;
  const size_t o = (size_t)b * P + j;
  float* r = rois_out + o * 5;
  r[0] = (float)b;
  r[1] = box.x; r[2] = box.y; r[3] = box.z; r[4] = box.w;
  labels_out[o] = lab;
  const float4 tt = fg ? t : make_float4(0.f, 0.f, 0.f, 0.f);
  const float4 iw = fg ? inside_w : make_float4(0.f, 0.f, 0.f, 0.f);
  reinterpret_cast<float4*>(targets_out)[o] = tt;
  reinterpret_cast<float4*>(inside_out)[o] = iw;
  reinterpret_cast<float4*>(outside_out)[o] =
      make_float4(iw.x > 0.f ? 1.f : 0.f, iw.y > 0.f ? 1.f : 0.f, iw.z > 0.f ? 1.f : 0.f, iw.w > 0.f ? 1.f : 0.f);
}

}  // namespace tlod

using namespace tlod;

extern "C" int tlod_bbox_overlaps_batch(const float* anchors, int anchors_batched,
                                        int anchor_stride, int anchor_offset, const float* gt,
                                        int gt_stride, float* overlaps, int batch, int n, int k,
                                        void* stream) {
  if (!anchors || !gt || !overlaps) return TLOD_ERR_NULL_POINTER;
  if (batch <= 0 || n < 0 || k < 0 || anchor_stride < 4 || gt_stride < 4 || anchor_offset < 0 ||
      anchor_offset + 4 > anchor_stride || batch > 65535)
    return TLOD_ERR_BAD_SHAPE;
  if ((long long)n * k == 0) return TLOD_OK;
  const long long per = (long long)n * k;
  dim3 grid((unsigned)((per + 255) / 256), batch);
  {
    LaunchScope scope("overlaps_kernel", (cudaStream_t)stream);
    overlaps_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(anchors, anchors_batched, anchor_stride,
                                                            anchor_offset, gt, gt_stride, overlaps, n, k);
  }
  return last_launch_status();
}

extern "C" int tlod_bbox_transform_batch(const float* ex_rois, int ex_batched, const float* gt_rois,
                                         float* targets, int batch, int n, void* stream) {
  if (!ex_rois || !gt_rois || !targets) return TLOD_ERR_NULL_POINTER;
  if (batch <= 0 || n < 0 || batch > 65535) return TLOD_ERR_BAD_SHAPE;
  if (n == 0) return TLOD_OK;
  dim3 grid((n + 255) / 256, batch);
  {
    LaunchScope scope("transform_batch_kernel", (cudaStream_t)stream);
    transform_batch_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(ex_rois, ex_batched, gt_rois, targets, n);
  }
  return last_launch_status();
}

extern "C" int tlod_bbox_transform_inv_clip(const float* boxes, int boxes_batched,
                                            const float* deltas, const float* im_info, float* out,
                                            int batch, int n, void* stream) {
  if (!boxes || !deltas || !out) return TLOD_ERR_NULL_POINTER;
  if (batch <= 0 || n < 0 || batch > 65535) return TLOD_ERR_BAD_SHAPE;
  if (n == 0) return TLOD_OK;
  dim3 grid((n + 255) / 256, batch);
  {
    LaunchScope scope("transform_inv_kernel", (cudaStream_t)stream);
    transform_inv_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(boxes, boxes_batched, deltas, im_info, out, n);
  }
  return last_launch_status();
}

extern "C" int tlod_clip_boxes(float* boxes, const float* im_info, int batch, int n, int k,
                               void* stream) {
  if (!boxes || !im_info) return TLOD_ERR_NULL_POINTER;
  if (batch <= 0 || n < 0 || k <= 0 || batch > 65535) return TLOD_ERR_BAD_SHAPE;
  const long long per = (long long)n * 4 * k;
  if (per == 0) return TLOD_OK;
  dim3 grid((unsigned)((per + 255) / 256), batch);
  {
    LaunchScope scope("clip_kernel", (cudaStream_t)stream);
    clip_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(boxes, im_info, per);
  }
  return last_launch_status();
}

extern "C" size_t tlod_anchor_labels_workspace_bytes(int batch, int k) {
  if (batch <= 0 || k <= 0) return 16;
  return ((size_t)batch * k * sizeof(int) + 255) / 256 * 256;
}

extern "C" int tlod_anchor_labels(const float* anchors, const float* gt, int gt_stride,
                                  float* labels, int* argmax, float* max_overlaps, int batch, int n,
                                  int k, float negative_overlap, float positive_overlap,
                                  int clobber_positives, void* workspace, size_t workspace_bytes,
                                  void* stream) {
  if (!anchors || !gt || !labels || !argmax) return TLOD_ERR_NULL_POINTER;
  if (batch <= 0 || n <= 0 || k <= 0 || gt_stride < 4 || batch > 65535) return TLOD_ERR_BAD_SHAPE;
  if (k > AT_MAXK) return TLOD_ERR_UNSUPPORTED;
  if (!workspace || workspace_bytes < tlod_anchor_labels_workspace_bytes(batch, k))
    return TLOD_ERR_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  int* gtmax = (int*)workspace;
  cudaError_t e = cudaMemsetAsync(gtmax, 0x80, (size_t)batch * k * sizeof(int), st);
  if (e != cudaSuccess) return (int)e;
  dim3 grid((n + 255) / 256, batch);
  {
    LaunchScope scope("gt_max_kernel", st);
    gt_max_kernel<<<dim3(k, batch, GTM_SEGS), 256, 0, st>>>(anchors, gt, gt_stride, gtmax, n, k);
  }
  int rc = last_launch_status();
  if (rc) return rc;
  {
    LaunchScope scope("anchor_labels_kernel", st);
    anchor_labels_kernel<<<grid, 256, 0, st>>>(anchors, gt, gt_stride, gtmax, labels, argmax,
                                               max_overlaps, n, k, negative_overlap, positive_overlap,
                                               clobber_positives);
  }
  return last_launch_status();
}

static int anchor_targets_finalize(const float* labels, const int* argmax, const float* anchors, const float* gt,
                                   int gt_stride, const int* inv_index, float* labels_out, float* targets_out,
                                   float* inside_w_out, float* outside_w_out, int batch, int n, int k,
                                   int num_anchors, int height, int width, float inside_weight,
                                   float positive_weight, float negative_weight, const float* weights_dev,
                                   void* stream) {
  if (!labels || !argmax || !anchors || !gt || !inv_index || !labels_out || !targets_out ||
      !inside_w_out || !outside_w_out)
    return TLOD_ERR_NULL_POINTER;
  if (batch <= 0 || n <= 0 || k <= 0 || num_anchors <= 0 || height <= 0 || width <= 0 ||
      gt_stride < 4 || batch > 65535)
    return TLOD_ERR_BAD_SHAPE;
  const int total = num_anchors * height * width;
  dim3 grid((total + 255) / 256, batch);
  {
    LaunchScope scope("anchor_finalize_kernel", (cudaStream_t)stream);
    anchor_finalize_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
        labels, argmax, anchors, gt, gt_stride, inv_index, labels_out, targets_out, inside_w_out,
        outside_w_out, n, k, num_anchors, height, width, inside_weight, positive_weight,
        negative_weight, weights_dev);
  }
  return last_launch_status();
}

extern "C" int tlod_anchor_targets_finalize(const float* labels, const int* argmax,
                                            const float* anchors, const float* gt, int gt_stride,
                                            const int* inv_index, float* labels_out,
                                            float* targets_out, float* inside_w_out,
                                            float* outside_w_out, int batch, int n, int k,
                                            int num_anchors, int height, int width,
                                            float inside_weight, float positive_weight,
                                            float negative_weight, void* stream) {
  return anchor_targets_finalize(labels, argmax, anchors, gt, gt_stride, inv_index, labels_out, targets_out,
                                 inside_w_out, outside_w_out, batch, n, k, num_anchors, height, width,
                                 inside_weight, positive_weight, negative_weight, nullptr, stream);
}

extern "C" int tlod_anchor_targets_finalize_dev(const float* labels, const int* argmax, const float* anchors,
                                                const float* gt, int gt_stride, const int* inv_index,
                                                float* labels_out, float* targets_out, float* inside_w_out,
                                                float* outside_w_out, int batch, int n, int k, int num_anchors,
                                                int height, int width, const float* weights_dev, void* stream) {
  if (!weights_dev) return TLOD_ERR_NULL_POINTER;
  return anchor_targets_finalize(labels, argmax, anchors, gt, gt_stride, inv_index, labels_out, targets_out,
                                 inside_w_out, outside_w_out, batch, n, k, num_anchors, height, width, 0.f, 0.f, 0.f,
                                 weights_dev, stream);
}

extern "C" int tlod_roi_gt_assign(const float* rois, int roi_stride, int roi_offset, const float* gt,
                                  int gt_stride, float* max_overlaps, int* assignment, float* labels,
                                  int batch, int n, int k, void* stream) {
  if (!rois || !gt || !max_overlaps || !assignment || !labels) return TLOD_ERR_NULL_POINTER;
  if (batch <= 0 || n <= 0 || k <= 0 || roi_stride < roi_offset + 4 || gt_stride < 5 || batch > 65535)
    return TLOD_ERR_BAD_SHAPE;
  cudaStream_t st = (cudaStream_t)stream;
  {
    LaunchScope scope("roi_gt_assign_kernel", st);
    roi_gt_assign_kernel<<<dim3((n + 255) / 256, batch), 256, 0, st>>>(rois, roi_stride, roi_offset, gt, gt_stride,
                                                                    max_overlaps, assignment, labels, n, k);
  }
  return last_launch_status();
}

extern "C" int tlod_proposal_targets(const float* rois, int roi_stride, int roi_offset, const float* gt,
                                     int gt_stride, const int* assignment, const float* labels,
                                     const int* keep, const int* fg_count, float* rois_out,
                                     float* labels_out, float* targets_out, float* inside_out,
                                     float* outside_out, int batch, int n, int k, int rois_per_image,
                                     const float* h_means, const float* h_stds, const float* h_inside_w,
                                     int normalize, void* stream) {
  if (!rois || !gt || !assignment || !labels || !keep || !fg_count || !rois_out || !labels_out ||
      !targets_out || !inside_out || !outside_out || !h_means || !h_stds || !h_inside_w)
    return TLOD_ERR_NULL_POINTER;
  if (batch <= 0 || n <= 0 || k <= 0 || rois_per_image <= 0 || roi_stride < roi_offset + 4 || gt_stride < 5 ||
      batch > 65535)
    return TLOD_ERR_BAD_SHAPE;
  if (((uintptr_t)targets_out | (uintptr_t)inside_out | (uintptr_t)outside_out) & 15) return TLOD_ERR_BAD_SHAPE;
  cudaStream_t st = (cudaStream_t)stream;
  const float4 m = make_float4(h_means[0], h_means[1], h_means[2], h_means[3]);
  const float4 sd = make_float4(h_stds[0], h_stds[1], h_stds[2], h_stds[3]);
  const float4 iw = make_float4(h_inside_w[0], h_inside_w[1], h_inside_w[2], h_inside_w[3]);
  {
    LaunchScope scope("proposal_targets_kernel", st);
    proposal_targets_kernel<<<dim3((rois_per_image + 255) / 256, batch), 256, 0, st>>>(
        rois, roi_stride, roi_offset, gt, gt_stride, assignment, labels, keep, fg_count, rois_out, labels_out,
        targets_out, inside_out, outside_out, n, k, rois_per_image, m, sd, iw, normalize);
  }
  return last_launch_status();
}
