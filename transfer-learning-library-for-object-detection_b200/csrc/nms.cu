// NMS and the fused, batched proposal layer for sm_100a.
//
// Reference semantics:
//   nms            lib/model/nms/src/nms_cuda_kernel.cu:31-39 (devIoU), :41-85 (mask),
//                  :132-144 (host greedy scan)
//   proposal layer lib/model/rpn/proposal_layer.py:49-163,
//                  lib/model/rpn/bbox_transform.py:77-103 (decode), :125-133 (clip)
//
// Pipeline for a batch (3 launches, no host synchronisation, no D2H mask copy):
//   1. proposal_topk_decode_kernel  one CTA per image: 64-bit radix select of the
//      pre_nms_topN best (score desc, index asc) keys, bitonic sort in shared
//      memory, decode + clip of only the selected anchors.
//   2. nms_mask_kernel              upper-triangular 64x64 tiles of the IoU > thresh
//      bitmask over all images (fp32 ALU bound; this is where the time goes).
//   3. nms_scan_kernel              one CTA per image: the greedy scan the reference
//      runs on the host, done on chip with a speculative prefetch of the diagonal
//      and super-diagonal mask words, then the padded (post_nms_topN, 5) output.
#include "common.cuh"

namespace tlod {

// ===========================================================================
// IoU > thresh, bit-exact with the reference's fp32 expression
//   interS / (Sa + Sb - interS) > thresh           (every op rounded, no FMA)
// The division is only executed when the quotient is within ~1e-6 of thresh.
// ===========================================================================
__device__ __forceinline__ bool iou_exceeds(float inter, float uni, float thresh, float thr_hi,
                                            float thr_lo) {
  // thr_hi = thresh * (1 + 2^-20), thr_lo = thresh * (1 - 2^-20); valid because
  // uni is a sum of areas in [1e-20, 1e30] (checked) and thresh in [1e-6, 1e6].
  if (uni >= 1e-20f && uni <= 1e30f) {
    if (inter > thr_hi * uni) return true;
    if (inter < thr_lo * uni) return false;
  }
  return __fdiv_rn(inter, uni) > thresh;
}

template <bool FILTER>
__device__ __forceinline__ bool box_suppresses(const float4 a, float area_a, const float4 b,
                                               float area_b, float thresh, float thr_hi,
                                               float thr_lo) {
  const float left = fmaxf(a.x, b.x), right = fminf(a.z, b.z);
  const float top = fmaxf(a.y, b.y), bottom = fminf(a.w, b.w);
  const float w = fmaxf(__fadd_rn(__fsub_rn(right, left), 1.f), 0.f);
  const float h = fmaxf(__fadd_rn(__fsub_rn(bottom, top), 1.f), 0.f);
  const float inter = __fmul_rn(w, h);
  const float uni = __fsub_rn(__fadd_rn(area_a, area_b), inter);
  if (FILTER) return iou_exceeds(inter, uni, thresh, thr_hi, thr_lo);
  return __fdiv_rn(inter, uni) > thresh;
}

__device__ __forceinline__ float box_area(const float4 b) {
  return __fmul_rn(__fadd_rn(__fsub_rn(b.z, b.x), 1.f), __fadd_rn(__fsub_rn(b.w, b.y), 1.f));
}

__device__ __forceinline__ float4 load_box(const float* __restrict__ p, int stride) {
  if (stride == 4) return __ldg(reinterpret_cast<const float4*>(p));
  return make_float4(__ldg(p), __ldg(p + 1), __ldg(p + 2), __ldg(p + 3));
}

// grid (col_blocks, row_blocks, batch); 64 threads; only col_block >= row_block does work.
// mask[(img * n + row) * ncb + col_block], bits for columns > row only.
template <bool FILTER>
__global__ void __launch_bounds__(64)
    nms_mask_kernel(const float* __restrict__ boxes, int n, int stride, float thresh,
                    unsigned long long* __restrict__ mask) {
  const int col_blk = blockIdx.x, row_blk = blockIdx.y, img = blockIdx.z;
  if (col_blk < row_blk) return;
  const int ncb = gridDim.x;
  const float* bx = boxes + (size_t)img * n * stride;
  __shared__ float4 cbox[64];
  __shared__ float carea[64];
  const int col_size = min(n - col_blk * 64, 64);
  const int row_size = min(n - row_blk * 64, 64);
  const int t = threadIdx.x;
  if (t < col_size) {
    const float4 b = load_box(bx + (size_t)(col_blk * 64 + t) * stride, stride);
    cbox[t] = b;
    carea[t] = box_area(b);
  }
  __syncthreads();
  if (t >= row_size) return;
  const int row = row_blk * 64 + t;
  const float4 a = load_box(bx + (size_t)row * stride, stride);
  const float area_a = box_area(a);
  const float thr_hi = thresh * (1.f + 9.5367431640625e-7f);
  const float thr_lo = thresh * (1.f - 9.5367431640625e-7f);
  unsigned long long bits = 0;
#pragma unroll 8
  for (int j = 0; j < col_size; ++j) {
    if (box_suppresses<FILTER>(a, area_a, cbox[j], carea[j], thresh, thr_hi, thr_lo))
      bits |= 1ULL << j;
  }
  if (row_blk == col_blk) bits &= ~((2ULL << t) - 1ULL);  // keep columns > row only
  mask[((size_t)img * n + row) * ncb + col_blk] = bits;
}

// ===========================================================================
// greedy scan (one CTA per image)
// ===========================================================================
constexpr int SCAN_THREADS = 256;

__device__ __forceinline__ unsigned long long shfl_u64(unsigned long long v, int src) {
  unsigned lo = __shfl_sync(0xffffffffu, (unsigned)v, src);
  unsigned hi = __shfl_sync(0xffffffffu, (unsigned)(v >> 32), src);
  return ((unsigned long long)hi << 32) | lo;
}

__device__ __forceinline__ unsigned long long warp_or_u64(unsigned long long v) {
  unsigned lo = __reduce_or_sync(0xffffffffu, (unsigned)v);
  unsigned hi = __reduce_or_sync(0xffffffffu, (unsigned)(v >> 32));
  return ((unsigned long long)hi << 32) | lo;
}

// OR of mask[(row0 + b) * ncb + j] over the set bits b of `bits`, at most 8 loads at a time
// in flight (the loads are independent; a plain loop would serialise on L2 latency).
__device__ __forceinline__ unsigned long long or_rows(const unsigned long long* __restrict__ m,
                                                      int ncb, int row0, int j,
                                                      unsigned long long bits) {
  unsigned long long acc = 0ULL;
  while (bits) {
    unsigned long long v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      v[u] = 0ULL;
      if (bits) {
        const int b = __ffsll((long long)bits) - 1;
        bits &= bits - 1ULL;
        v[u] = m[(size_t)(row0 + b) * ncb + j];
      }
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) acc |= v[u];
  }
  return acc;
}

// keep_out[img * keep_stride + r] = r-th kept index (ascending), num_out[img] = count
// (<= max_keep).  If rois_out != NULL also writes the reference's padded
// (post, 5) block for the image: column 0 = image index, rows [0, count) = boxes.
//
// Chunk c = boxes [64c, 64c+64).  Warp 0 resolves the chunks in order; the suppression
// word of chunk c is
//     remv[c]                      rows of survivors of chunks <= c-3   (other warps, below)
//   | OR survivors(c-1) of mask[row][c]   "s1", prefetched by warp 0 before it knew the survivors
//   | OR survivors(c-2) of mask[row][c]   "s2", likewise
// so the only serial work per chunk is the in-register resolve of the 64x64 diagonal block.
// The other 7 warps fold the rows of chunk c's survivors into remv[j >= c+3]; their loads are
// issued in iteration c and consumed in iteration c+1, two barriers before warp 0 needs them.
__global__ void __launch_bounds__(SCAN_THREADS)
    nms_scan_kernel(const unsigned long long* __restrict__ mask, int n, int max_keep,
                    int* __restrict__ keep_out, int keep_stride, int* __restrict__ num_out,
                    const float* __restrict__ boxes, int box_stride, float* __restrict__ rois_out,
                    int post) {
  extern __shared__ unsigned long long remv[];  // ncb words
  __shared__ unsigned long long kept_sh[2];
  __shared__ int done_sh[2];
  constexpr int HELPERS = SCAN_THREADS - 32;
  const int img = blockIdx.x;
  const int ncb = (n + 63) >> 6;
  const unsigned long long* m = mask + (size_t)img * n * ncb;
  int* keep = keep_out + (size_t)img * keep_stride;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  for (int j = tid; j < ncb; j += SCAN_THREADS) remv[j] = 0ULL;
  __syncthreads();

  // ---- warp 0 state ----
  int nkeep = 0;
  unsigned long long d0 = 0, d1 = 0;      // diagonal words of the current chunk (rows lane, lane+32)
  unsigned long long s1a = 0, s1b = 0;    // rows of chunk c-1 at word c
  unsigned long long s2a = 0, s2b = 0;    // rows of chunk c-2 at word c
  unsigned long long kept1 = 0ULL, kept2 = 0ULL;  // survivors of chunks c-1, c-2
  if (wid == 0 && ncb > 0) {
    if (lane < n) d0 = m[(size_t)lane * ncb];
    if (lane + 32 < n) d1 = m[(size_t)(lane + 32) * ncb];
  }
  // ---- helper state: word owned this round and its loads in flight ----
  int pend_j = -1;
  unsigned long long pend[8];
#pragma unroll
  for (int u = 0; u < 8; ++u) pend[u] = 0ULL;

  for (int c = 0; c < ncb; ++c) {
    if (wid == 0) {
      // prefetch everything chunk c+1 will need that does not depend on decisions
      unsigned long long nd0 = 0, nd1 = 0, n1a = 0, n1b = 0, n2a = 0, n2b = 0;
      if (c + 1 < ncb) {  // then chunks <= c are full: all their rows exist
        const int w1 = c + 1;
        const int r0 = 64 * w1 + lane, r1 = r0 + 32;
        if (r0 < n) nd0 = m[(size_t)r0 * ncb + w1];
        if (r1 < n) nd1 = m[(size_t)r1 * ncb + w1];
        n1a = m[(size_t)(64 * c + lane) * ncb + w1];
        n1b = m[(size_t)(64 * c + 32 + lane) * ncb + w1];
        if (c >= 1) {
          n2a = m[(size_t)(64 * (c - 1) + lane) * ncb + w1];
          n2b = m[(size_t)(64 * (c - 1) + 32 + lane) * ncb + w1];
        }
      }
      unsigned long long urgent = 0ULL;
      if ((kept1 >> lane) & 1ULL) urgent |= s1a;
      if ((kept1 >> (lane + 32)) & 1ULL) urgent |= s1b;
      if ((kept2 >> lane) & 1ULL) urgent |= s2a;
      if ((kept2 >> (lane + 32)) & 1ULL) urgent |= s2b;
      urgent = warp_or_u64(urgent);
      const unsigned long long cur = remv[c] | urgent;
      const int rows = min(64, n - 64 * c);
      const unsigned long long valid = rows == 64 ? ~0ULL : ((1ULL << rows) - 1ULL);
      unsigned long long alive = ~cur & valid;
      unsigned long long kept = 0ULL;
      while (alive) {
        const int b = __ffsll((long long)alive) - 1;
        kept |= 1ULL << b;
        const unsigned long long w = shfl_u64(b < 32 ? d0 : d1, b & 31);
        alive &= ~w;
        alive &= ~(1ULL << b);
      }
      // emit indices (ascending) up to max_keep
      if ((kept >> lane) & 1ULL) {
        const int r = nkeep + __popcll(kept & ((1ULL << lane) - 1ULL));
        if (r < max_keep) keep[r] = 64 * c + lane;
      }
      if ((kept >> (lane + 32)) & 1ULL) {
        const int r = nkeep + __popcll(kept & ((1ULL << (lane + 32)) - 1ULL));
        if (r < max_keep) keep[r] = 64 * c + lane + 32;
      }
      nkeep += __popcll(kept);
      if (lane == 0) {
        kept_sh[c & 1] = kept;
        done_sh[c & 1] = (nkeep >= max_keep) ? 1 : 0;
      }
      kept2 = kept1;
      kept1 = kept;
      d0 = nd0; d1 = nd1;
      s1a = n1a; s1b = n1b; s2a = n2a; s2b = n2b;
    }
    __syncthreads();
    if (done_sh[c & 1]) break;
    if (wid != 0) {
      // consume the loads issued one iteration ago (rows of chunk c-1)
      if (pend_j >= 0) {
        unsigned long long acc = 0ULL;
#pragma unroll
        for (int u = 0; u < 8; ++u) acc |= pend[u];
        remv[pend_j] |= acc;
        pend_j = -1;
      }
      // rows of chunk c's survivors -> words j >= c+3.  A word is always handled by the
      // same thread (j mod HELPERS), so the read-modify-writes of remv never race.
      unsigned long long k = kept_sh[c & 1];
      const int t = tid - 32;
      const int first = c + 3;
      int j = first + ((t - first) % HELPERS + HELPERS) % HELPERS;
      if (k && j < ncb) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          pend[u] = 0ULL;
          if (k) {
            const int b = __ffsll((long long)k) - 1;
            k &= k - 1ULL;
            pend[u] = m[(size_t)(64 * c + b) * ncb + j];
          }
        }
        pend_j = j;
        if (k) remv[j] |= or_rows(m, ncb, 64 * c, j, k);  // more than 8 survivors in the chunk
        const unsigned long long all = kept_sh[c & 1];
        for (j += HELPERS; j < ncb; j += HELPERS) remv[j] |= or_rows(m, ncb, 64 * c, j, all);
      }
    }
  }
  __syncthreads();
  if (tid == 0) {
    // nkeep lives in warp 0 only
    num_out[img] = min(nkeep, max_keep);
    done_sh[0] = min(nkeep, max_keep);
  }
  __syncthreads();
  if (rois_out) {
    const int cnt = done_sh[0];
    float* o = rois_out + (size_t)img * post * 5;
    const float* bx = boxes + (size_t)img * n * box_stride;
    for (int r = tid; r < post; r += SCAN_THREADS) {
      float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < cnt) b = load_box(bx + (size_t)keep[r] * box_stride, box_stride);
      o[r * 5 + 0] = (float)img;
      o[r * 5 + 1] = b.x;
      o[r * 5 + 2] = b.y;
      o[r * 5 + 3] = b.z;
      o[r * 5 + 4] = b.w;
    }
  }
}

// ===========================================================================
// top-k select + sort + decode (one CTA per image)
// ===========================================================================
constexpr int TK_THREADS = 1024;
constexpr int TK_BINS = 2048;
constexpr int TK_SMEM_SORT = 16384;  // 64-bit keys sortable in shared memory

// Monotone key: larger score -> larger key.  -0 == +0, every NaN sorts first
// (torch.sort puts NaN first in descending order).
__device__ __forceinline__ unsigned score_key(float f) {
  f = f + 0.0f;
  unsigned b = __float_as_uint(f);
  if ((b & 0x7fffffffu) > 0x7f800000u) b = 0x7fc00000u;
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

struct TKShared {
  int hist[TK_BINS];
  int warp_sums[TK_THREADS / 32];
  int found_digit;
  int found_need;
  int found_count;
  int counter;
  float anchors[64 * 4];
};

// composite key of flat anchor index i: smaller = better (higher score, then lower index)
__device__ __forceinline__ unsigned long long comp_key(const float* __restrict__ fg, int K, int A,
                                                       int mem_idx, int* flat_idx) {
  const int a = mem_idx / K;
  const int k = mem_idx - a * K;
  const int i = k * A + a;
  *flat_idx = i;
  const unsigned key = score_key(__ldg(fg + mem_idx));
  return ((unsigned long long)(~key) << 32) | (unsigned)i;
}

__device__ inline void bitonic_sort_u64(unsigned long long* buf, int n_pad) {
  for (int k = 2; k <= n_pad; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < (n_pad >> 1); t += blockDim.x) {
        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        const int l = i | j;
        const unsigned long long a = buf[i], b = buf[l];
        const bool up = (i & k) == 0;
        if ((a > b) == up) {
          buf[i] = b;
          buf[l] = a;
        }
      }
      __syncthreads();
    }
  }
}

// scores (B, 2A, H, W); deltas (B, 4A, H, W); boxes_out (B, n_sorted, 4).
// gbuf: global scratch (B, n_pad) u64, used only when n_pad > TK_SMEM_SORT.
__global__ void __launch_bounds__(TK_THREADS, 1)
    proposal_topk_decode_kernel(const float* __restrict__ scores, const float* __restrict__ deltas,
                                const float* __restrict__ im_info,
                                const float* __restrict__ anchors, int A, int H, int W,
                                int feat_stride, int n_sorted, int n_pad,
                                unsigned long long* __restrict__ gbuf,
                                float* __restrict__ boxes_out, int* __restrict__ order_out) {
  extern __shared__ __align__(16) unsigned char tk_smem[];
  TKShared& sh = *reinterpret_cast<TKShared*>(tk_smem);
  unsigned long long* sbuf = reinterpret_cast<unsigned long long*>(tk_smem + sizeof(TKShared));
  const int img = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int K = H * W, N = K * A;
  const float* fg = scores + ((size_t)img * 2 * A + A) * K;  // channels [A, 2A)
  unsigned long long* buf = (n_pad <= TK_SMEM_SORT) ? sbuf : gbuf + (size_t)img * n_pad;

  for (int i = tid; i < A * 4; i += TK_THREADS) sh.anchors[i] = __ldg(anchors + i);

  // ---- 64-bit radix select: threshold `thr` with exactly n_sorted keys <= thr ----
  unsigned long long thr = ~0ULL;
  if (n_sorted < N) {
    unsigned long long prefix = 0ULL;  // value of the bits above the current digit
    int need = n_sorted;
    const int shifts[5] = {53, 42, 32, 11, 0};
    const int widths[5] = {11, 11, 10, 11, 11};
    int top = 64;  // bits [top, 64) are fixed to `prefix`
    bool finished = false;
    for (int pass = 0; pass < 5 && !finished; ++pass) {
      const int shift = shifts[pass], bits = widths[pass];
      if (pass == 3) {
        // bits [22, 32) of the index half are zero for every key (N < 2^22)
        prefix <<= 10;
        top = 22;
      }
      for (int i = tid; i < TK_BINS; i += TK_THREADS) sh.hist[i] = 0;
      __syncthreads();
      for (int mi = tid; mi < N; mi += TK_THREADS) {
        int fi;
        const unsigned long long ck = comp_key(fg, K, A, mi, &fi);
        if (top == 64 || (ck >> top) == prefix) atomicAdd(&sh.hist[(ck >> shift) & ((1u << bits) - 1u)], 1);
      }
      __syncthreads();
      // ascending scan over bins: 2 bins per thread
      const int b0 = 2 * tid, b1 = 2 * tid + 1;
      const int h0 = sh.hist[b0], h1 = sh.hist[b1];
      int incl = h0 + h1;
      for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
      }
      if (lane == 31) sh.warp_sums[wid] = incl;
      __syncthreads();
      int before = incl - (h0 + h1);
      for (int w = 0; w < wid; ++w) before += sh.warp_sums[w];
      if (before < need && before + h0 >= need) {
        sh.found_digit = b0; sh.found_need = need - before; sh.found_count = h0;
      } else if (before + h0 < need && before + h0 + h1 >= need) {
        sh.found_digit = b1; sh.found_need = need - before - h0; sh.found_count = h1;
      }
      __syncthreads();
      const int digit = sh.found_digit;
      need = sh.found_need;
      prefix = (prefix << bits) | (unsigned long long)digit;
      top = shift;
      if (sh.found_count == need || shift == 0) {
        // every key with this prefix is taken: threshold = prefix followed by ones
        thr = (shift == 0) ? prefix : ((prefix << shift) | ((1ULL << shift) - 1ULL));
        finished = true;
      }
      __syncthreads();
    }
  }

  // ---- compaction (order is irrelevant: keys are unique and get sorted) ----
  if (tid == 0) sh.counter = 0;
  __syncthreads();
  for (int mi = tid; mi < N; mi += TK_THREADS) {
    int fi;
    const unsigned long long ck = comp_key(fg, K, A, mi, &fi);
    if (ck <= thr) {
      const int pos = atomicAdd(&sh.counter, 1);
      if (pos < n_pad) buf[pos] = ck;
    }
  }
  for (int i = n_sorted + tid; i < n_pad; i += TK_THREADS) buf[i] = ~0ULL;
  __syncthreads();
  bitonic_sort_u64(buf, n_pad);

  // ---- decode + clip only the selected anchors (bbox_transform.py:77-103, :125-133) ----
  const float im_h = __ldg(im_info + img * 3 + 0), im_w = __ldg(im_info + img * 3 + 1);
  const float max_x = __fsub_rn(im_w, 1.f), max_y = __fsub_rn(im_h, 1.f);
  const float* dl = deltas + (size_t)img * 4 * A * K;
  for (int r = tid; r < n_sorted; r += TK_THREADS) {
    const int i = (int)(unsigned)buf[r];
    const int a = i % A, k = i / A;
    const int x = k % W, y = k / W;
    const float sx = (float)(x * feat_stride), sy = (float)(y * feat_stride);
    const float ax1 = __fadd_rn(sh.anchors[a * 4 + 0], sx), ay1 = __fadd_rn(sh.anchors[a * 4 + 1], sy);
    const float ax2 = __fadd_rn(sh.anchors[a * 4 + 2], sx), ay2 = __fadd_rn(sh.anchors[a * 4 + 3], sy);
    const float dx = __ldg(dl + (size_t)(a * 4 + 0) * K + k), dy = __ldg(dl + (size_t)(a * 4 + 1) * K + k);
    const float dw = __ldg(dl + (size_t)(a * 4 + 2) * K + k), dh = __ldg(dl + (size_t)(a * 4 + 3) * K + k);
    const float w = __fadd_rn(__fsub_rn(ax2, ax1), 1.0f), h = __fadd_rn(__fsub_rn(ay2, ay1), 1.0f);
    const float cx = __fadd_rn(ax1, __fmul_rn(0.5f, w)), cy = __fadd_rn(ay1, __fmul_rn(0.5f, h));
    const float pcx = __fadd_rn(__fmul_rn(dx, w), cx), pcy = __fadd_rn(__fmul_rn(dy, h), cy);
    const float pw = __fmul_rn(expf(dw), w), ph = __fmul_rn(expf(dh), h);
    const float hw = __fmul_rn(0.5f, pw), hh = __fmul_rn(0.5f, ph);
    float4 o;
    o.x = fminf(fmaxf(__fsub_rn(pcx, hw), 0.f), max_x);
    o.y = fminf(fmaxf(__fsub_rn(pcy, hh), 0.f), max_y);
    o.z = fminf(fmaxf(__fadd_rn(pcx, hw), 0.f), max_x);
    o.w = fminf(fmaxf(__fadd_rn(pcy, hh), 0.f), max_y);
    reinterpret_cast<float4*>(boxes_out)[(size_t)img * n_sorted + r] = o;
    if (order_out) order_out[(size_t)img * n_sorted + r] = i;
  }
}

// ===========================================================================
// host side
// ===========================================================================
static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
static inline int next_pow2(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

static int launch_mask_scan(const float* boxes, int batch, int n, int stride, float thresh,
                            int max_keep, unsigned long long* mask, int* keep, int keep_stride,
                            int* num, float* rois_out, int post, cudaStream_t st) {
  const int ncb = (n + 63) / 64;
  if (ncb > 65535 || batch > 65535) return TLOD_ERR_UNSUPPORTED;
  dim3 grid(ncb, ncb, batch);
  const bool filter = thresh >= 1e-6f && thresh <= 1e6f;
  {
    LaunchScope scope("nms_mask_kernel", st);
    if (filter)
      nms_mask_kernel<true><<<grid, 64, 0, st>>>(boxes, n, stride, thresh, mask);
    else
      nms_mask_kernel<false><<<grid, 64, 0, st>>>(boxes, n, stride, thresh, mask);
  }
  int rc = last_launch_status();
  if (rc) return rc;
  {
    LaunchScope scope("nms_scan_kernel", st);
    nms_scan_kernel<<<batch, SCAN_THREADS, (size_t)ncb * 8, st>>>(mask, n, max_keep, keep, keep_stride,
                                                                num, boxes, stride, rois_out, post);
  }
  return last_launch_status();
}

}  // namespace tlod

using namespace tlod;

extern "C" size_t tlod_nms_workspace_bytes(int n) {
  if (n <= 0) return 16;
  const size_t ncb = (size_t)(n + 63) / 64;
  return align_up((size_t)n * ncb * 8, 256) + 256;
}

extern "C" int tlod_nms(const float* boxes, int n, int box_stride, float thresh, int max_keep,
                        int* keep_out, int* num_out, void* workspace, size_t workspace_bytes,
                        void* stream) {
  if (!keep_out || !num_out) return TLOD_ERR_NULL_POINTER;
  if (n < 0 || box_stride < 4) return TLOD_ERR_BAD_SHAPE;
  cudaStream_t st = (cudaStream_t)stream;
  if (n == 0) {
    cudaError_t e = cudaMemsetAsync(num_out, 0, sizeof(int), st);
    return (int)e;
  }
  if (!boxes) return TLOD_ERR_NULL_POINTER;
  if (!workspace || workspace_bytes < tlod_nms_workspace_bytes(n)) return TLOD_ERR_WORKSPACE;
  if (n > (1 << 20)) return TLOD_ERR_UNSUPPORTED;
  if (max_keep <= 0 || max_keep > n) max_keep = n;
  return launch_mask_scan(boxes, 1, n, box_stride, thresh, max_keep,
                          (unsigned long long*)workspace, keep_out, n, num_out, nullptr, 0, st);
}

extern "C" int tlod_proposals_n_sorted(int batch, int num_anchors, int height, int width,
                                       int pre_nms_topN) {
  const long long N = (long long)num_anchors * height * width;
  const long long numel = N * batch;
  long long n = (pre_nms_topN > 0 && pre_nms_topN < numel) ? pre_nms_topN : N;  // proposal_layer.py:138
  if (n > N) n = N;
  return (int)n;
}

struct ProposalWs {
  size_t boxes, mask, keep, num, sortbuf, total;
};
static ProposalWs proposal_ws(int batch, int n_sorted, int post) {
  ProposalWs w;
  const size_t ncb = (size_t)(n_sorted + 63) / 64;
  const int n_pad = next_pow2(n_sorted);
  size_t off = 0;
  w.boxes = off; off = align_up(off + (size_t)batch * n_sorted * 16, 256);
  w.mask = off;  off = align_up(off + (size_t)batch * n_sorted * ncb * 8, 256);
  w.keep = off;  off = align_up(off + (size_t)batch * (post > 0 ? post : n_sorted) * 4, 256);
  w.num = off;   off = align_up(off + (size_t)batch * 4, 256);
  w.sortbuf = off;
  if (n_pad > TK_SMEM_SORT) off = align_up(off + (size_t)batch * n_pad * 8, 256);
  w.total = off + 256;
  return w;
}

extern "C" size_t tlod_proposals_workspace_bytes(int batch, int num_anchors, int height, int width,
                                                 int pre_nms_topN, int post_nms_topN) {
  if (batch <= 0 || num_anchors <= 0 || height <= 0 || width <= 0) return 0;
  const int n = tlod_proposals_n_sorted(batch, num_anchors, height, width, pre_nms_topN);
  return proposal_ws(batch, n, post_nms_topN).total;
}

extern "C" int tlod_proposals(const float* scores, const float* deltas, const float* im_info,
                              const float* anchors, float* rois_out, int batch, int num_anchors,
                              int height, int width, int feat_stride, int pre_nms_topN,
                              int post_nms_topN, float nms_thresh, int* order_out,
                              float* sorted_boxes_out, int* num_out, void* workspace,
                              size_t workspace_bytes, void* stream) {
  if (!scores || !deltas || !im_info || !anchors || !rois_out) return TLOD_ERR_NULL_POINTER;
  if (batch <= 0 || num_anchors <= 0 || height <= 0 || width <= 0 || post_nms_topN <= 0)
    return TLOD_ERR_BAD_SHAPE;
  if (num_anchors > 64) return TLOD_ERR_UNSUPPORTED;
  const long long N = (long long)num_anchors * height * width;
  if (N >= (1 << 22)) return TLOD_ERR_UNSUPPORTED;
  const int n = tlod_proposals_n_sorted(batch, num_anchors, height, width, pre_nms_topN);
  const ProposalWs w = proposal_ws(batch, n, post_nms_topN);
  if (!workspace || workspace_bytes < w.total || ((uintptr_t)workspace & 15)) return TLOD_ERR_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  unsigned char* base = (unsigned char*)workspace;
  float* boxes = sorted_boxes_out ? sorted_boxes_out : (float*)(base + w.boxes);
  if ((uintptr_t)boxes & 15) return TLOD_ERR_BAD_SHAPE;
  unsigned long long* mask = (unsigned long long*)(base + w.mask);
  int* keep = (int*)(base + w.keep);
  int* num = num_out ? num_out : (int*)(base + w.num);
  const int n_pad = next_pow2(n);
  size_t smem = sizeof(TKShared) + (n_pad <= TK_SMEM_SORT ? (size_t)n_pad * 8 : 0);
  if (smem > (size_t)device_info().max_smem_optin) return TLOD_ERR_UNSUPPORTED;
  cudaError_t e = cudaFuncSetAttribute(proposal_topk_decode_kernel,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  {
    LaunchScope scope("proposal_topk_decode_kernel", st);
    proposal_topk_decode_kernel<<<batch, TK_THREADS, smem, st>>>(
        scores, deltas, im_info, anchors, num_anchors, height, width, feat_stride, n, n_pad,
        (unsigned long long*)(base + w.sortbuf), boxes, order_out);
  }
  int rc = last_launch_status();
  if (rc) return rc;
  return launch_mask_scan(boxes, batch, n, 4, nms_thresh, post_nms_topN, mask, keep, post_nms_topN,
                          num, rois_out, post_nms_topN, st);
}
