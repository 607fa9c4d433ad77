// NMS and the fused, batched proposal layer for sm_100a.
//
// Reference semantics:
//   nms            lib/model/nms/src/nms_cuda_kernel.cu:31-39 (devIoU), :41-85 (mask),
//                  :132-144 (host greedy scan)
//   proposal layer lib/model/rpn/proposal_layer.py:49-163,
//                  lib/model/rpn/bbox_transform.py:77-103 (decode), :125-133 (clip)
//
// Pipeline for a batch (no host synchronisation, no D2H mask copy):
//   1. proposal_sort_runs_kernel + proposal_rank_decode_kernel: sort by ranking of the
//      (score desc, index asc) keys over many CTAs; only the pre_nms_topN best anchors are
//      decoded + clipped, each written straight to its rank.
//   2. nms_mask_kernel   lower-triangular 64x64 tiles of the IoU > thresh bitmask (fp32 ALU
//      bound; row x holds the EARLIER boxes that overlap box x), computed in phases: only the
//      leading S x S triangle the scan can reach before post_nms_topN boxes have survived,
//      extended x4 only if the scan asks for it.
//   3. nms_scan_kernel   one CTA per image: the greedy scan the reference runs on the host,
//      done on chip: warp 0 resolves each 64-box chunk in a few warp-wide rounds; eleven helper
//      warps, one chunk each, AND the chunk's rows with the survivors of the earlier chunks;
//      then the padded (post_nms_topN, 5) output.
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

// Mask layout, per image: word-major, mask[word j][row r] (j <= r / 64: lower triangular) with rows padded to whole chunks --
// the 64 rows of a chunk at one word are 512 contiguous bytes, so the scan fetches them with
// fully coalesced loads.  (A row-major mask made every scan load a gather of one 32-byte sector
// per lane: 32 tag lookups per instruction on the single SM that scans an image, ~2000 cycles
// per chunk.)
__host__ __device__ inline int nms_rows_padded(int n) { return ((n + 63) >> 6) << 6; }
__host__ __device__ inline size_t nms_mask_words(int n) {
  return (size_t)((n + 63) >> 6) * (size_t)nms_rows_padded(n);
}

// Per-image scan state carried from one phase to the next.
struct NmsState {
  int nkeep;  // survivors so far (<= max_keep)
  int done;   // max_keep reached or every box examined: later phases exit at once
};

// Phase p covers the boxes [S_{p-1}, S_p).  Its mask kernel fills, for the row blocks of
// the phase, the 64x64 tiles of all column blocks at or before the diagonal; the scan kernel then
// continues the greedy scan over those boxes.  S_1 is sized so that the scan normally reaches
// max_keep inside phase 1 (the IoU work is then S_1^2/2 pairs instead of n^2/2); if it does
// not, the next phase is four times larger.  Later phases find `done` set and return.
// (Tried: merging the later phases into one -- 6000 -> 300 as 1024 | 6000 instead of 1024 | 4096 | 6000 --
// saves two idle launches, ~5 us of 60 at IoU 0.7, but doubles the IoU work whenever phase 2 is
// entered: the IoU-0.3 sweep at batch 64 went from 1.08 to 1.97 ms.)
__host__ __device__ inline int nms_phase_end(int n, int max_keep, int phase /* 1-based */) {
  long long s = 2LL * max_keep;
  if (s < 1024) s = 1024;
  s = (s + 63) / 64 * 64;
  for (int p = 1; p < phase && s < n; ++p) s *= 4;
  return s < n ? (int)s : n;
}

// Persistent grid (gridDim.x CTAs per image, blockIdx.y = image); 256 threads: thread t owns
// row t & 63 of a 64x64 tile and 16 of its 64 columns (quarter t >> 6).  The tiles of a phase are the pairs
// (row block rb in [cb0, cb1), col block cb <= rb), numbered rb-major (the enumeration below finds
// the pair as (larger, smaller) and swaps the roles).
// mask[img][col_block][row] (np = padded rows); in the diagonal tile the word holds every other box of
// the chunk that overlaps the row's box (both directions).
template <bool FILTER>
__global__ void __launch_bounds__(256)
    nms_mask_kernel(const float* __restrict__ boxes, int n, int stride, float thresh,
                    unsigned long long* __restrict__ mask, int np, int cb0, int cb1, int first_phase,
                    NmsState* __restrict__ state) {
  const int img = blockIdx.y;
  const int t = threadIdx.x;
  if (first_phase) {
    if (blockIdx.x == 0 && t == 0) {
      state[img].nkeep = 0;
      state[img].done = 0;
    }
  } else if (state[img].done) {
    return;
  }
  const float* bx = boxes + (size_t)img * n * stride;
  __shared__ float4 cbox[64];
  __shared__ float carea[64];
  const float thr_hi = thresh * (1.f + 9.5367431640625e-7f);
  const float thr_lo = thresh * (1.f - 9.5367431640625e-7f);
  const long long tri0 = (long long)cb0 * (cb0 + 1) / 2;
  const long long ntiles = (long long)cb1 * (cb1 + 1) / 2 - tri0;
  // thread t: row r = t & 63 of the tile, column quarter q = t >> 6 (16 columns).  q is uniform in
  // a warp, so the column boxes are read as broadcasts (one wavefront per read; with q varying
  // inside the warp the four quarters sat on the same banks: 4-way conflicts, LSU pipe 80 % busy).
  const int r = t & 63, q = t >> 6;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    // tile + tri0 = cb (cb + 1) / 2 + rb, 0 <= rb <= cb
    const long long g = tile + tri0;
    int col_blk = (int)((sqrtf(8.f * (float)g + 1.f) - 1.f) * 0.5f);
    while ((long long)col_blk * (col_blk + 1) / 2 > g) --col_blk;
    while ((long long)(col_blk + 1) * (col_blk + 2) / 2 <= g) ++col_blk;
    int row_blk = (int)(g - (long long)col_blk * (col_blk + 1) / 2);
    {  // the scan wants the LOWER triangle: rows of the phase's chunks against the chunks at or before them
      const int big = col_blk;
      col_blk = row_blk;
      row_blk = big;
    }
    const int col_size = min(n - col_blk * 64, 64);
    const int row_size = min(n - row_blk * 64, 64);
    __syncthreads();  // previous tile's cbox readers
    if (t < col_size) {
      const float4 b = load_box(bx + (size_t)(col_blk * 64 + t) * stride, stride);
      cbox[t] = b;
      carea[t] = box_area(b);
    }
    __syncthreads();
    const int row = row_blk * 64 + (r < row_size ? r : 0);
    const float4 a = load_box(bx + (size_t)row * stride, stride);
    const float area_a = box_area(a);
    unsigned bits = 0;
#pragma unroll 8
    for (int j = 0; j < 16; ++j) {
      const int c = q * 16 + j;
      if (c < col_size && box_suppresses<FILTER>(a, area_a, cbox[c], carea[c], thresh, thr_hi, thr_lo))
        bits |= 1u << j;
    }
    if (r < row_size) {
      // diagonal tile: the full symmetric word minus the box itself (IoU is symmetric bit for
      // bit: commutative adds, min / max); the scan derives "earlier boxes that overlap me" from it
      if (row_blk == col_blk && (r >> 4) == q) bits &= ~(1u << (r & 15));
      // the row's 64-bit word is written as its four 16-bit quarters (little endian)
      unsigned short* w16 = reinterpret_cast<unsigned short*>(
          mask + (size_t)img * nms_mask_words(n) + (size_t)col_blk * np + row);
      w16[q] = (unsigned short)bits;
    }
  }
}

// ===========================================================================
// greedy scan (one CTA per image and phase)
// ===========================================================================
constexpr int SCAN_THREADS = 512;

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

// Greedy resolve of one 64-box chunk inside a warp.  `cand` = boxes not suppressed from
// outside the chunk; f0 / f1 = this lane's words (boxes lane, lane + 32) of the diagonal block:
// every other box of the chunk that overlaps it.  Only the EARLIER overlappers matter to a box, so
// all tests are lane-local and a round costs four ballots (the first version reduced row words
// across the lanes: four REDUX per round, 900 cycles per chunk).  Every round keeps ALL undecided
// boxes that no earlier undecided box overlaps (the lowest one always qualifies) and removes the
// boxes they overlap; the result equals the sequential scan (nms_cuda_kernel.cu:132-144) and the
// number of rounds is the longest suppression chain.
__device__ __forceinline__ unsigned long long ballot_u64(bool p0, bool p1) {
  const unsigned lo = __ballot_sync(0xffffffffu, p0), hi = __ballot_sync(0xffffffffu, p1);
  return ((unsigned long long)hi << 32) | lo;
}
__device__ __forceinline__ unsigned long long resolve_chunk(unsigned long long cand,
                                                            unsigned long long f0,
                                                            unsigned long long f1, int lane) {
  const unsigned long long low0 = f0 & ((1ULL << lane) - 1ULL);
  const unsigned long long low1 = f1 & ((1ULL << (lane + 32)) - 1ULL);
  unsigned long long und = cand, kept = 0ULL;
  while (und) {
    const bool u0 = (und >> lane) & 1ULL, u1 = (und >> (lane + 32)) & 1ULL;
    const unsigned long long safe = ballot_u64(u0 && !(low0 & und), u1 && !(low1 & und));
    kept |= safe;
    if (safe == und) break;
    const unsigned long long victims = ballot_u64(u0 && (low0 & safe), u1 && (low1 & safe));
    und &= ~(safe | victims);
  }
  return kept;
}

// Continues the scan over the chunks [c0, c1) of one phase.
// keep_out[img * keep_stride + r] = r-th kept index (ascending), num_out[img] = count
// (<= max_keep).  If rois_out != NULL also writes the reference's padded
// (post, 5) block for the image: column 0 = image index, rows [0, count) = boxes.
//
// The mask is LOWER triangular: the word of row x at word w <= x / 64 holds the boxes of chunk w
// that overlap box x (the mask kernel computes the transposed tiles; IoU is symmetric bit for
// bit).  Box x of chunk c is suppressed from outside its chunk iff
//     OR_{w < c} ( mask[w][x] & kept[w] ) != 0          kept[w] = survivors of chunk w
// -- a test every lane does on its own two rows: no cross-lane OR of 64-bit words (REDUX was the
// longest item on the per-chunk chain of the first design, which pushed each chunk's surviving
// rows into the words of all later chunks), no per-survivor selects, and nothing has to be
// carried in from earlier phases except kept[] itself.
//   * Warp 0 resolves the chunks in order.  The term w = c - 1 it evaluates itself (two ANDs and
//     two ballots); its own mask words arrive three chunks ahead through a cp.async ring.  The keep
//     list is written by another warp (the emitter) from kept[].
//   * The words w <= c - 2 of a chunk belong to ONE helper warp (helper h: chunks c0 + h, c0 + h + 11,
//     ...).  It requests the chunk's rows sixteen words at a time and ANDs them with kept[w]; the
//     newest words it polls for, the very last one (w = c - 2) while warp 0 works on chunk c - 1,
//     and it publishes one 64-bit suppression word.  A helper is busy with its chunk for a few
//     thousand cycles (an L2 round trip per batch); the eleven of them together deliver one chunk
//     per ~300 cycles.
//   * There is no CTA barrier in the loop: warp 0 publishes kept[c], the helpers their words,
//     through shared memory as self-validating tagged words (below).
// Measured (profiles/r02_nms_scan_variants.txt): 12000 -> 2000 boxes, 46 chunks: 55 -> 35 us per
// image.  The designs that prefetched two or three chunks ahead INSIDE a warp (the round-1 kernel and
// four rewrites) all ran at ~1.2 us per chunk whatever the distance: a warp has six scoreboards, so a
// wait for the rows requested chunks ago also waits for the ones requested just now.
// helper warps: 11 of the 16 -- warp 1 writes the keep list, warps 4, 8 and 12 stay idle so that warp 0 has its scheduler (SM
// sub-partition = warp index mod 4) to itself: next to three polling helpers its ~40 instructions of
// publish + emit took 443 cycles per chunk.  Helper h owns the chunks c0 + h, c0 + h + 11, ...
constexpr int SCAN_HW = SCAN_THREADS / 32 - SCAN_THREADS / 128 - 1;  // 11: warp 1 writes the keep list
constexpr int SCAN_BATCH = 16;                  // words a helper has in flight at a time (32 loads)
constexpr int SCAN_RING = 8;                    // partial-word slots (a helper is at most 2 chunks ahead of warp 0)

// Everything the warps hand to each other travels as SELF-VALIDATING 64-bit words: (tag << 32) | half,
// tag = chunk + 1.  A 64-bit shared-memory store is single-copy atomic, so a reader that sees the
// tag has the payload.
__device__ __forceinline__ unsigned long long scan_pack(int tag, unsigned half) {
  return ((unsigned long long)(unsigned)tag << 32) | half;
}
struct ScanShared {
  unsigned long long partial[SCAN_RING][2];  // helper -> warp 0: suppression word of a chunk, two tagged halves
  volatile int stop;                            // max_keep reached or the phase is over: helpers leave
  alignas(16) unsigned long long stage[4][32][4];  // warp 0's own mask words, four chunks in flight (cp.async)
  int nkeep;                                    // survivors so far (capped)
  volatile int c_end;                           // chunks [c0, c_end) were resolved in this phase
};

// kept[w] (survivors of chunk w) once warp 0 has published it; false: the scan has stopped
__device__ __forceinline__ bool scan_kept(const volatile unsigned long long* kp, const volatile int* stop, int w,
                                          unsigned long long* out) {
  unsigned long long a = kp[2 * w], b = kp[2 * w + 1];
  while ((int)(a >> 32) != w + 1 || (int)(b >> 32) != w + 1) {
    if (*stop) return false;
    a = kp[2 * w];
    b = kp[2 * w + 1];
  }
  *out = (b << 32) | (a & 0xffffffffULL);
  return true;
}

__device__ __forceinline__ unsigned long long scan_ldg(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.global.nc.u64 %0, [%1];" : "=l"(v) : "l"(p));
  return v;
}

__device__ __forceinline__ void scan_cp_async8(unsigned dst, const unsigned long long* src, bool valid) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst), "l"(src), "r"(valid ? 8 : 0) : "memory");
}

// One helper warp, one chunk: which of the chunk's boxes (lane: rows lane and lane + 32) are suppressed by
// survivors of the chunks <= c - 2.  SCAN_BATCH words are requested at a time and ANDed with kept[w] in
// ascending order; only the last few words can still be unresolved (then the lane polls).  The latency
// of the batches is hidden by the other 10 helpers working on the next chunks -- NOT by prefetching
// several chunks ahead inside one warp: a warp has six scoreboards, so loads issued for later chunks
// share a scoreboard with the ones being waited for, and every wait became a full L2 round trip
// whatever the prefetch distance (the common ~1.2 us per chunk of all earlier versions).
__device__ __forceinline__ bool scan_chunk(ScanShared& sh, const unsigned long long* __restrict__ m, int np,
                                           const volatile unsigned long long* kp, int c, int lane) {
  unsigned long long a0 = 0ULL, a1 = 0ULL;
  const unsigned long long* rows = m + 64 * c + lane;
  for (int wb = 0; wb <= c - 2; wb += SCAN_BATCH) {
    if (sh.stop) return false;
    unsigned long long v[SCAN_BATCH][2];
    // `asm volatile` loads: issued HERE, before the polls below.  With plain loads the compiler sinks
    // them to their first use, behind the wait for kept[w] -- one L2 round trip per chunk on the chain.
#pragma unroll
    for (int k = 0; k < SCAN_BATCH; ++k) {
      const int w = wb + k <= c - 2 ? wb + k : 0;
      v[k][0] = scan_ldg(rows + (size_t)w * np);
      v[k][1] = scan_ldg(rows + (size_t)w * np + 32);
    }
    // optimistic: the batch's kept[] words are read together and their tags checked afterwards; a word
    // that is not there yet (the newest one or two of the chunk) is polled for, in ascending order
    unsigned long long lo_w[SCAN_BATCH], hi_w[SCAN_BATCH];
#pragma unroll
    for (int k = 0; k < SCAN_BATCH; ++k) {
      const int w = wb + k <= c - 2 ? wb + k : 0;
      lo_w[k] = kp[2 * w];
      hi_w[k] = kp[2 * w + 1];
    }
#pragma unroll
    for (int k = 0; k < SCAN_BATCH; ++k) {
      const int w = wb + k;
      if (w <= c - 2) {
        unsigned long long km = (hi_w[k] << 32) | (lo_w[k] & 0xffffffffULL);
        if ((int)(lo_w[k] >> 32) != w + 1 || (int)(hi_w[k] >> 32) != w + 1) {
          if (!scan_kept(kp, &sh.stop, w, &km)) return false;
        }
        a0 |= v[k][0] & km;
        a1 |= v[k][1] & km;
      }
    }
  }
  const unsigned lo = __ballot_sync(0xffffffffu, a0 != 0ULL), hi = __ballot_sync(0xffffffffu, a1 != 0ULL);
  if (lane == 0) {
    volatile unsigned long long* p = sh.partial[c % SCAN_RING];
    p[0] = scan_pack(c + 1, lo);
    p[1] = scan_pack(c + 1, hi);
  }
  return true;
}

__global__ void __launch_bounds__(SCAN_THREADS)
    nms_scan_kernel(const unsigned long long* __restrict__ mask, int n, int max_keep,
                    int* __restrict__ keep_out, int keep_stride, int* __restrict__ num_out,
                    const float* __restrict__ boxes, int box_stride, float* __restrict__ rois_out,
                    int post, int c0, int c1, int last_phase, NmsState* __restrict__ state, int np, int ncb) {
  extern __shared__ unsigned long long kp_smem[];  // 2 * ncb tagged halves: survivors per chunk
  volatile unsigned long long* kp = kp_smem;
  __shared__ ScanShared sh;
  const int img = blockIdx.x;
  if (state[img].done) return;
  const unsigned long long* m = mask + (size_t)img * nms_mask_words(n);  // m[word * np + row], word <= row / 64
  int* keep = keep_out + (size_t)img * keep_stride;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int nkeep0 = state[img].nkeep;

  // ---- kept[] of the earlier phases, from the keep list (chunks of this phase: tag 0 = not resolved) ----
  for (int j = tid; j < 2 * c1; j += SCAN_THREADS) kp_smem[j] = (j >> 1) < c0 ? scan_pack((j >> 1) + 1, 0u) : 0ULL;
  if (tid < SCAN_RING * 2) (&sh.partial[0][0])[tid] = 0ULL;
  if (tid == 0) sh.stop = 0;
  __syncthreads();
  for (int k = tid; k < nkeep0; k += SCAN_THREADS) {
    const int idx = keep[k];
    atomicOr(&kp_smem[2 * (idx >> 6) + ((idx >> 5) & 1)], 1ULL << (idx & 31));
  }
  __syncthreads();

  if (wid == 0) {
    // ---- warp 0: the chain ----
    int nkeep = nkeep0;
    unsigned long long kept1 = 0ULL;
    if (c0 > 0) kept1 = (kp[2 * c0 - 1] << 32) | (kp[2 * c0 - 2] & 0xffffffffULL);
    const unsigned stage_addr = (unsigned)__cvta_generic_to_shared(&sh.stage[0][0][0]);
    // Warp 0's own mask words -- T = this lane's rows at word c - 1, D = at the diagonal word c -- travel
    // through a 4-slot shared-memory ring by cp.async, three chunks ahead.  Plain loads into rotating
    // registers do not work here: a warp has six scoreboards, the compiler lets the load requested in this
    // iteration share one with the load consumed in it, and every third chunk then waited a full L2 round
    // trip (clock64: 650 of ~900 cycles).  cp.async.wait_group waits for the OLDEST group only.
    auto request = [&](int cc) {
      const int r0 = 64 * cc + lane, r1 = r0 + 32;
      const unsigned dst = stage_addr + (unsigned)((cc & 3) * 1024 + lane * 32);
      const bool in = cc < c1;
      const unsigned long long* d0 = m + (size_t)(in ? cc : 0) * np + (in && r0 < n ? r0 : 0);
      const unsigned long long* d1 = m + (size_t)(in ? cc : 0) * np + (in && r1 < n ? r1 : 0);
      const unsigned long long* t0 = m + (size_t)(in && cc > 0 ? cc - 1 : 0) * np + (in && r0 < n ? r0 : 0);
      const unsigned long long* t1 = m + (size_t)(in && cc > 0 ? cc - 1 : 0) * np + (in && r1 < n ? r1 : 0);
      // src-size 0 = zero fill (rows past n, chunk 0 has no word c - 1, chunks past the phase)
      scan_cp_async8(dst, t0, in && cc > 0 && r0 < n);
      scan_cp_async8(dst + 8, t1, in && cc > 0 && r1 < n);
      scan_cp_async8(dst + 16, d0, in && r0 < n);
      scan_cp_async8(dst + 24, d1, in && r1 < n);
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    request(c0);
    request(c0 + 1);
    request(c0 + 2);
    int c_end = c1;
    for (int c = c0; c < c1; ++c) {
      request(c + 3);
      // suppression from the chunks <= c - 2: the helper's word
      unsigned long long cur = 0ULL;
      {
        const volatile unsigned long long* p = sh.partial[c % SCAN_RING];
        unsigned long long x = p[0], y = p[1];
        while ((int)(x >> 32) != c + 1 || (int)(y >> 32) != c + 1) {
          x = p[0];
          y = p[1];
        }
        cur = (y << 32) | (x & 0xffffffffULL);
      }
      asm volatile("cp.async.wait_group 3;" ::: "memory");  // the group of chunk c (three younger ones may be pending)
      unsigned long long T0, T1, D0, D1;
      {
        const unsigned src = stage_addr + (unsigned)((c & 3) * 1024 + lane * 32);
        asm volatile("ld.shared.v2.u64 {%0, %1}, [%2];" : "=l"(T0), "=l"(T1) : "r"(src));
        asm volatile("ld.shared.v2.u64 {%0, %1}, [%2];" : "=l"(D0), "=l"(D1) : "r"(src + 16));
      }
      // ... and from chunk c - 1
      cur |= ballot_u64((T0 & kept1) != 0ULL, (T1 & kept1) != 0ULL);
      const int rows = min(64, n - 64 * c);
      const unsigned long long valid = rows == 64 ? ~0ULL : ((1ULL << rows) - 1ULL);
      const unsigned long long kk = resolve_chunk(~cur & valid, D0, D1, lane);
      if (lane == 0) {
        kp[2 * c] = scan_pack(c + 1, (unsigned)kk);
        kp[2 * c + 1] = scan_pack(c + 1, (unsigned)(kk >> 32));
      }
      // (the keep list is written by the emitter warp, from kept[])
      nkeep += __popcll(kk);
      kept1 = kk;
      if (nkeep >= max_keep) {
        c_end = c + 1;
        break;
      }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    if (lane == 0) {
      sh.nkeep = min(nkeep, max_keep);
      sh.c_end = c_end;
      sh.stop = 1;
    }
  } else if (wid == 1) {
    // ---- emitter: the keep list (ascending, up to max_keep) from kept[], off warp 0's chain ----
    int nkeep = nkeep0;
    for (int c = c0; c < c1; ++c) {
      unsigned long long a = kp[2 * c], b = kp[2 * c + 1];
      bool there = true;
      while ((int)(a >> 32) != c + 1 || (int)(b >> 32) != c + 1) {
        if (sh.stop && c >= sh.c_end) {  // stop is written after c_end and after the last kept[] word
          there = false;
          break;
        }
        a = kp[2 * c];
        b = kp[2 * c + 1];
      }
      if (!there) break;
      const unsigned long long kk = (b << 32) | (a & 0xffffffffULL);
      if ((kk >> lane) & 1ULL) {
        const int r = nkeep + __popcll(kk & ((1ULL << lane) - 1ULL));
        if (r < max_keep) keep[r] = 64 * c + lane;
      }
      if ((kk >> (lane + 32)) & 1ULL) {
        const int r = nkeep + __popcll(kk & ((1ULL << (lane + 32)) - 1ULL));
        if (r < max_keep) keep[r] = 64 * c + lane + 32;
      }
      nkeep += __popcll(kk);
      if (nkeep >= max_keep) break;
    }
  } else if ((wid & 3) != 0) {
    // ---- helpers ----
    for (int c = c0 + (wid - 2 - wid / 4); c < c1; c += SCAN_HW)
      if (!scan_chunk(sh, m, np, kp, c, lane)) break;
  }
  __syncthreads();
  const int cnt = sh.nkeep;
  const bool finished = cnt >= max_keep || last_phase;
  if (tid == 0) {
    state[img].nkeep = cnt;
    state[img].done = finished ? 1 : 0;
    if (finished) num_out[img] = cnt;
  }
  if (finished && rois_out) {
    float* o = rois_out + (size_t)img * post * 5;
    const float* bx = boxes + (size_t)img * n * box_stride;
    // four rows per thread and pass: the index and box loads of a pass are all in flight together
    for (int r0 = tid; r0 < post; r0 += 4 * SCAN_THREADS) {
      int idx[4];
      float4 b[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int r = r0 + q * SCAN_THREADS;
        idx[q] = r < cnt ? keep[r] : -1;
      }
#pragma unroll
      for (int q = 0; q < 4; ++q)
        b[q] = idx[q] >= 0 ? load_box(bx + (size_t)idx[q] * box_stride, box_stride) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int r = r0 + q * SCAN_THREADS;
        if (r < post) {
          o[r * 5 + 0] = (float)img;
          o[r * 5 + 1] = b[q].x;
          o[r * 5 + 2] = b[q].y;
          o[r * 5 + 3] = b[q].z;
          o[r * 5 + 4] = b[q].w;
        }
      }
    }
  }
}

// ===========================================================================
// top-k + sort + decode: sort by ranking
// ===========================================================================
// Keys are unique 64-bit composites (~score_key << 32 | flat anchor index): smaller = better
// (higher score, then lower index = the stable descending sort of proposal_layer.py:125).
//   1. proposal_sort_runs_kernel: every CTA bitonic-sorts one run of TK_RUN keys in shared
//      memory and writes it to the workspace.
//   2. proposal_rank_decode_kernel: every key finds its global rank = its position in its own
//      run + the number of smaller keys in every other run (binary searches over the
//      L2-resident runs); if rank < n_sorted the anchor is decoded, clipped and written
//      straight to row `rank`.  No merge passes, no single-CTA serial section.
constexpr int TK_THREADS = 1024;
constexpr int TK_RUN = 2048;

// Monotone key: larger score -> larger key.  -0 == +0, every NaN sorts first
// (torch.sort puts NaN first in descending order).
__device__ __forceinline__ unsigned score_key(float f) {
  f = f + 0.0f;
  unsigned b = __float_as_uint(f);
  if ((b & 0x7fffffffu) > 0x7f800000u) b = 0x7fc00000u;
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

// Bitonic stages j = jmax .. 1 (jmax <= 32) of merge size k on the two keys (elements 2t and
// 2t + 1) a thread holds: partners for j >= 2 are in the same warp (lane ^ j/2).
__device__ __forceinline__ void bitonic_reg_phase(unsigned long long& a, unsigned long long& b, int t,
                                                  int k, int jmax) {
  const bool asc = ((2 * t) & k) == 0;
  for (int j = jmax; j >= 2; j >>= 1) {
    const unsigned long long pa = __shfl_xor_sync(0xffffffffu, a, j >> 1);
    const unsigned long long pb = __shfl_xor_sync(0xffffffffu, b, j >> 1);
    const bool keep_min = (((2 * t) & j) == 0) == asc;
    a = keep_min ? (a < pa ? a : pa) : (a > pa ? a : pa);
    b = keep_min ? (b < pb ? b : pb) : (b > pb ? b : pb);
  }
  if ((a > b) == asc) {
    const unsigned long long x = a;
    a = b;
    b = x;
  }
}

// scores (B, 2A, H, W) -> runs (B, nruns, TK_RUN) sorted ascending, padded with ~0.
// Stages with j <= 32 run in registers / shuffles; only the 15 stages with j >= 64 go through
// shared memory.
__global__ void __launch_bounds__(TK_THREADS)
    proposal_sort_runs_kernel(const float* __restrict__ scores, int A, int K,
                              unsigned long long* __restrict__ runs) {
  static_assert(TK_RUN == 2 * TK_THREADS, "two keys per thread");
  __shared__ unsigned long long buf[TK_RUN];
  const int run = blockIdx.x, img = blockIdx.y, t = threadIdx.x;
  const int N = K * A;
  const float* fg = scores + ((size_t)img * 2 * A + A) * K;  // channels [A, 2A)
  unsigned long long key[2];
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    const int mi = run * TK_RUN + 2 * t + e;  // memory order: a * K + k
    key[e] = ~0ULL;
    if (mi < N) {
      const int a = mi / K, k = mi - a * K;
      key[e] = ((unsigned long long)(~score_key(__ldg(fg + mi))) << 32) | (unsigned)(k * A + a);
    }
  }
  unsigned long long a = key[0], b = key[1];
  for (int k = 2; k <= 64; k <<= 1) bitonic_reg_phase(a, b, t, k, k >> 1);
  buf[2 * t] = a;
  buf[2 * t + 1] = b;
  __syncthreads();
  for (int k = 128; k <= TK_RUN; k <<= 1) {
    for (int j = k >> 1; j >= 64; j >>= 1) {
      const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
      const int l = i | j;
      const unsigned long long x = buf[i], y = buf[l];
      if ((x > y) == ((i & k) == 0)) {
        buf[i] = y;
        buf[l] = x;
      }
      __syncthreads();
    }
    a = buf[2 * t];
    b = buf[2 * t + 1];
    bitonic_reg_phase(a, b, t, k, 32);
    if (k < TK_RUN) {
      buf[2 * t] = a;
      buf[2 * t + 1] = b;
      __syncthreads();
    }
  }
  unsigned long long* out = runs + ((size_t)img * gridDim.x + run) * TK_RUN;
  reinterpret_cast<ulonglong2*>(out)[t] = make_ulonglong2(a, b);
}

// number of keys < key in a sorted run of TK_RUN keys (keys are unique)
__device__ __forceinline__ int run_lower_bound(const unsigned long long* __restrict__ r,
                                               unsigned long long key) {
  int lo = 0;
#pragma unroll
  for (int step = TK_RUN / 2; step > 0; step >>= 1)
    if (__ldg(r + lo + step - 1) < key) lo += step;
  return lo + (__ldg(r + lo) < key ? 1 : 0);
}

// Same count with the first 6 of the 11 steps served from shared memory: smp[b] = last key of
// the b-th block of 32 keys of the run.
constexpr int TK_SMP = TK_RUN / 32;  // samples per run
constexpr int TK_SMP_RUNS = 64;      // runs whose samples fit the shared-memory table
// number of 32-key blocks of a run whose last key is < key (6 steps in the shared sample table)
__device__ __forceinline__ int run_blocks_below(const unsigned long long* __restrict__ smp, unsigned long long key) {
  int blk = 0;
#pragma unroll
  for (int step = TK_SMP / 2; step > 0; step >>= 1)
    if (smp[blk + step - 1] < key) blk += step;
  blk += (smp[blk] < key) ? 1 : 0;
  return blk;
}
__device__ __forceinline__ int run_lower_bound_2level(const unsigned long long* __restrict__ r,
                                                      const unsigned long long* __restrict__ smp,
                                                      unsigned long long key) {
  const int blk = run_blocks_below(smp, key);
  if (blk == TK_SMP) return TK_RUN;
  const unsigned long long* q = r + blk * 32;
  int lo = 0;  // keys of the block that are < key (its last key is >= key)
#pragma unroll
  for (int step = 16; step > 0; step >>= 1)
    if (__ldg(q + lo + step - 1) < key) lo += step;
  return blk * 32 + lo;
}

// grid (2 * nruns, batch): one key per thread.
__global__ void __launch_bounds__(TK_THREADS)
    proposal_rank_decode_kernel(const unsigned long long* __restrict__ runs,
                                const float* __restrict__ deltas, const float* __restrict__ im_info,
                                const float* __restrict__ anchors, int A, int H, int W, int feat_stride,
                                int n_sorted, float* __restrict__ boxes_out, int* __restrict__ order_out) {
  __shared__ float sh_anchors[64 * 4];
  __shared__ unsigned long long smp[TK_SMP_RUNS * TK_SMP];
  const int run = blockIdx.x >> 1, img = blockIdx.y, tid = threadIdx.x, nruns = gridDim.x >> 1;
  const int K = H * W;
  const unsigned long long* base = runs + (size_t)img * nruns * TK_RUN;
  const int p = (blockIdx.x & 1) * TK_THREADS + tid;
  if ((blockIdx.x & 1) * TK_THREADS >= n_sorted) return;  // whole CTA: rank >= p >= n_sorted
  for (int i = tid; i < A * 4; i += TK_THREADS) sh_anchors[i] = __ldg(anchors + i);
  const bool two_level = nruns <= TK_SMP_RUNS;
  if (two_level)
    for (int i = tid; i < nruns * TK_SMP; i += TK_THREADS)
      smp[i] = __ldg(base + (size_t)(i / TK_SMP) * TK_RUN + (i % TK_SMP) * 32 + 31);
  __syncthreads();
  const float im_h = __ldg(im_info + img * 3 + 0), im_w = __ldg(im_info + img * 3 + 1);
  const float max_x = __fsub_rn(im_w, 1.f), max_y = __fsub_rn(im_h, 1.f);
  const float* dl = deltas + (size_t)img * 4 * A * K;
  {
    if (p >= n_sorted) return;  // rank >= p
    const unsigned long long key = __ldg(base + (size_t)run * TK_RUN + p);
    if (key == ~0ULL) return;   // padding
    int rank = p;
    if (two_level) {
      // A lower bound of the rank from the shared sample tables alone (whole 32-key blocks below the
      // key): most keys of a run cannot reach the top n_sorted (6000 or 12000 of 33300) and leave here,
      // before any of the searches that go to L2.  Keys of a warp are neighbours in their run, so
      // warps leave as a whole.
      // Pays when few keys survive (TEST: 6000 of 33300, -4 us of 21); with 12000 of 33300 the extra
      // pass costs more than it saves.
      if (3LL * n_sorted < (long long)nruns * TK_RUN) {
        int lb = p;
        for (int r2 = 0; r2 < nruns; ++r2)
          if (r2 != run) lb += 32 * run_blocks_below(smp + r2 * TK_SMP, key);
        if (lb >= n_sorted) return;
      }
      // independent searches, eight in flight
      int r2 = 0;
      for (; r2 + 8 <= nruns; r2 += 8) {
        int c[8];
#pragma unroll
        for (int u = 0; u < 8; ++u)
          c[u] = (r2 + u == run) ? 0
                                 : run_lower_bound_2level(base + (size_t)(r2 + u) * TK_RUN,
                                                          smp + (r2 + u) * TK_SMP, key);
#pragma unroll
        for (int u = 0; u < 8; ++u) rank += c[u];
      }
      for (; r2 < nruns; ++r2)
        if (r2 != run) rank += run_lower_bound_2level(base + (size_t)r2 * TK_RUN, smp + r2 * TK_SMP, key);
    } else {
      for (int r2 = 0; r2 < nruns; ++r2)
        if (r2 != run) rank += run_lower_bound(base + (size_t)r2 * TK_RUN, key);
    }
    if (rank >= n_sorted) return;
    // ---- decode + clip (bbox_transform.py:77-103, :125-133), every op individually rounded ----
    const int i = (int)(unsigned)key;
    const int a = i % A, k = i / A;
    const int x = k % W, y = k / W;
    const float sx = (float)(x * feat_stride), sy = (float)(y * feat_stride);
    const float ax1 = __fadd_rn(sh_anchors[a * 4 + 0], sx), ay1 = __fadd_rn(sh_anchors[a * 4 + 1], sy);
    const float ax2 = __fadd_rn(sh_anchors[a * 4 + 2], sx), ay2 = __fadd_rn(sh_anchors[a * 4 + 3], sy);
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
    reinterpret_cast<float4*>(boxes_out)[(size_t)img * n_sorted + rank] = o;
    if (order_out) order_out[(size_t)img * n_sorted + rank] = i;
  }
}

// ===========================================================================
// test-time per-class NMS, all classes of an image at once
// ===========================================================================
// methods/DAF/DAF_test.py:302-320 (same loop in every *_test.py): for each class j >= 1 keep the
// detections with score > thresh, sort them by score, run nms(cfg.TEST.NMS) -- 8 sort + nms +
// .cpu() rounds per image in the reference.  Here one CTA per class selects and sorts its
// candidates (score descending, row index ascending on ties) and writes them as one "image" of a
// batched NMS; rows that fail the threshold become far-away unit boxes behind the real ones, so
// the batched mask / scan kernels run unchanged and the real survivors are the prefix of the keep
// list with index < count.
constexpr int CN_THREADS = 1024;
constexpr int CN_MAXR = 2048;

__global__ void __launch_bounds__(CN_THREADS)
    class_sort_kernel(const float* __restrict__ scores, const float* __restrict__ boxes, int R, int K,
                      int first_class, int box_cols, float score_thresh, int Rp,
                      float* __restrict__ dets_out, int* __restrict__ order_out, int* __restrict__ count_out) {
  __shared__ unsigned long long buf[CN_MAXR];
  __shared__ int cnt;
  const int cls = first_class + blockIdx.x, tid = threadIdx.x;
  if (tid == 0) cnt = 0;
  __syncthreads();
  int mine = 0;
  for (int r = tid; r < Rp; r += CN_THREADS) {
    unsigned long long key = ~0ULL;
    if (r < R) {
      const float sc = __ldg(scores + (size_t)r * K + cls);
      if (sc > score_thresh) {  // NaN fails, as in the reference's `scores[:, j] > thresh`
        key = ((unsigned long long)(~score_key(sc)) << 32) | (unsigned)r;
        ++mine;
      }
    }
    buf[r] = key;
  }
  if (mine) atomicAdd(&cnt, mine);
  __syncthreads();
  for (int k = 2; k <= Rp; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = tid; t < (Rp >> 1); t += CN_THREADS) {
        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        const int l = i | j;
        const unsigned long long x = buf[i], y = buf[l];
        if ((x > y) == ((i & k) == 0)) {
          buf[i] = y;
          buf[l] = x;
        }
      }
      __syncthreads();
    }
  }
  const int n = cnt;
  float* d = dets_out + (size_t)blockIdx.x * Rp * 5;
  for (int r = tid; r < Rp; r += CN_THREADS) {
    float4 b;
    float sc;
    int src = -1;
    if (r < n) {
      src = (int)(unsigned)buf[r];
      const float* bp = boxes + (size_t)src * box_cols + (box_cols > 4 ? cls * 4 : 0);
      b = make_float4(__ldg(bp), __ldg(bp + 1), __ldg(bp + 2), __ldg(bp + 3));
      sc = __ldg(scores + (size_t)src * K + cls);
    } else {  // filler: unit boxes far from everything and from each other (IoU 0 with all)
      const float x = 3.0e7f + 1024.f * (float)r;
      b = make_float4(x, 3.0e7f, x, 3.0e7f);
      sc = -INFINITY;
    }
    d[r * 5 + 0] = b.x; d[r * 5 + 1] = b.y; d[r * 5 + 2] = b.z; d[r * 5 + 3] = b.w; d[r * 5 + 4] = sc;
    order_out[(size_t)blockIdx.x * Rp + r] = src;
  }
  if (tid == 0) count_out[blockIdx.x] = n;
}

// valid_out[c] = survivors that are real detections = keep entries < count (they form a prefix)
__global__ void class_valid_kernel(const int* __restrict__ keep, const int* __restrict__ num,
                                   const int* __restrict__ count, int Rp, int* __restrict__ valid_out) {
  const int c = blockIdx.x;
  int v = 0;
  for (int i = threadIdx.x; i < num[c]; i += blockDim.x) v += keep[(size_t)c * Rp + i] < count[c] ? 1 : 0;
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  __shared__ int s[32];
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += s[w];
    valid_out[c] = t;
  }
}

// ===========================================================================
// host side
// ===========================================================================
static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static int launch_mask_scan(const float* boxes, int batch, int n, int stride, float thresh,
                            int max_keep, unsigned long long* mask, int* keep, int keep_stride,
                            int* num, float* rois_out, int post, NmsState* state, cudaStream_t st) {
  const int ncb = (n + 63) / 64, np = nms_rows_padded(n);
  if (batch > 65535) return TLOD_ERR_UNSUPPORTED;
  const bool filter = thresh >= 1e-6f && thresh <= 1e6f;
  int prev = 0;
  for (int phase = 1; prev < n; ++phase) {
    const int end = nms_phase_end(n, max_keep, phase);
    const int c0 = prev / 64, c1 = (end + 63) / 64;  // prev is a multiple of 64
    const long long ntiles = (long long)c1 * (c1 + 1) / 2 - (long long)c0 * (c0 + 1) / 2;
    long long per_img = ((long long)device_info().sm_count * 8 + batch - 1) / batch;  // 8 CTAs per SM
    if (per_img > ntiles) per_img = ntiles;
    if (per_img < 1) per_img = 1;
    dim3 grid((unsigned)per_img, batch);
    {
      LaunchScope scope("nms_mask_kernel", st);
      if (filter)
        nms_mask_kernel<true><<<grid, 256, 0, st>>>(boxes, n, stride, thresh, mask, np, c0, c1, phase == 1,
                                                    state);
      else
        nms_mask_kernel<false><<<grid, 256, 0, st>>>(boxes, n, stride, thresh, mask, np, c0, c1, phase == 1,
                                                     state);
    }
    int rc = last_launch_status();
    if (rc) return rc;
    {
      // ncb * 16 bytes of dynamic shared memory: beyond 48 KB (n > 196 K boxes) opt in, beyond the
      // device limit refuse instead of failing the launch
      const size_t scan_smem = (size_t)ncb * 16;  // kept[]: two tagged halves per chunk
      if (scan_smem > 48 * 1024) {
        if (scan_smem > (size_t)device_info().max_smem_optin) return TLOD_ERR_UNSUPPORTED;
        cudaError_t e = cudaFuncSetAttribute(nms_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)scan_smem);
        if (e != cudaSuccess) return (int)e;
      }
      LaunchScope scope("nms_scan_kernel", st);
      nms_scan_kernel<<<batch, SCAN_THREADS, scan_smem, st>>>(mask, n, max_keep, keep, keep_stride, num, boxes,
                                                              stride, rois_out, post, c0, c1, end >= n, state, np,
                                                              ncb);
    }
    rc = last_launch_status();
    if (rc) return rc;
    prev = end;
  }
  return TLOD_OK;
}

}  // namespace tlod

using namespace tlod;

extern "C" size_t tlod_nms_workspace_bytes(int n) {
  if (n <= 0) return 16;
  return align_up(nms_mask_words(n) * 8, 256) + 256;  // mask + NmsState
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
  if ((uintptr_t)workspace & 31) return TLOD_ERR_WORKSPACE;
  NmsState* state = (NmsState*)((unsigned char*)workspace + align_up(nms_mask_words(n) * 8, 256));
  return launch_mask_scan(boxes, 1, n, box_stride, thresh, max_keep,
                          (unsigned long long*)workspace, keep_out, n, num_out, nullptr, 0, state, st);
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
  size_t boxes, mask, keep, num, state, sortbuf, total;
};
static ProposalWs proposal_ws(int batch, int n_sorted, int post, int nruns) {
  ProposalWs w;
  size_t off = 0;
  w.boxes = off; off = align_up(off + (size_t)batch * n_sorted * 16, 256);
  w.mask = off;  off = align_up(off + (size_t)batch * nms_mask_words(n_sorted) * 8, 256);
  w.keep = off;  off = align_up(off + (size_t)batch * (post > 0 ? post : n_sorted) * 4, 256);
  w.num = off;   off = align_up(off + (size_t)batch * 4, 256);
  w.state = off; off = align_up(off + (size_t)batch * sizeof(NmsState), 256);
  w.sortbuf = off;
  off = align_up(off + (size_t)batch * nruns * TK_RUN * 8, 256);
  w.total = off + 256;
  return w;
}

extern "C" size_t tlod_proposals_workspace_bytes(int batch, int num_anchors, int height, int width,
                                                 int pre_nms_topN, int post_nms_topN) {
  if (batch <= 0 || num_anchors <= 0 || height <= 0 || width <= 0) return 0;
  const int n = tlod_proposals_n_sorted(batch, num_anchors, height, width, pre_nms_topN);
  const int nruns = (num_anchors * height * width + TK_RUN - 1) / TK_RUN;
  return proposal_ws(batch, n, post_nms_topN, nruns).total;
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
  const int nruns = (int)((N + TK_RUN - 1) / TK_RUN);
  if (nruns > 32767 || batch > 65535) return TLOD_ERR_UNSUPPORTED;
  const ProposalWs w = proposal_ws(batch, n, post_nms_topN, nruns);
  if (!workspace || workspace_bytes < w.total || ((uintptr_t)workspace & 31)) return TLOD_ERR_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  unsigned char* base = (unsigned char*)workspace;
  float* boxes = sorted_boxes_out ? sorted_boxes_out : (float*)(base + w.boxes);
  if ((uintptr_t)boxes & 15) return TLOD_ERR_BAD_SHAPE;
  unsigned long long* mask = (unsigned long long*)(base + w.mask);
  int* keep = (int*)(base + w.keep);
  int* num = num_out ? num_out : (int*)(base + w.num);
  unsigned long long* runs = (unsigned long long*)(base + w.sortbuf);
  {
    LaunchScope scope("proposal_sort_runs_kernel", st);
    proposal_sort_runs_kernel<<<dim3(nruns, batch), TK_THREADS, 0, st>>>(scores, num_anchors, height * width, runs);
  }
  int rc = last_launch_status();
  if (rc) return rc;
  {
    LaunchScope scope("proposal_rank_decode_kernel", st);
    proposal_rank_decode_kernel<<<dim3(2 * nruns, batch), TK_THREADS, 0, st>>>(
        runs, deltas, im_info, anchors, num_anchors, height, width, feat_stride, n, boxes, order_out);
  }
  rc = last_launch_status();
  if (rc) return rc;
  return launch_mask_scan(boxes, batch, n, 4, nms_thresh, post_nms_topN, mask, keep, post_nms_topN,
                          num, rois_out, post_nms_topN, (NmsState*)(base + w.state), st);
}

static int cn_pow2(int r) {
  int p = 64;
  while (p < r) p <<= 1;
  return p;
}

extern "C" int tlod_class_nms_padded_rows(int num_rois) { return num_rois > 0 ? cn_pow2(num_rois) : 0; }

extern "C" size_t tlod_class_nms_workspace_bytes(int num_rois, int num_classes) {
  if (num_rois <= 0 || num_classes <= 0) return 256;
  const int Rp = cn_pow2(num_rois);
  return align_up((size_t)num_classes * nms_mask_words(Rp) * 8, 256) +
         align_up((size_t)num_classes * sizeof(NmsState), 256) + 256;
}

extern "C" int tlod_class_nms(const float* scores, const float* boxes, int num_rois, int num_classes,
                              int first_class, int box_cols, float score_thresh, float nms_thresh,
                              float* dets_out, int* order_out, int* keep_out, int* num_out, int* count_out,
                              int* valid_out, void* workspace, size_t workspace_bytes, void* stream) {
  if (!scores || !boxes || !dets_out || !order_out || !keep_out || !num_out || !count_out || !valid_out)
    return TLOD_ERR_NULL_POINTER;
  if (num_rois <= 0 || num_classes <= 0 || first_class < 0 || first_class >= num_classes ||
      (box_cols != 4 && box_cols != 4 * num_classes))
    return TLOD_ERR_BAD_SHAPE;
  if (num_rois > CN_MAXR) return TLOD_ERR_UNSUPPORTED;
  if (!workspace || workspace_bytes < tlod_class_nms_workspace_bytes(num_rois, num_classes) ||
      ((uintptr_t)workspace & 31))
    return TLOD_ERR_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  const int Rp = cn_pow2(num_rois), nc = num_classes - first_class;
  unsigned char* base = (unsigned char*)workspace;
  unsigned long long* mask = (unsigned long long*)base;
  NmsState* state = (NmsState*)(base + align_up((size_t)num_classes * nms_mask_words(Rp) * 8, 256));
  {
    LaunchScope scope("class_sort_kernel", st);
    class_sort_kernel<<<nc, CN_THREADS, 0, st>>>(scores, boxes, num_rois, num_classes, first_class, box_cols,
                                                 score_thresh, Rp, dets_out, order_out, count_out);
  }
  int rc = last_launch_status();
  if (rc) return rc;
  rc = launch_mask_scan(dets_out, nc, Rp, 5, nms_thresh, Rp, mask, keep_out, Rp, num_out, nullptr, 0, state, st);
  if (rc) return rc;
  {
    LaunchScope scope("class_valid_kernel", st);
    class_valid_kernel<<<nc, 128, 0, st>>>(keep_out, num_out, count_out, Rp, valid_out);
  }
  return last_launch_status();
}
