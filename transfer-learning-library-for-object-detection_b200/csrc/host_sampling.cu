// Host side of the anchor-target layer's random subsampling
// (lib/model/rpn/anchor_target_layer.py:118-145), bit-compatible with numpy's global
// `np.random` stream.
//
// The reference disables surplus foreground / background anchors with
// `np.random.permutation(n)` -- numpy's legacy RandomState: MT19937, `permutation(n)` =
// Fisher-Yates over arange(n) from the top (i = n-1 .. 1, j = random_interval(i), swap), and
// `random_interval(max)` = 32-bit draws masked to the next 2^k - 1 with rejection (64-bit draws
// above 2^32 - 1).  The position of the stream after a call depends on the data, so the draws
// cannot move to the device without changing which anchors every later call selects.  numpy's own
// shuffle costs ~20 ns per element (generic byte-swap path, an unpredictable rejection branch per
// draw): 345 us for the ~17 000 background anchors of one 600x1200 image, which made this host
// loop the longest chain of the whole training step.  Here the same draws run branch-free on
// int32 indices, fused with the label scans, directly on numpy's own MT19937 state
// (np.random.mtrand._rand._bit_generator.ctypes.state_address: key[624] followed by pos), so
// every other consumer of the stream sees exactly what it would have seen after the reference's
// calls.
#include <stdint.h>
#include <string.h>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif

#include <vector>

#include "tlod_b200.h"

namespace {

// k[0..624) -> the next 624 words of the stream, in place.  Branch-free twist; neither loop
// carries a dependency shorter than 227 words, so the host compiler vectorises both.
static void mt_twist(uint32_t* k) {
  const int N = 624, M = 397;
  const uint32_t MATRIX_A = 0x9908b0dfu, UPPER = 0x80000000u, LOWER = 0x7fffffffu;
#pragma GCC ivdep
  for (int kk = 0; kk < N - M; ++kk) {
    const uint32_t y = (k[kk] & UPPER) | (k[kk + 1] & LOWER);
    k[kk] = k[kk + M] ^ (y >> 1) ^ ((0u - (y & 1u)) & MATRIX_A);
  }
#pragma GCC ivdep
  for (int kk = N - M; kk < N - 1; ++kk) {
    const uint32_t y = (k[kk] & UPPER) | (k[kk + 1] & LOWER);
    k[kk] = k[kk + (M - N)] ^ (y >> 1) ^ ((0u - (y & 1u)) & MATRIX_A);
  }
  const uint32_t y = (k[N - 1] & UPPER) | (k[0] & LOWER);
  k[N - 1] = k[M - 1] ^ (y >> 1) ^ ((0u - (y & 1u)) & MATRIX_A);
}

struct Mt19937 {
  uint32_t* key;  // 624 words, numpy's layout: numpy's own state, or a block of `ahead`
  int pos;
  // Key blocks generated ahead of time (tlod_mt_pregen: block 0 = numpy's key, block j = its j-th
  // successor), consumed instead of twisting on the critical path; `home` is numpy's key, which
  // receives the current block in writeback().
  uint32_t* home = nullptr;
  uint32_t* ahead_end = nullptr;

  // `ahead` is used only if its block 0 still equals numpy's key (nobody drew in between)
  static Mt19937 attach(uint32_t* numpy_key, int numpy_pos, uint32_t* ahead, int ahead_blocks) {
    Mt19937 r{numpy_key, numpy_pos};
    r.home = numpy_key;
    if (ahead && ahead_blocks > 0 && memcmp(ahead, numpy_key, 624 * sizeof(uint32_t)) == 0) {
      r.key = ahead;
      r.ahead_end = ahead + (size_t)(ahead_blocks + 1) * 624;
    }
    return r;
  }
  void writeback() {
    if (home && key != home) memcpy(home, key, 624 * sizeof(uint32_t));
    key = home ? home : key;
    ahead_end = nullptr;
  }

  void refill() {
    pos = 0;
    if (ahead_end) {
      if (key + 1248 <= ahead_end) {  // the successor was generated ahead of time
        key += 624;
        return;
      }
      writeback();  // out of blocks: go on in numpy's own memory
    }
    mt_twist(key);
  }
  uint32_t next32() {
    if (pos == 624) refill();
    uint32_t y = key[pos++];
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
  }
  uint64_t next64() {  // numpy: high word first
    const uint64_t hi = next32();
    return (hi << 32) | next32();
  }
  uint64_t interval(uint64_t max) {
    if (max == 0) return 0;
    uint64_t mask = max;
    mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4;
    mask |= mask >> 8; mask |= mask >> 16; mask |= mask >> 32;
    uint64_t v;
    if (max <= 0xffffffffull) {
      while ((v = (next32() & mask)) > max) {}
    } else {
      while ((v = (next64() & mask)) > max) {}
    }
    return v;
  }
};

// np.random.permutation(n) on x[0..n) = 0..n-1.  Same draws as numpy's
// `for i in n-1 .. 1: j = random_interval(i); swap(x[i], x[j])` with
// `random_interval(i)`: `do v = next32() & mask(i) while (v > i)` -- a rejected draw is turned
// into a self-swap and does not advance i, so the loop has no data-dependent branch.
template <typename T>
void legacy_permutation(Mt19937& rng, T* x, long long n) {
  for (long long i = 0; i < n; ++i) x[i] = (T)i;
  if (n < 2) return;
  if ((uint64_t)(n - 1) > 0xffffffffull) {  // 64-bit draws: the plain loop
    for (long long i = n - 1; i >= 1; --i) {
      const long long j = (long long)rng.interval((uint64_t)i);
      const T t = x[i];
      x[i] = x[j];
      x[j] = t;
    }
    return;
  }
  uint32_t i = (uint32_t)(n - 1);
  uint32_t mask = i;
  mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
  // draws are consumed from a block of tempered words (the tempering vectorises); rng.pos is the
  // position inside numpy's key array, as numpy keeps it
  uint32_t block[624];
  int avail = 0, at = 0;
  while (i >= 1) {
    if (at == avail) {
      if (rng.pos == 624) rng.refill();
      avail = 624 - rng.pos;
      const uint32_t* k = rng.key + rng.pos;
      for (int q = 0; q < avail; ++q) {
        uint32_t y = k[q];
        y ^= (y >> 11);
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= (y >> 18);
        block[q] = y;
      }
      at = 0;
    }
    // consume until the block runs out or the permutation is complete
    int q = at;
    while (q < avail && i >= 1) {
      if (i <= (mask >> 1)) mask >>= 1;  // log2(n) times
      const uint32_t v = block[q++] & mask;
      const uint32_t acc = v <= i;
      const uint32_t j = acc ? v : i;
      const T a = x[i], b = x[j];
      x[i] = b;
      x[j] = a;
      i -= acc;
    }
    rng.pos += q - at;
    at = q;
  }
}

// The anchor-target layer only uses the permutation to DISABLE all but `keep` of n candidates
// (`labels[inds[perm[:n - keep]]] = -1`): what survives are the last `keep` entries of the
// permutation.  Fisher-Yates from the top finalises position i in iteration i, so those entries
// are known after the first `keep` accepted draws; the remaining n - 1 - keep iterations only have
// to advance the stream by exactly the words numpy would have consumed (rejections included), which
// is a streaming count over the tempered words with no random memory access.  out[0 .. keep) receives
// the surviving entries (the last `keep` of the permutation, in order); the stream position
// afterwards is the one `np.random.permutation(n)` leaves.  `ident` is a per-thread identity array
// (ident[i] == i between calls): only the <= 2 * keep entries the swaps touch are written and put
// back, instead of initialising n entries per call.
void legacy_permutation_keep_last(Mt19937& rng, std::vector<int>& ident, std::vector<int>& touched, int n,
                                  int keep, int* out) {
  for (int i = (int)ident.size(); i < n; ++i) ident.push_back(i);
  int* x = ident.data();
  if (keep > n) keep = n;
  if (n < 2) {
    for (int k = 0; k < keep; ++k) out[k] = n - keep + k;
    return;
  }
  touched.clear();
  const uint32_t thr = (uint32_t)(n - 1 - keep);  // iterations i > thr swap; i <= thr only count
  uint32_t i = (uint32_t)(n - 1);
  uint32_t mask = i;
  mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
  uint32_t block[624];
  int avail = 0, at = 0;
  while (i >= 1) {
    if (at == avail) {
      if (rng.pos == 624) rng.refill();
      avail = 624 - rng.pos;
      const uint32_t* k = rng.key + rng.pos;
      for (int q = 0; q < avail; ++q) {
        uint32_t y = k[q];
        y ^= (y >> 11);
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= (y >> 18);
        block[q] = y;
      }
      at = 0;
    }
    int q = at;
    while (q < avail && i > thr && i >= 1) {  // the part of the permutation that is read afterwards
      if (i <= (mask >> 1)) mask >>= 1;
      const uint32_t v = block[q++] & mask;
      const uint32_t acc = v <= i;
      const uint32_t j = acc ? v : i;
      const int a = x[i], b = x[j];
      x[i] = b;
      x[j] = a;
      touched.push_back((int)j);
      i -= acc;
    }
    while (q < avail && i >= 1) {  // stream position only: one epoch of constant mask at a time
      while (i <= (mask >> 1)) mask >>= 1;
      const uint32_t floor_i = mask >> 1;  // the mask halves once i reaches it
      const uint32_t m = mask;
      uint32_t ii = i;
      int qq = q;
      // Sixteen draws at a time: draw t is accepted iff v_t <= i_t, and i_t lies in [ii - 16, ii]
      // for every t of the block, so v <= ii - 16 is an acceptance and v > ii a rejection whatever
      // came before.  Only a draw in between (16 of ~2^14 values) depends on its predecessors: such
      // a block is resolved one draw at a time.  The counting loop has no carried dependency and
      // vectorises; the scalar loop it replaces was a 3-cycle chain per draw, ~40 us per image.
      while (qq + 16 <= avail && ii > floor_i + 16) {
        const uint32_t lo = ii - 16;
        uint32_t sure = 0, unsure = 0;
#if defined(__SSE2__)
        // all values are below 2^31 (n is an int): signed compares
        const __m128i vm = _mm_set1_epi32((int)m), vlo = _mm_set1_epi32((int)lo), vii = _mm_set1_epi32((int)ii);
        int above_lo = 0, above_ii = 0;
        for (int t = 0; t < 4; ++t) {
          const __m128i v = _mm_and_si128(_mm_loadu_si128(reinterpret_cast<const __m128i*>(block + qq + 4 * t)), vm);
          above_lo |= _mm_movemask_ps(_mm_castsi128_ps(_mm_cmpgt_epi32(v, vlo))) << (4 * t);
          above_ii |= _mm_movemask_ps(_mm_castsi128_ps(_mm_cmpgt_epi32(v, vii))) << (4 * t);
        }
        uint32_t pc = (uint32_t)above_lo;  // 16-bit population count (no POPCNT in the baseline ISA)
        pc = pc - ((pc >> 1) & 0x5555u);
        pc = (pc & 0x3333u) + ((pc >> 2) & 0x3333u);
        pc = (pc + (pc >> 4)) & 0x0f0fu;
        pc = (pc + (pc >> 8)) & 0x1fu;
        sure = 16u - pc;
        unsure = (uint32_t)(above_lo & ~above_ii);
#else
        for (int t = 0; t < 16; ++t) {
          const uint32_t v = block[qq + t] & m;
          sure += v <= lo;
          unsure += (v > lo) & (v <= ii);
        }
#endif
        if (unsure) break;
        ii -= sure;
        qq += 16;
      }
      const int lim = qq + 16 < avail ? qq + 16 : avail;
      while (qq < lim && ii > floor_i) ii -= (block[qq++] & m) <= ii;
      i = ii;
      q = qq;
    }
    rng.pos += q - at;
    at = q;
  }
  for (int k = 0; k < keep; ++k) out[k] = x[n - keep + k];
  for (int j : touched) x[j] = j;
  for (int k = n - keep; k < n; ++k) x[k] = k;
}

// lab[k] = -1 wherever lab[k] == +-0
void disable_zeros(float* lab, int n) {
  int k = 0;
#if defined(__SSE2__)
  const __m128i zero = _mm_setzero_si128();
  const __m128i minus1 = _mm_castps_si128(_mm_set1_ps(-1.f));
  for (; k + 4 <= n; k += 4) {
    const __m128i x = _mm_loadu_si128(reinterpret_cast<const __m128i*>(lab + k));
    const __m128i z = _mm_cmpeq_epi32(_mm_slli_epi32(x, 1), zero);
    _mm_storeu_si128(reinterpret_cast<__m128i*>(lab + k), _mm_or_si128(_mm_andnot_si128(z, x), _mm_and_si128(z, minus1)));
  }
#endif
  for (; k < n; ++k)
    if (lab[k] == 0.f) lab[k] = -1.f;
}

}  // namespace

extern "C" int tlod_numpy_permutation(unsigned int* mt_key, int* mt_pos, long long n, long long* out) {
  if (!mt_key || !mt_pos || (n > 0 && !out)) return TLOD_ERR_NULL_POINTER;
  if (n < 0 || *mt_pos < 0 || *mt_pos > 624) return TLOD_ERR_BAD_SHAPE;
  Mt19937 rng{mt_key, *mt_pos};
  legacy_permutation(rng, out, n);
  *mt_pos = rng.pos;
  return TLOD_OK;
}

extern "C" int tlod_mt_pregen(const unsigned int* mt_key, unsigned int* ahead, int blocks) {
  if (!mt_key || !ahead) return TLOD_ERR_NULL_POINTER;
  if (blocks < 0) return TLOD_ERR_BAD_SHAPE;
  memcpy(ahead, mt_key, 624 * sizeof(uint32_t));
  for (int j = 0; j < blocks; ++j) {
    uint32_t* next = ahead + (size_t)(j + 1) * 624;
    memcpy(next, next - 624, 624 * sizeof(uint32_t));
    mt_twist(next);
  }
  return TLOD_OK;
}

extern "C" int tlod_anchor_subsample_host(float* labels, int batch, int n, int num_fg, int rpn_batchsize,
                                          unsigned int* mt_key, int* mt_pos, int* num_examples_last) {
  return tlod_anchor_subsample_host_ahead(labels, batch, n, num_fg, rpn_batchsize, mt_key, mt_pos, nullptr, 0,
                                          num_examples_last);
}

extern "C" int tlod_anchor_subsample_host_ahead(float* labels, int batch, int n, int num_fg, int rpn_batchsize,
                                                unsigned int* mt_key, int* mt_pos, unsigned int* ahead,
                                                int ahead_blocks, int* num_examples_last) {
  if (!labels || !mt_key || !mt_pos || !num_examples_last) return TLOD_ERR_NULL_POINTER;
  if (batch <= 0 || n < 0 || *mt_pos < 0 || *mt_pos > 624 || ahead_blocks < 0) return TLOD_ERR_BAD_SHAPE;
  Mt19937 rng = Mt19937::attach(mt_key, *mt_pos, ahead, ahead_blocks);
  // scratch kept per thread: three 70 KB vectors allocated and freed per call made glibc trim and
  // regrow the heap every time (~50 page faults, more than the scan below costs)
  static thread_local std::vector<int> fg, bg, ident, touched, kept;
  if ((int)fg.size() < n) fg.resize(n);
  if ((int)bg.size() < n) bg.resize(n);
  int examples = 0;
  for (int i = 0; i < batch; ++i) {
    float* lab = labels + (size_t)i * n;
    // one pass: the foreground subsampling only turns 1 into -1, so the background list
    // (labels == 0) can be collected before it
    int n_bg = 0, n_fg = 0;
    int* __restrict__ bgp = bg.data();
    int* __restrict__ fgp = fg.data();
    // branch-free compaction of both lists, on the bit patterns (+-0.0 and 1.0f: an integer compare is
    // a third of the instructions of the NaN-aware float one)
    const uint32_t* __restrict__ bits = reinterpret_cast<const uint32_t*>(lab);
    for (int k = 0; k < n; ++k) {
      const uint32_t v = bits[k];
      bgp[n_bg] = k;
      n_bg += (v << 1) == 0u;
      fgp[n_fg] = k;
      n_fg += v == 0x3f800000u;
    }
    if (n_fg > num_fg) {  // :124-132: disable perm[:n_fg - num_fg] = keep the last num_fg entries
      const int keep = num_fg > 0 ? num_fg : 0;
      kept.resize(keep);
      legacy_permutation_keep_last(rng, ident, touched, n_fg, keep, kept.data());
      for (int k = 0; k < n_fg; ++k) lab[fg[k]] = -1.f;
      for (int k = 0; k < keep; ++k) lab[fg[kept[k]]] = 1.f;
      n_fg = num_fg;
    }
    const int num_bg = rpn_batchsize - n_fg;  // :135
    int bg_kept = n_bg;
    if (n_bg > num_bg) {  // :138-145
      const int keep = num_bg > 0 ? num_bg : 0;
      kept.resize(keep);
      legacy_permutation_keep_last(rng, ident, touched, n_bg, keep, kept.data());
      disable_zeros(lab, n);  // every background label -> -1 (one streaming pass), then the kept ones back
      for (int k = 0; k < keep; ++k) lab[bgp[kept[k]]] = 0.f;
      bg_kept = keep;
    }
    // :156 -- the LAST image's count of labels >= 0 (stale loop variable in the reference)
    examples = n_fg + bg_kept;
  }
  rng.writeback();
  *mt_pos = rng.pos;
  *num_examples_last = examples;
  return TLOD_OK;
}

// Host side of _ProposalTargetLayer's fg / bg sampling
// (lib/model/rpn/proposal_target_layer_cascade.py:140-181) on numpy's global stream:
// np.random.permutation(fg_num) (always drawn in full), np.random.rand(k) (legacy random_sample:
// two 32-bit draws per double, (a >> 5) * 2^26 + (b >> 6)) / 2^53) and the index arithmetic
// floor(rand * num).  keep_out (batch, rois_per_image): sampled candidate indices, foreground
// first; fg_count_out (batch).  Returns TLOD_ERR_BAD_SHAPE if an image has neither foreground nor
// background candidates (the reference raises ValueError there).
extern "C" int tlod_proposal_sample_host(const float* max_overlaps, int batch, int n, int rois_per_image,
                                         int fg_rois_per_image, float fg_thresh, float bg_thresh_hi,
                                         float bg_thresh_lo, unsigned int* mt_key, int* mt_pos, int* keep_out,
                                         int* fg_count_out) {
  if (!max_overlaps || !mt_key || !mt_pos || !keep_out || !fg_count_out) return TLOD_ERR_NULL_POINTER;
  if (batch <= 0 || n < 0 || rois_per_image <= 0 || *mt_pos < 0 || *mt_pos > 624) return TLOD_ERR_BAD_SHAPE;
  Mt19937 rng{mt_key, *mt_pos};
  static thread_local std::vector<int> fg, bg, perm;
  if ((int)fg.capacity() < n) fg.reserve(n);
  if ((int)bg.capacity() < n) bg.reserve(n);
  auto rand_index = [&](int num) {  // floor(np.random.rand() * num)
    const uint32_t a = rng.next32() >> 5, b = rng.next32() >> 6;
    const double u = ((double)a * 67108864.0 + (double)b) / 9007199254740992.0;
    return (int)(long long)(u * (double)num);  // u * num >= 0: truncation == floor
  };
  for (int i = 0; i < batch; ++i) {
    const float* mo = max_overlaps + (size_t)i * n;
    fg.clear();
    bg.clear();
    for (int k = 0; k < n; ++k) {
      const float v = mo[k];
      if (v >= fg_thresh) fg.push_back(k);
      if (v < bg_thresh_hi && v >= bg_thresh_lo) bg.push_back(k);
    }
    const int fg_num = (int)fg.size(), bg_num = (int)bg.size();
    int* keep = keep_out + (size_t)i * rois_per_image;
    int fg_this;
    if (fg_num > 0 && bg_num > 0) {
      fg_this = fg_rois_per_image < fg_num ? fg_rois_per_image : fg_num;
      perm.resize(fg_num);
      legacy_permutation(rng, perm.data(), fg_num);
      for (int k = 0; k < fg_this; ++k) keep[k] = fg[perm[k]];
      for (int k = fg_this; k < rois_per_image; ++k) keep[k] = bg[rand_index(bg_num)];
    } else if (fg_num > 0) {
      for (int k = 0; k < rois_per_image; ++k) keep[k] = fg[rand_index(fg_num)];
      fg_this = rois_per_image;
    } else if (bg_num > 0) {
      for (int k = 0; k < rois_per_image; ++k) keep[k] = bg[rand_index(bg_num)];
      fg_this = 0;
    } else {
      *mt_pos = rng.pos;
      return TLOD_ERR_BAD_SHAPE;
    }
    fg_count_out[i] = fg_this;
  }
  *mt_pos = rng.pos;
  return TLOD_OK;
}
