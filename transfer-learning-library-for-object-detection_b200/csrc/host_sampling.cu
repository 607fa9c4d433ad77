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

#include <vector>

#include "tlod_b200.h"

namespace {

struct Mt19937 {
  uint32_t* key;  // 624 words, numpy's layout
  int pos;

  void refill() {
    // branch-free twist; neither loop carries a dependency shorter than 227 words, so the host
    // compiler vectorises both (this refill runs once per 624 draws: ~40 times per image)
    const int N = 624, M = 397;
    const uint32_t MATRIX_A = 0x9908b0dfu, UPPER = 0x80000000u, LOWER = 0x7fffffffu;
    uint32_t* k = key;
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
    pos = 0;
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
// is a streaming count over the tempered words with no random memory access.  x[n - keep .. n) holds
// the surviving entries on return (x[0 .. n - keep) is unspecified); the stream position afterwards
// is the one `np.random.permutation(n)` leaves.
void legacy_permutation_keep_last(Mt19937& rng, int* x, int n, int keep) {
  for (int i = 0; i < n; ++i) x[i] = i;
  if (n < 2) return;
  if (keep > n) keep = n;
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
      i -= acc;
    }
    while (q < avail && i >= 1) {  // stream position only: one epoch of constant mask at a time
      while (i <= (mask >> 1)) mask >>= 1;
      const uint32_t floor_i = mask >> 1;  // the mask halves once i reaches it
      const uint32_t m = mask;
      uint32_t ii = i;
      int qq = q;
      while (qq < avail && ii > floor_i) ii -= (block[qq++] & m) <= ii;
      i = ii;
      q = qq;
    }
    rng.pos += q - at;
    at = q;
  }
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

extern "C" int tlod_anchor_subsample_host(float* labels, int batch, int n, int num_fg, int rpn_batchsize,
                                          unsigned int* mt_key, int* mt_pos, int* num_examples_last) {
  if (!labels || !mt_key || !mt_pos || !num_examples_last) return TLOD_ERR_NULL_POINTER;
  if (batch <= 0 || n < 0 || *mt_pos < 0 || *mt_pos > 624) return TLOD_ERR_BAD_SHAPE;
  Mt19937 rng{mt_key, *mt_pos};
  std::vector<int> fg, bg, perm;
  fg.reserve(n);
  bg.resize(n);
  perm.reserve(n);
  int examples = 0;
  for (int i = 0; i < batch; ++i) {
    float* lab = labels + (size_t)i * n;
    // one pass: the foreground subsampling only turns 1 into -1, so the background list
    // (labels == 0) can be collected before it
    fg.clear();
    int n_bg = 0;
    int* bgp = bg.data();
    for (int k = 0; k < n; ++k) {
      const float v = lab[k];
      bgp[n_bg] = k;
      n_bg += v == 0.f;
      if (v == 1.f) fg.push_back(k);
    }
    int n_fg = (int)fg.size();
    if (n_fg > num_fg) {  // :124-132: disable perm[:n_fg - num_fg] = keep the last num_fg entries
      perm.resize(n_fg);
      const int keep = num_fg > 0 ? num_fg : 0;
      legacy_permutation_keep_last(rng, perm.data(), n_fg, keep);
      for (int k = 0; k < n_fg; ++k) lab[fg[k]] = -1.f;
      for (int k = n_fg - keep; k < n_fg; ++k) lab[fg[perm[k]]] = 1.f;
      n_fg = num_fg;
    }
    const int num_bg = rpn_batchsize - n_fg;  // :135
    int bg_kept = n_bg;
    if (n_bg > num_bg) {  // :138-145
      perm.resize(n_bg);
      const int keep = num_bg > 0 ? num_bg : 0;
      legacy_permutation_keep_last(rng, perm.data(), n_bg, keep);
      for (int k = 0; k < n_bg; ++k) lab[bgp[k]] = -1.f;
      for (int k = n_bg - keep; k < n_bg; ++k) lab[bgp[perm[k]]] = 0.f;
      bg_kept = keep;
    }
    // :156 -- the LAST image's count of labels >= 0 (stale loop variable in the reference)
    examples = n_fg + bg_kept;
  }
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
  std::vector<int> fg, bg, perm;
  fg.reserve(n);
  bg.reserve(n);
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
