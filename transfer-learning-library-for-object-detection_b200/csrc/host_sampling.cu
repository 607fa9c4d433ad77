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
    const int N = 624, M = 397;
    const uint32_t MATRIX_A = 0x9908b0dfu, UPPER = 0x80000000u, LOWER = 0x7fffffffu;
    int kk = 0;
    uint32_t y;
    for (; kk < N - M; ++kk) {
      y = (key[kk] & UPPER) | (key[kk + 1] & LOWER);
      key[kk] = key[kk + M] ^ (y >> 1) ^ ((y & 1u) ? MATRIX_A : 0u);
    }
    for (; kk < N - 1; ++kk) {
      y = (key[kk] & UPPER) | (key[kk + 1] & LOWER);
      key[kk] = key[kk + (M - N)] ^ (y >> 1) ^ ((y & 1u) ? MATRIX_A : 0u);
    }
    y = (key[N - 1] & UPPER) | (key[0] & LOWER);
    key[N - 1] = key[M - 1] ^ (y >> 1) ^ ((y & 1u) ? MATRIX_A : 0u);
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
    if (n_fg > num_fg) {  // :124-132
      perm.resize(n_fg);
      legacy_permutation(rng, perm.data(), n_fg);
      for (int k = 0; k < n_fg - num_fg; ++k) lab[fg[perm[k]]] = -1.f;
      n_fg = num_fg;
    }
    const int num_bg = rpn_batchsize - n_fg;  // :135
    int bg_kept = n_bg;
    if (n_bg > num_bg) {  // :138-145
      perm.resize(n_bg);
      legacy_permutation(rng, perm.data(), n_bg);
      for (int k = 0; k < n_bg - num_bg; ++k) lab[bgp[perm[k]]] = -1.f;
      bg_kept = num_bg > 0 ? num_bg : 0;
    }
    // :156 -- the LAST image's count of labels >= 0 (stale loop variable in the reference)
    examples = n_fg + bg_kept;
  }
  *mt_pos = rng.pos;
  *num_examples_last = examples;
  return TLOD_OK;
}
