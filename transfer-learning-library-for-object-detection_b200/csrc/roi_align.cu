// RoIAlign forward / backward for sm_100a.
//
// Reference semantics: lib/model/roi_align/src/roi_align_kernel.cu:15-70 (fwd),
// :94-143 (bwd); glue lib/model/roi_align/src/roi_align_cuda.c.
//
// Three pieces:
//
//  * the *plan* (tlod_roi_align_plan): everything that depends only on the RoIs and the
//    map geometry, computed once per `rois` tensor and shared by the forward and the
//    backward call: per RoI the sampling tables of its AH rows and AW columns (cell offset,
//    the two bilinear weights), the RoI indices stably sorted by image, per-image offsets,
//    and for the backward the column "merge chain" that turns the 16 (possibly
//    coinciding) column cells of a sample row into at most 16 *distinct* cells.
//
//  * plane-resident forward: a CTA keeps the 16 feature planes of one (image, 16-channel
//    slab) in shared memory (16 x H*W fp32 = 178-182 KB for the 37x75 / 38x75 maps) and
//    streams the image's RoIs through them.  HBM traffic is the algorithmic minimum: each
//    plane is read once, the (R, C, AH, AW) tensor is written once with full 32-byte
//    sectors.  Lane = 2*channel + slot; slot = half of an output row (4 samples).  The
//    plane stride is 2 (mod 4) floats so the 16 channels land on 16 banks of one parity,
//    and the two slots always read cells of opposite column parity, so every shared-memory
//    gather is bank-conflict free for any RoI geometry.
//
//  * band-resident backward (no global atomics, no memset): a CTA owns the gradient planes
//    of (image, 128-channel group, row band) in shared memory.  Lane = channel: a plane is
//    only ever touched by one thread, so the scatter is plain load-add-store on shared
//    memory, with a fixed summation order (bitwise reproducible); the band is written to
//    HBM once with coalesced stores.  The reference issues 4*R*C*AH*AW global fp32 REDs
//    (roi_align_kernel.cu:131-134).
//
//  * generic kernels for shapes the resident layouts cannot hold (channels % 16 != 0,
//    planes too large for shared memory, aligned size > 16, no plan): one CTA per
//    (RoI, channel block), geometry hoisted to shared memory, coalesced output, fp32
//    atomics in the backward.
#include <cuda.h>

#include <cstdlib>

#include "common.cuh"

namespace tlod {

// ===========================================================================
// generic kernels
// ===========================================================================
template <bool BACKWARD>
__global__ void __launch_bounds__(256)
    roi_align_generic_kernel(const float* __restrict__ src, const float* __restrict__ rois,
                             float* __restrict__ dst, int B, int C, int H, int W, int AH, int AW,
                             float scale, int chans_per_block) {
  extern __shared__ AxisTab tab[];  // [AH rows][AW cols]
  const int n = blockIdx.x;
  const int c0 = blockIdx.y * chans_per_block;
  const float* r = rois + (size_t)n * 5;
  const int b = (int)r[0];
  for (int p = threadIdx.x; p < AH + AW; p += blockDim.x) {
    if (p < AH)
      tab[p] = make_tab(align_axis(r[2], r[4], scale, AH, H, p), W);
    else
      tab[p] = make_tab(align_axis(r[1], r[3], scale, AW, W, p - AH), 1);
  }
  __syncthreads();
  const int S = AH * AW;
  const int cb = min(chans_per_block, C - c0);
  const bool image_ok = (b >= 0 && b < B);
  const size_t roi_base = ((size_t)n * C + c0) * S;
  const size_t img_base = ((size_t)(image_ok ? b : 0) * C + c0) * H * W;
  for (int o = threadIdx.x; o < cb * S; o += blockDim.x) {
    const int c = o / S;
    const int i = o - c * S;
    const int ph = i / AW;
    const int pw = i - ph * AW;
    const AxisTab row = tab[ph];
    const AxisTab col = tab[AH + pw];
    const bool ok = image_ok && row.off >= 0 && col.off >= 0;
    const size_t cell = img_base + (size_t)c * H * W + (ok ? row.off + col.off : 0);
    if (!BACKWARD) {
      float v = 0.f;
      if (ok) {
        const float* p = src + cell;
        v = __ldg(p) * (row.w0 * col.w0);
        v = fmaf(__ldg(p + 1), row.w0 * col.w1, v);
        v = fmaf(__ldg(p + W), row.w1 * col.w0, v);
        v = fmaf(__ldg(p + W + 1), row.w1 * col.w1, v);
      }
      dst[roi_base + o] = v;
    } else if (ok) {
      const float g = __ldg(src + roi_base + o);
      float* p = dst + cell;
      atomicAdd(p, g * (row.w0 * col.w0));
      atomicAdd(p + 1, g * (row.w0 * col.w1));
      atomicAdd(p + W, g * (row.w1 * col.w0));
      atomicAdd(p + W + 1, g * (row.w1 * col.w1));
    }
  }
}

// ===========================================================================
// plan
// ===========================================================================
constexpr int PL_MAXB = 1024;  // images per call on the planned paths
constexpr int PL_MAXA = 16;    // aligned_h / aligned_w limit on the planned paths
constexpr int PL_THREADS = 256;

// Column chain of one RoI for the backward pass (aligned_w == 8).  Walking the 8 samples of
// a row left to right, (a0, a1) accumulate the values of cells (cur, cur + 1):
//     a0' = g*cw0[t] + ms[t]*a0 + mh[t]*a1        a1' = g*cw1[t] + ms[t]*a1
// (same cell: ms=1; moved right by one: mh=1; jumped: both 0).  Before sample t (t = 1..7)
// and after the last one (t = 8) the accumulators that fall out of the window are final:
// a0 -> cell ex[t], a1 -> cell ex[t] + 1 ("sites" 2(t-1) and 2(t-1)+1; a site that is not
// emitted has offset -1).  All cells emitted for one row are distinct, so their
// read-modify-writes are independent.
struct __align__(16) BwdCols {
  float cw0[8], cw1[8], ms[8], mh[8];
  int soff[16];  // site 2(t-1)+k (t = 1..8, k = 0/1): byte offset (ex[t] + k) * 4 in the row, -1 = off
  int all_jump;  // every sample is valid and starts a new pair of cells: the chain is the identity
  int pad[3];
};
static_assert(sizeof(BwdCols) == 208, "BwdCols layout");

struct PlanLayout {
  size_t cum, list, yrow, tabs, bwdx, total;
};
__host__ __device__ inline size_t pl_align(size_t v) { return (v + 255) / 256 * 256; }
__host__ __device__ inline PlanLayout plan_layout(int B, int R) {
  PlanLayout L;
  size_t off = 0;
  L.cum = off;  off = pl_align(off + (size_t)(B + 2) * 4);
  L.list = off; off = pl_align(off + (size_t)R * 4);
  L.yrow = off; off = pl_align(off + (size_t)R * 32);
  L.tabs = off; off = pl_align(off + (size_t)R * 512);
  L.bwdx = off; off = pl_align(off + (size_t)R * sizeof(BwdCols));
  L.total = off;
  return L;
}

struct PlanPtrs {
  int* cum;        // [B + 2] exclusive prefix of RoIs per image; slot B = invalid image index
  int* list;       // [R] RoI indices sorted by image, stable
  short* yrow;     // [R][16] first sampled row of each output row, -1 if none
  float4* tabs;    // [R][32] rows 0..15: {row*W | -1, w0, w1, row}; cols 16..31: {col | -1, w0, w1, -}
  BwdCols* bwdx;   // [R]
};
__host__ __device__ inline PlanPtrs plan_ptrs(void* base, int B, int R) {
  const PlanLayout L = plan_layout(B, R);
  unsigned char* p = (unsigned char*)base;
  PlanPtrs q;
  q.cum = (int*)(p + L.cum);
  q.list = (int*)(p + L.list);
  q.yrow = (short*)(p + L.yrow);
  q.tabs = (float4*)(p + L.tabs);
  q.bwdx = (BwdCols*)(p + L.bwdx);
  return q;
}

__device__ __forceinline__ int roi_image(const float* __restrict__ rois, int i, int B) {
  const int b = (int)__ldg(rois + (size_t)i * 5);
  return (b < 0 || b >= B) ? B : b;
}

// CTA 0: counting sort of the RoI indices by image (stable).  CTAs >= 1: one warp per RoI
// computes its 32 table entries and the backward column chain.
__global__ void __launch_bounds__(PL_THREADS)
    roi_align_plan_kernel(const float* __restrict__ rois, PlanPtrs pl, int B, int H, int W, int R,
                          int AH, int AW, float scale) {
  const int tid = threadIdx.x, lane = lane_id(), wid = warp_id();
  if (blockIdx.x == 0) {
    __shared__ int cnt[PL_MAXB + 2];
    __shared__ int next[PL_MAXB + 2];
  
    for (int i = tid; i <= B; i += PL_THREADS) cnt[i] = 0;
    __syncthreads();
    for (int i = tid; i < R; i += PL_THREADS) atomicAdd(&cnt[roi_image(rois, i, B)], 1);
    __syncthreads();
    if (wid == 0) {
      int run = 0;
      for (int base = 0; base <= B; base += 32) {
        const int i = base + lane;
        const int v = (i <= B) ? cnt[i] : 0;
        int incl = v;
        for (int d = 1; d < 32; d <<= 1) {
          const int t = __shfl_up_sync(0xffffffffu, incl, d);
          if (lane >= d) incl += t;
        }
        if (i <= B) {
          next[i] = run + incl - v;
          pl.cum[i] = run + incl - v;
        }
        run += __shfl_sync(0xffffffffu, incl, 31);
      }
      if (lane == 0) {
        pl.cum[B + 1] = run;

      }
    }
    __syncthreads();
    for (int base = 0; base < R; base += PL_THREADS) {
      const int i = base + tid;
      const int b = (i < R) ? roi_image(rois, i, B) : -1 - tid;  // unique dummy keys
      for (int w = 0; w < PL_THREADS / 32; ++w) {
        if (wid == w) {
          const unsigned peers = __match_any_sync(0xffffffffu, b);
          const int leader = __ffs(peers) - 1;
          int pos = 0;
          if (lane == leader && b >= 0) {
            pos = next[b];
            next[b] = pos + __popc(peers);
          }
          pos = __shfl_sync(0xffffffffu, pos, leader);
          if (b >= 0) pl.list[pos + __popc(peers & ((1u << lane) - 1u))] = i;
        }
        __syncthreads();
      }
    }
    return;
  }

  const int n = (blockIdx.x - 1) * (PL_THREADS / 32) + wid;
  if (n >= R) return;
  const float* r = rois + (size_t)n * 5;
  AxisTab t;
  int start = -1;
  t.off = -1; t.w0 = 0.f; t.w1 = 0.f;
  if (lane < 16) {
    if (lane < AH) {
      const AlignAxis a = align_axis(__ldg(r + 2), __ldg(r + 4), scale, AH, H, lane);
      t = make_tab(a, W);
      start = a.valid ? a.start : -1;
    }
    pl.yrow[(size_t)n * 16 + lane] = (short)start;
  } else if (lane - 16 < AW) {
    const AlignAxis a = align_axis(__ldg(r + 1), __ldg(r + 3), scale, AW, W, lane - 16);
    t = make_tab(a, 1);
    start = a.valid ? a.start : -1;
  }
  // invalid samples carry zero weights so that the kernels need no select
  if (t.off < 0) { t.w0 = 0.f; t.w1 = 0.f; }
  pl.tabs[(size_t)n * 32 + lane] = make_float4(__int_as_float(t.off), t.w0, t.w1, __int_as_float(start));

  if (AW == 8) {
    // every lane walks the chain (cheap); lane t keeps the entries of sample t
    int cur = -1;
    unsigned en0 = 0u, en1 = 0u;
    int all_jump = 1;
    float my_cw0 = 0.f, my_cw1 = 0.f, my_ms = 1.f, my_mh = 0.f;
    int my_ex = 0, ex8 = 0;
#pragma unroll
    for (int s = 0; s < 8; ++s) {
      const int x = __shfl_sync(0xffffffffu, t.off, 16 + s);
      const float w0 = __shfl_sync(0xffffffffu, t.w0, 16 + s);
      const float w1 = __shfl_sync(0xffffffffu, t.w1, 16 + s);
      float ms = 1.f, mh = 0.f;
      const int ex = cur < 0 ? 0 : cur;
      if (x >= 0) {
        if (cur < 0) {
          ms = 0.f;
        } else if (x == cur) {
          all_jump = 0;
        } else if (x == cur + 1) {
          ms = 0.f; mh = 1.f;
          en0 |= 1u << s;
          all_jump = 0;
        } else {
          ms = 0.f;
          en0 |= 1u << s;
          en1 |= 1u << s;
        }
        cur = x;
      } else {
        all_jump = 0;  // the fast path assumes sample t's cells are emitted at transition t+1
      }
      if (lane == s) { my_cw0 = w0; my_cw1 = w1; my_ms = ms; my_mh = mh; my_ex = ex; }
    }
    if (cur >= 0) { en0 |= 1u << 8; en1 |= 1u << 8; ex8 = cur; }
    BwdCols* bc = pl.bwdx + n;
    // lane t (0..7) holds ex of transition t; sites belong to transitions 1..8
    const int ex_next = __shfl_down_sync(0xffffffffu, my_ex, 1);  // lane t: ex[t + 1]
    if (lane < 8) {
      bc->cw0[lane] = my_cw0; bc->cw1[lane] = my_cw1; bc->ms[lane] = my_ms; bc->mh[lane] = my_mh;
      const int t = lane + 1;
      const int ex = (t == 8) ? ex8 : ex_next;
      bc->soff[2 * lane] = ((en0 >> t) & 1u) ? ex * 4 : -1;
      bc->soff[2 * lane + 1] = ((en1 >> t) & 1u) ? (ex + 1) * 4 : -1;
    } else if (lane == 8) {
      bc->all_jump = all_jump;
    }
  }
}

// ===========================================================================
// plane-resident forward
// ===========================================================================
constexpr int PR_CH = 16;        // channels per slab == warps per CTA
constexpr int PR_THREADS = 512;  // 16 warps

struct PRShared {
  float4 wtab[PR_THREADS / 32][32];  // per-warp copy of the current RoI's tables
  int cur[4];                        // broadcast slots: image, slab, rank_lo, rank_hi
};

// Shared-memory plane geometry: rows padded to an even stride, planes to 2 (mod 4) floats.
// Then (i) the 16 channels of a slab start on 16 distinct banks of one parity and (ii) every
// row starts on an even word, so a lane that reads column x and a lane that reads column
// x + 1 -- of any two rows -- are always on banks of opposite parity.
__host__ __device__ inline int pr_row_stride(int w) { return w + (w & 1); }
__host__ __device__ inline int pr_plane_stride(int h, int w) {
  int p = h * pr_row_stride(w);
  while ((p & 3) != 2) ++p;
  return p;
}

__device__ __forceinline__ void st_global_v8(float* p, const float (&o)[8]) {
  asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "f"(o[0]),
               "f"(o[1]), "f"(o[2]), "f"(o[3]), "f"(o[4]), "f"(o[5]), "f"(o[6]), "f"(o[7])
               : "memory");
}

__device__ __forceinline__ float lds_f32(unsigned addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
template <int IMM>
__device__ __forceinline__ float lds_f32_imm(unsigned addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(v) : "r"(addr), "n"(IMM));
  return v;
}

// One warp, one RoI, 16 channels, AW == 8: lane (c, slot) produces the whole 8-wide output
// rows ph = 2j + slot of channel c and writes each with one 256-bit store (a lane pair
// covers 64 contiguous bytes).  Slot 0 reads its cell pairs as (x, x+1), slot 1 as (x+1, x):
// with the even row stride that alone makes every gather bank-conflict free.
// Addresses are 32-bit shared-window byte addresses; WP > 0 is the compile-time row stride
// (the second row of a sample is then an immediate offset), WP == 0 the run-time one.
template <int WP>
__device__ __forceinline__ void pr_fwd_roi_w8(unsigned plane_addr, int Wp_rt,
                                              const float4* __restrict__ wtab, int AH, int slot,
                                              float* __restrict__ out_c /* channel c of the RoI */,
                                              int dbg) {
  unsigned ca[8], cb[8];
  float wp[8], wq[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const float4 m = wtab[16 + q];
    int x = __float_as_int(m.x);
    x = x < 0 ? 0 : x;
    ca[q] = plane_addr + 4u * (unsigned)(x + slot);
    cb[q] = plane_addr + 4u * (unsigned)(x + (slot ^ 1));
    wp[q] = slot ? m.z : m.y;
    wq[q] = slot ? m.y : m.z;
  }
  const unsigned row_bytes = 4u * (unsigned)(WP > 0 ? WP : Wp_rt);
  for (int j = 0; 2 * j < AH; ++j) {
    const int ph = min(2 * j + slot, AH - 1);
    const float4 r = wtab[ph];
    const unsigned ro = (unsigned)__float_as_int(r.x);  // byte offset of the first sampled row
    float o[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const unsigned a = ca[q] + ro, b = cb[q] + ro;
      float p00, p01, p10, p11;
      if (WP > 0) {
        p00 = lds_f32(a); p01 = lds_f32(b);
        p10 = lds_f32_imm<4 * WP>(a); p11 = lds_f32_imm<4 * WP>(b);
      } else {
        p00 = lds_f32(a); p01 = lds_f32(b);
        p10 = lds_f32(a + row_bytes); p11 = lds_f32(b + row_bytes);
      }
      const float t0 = fmaf(p01, wq[q], p00 * wp[q]);
      const float t1 = fmaf(p11, wq[q], p10 * wp[q]);
      o[q] = fmaf(t1, r.z, t0 * r.y);
    }
    if (2 * j + slot < AH && dbg != 1) st_global_v8(out_c + ph * 8, o);
  }
}

// Any AH, AW <= 16: lane (c, slot) produces samples i = 2k + slot of channel c.
__device__ __forceinline__ void pr_fwd_roi_any(const float* __restrict__ plane, int Wp,
                                               const float4* __restrict__ wtab, int AH, int AW,
                                               int slot, float* __restrict__ out_c) {
  const int S = AH * AW;
  for (int k = 0; 2 * k < S; ++k) {
    const int i = min(2 * k + slot, S - 1);
    const int ph = i / AW, pw = i - ph * AW;
    const float4 row = wtab[ph], col = wtab[16 + pw];
    const int co = __float_as_int(col.x);
    const int off = __float_as_int(row.x) + (co < 0 ? 0 : co);
    const int other = __shfl_xor_sync(0xffffffffu, off, 1);
    const int flip = slot & (((off ^ other) & 1) ^ 1);
    const float* p = plane + off;
    const float a = p[flip];
    const float b = p[flip ^ 1];
    const float cc = p[Wp + flip];
    const float d = p[Wp + (flip ^ 1)];
    const float wa = flip ? col.z : col.y;
    const float wb = flip ? col.y : col.z;
    const float t0 = fmaf(b, wb, a * wa);
    const float t1 = fmaf(d, wb, cc * wa);
    if (2 * k + slot < S) out_c[i] = fmaf(t1, row.z, t0 * row.y);
  }
}

// W8: 0 = any aligned width (<= 16); 1 = aligned_w == 8; 76 = aligned_w == 8 and a padded row
// stride of 76 floats (W = 75 or 76: the 600x1200 / stride-16 maps) as a compile-time constant.
template <int W8>
__global__ void __launch_bounds__(PR_THREADS, 1)
    roi_align_fwd_planes_kernel(const float* __restrict__ features, float* __restrict__ output,
                                PlanPtrs pl, int B, int C, int H, int W, int R, int AH, int AW,
                                int Pp, int dbg) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* planes = reinterpret_cast<float*>(smem_raw);
  PRShared& sh = *reinterpret_cast<PRShared*>(smem_raw + (size_t)PR_CH * Pp * sizeof(float));
  const int tid = threadIdx.x, lane = lane_id(), wid = warp_id();
  const int nslabs = C / PR_CH;
  const int P = H * W, S = AH * AW;
  const int Wp = pr_row_stride(W);
  const int c = lane >> 1, slot = lane & 1;
  const int* __restrict__ cum = pl.cum;

  // this CTA's contiguous range of (image, slab, RoI-rank) units
  const long long U = (long long)nslabs * R;
  long long u = U * blockIdx.x / gridDim.x;
  const long long u_end = U * (blockIdx.x + 1) / gridDim.x;
  int staged_img = -1, staged_slab = -1;

  while (u < u_end) {
    if (tid == 0) {
      // image b with nslabs*cum[b] <= u < nslabs*cum[b+1]
      int lo = 0, hi = B;
      while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if ((long long)nslabs * __ldg(cum + mid) <= u) lo = mid; else hi = mid - 1;
      }
      while (lo < B && __ldg(cum + lo + 1) == __ldg(cum + lo)) ++lo;  // images without RoIs
      const int cnt = __ldg(cum + lo + 1) - __ldg(cum + lo);
      const long long rem = u - (long long)nslabs * __ldg(cum + lo);
      const int slab = (int)(rem / cnt);
      const int r_lo = (int)(rem - (long long)slab * cnt);
      const long long room = u_end - u;
      int r_hi = cnt;
      if (room < (long long)(r_hi - r_lo)) r_hi = r_lo + (int)room;
      sh.cur[0] = lo; sh.cur[1] = slab; sh.cur[2] = r_lo; sh.cur[3] = r_hi;
    }
    __syncthreads();
    const int img = sh.cur[0], slab = sh.cur[1], r_lo = sh.cur[2], r_hi = sh.cur[3];
    const int c0 = slab * PR_CH;

    // ---- stage the 16 planes of (img, slab): warp w copies channel w, 32 loads in flight ----
    if (img < B && (img != staged_img || slab != staged_slab)) {
      const float* g = features + ((size_t)img * C + c0 + wid) * P;
      float* s = planes + wid * Pp;
      const int pad = Wp - W;
      int y = lane / W, x = lane - y * W;  // position of element i = lane + 32 k
      for (int i = lane; i < P; i += 32 * 32) {
        float v[32];
        int d[32];
#pragma unroll
        for (int k = 0; k < 32; ++k) {
          v[k] = (i + k * 32 < P) ? __ldg(g + i + k * 32) : 0.f;
          d[k] = i + k * 32 + y * pad;
          x += 32;
          while (x >= W) { x -= W; ++y; }
        }
#pragma unroll
        for (int k = 0; k < 32; ++k)
          if (i + k * 32 < P) s[d[k]] = v[k];
      }
      staged_img = img;
      staged_slab = slab;
    }
    __syncthreads();

    const int base = __ldg(cum + img);
    const float* plane = planes + c * Pp;
    const unsigned plane_addr = (unsigned)__cvta_generic_to_shared(plane);
    float4* wtab = sh.wtab[wid];
    int e = r_lo + wid;
    int n = (e < r_hi) ? __ldg(pl.list + base + e) : 0;
    float4 t = (e < r_hi && img < B) ? __ldg(pl.tabs + (size_t)n * 32 + lane) : make_float4(0, 0, 0, 0);
    for (; e < r_hi; e += PR_THREADS / 32) {
      // prefetch the next RoI of this warp
      const int e2 = e + PR_THREADS / 32;
      const int n2 = (e2 < r_hi) ? __ldg(pl.list + base + e2) : 0;
      float4 t2 = make_float4(0, 0, 0, 0);
      if (e2 < r_hi && img < B) t2 = __ldg(pl.tabs + (size_t)n2 * 32 + lane);

      float* out_roi = output + ((size_t)n * C + c0) * S;
      if (img == B) {  // invalid image index: zeros
        for (int i = lane; i < PR_CH * S; i += 32) out_roi[i] = 0.f;
      } else {
        // rows: offset of the first sampled row in the padded plane (0 if out of range: the
        // weights are zero then); bytes for the aligned_w == 8 path, cells otherwise
        if (lane < 16) {
          const int start = __float_as_int(t.w);
          t.x = __int_as_float(start < 0 ? 0 : start * Wp * (W8 ? 4 : 1));
        }
        wtab[lane] = t;
        __syncwarp();
        if (W8 == 0)
          pr_fwd_roi_any(plane, Wp, wtab, AH, AW, slot, out_roi + (size_t)c * S);
        else
          pr_fwd_roi_w8<(W8 > 1 ? W8 : 0)>(plane_addr, Wp, wtab, AH, slot, out_roi + (size_t)c * S, dbg);
        __syncwarp();
      }
      n = n2;
      t = t2;
    }
    __syncthreads();
    u += (r_hi - r_lo);
  }
}

// ===========================================================================
// band-resident backward (no atomics)
// ===========================================================================
// Work unit ("item") = one 8-wide gradient row (RoI n, output row ph): it scatters into plane
// rows y0 and y0 + 1.  A CTA walks the items whose rows intersect its band.  The 32-byte
// gradient rows of the CTA's 128 channels (4 KB, 256 B apart in HBM) are fetched by TMA
// (4-D tensor map over (R, C, AH, 8), box 8 x 1 x 128 x 1, 32-byte swizzle) into a ring of
// shared-memory stages by a producer warp; the four consumer warps read their lane's row with
// two conflict-free LDS.128 and release the stage.
constexpr int BW_WARPS = 4;                    // consumer warps
constexpr int BW_THREADS = BW_WARPS * 32 + 32;  // + 1 producer warp
constexpr int BW_CH = BW_WARPS * 32;            // channels per CTA
constexpr int BW_CHUNK = BW_WARPS * 32;         // RoIs examined per round
constexpr int BW_STAGES = 6;
constexpr int BW_TILE_BYTES = BW_CH * 32;
constexpr int BW_META_BYTES = 256;  // BwdCols (208) + row table entry (16), padded
constexpr int BW_STAGE_BYTES = BW_TILE_BYTES + BW_META_BYTES;

struct BWShared {
  int items[BW_CHUNK * 16];  // (roi << 4) | ph, in (list order, ph) order
  int warp_sums[BW_WARPS];
  int nitems;
  unsigned long long full_bar[BW_STAGES];
  unsigned long long empty_bar[BW_STAGES];
};

__device__ __forceinline__ unsigned smem_u32(const void* p) {
  return (unsigned)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  const unsigned addr = smem_u32(bar);
  unsigned done = 0;
  // bounded: a lost TMA transaction must fault the launch, not hang the device
  for (unsigned spins = 0; !done; ++spins) {
    if (spins > (1u << 24)) __trap();
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, unsigned bytes,
                                          unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const void* tmap, unsigned long long* bar,
                                            int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"((unsigned long long)tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

struct BwdState {
  float cw0[8], cw1[8], ms[8], mh[8];
  unsigned sa[16];  // shared-window address of site j in band row 0 of this lane's plane; ~0u = off
  int all_jump;
  int n;
};

// per-RoI column state from the stage's metadata block (all lanes read the same 208 bytes)
__device__ __forceinline__ void bw_load_cols(BwdState& s, const float4* __restrict__ q, unsigned plane_addr) {
  float4 v[13];
#pragma unroll
  for (int i = 0; i < 13; ++i) v[i] = q[i];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    s.cw0[4 * i] = v[i].x; s.cw0[4 * i + 1] = v[i].y; s.cw0[4 * i + 2] = v[i].z; s.cw0[4 * i + 3] = v[i].w;
    s.cw1[4 * i] = v[2 + i].x; s.cw1[4 * i + 1] = v[2 + i].y; s.cw1[4 * i + 2] = v[2 + i].z; s.cw1[4 * i + 3] = v[2 + i].w;
    s.ms[4 * i] = v[4 + i].x; s.ms[4 * i + 1] = v[4 + i].y; s.ms[4 * i + 2] = v[4 + i].z; s.ms[4 * i + 3] = v[4 + i].w;
    s.mh[4 * i] = v[6 + i].x; s.mh[4 * i + 1] = v[6 + i].y; s.mh[4 * i + 2] = v[6 + i].z; s.mh[4 * i + 3] = v[6 + i].w;
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int o[4] = {__float_as_int(v[8 + i].x), __float_as_int(v[8 + i].y), __float_as_int(v[8 + i].z),
                      __float_as_int(v[8 + i].w)};
#pragma unroll
    for (int k = 0; k < 4; ++k) s.sa[4 * i + k] = o[k] < 0 ? 0xffffffffu : plane_addr + (unsigned)o[k];
  }
  s.all_jump = __float_as_int(v[12].x);
}

// shared-memory access at [addr + IMM]; the _if forms are predicated on addr != ~0u (branch-free)
template <int IMM>
__device__ __forceinline__ float lds_at(unsigned addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(v) : "r"(addr), "n"(IMM));
  return v;
}
template <int IMM>
__device__ __forceinline__ void sts_at(unsigned addr, float v) {
  asm volatile("st.shared.f32 [%0+%1], %2;" ::"r"(addr), "n"(IMM), "f"(v) : "memory");
}
template <int IMM>
__device__ __forceinline__ float lds_at_if(unsigned addr) {
  float v = 0.f;
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %1, 0xffffffff;\n\t@p ld.shared.f32 %0, [%1+%2];\n\t}"
               : "+f"(v) : "r"(addr), "n"(IMM));
  return v;
}
template <int IMM>
__device__ __forceinline__ void sts_at_if(unsigned addr, float v) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %0, 0xffffffff;\n\t@p st.shared.f32 [%0+%1], %2;\n\t}"
               :: "r"(addr), "n"(IMM), "f"(v) : "memory");
}

// Values of the 16 column sites of one gradient row: the chain of BwdCols (ALLJ: identity).
template <bool ALLJ>
__device__ __forceinline__ void bw_sites(const float (&g)[8], const BwdState& s, float (&e)[16]) {
  if (ALLJ) {
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      e[2 * t] = g[t] * s.cw0[t];
      e[2 * t + 1] = g[t] * s.cw1[t];
    }
  } else {
    float a0 = 0.f, a1 = 0.f;
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      if (t > 0) { e[2 * (t - 1)] = a0; e[2 * (t - 1) + 1] = a1; }
      const float na0 = fmaf(s.ms[t], a0, fmaf(s.mh[t], a1, g[t] * s.cw0[t]));
      a1 = fmaf(s.ms[t], a1, g[t] * s.cw1[t]);
      a0 = na0;
    }
    e[14] = a0; e[15] = a1;
  }
}

// Add rw0 * e into band row IMM0 / 4W and rw1 * e into band row IMM1 / 4W (byte offsets from band
// row 0 as compile-time immediates: no address arithmetic per access).  All enabled cells of a
// row are distinct, so the loads of the old values are issued first -- before the site values are
// even computed, so that the serial column chain runs under their latency -- then the stores.
// ALLJ: all 16 sites are enabled (no predicates).
template <bool ALLJ, bool DO0, bool DO1, int IMM0, int IMM1>
__device__ __forceinline__ void bw_rmw_fixed(float rw0, float rw1, const BwdState& s, const float (&g)[8]) {
  float o0[16], o1[16], e[16];
  if (DO0) {
#pragma unroll
    for (int j = 0; j < 16; ++j) o0[j] = ALLJ ? lds_at<IMM0>(s.sa[j]) : lds_at_if<IMM0>(s.sa[j]);
  }
  if (DO1) {
#pragma unroll
    for (int j = 0; j < 16; ++j) o1[j] = ALLJ ? lds_at<IMM1>(s.sa[j]) : lds_at_if<IMM1>(s.sa[j]);
  }
  bw_sites<ALLJ>(g, s, e);
  if (DO0) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float v = fmaf(rw0, e[j], o0[j]);
      if (ALLJ) sts_at<IMM0>(s.sa[j], v); else sts_at_if<IMM0>(s.sa[j], v);
    }
  }
  if (DO1) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float v = fmaf(rw1, e[j], o1[j]);
      if (ALLJ) sts_at<IMM1>(s.sa[j], v); else sts_at_if<IMM1>(s.sa[j], v);
    }
  }
}

// run-time row offsets (generic map widths / band heights): one add per access
template <bool ALLJ>
__device__ __forceinline__ void bw_rmw_dyn(unsigned off0, unsigned off1, bool do0, bool do1, float rw0,
                                           float rw1, const BwdState& s, const float (&g)[8]) {
  float o0[16], o1[16], e[16];
  if (do0) {
#pragma unroll
    for (int j = 0; j < 16; ++j) o0[j] = ALLJ ? lds_at<0>(s.sa[j] + off0) : (s.sa[j] != 0xffffffffu ? lds_at<0>(s.sa[j] + off0) : 0.f);
  }
  if (do1) {
#pragma unroll
    for (int j = 0; j < 16; ++j) o1[j] = ALLJ ? lds_at<0>(s.sa[j] + off1) : (s.sa[j] != 0xffffffffu ? lds_at<0>(s.sa[j] + off1) : 0.f);
  }
  bw_sites<ALLJ>(g, s, e);
  if (do0) {
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if (ALLJ || s.sa[j] != 0xffffffffu) sts_at<0>(s.sa[j] + off0, fmaf(rw0, e[j], o0[j]));
  }
  if (do1) {
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if (ALLJ || s.sa[j] != 0xffffffffu) sts_at<0>(s.sa[j] + off1, fmaf(rw1, e[j], o1[j]));
  }
}

// W_T > 0: compile-time map width and band height BAND (1 or 2); W_T == 0: run-time geometry.
template <bool ALLJ, int W_T, int BAND>
__device__ __forceinline__ void bw_scatter(int W, int y_lo, int y_hi, const float4 rowt, const BwdState& s,
                                           const float (&g)[8]) {
  const int r0 = __float_as_int(rowt.w) - y_lo;  // band row of the first of the two plane rows
  if (W_T > 0) {
    constexpr int RB = 4 * W_T;
    if (BAND == 2) {
      if (r0 == 0) {
        if (y_hi - y_lo == 2) bw_rmw_fixed<ALLJ, true, true, 0, RB>(rowt.y, rowt.z, s, g);
        else bw_rmw_fixed<ALLJ, true, false, 0, RB>(rowt.y, rowt.z, s, g);  // last, one-row band
      } else if (r0 < 0) {
        bw_rmw_fixed<ALLJ, false, true, 0, 0>(rowt.y, rowt.z, s, g);
      } else if (r0 < y_hi - y_lo) {
        bw_rmw_fixed<ALLJ, true, false, RB, RB>(rowt.y, rowt.z, s, g);
      }
    } else {
      if (r0 == 0) bw_rmw_fixed<ALLJ, true, false, 0, 0>(rowt.y, rowt.z, s, g);
      else bw_rmw_fixed<ALLJ, false, true, 0, 0>(rowt.y, rowt.z, s, g);
    }
  } else {
    const int rows = y_hi - y_lo;
    bw_rmw_dyn<ALLJ>(4u * (unsigned)(r0 * W), 4u * (unsigned)((r0 + 1) * W), r0 >= 0 && r0 < rows,
                     r0 + 1 >= 0 && r0 + 1 < rows, rowt.y, rowt.z, s, g);
  }
}

template <int W_T, int BAND>
__device__ __forceinline__ void bw_item(int W, int y_lo, int y_hi, const float (&g)[8], const float4 rowt,
                                        const BwdState& s) {
  if (s.all_jump)
    bw_scatter<true, W_T, BAND>(W, y_lo, y_hi, rowt, s, g);
  else
    bw_scatter<false, W_T, BAND>(W, y_lo, y_hi, rowt, s, g);
}

template <int W_T, int BAND>
__global__ void __launch_bounds__(BW_THREADS)
    roi_align_bwd_planes_kernel(const __grid_constant__ CUtensorMap tmap, float* __restrict__ bottom_grad,
                                PlanPtrs pl, int B, int C, int H, int W, int AH, int ngroups,
                                int nbands, int band_rows, int Sb) {
  extern __shared__ __align__(1024) unsigned char smem_bw[];
  // [stages][planes][BWShared]: the TMA stages need 256-byte alignment for the 32-byte swizzle
  unsigned char* stages = smem_bw;
  float* planes = reinterpret_cast<float*>(smem_bw + BW_STAGES * BW_STAGE_BYTES);
  BWShared& sh = *reinterpret_cast<BWShared*>(smem_bw + BW_STAGES * BW_STAGE_BYTES +
                                              ((size_t)BW_CH * Sb * sizeof(float) + 15) / 16 * 16);
  const int tid = threadIdx.x, lane = lane_id(), wid = warp_id();
  const int band = blockIdx.x % nbands;
  const int rest = blockIdx.x / nbands;
  const int grp = rest % ngroups, img = rest / ngroups;
  const int y_lo = band * band_rows, y_hi = min(H, y_lo + band_rows);
  const bool producer = wid == BW_WARPS;
  const int cw = grp * BW_CH + wid * 32;       // first channel of this consumer warp
  const bool active = !producer && cw < C;     // C % 32 == 0
  const int n_active = min(BW_WARPS, (C - grp * BW_CH) / 32);
  const unsigned plane_addr = smem_u32(planes + (size_t)((wid & (BW_WARPS - 1)) * 32 + lane) * Sb);

  if (tid == 0) {
    for (int s = 0; s < BW_STAGES; ++s) {
      mbar_init(&sh.full_bar[s], 1);
      mbar_init(&sh.empty_bar[s], n_active);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = tid; i < BW_CH * Sb; i += BW_THREADS) planes[i] = 0.f;
  __syncthreads();

  const int base = __ldg(pl.cum + img);
  const int n_img = __ldg(pl.cum + img + 1) - base;
  BwdState st;
  st.n = -1;
  unsigned it0 = 0;  // items issued / consumed before this chunk (stage = it % BW_STAGES)

  for (int chunk = 0; chunk < n_img; chunk += BW_CHUNK) {
    // ---- items of this chunk: (RoI, ph) whose rows y0 / y0+1 intersect the band ----
    unsigned mask = 0u;
    int n = 0;
    if (!producer && chunk + tid < n_img) {
      n = __ldg(pl.list + base + chunk + tid);
      const int4* yr = reinterpret_cast<const int4*>(pl.yrow + (size_t)n * 16);
      const int4 a = __ldg(yr), b = __ldg(yr + 1);
      const int w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        const int ys = (int)(short)((k & 1) ? (w[k >> 1] >> 16) : (w[k >> 1] & 0xffff));
        if (ys >= 0 && ys + 1 >= y_lo && ys < y_hi) mask |= 1u << k;
      }
    }
    const int cnt = __popc(mask);
    int incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += t;
    }
    if (!producer && lane == 31) sh.warp_sums[wid] = incl;
    __syncthreads();  // also orders the previous round's item reads before this round's writes
    if (!producer) {
      int pos = incl - cnt;
#pragma unroll
      for (int w = 0; w < BW_WARPS; ++w)
        if (w < wid) pos += sh.warp_sums[w];
      if (tid == BW_WARPS * 32 - 1) sh.nitems = pos + cnt;
      while (mask) {
        const int k = __ffs(mask) - 1;
        mask &= mask - 1u;
        sh.items[pos++] = (n << 4) | k;
      }
    }
    __syncthreads();
    const int nitems = sh.nitems;

    if (producer) {
      if (lane == 0) {
        for (int i = 0; i < nitems; ++i) {
          const unsigned it = it0 + (unsigned)i;
          const int s = it % BW_STAGES;
          const unsigned ph = (it / BW_STAGES) & 1u;
          mbar_wait(&sh.empty_bar[s], ph ^ 1u);
          mbar_arrive_expect_tx(&sh.full_bar[s], BW_TILE_BYTES + (unsigned)sizeof(BwdCols) + 16u);
          const int item = sh.items[i];
          unsigned char* stg = stages + s * BW_STAGE_BYTES;
          tma_load_4d(stg, &tmap, &sh.full_bar[s], 0, item & 15, grp * BW_CH, item >> 4);
          bulk_load(stg + BW_TILE_BYTES, pl.bwdx + (item >> 4), (unsigned)sizeof(BwdCols), &sh.full_bar[s]);
          bulk_load(stg + BW_TILE_BYTES + sizeof(BwdCols), pl.tabs + (size_t)(item >> 4) * 32 + (item & 15),
                    16u, &sh.full_bar[s]);
        }
      }
    } else if (active) {
      // ---- scatter: every consumer warp walks all items for its own 32 channels ----
      // lane's 32-byte row inside a stage, 16-byte halves swapped by the 32-byte swizzle
      const int r = wid * 32 + lane;
      const int sw = ((r >> 2) & 1) << 4;
      for (int i = 0; i < nitems; ++i) {
        const unsigned it = it0 + (unsigned)i;
        const int s = it % BW_STAGES;
        const int nn = sh.items[i] >> 4;
        mbar_wait(&sh.full_bar[s], (it / BW_STAGES) & 1u);
        const unsigned char* stg = stages + s * BW_STAGE_BYTES;
        const float4 ga = *reinterpret_cast<const float4*>(stg + r * 32 + sw);
        const float4 gb = *reinterpret_cast<const float4*>(stg + r * 32 + (sw ^ 16));
        const float4* meta = reinterpret_cast<const float4*>(stg + BW_TILE_BYTES);
        const float4 rowt = meta[13];
        if (nn != st.n) { bw_load_cols(st, meta, plane_addr); st.n = nn; }
        __syncwarp();
        if (lane == 0) mbar_arrive(&sh.empty_bar[s]);
        const float g[8] = {ga.x, ga.y, ga.z, ga.w, gb.x, gb.y, gb.z, gb.w};
        bw_item<W_T, BAND>(W, y_lo, y_hi, g, rowt, st);
      }
    }
    it0 += (unsigned)nitems;
  }
  __syncthreads();
  // ---- write the band: warp w stores its 32 channels, lanes along the row cells ----
  if (active) {
    const int n_cells = (y_hi - y_lo) * W;
    for (int ch = 0; ch < 32; ++ch) {
      float* g = bottom_grad + ((size_t)img * C + cw + ch) * H * W + (size_t)y_lo * W;
      const float* sp = planes + (size_t)(wid * 32 + ch) * Sb;
      for (int i = lane; i < n_cells; i += 32) g[i] = sp[i];
    }
  }
}

// ===========================================================================
// host side
// ===========================================================================
static int check_common(const void* a, const void* b, const void* c, int batch, int channels,
                        int height, int width, int num_rois, int ah, int aw) {
  if (!a || !b || !c) return TLOD_ERR_NULL_POINTER;
  if (batch <= 0 || channels <= 0 || height < 2 || width < 2 || num_rois < 0 || ah < 2 || aw < 2)
    return TLOD_ERR_BAD_SHAPE;
  if ((long long)batch * channels * height * width >= (1LL << 31)) return TLOD_ERR_INT32_OVERFLOW;
  return TLOD_OK;
}

static bool plan_supported(int batch, int height, int width, int ah, int aw) {
  return batch <= PL_MAXB && ah <= PL_MAXA && aw <= PL_MAXA && height < 32768 &&
         (long long)height * width < (1LL << 30);
}

static size_t pr_smem_bytes(int h, int w) {
  return (size_t)PR_CH * pr_plane_stride(h, w) * sizeof(float) + sizeof(PRShared);
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time libcuda dependency)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = []() -> EncodeTiledFn {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPointByVersion("cuTensorMapEncodeTiled", &p, 12000, cudaEnableDefault, &q) !=
            cudaSuccess || q != cudaDriverEntryPointSuccess)
      return nullptr;
    return (EncodeTiledFn)p;
  }();
  return fn;
}

// (R, C, AH, 8) fp32 gradient tensor; box = one 8-wide row of BW_CH consecutive channels
static bool make_grad_tmap(CUtensorMap* map, const float* top_grad, int R, int C, int AH) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) return false;
  const cuuint64_t dims[4] = {8, (cuuint64_t)AH, (cuuint64_t)C, (cuuint64_t)R};
  const cuuint64_t strides[3] = {32, (cuuint64_t)AH * 32, (cuuint64_t)C * AH * 32};
  const cuuint32_t box[4] = {8, 1, (cuuint32_t)BW_CH, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(top_grad), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static int generic_launch(bool backward, const float* src, const float* rois, float* dst, int batch,
                          int channels, int height, int width, int num_rois, int ah, int aw,
                          float scale, cudaStream_t st) {
  if (num_rois == 0) return TLOD_OK;
  const int S = ah * aw;
  int cpb = (2048 + S - 1) / S;  // ~2048 outputs per CTA
  if (cpb > channels) cpb = channels;
  while ((channels + cpb - 1) / cpb > 65535) ++cpb;
  dim3 grid(num_rois, (channels + cpb - 1) / cpb);
  size_t smem = sizeof(AxisTab) * (size_t)(ah + aw);
  {
    LaunchScope scope(backward ? "roi_align_bwd_generic_kernel" : "roi_align_fwd_generic_kernel", st);
    if (backward)
      roi_align_generic_kernel<true><<<grid, 256, smem, st>>>(src, rois, dst, batch, channels, height,
                                                              width, ah, aw, scale, cpb);
    else
      roi_align_generic_kernel<false><<<grid, 256, smem, st>>>(src, rois, dst, batch, channels,
                                                               height, width, ah, aw, scale, cpb);
  }
  return last_launch_status();
}

}  // namespace tlod

using namespace tlod;

extern "C" size_t tlod_roi_align_plan_bytes(int batch, int num_rois) {
  if (batch <= 0 || num_rois < 0) return 0;
  return plan_layout(batch, num_rois).total + 256;
}

extern "C" int tlod_roi_align_plan(const float* rois, int batch, int height, int width, int num_rois,
                                   int aligned_h, int aligned_w, float spatial_scale, void* plan,
                                   size_t plan_bytes, void* stream) {
  if (!plan || (num_rois > 0 && !rois)) return TLOD_ERR_NULL_POINTER;
  if (batch <= 0 || height < 2 || width < 2 || num_rois < 0 || aligned_h < 2 || aligned_w < 2)
    return TLOD_ERR_BAD_SHAPE;
  if (!plan_supported(batch, height, width, aligned_h, aligned_w)) return TLOD_ERR_UNSUPPORTED;
  if (plan_bytes < tlod_roi_align_plan_bytes(batch, num_rois) || ((uintptr_t)plan & 255))
    return TLOD_ERR_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  const PlanPtrs pl = plan_ptrs(plan, batch, num_rois);
  const int grid = 1 + (num_rois + PL_THREADS / 32 - 1) / (PL_THREADS / 32);
  {
    LaunchScope scope("roi_align_plan_kernel", st);
    roi_align_plan_kernel<<<grid, PL_THREADS, 0, st>>>(rois, pl, batch, height, width, num_rois,
                                                       aligned_h, aligned_w, spatial_scale);
  }
  return last_launch_status();
}

extern "C" int tlod_roi_align_forward(const float* features, const float* rois, float* output,
                                      int batch, int channels, int height, int width, int num_rois,
                                      int aligned_h, int aligned_w, float spatial_scale,
                                      const void* plan, size_t plan_bytes, void* stream) {
  int rc = check_common(features, rois, output, batch, channels, height, width, num_rois, aligned_h,
                        aligned_w);
  if (rc != TLOD_OK) return rc;
  if (num_rois == 0) return TLOD_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const bool planned = plan != nullptr && plan_supported(batch, height, width, aligned_h, aligned_w);
  if (plan && (plan_bytes < tlod_roi_align_plan_bytes(batch, num_rois) || ((uintptr_t)plan & 255)))
    return TLOD_ERR_WORKSPACE;
  if (planned && channels % PR_CH == 0 &&
      pr_smem_bytes(height, width) <= (size_t)device_info().max_smem_optin) {
    const int Pp = pr_plane_stride(height, width);
    const size_t smem = pr_smem_bytes(height, width);
    const long long units = (long long)(channels / PR_CH) * num_rois;
    int grid = device_info().sm_count;
    if ((long long)grid > units) grid = (int)units;
    const bool w8 = aligned_w == 8 && ((uintptr_t)output & 31) == 0;
    auto kern = !w8 ? roi_align_fwd_planes_kernel<0>
                    : (pr_row_stride(width) == 76 ? roi_align_fwd_planes_kernel<76>
                                                  : roi_align_fwd_planes_kernel<1>);
    static const int dbg = getenv("TLOD_FWD_DEBUG") ? atoi(getenv("TLOD_FWD_DEBUG")) : 0;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    const PlanPtrs pl = plan_ptrs(const_cast<void*>(plan), batch, num_rois);
    {
      LaunchScope scope("roi_align_fwd_planes_kernel", st);
      kern<<<grid, PR_THREADS, smem, st>>>(features, output, pl, batch, channels, height, width,
                                           num_rois, aligned_h, aligned_w, Pp, dbg);
    }
    return last_launch_status();
  }
  return generic_launch(false, features, rois, output, batch, channels, height, width, num_rois,
                        aligned_h, aligned_w, spatial_scale, st);
}

extern "C" int tlod_roi_align_backward(const float* top_grad, const float* rois, float* bottom_grad,
                                       int batch, int channels, int height, int width,
                                       int num_rois, int aligned_h, int aligned_w,
                                       float spatial_scale, const void* plan, size_t plan_bytes,
                                       void* stream) {
  int rc = check_common(top_grad, rois, bottom_grad, batch, channels, height, width, num_rois,
                        aligned_h, aligned_w);
  if (rc != TLOD_OK) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t grad_bytes = (size_t)batch * channels * height * width * sizeof(float);
  if (num_rois == 0) return (int)cudaMemsetAsync(bottom_grad, 0, grad_bytes, st);
  if (plan && (plan_bytes < tlod_roi_align_plan_bytes(batch, num_rois) || ((uintptr_t)plan & 255)))
    return TLOD_ERR_WORKSPACE;
  const bool planned = plan != nullptr && plan_supported(batch, height, width, aligned_h, aligned_w);

  // band-resident path: AW == 8 (one 32-byte gradient row per item), C % 32 == 0
  if (planned && channels % 32 == 0 && aligned_w == 8 && ((uintptr_t)top_grad & 15) == 0) {
    const size_t fixed = sizeof(BWShared) + BW_STAGES * BW_STAGE_BYTES + 64;
    size_t budget = 112 * 1024 - fixed;  // two CTAs per SM
    int band_rows = (int)((budget / (BW_CH * sizeof(float)) - 1) / width);
    if (band_rows < 2) {  // wide maps: one CTA per SM
      budget = (size_t)device_info().max_smem_optin - fixed - 1024;
      band_rows = (int)((budget / (BW_CH * sizeof(float)) - 1) / width);
    }
    if (band_rows > height) band_rows = height;
    // small problems: thinner bands (more CTAs) until the grid covers two CTAs per SM -- a CTA's
    // item loop is serial, so an under-filled machine costs more than the re-read of the
    // gradient rows that straddle two bands
    {
      const long long per_band = (long long)batch * ((channels + BW_CH - 1) / BW_CH);
      while (band_rows > 1 &&
             per_band * ((height + band_rows - 1) / band_rows) < 2LL * device_info().sm_count)
        band_rows = (band_rows + 1) / 2;
    }
    CUtensorMap tmap;
    if (band_rows >= 1 && make_grad_tmap(&tmap, top_grad, num_rois, channels, aligned_h)) {
      const int Sb = (band_rows * width) | 1;  // odd stride: lane = channel is conflict free
      const int nbands = (height + band_rows - 1) / band_rows;
      const int ngroups = (channels + BW_CH - 1) / BW_CH;
      const size_t smem = BW_STAGES * BW_STAGE_BYTES + ((size_t)BW_CH * Sb * sizeof(float) + 15) / 16 * 16 +
                          sizeof(BWShared);
      const long long grid = (long long)batch * ngroups * nbands;
      if (grid <= 2147483647LL && smem <= (size_t)device_info().max_smem_optin) {
        // the 600x1200 / stride-16 maps (W = 75) with 1- or 2-row bands get compile-time row offsets
        auto kern = roi_align_bwd_planes_kernel<0, 0>;
        if (width == 75 && band_rows == 2) kern = roi_align_bwd_planes_kernel<75, 2>;
        if (width == 75 && band_rows == 1) kern = roi_align_bwd_planes_kernel<75, 1>;
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        const PlanPtrs pl = plan_ptrs(const_cast<void*>(plan), batch, num_rois);
        {
          LaunchScope scope("roi_align_bwd_planes_kernel", st);
          kern<<<(int)grid, BW_THREADS, smem, st>>>(tmap, bottom_grad, pl, batch, channels, height, width,
                                                   aligned_h, ngroups, nbands, band_rows, Sb);
        }
        return last_launch_status();
      }
    }
  }
  cudaError_t e = cudaMemsetAsync(bottom_grad, 0, grad_bytes, st);
  if (e != cudaSuccess) return (int)e;
  return generic_launch(true, top_grad, rois, bottom_grad, batch, channels, height, width, num_rois,
                        aligned_h, aligned_w, spatial_scale, st);
}
