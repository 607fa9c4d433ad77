// RoIAlign forward / backward for sm_100a.
//
// Reference semantics: lib/model/roi_align/src/roi_align_kernel.cu:15-70 (fwd),
// :94-143 (bwd); glue lib/model/roi_align/src/roi_align_cuda.c.
//
// The pieces:
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
//    sectors (8x8 outputs: through per-warp staging tiles and TMA tensor stores).  Lane =
//    (channel, slot); slot = parity of the output row.  The plane stride is 2 (mod 4) floats
//    so the 16 channels land on 16 banks of one parity, and the two slots always read cells
//    of opposite column parity, so every shared-memory gather is bank-conflict free for any
//    RoI geometry.
//
//  * row-resident backward (no global atomics, no memset): roi_align_bwd.cu.  The plan holds
//    its per-(image, plane row) lists of gradient rows.
//
//  * generic kernels for shapes the resident layouts cannot hold (channels % 16 != 0,
//    planes too large for shared memory, aligned size > 16, no plan): one CTA per
//    (RoI, channel block), geometry hoisted to shared memory, coalesced output, fp32
//    atomics in the backward.
#include <cstring>

#include "async_copy.cuh"
#include "roi_align_plan.cuh"

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
constexpr int PL_THREADS = 256;

__device__ __forceinline__ int roi_image(const float* __restrict__ rois, int i, int B) {
  const int b = (int)__ldg(rois + (size_t)i * 5);
  return (b < 0 || b >= B) ? B : b;
}

// CTA 0: counting sort of the RoI indices by image (stable).  CTAs >= 1: one warp per RoI
// computes its 32 table entries and the backward column chain, and counts the RoI's sample
// rows into the (image, plane row) bins of the backward row lists (pl.rowcnt, zeroed before).
__global__ void __launch_bounds__(PL_THREADS)
    roi_align_plan_kernel(const float* __restrict__ rois, PlanPtrs pl, int B, int H, int W, int R,
                          int AH, int AW, float scale, int count_rows) {
  const int tid = threadIdx.x, lane = lane_id(), wid = warp_id();
  if (blockIdx.x == 0) {
    __shared__ int cnt[PL_MAXB + 2];
    __shared__ int next[PL_MAXB + 2];
    __shared__ int unsorted;
    if (tid == 0) unsorted = 0;
    for (int i = tid; i <= B; i += PL_THREADS) cnt[i] = 0;
    __syncthreads();
    for (int i = tid; i < R; i += PL_THREADS) {
      const int b = roi_image(rois, i, B);
      atomicAdd(&cnt[b], 1);
      if (i + 1 < R && roi_image(rois, i + 1, B) < b) unsorted = 1;
    }
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
    if (!unsorted) {  // RoIs already grouped by image (a proposal layer's output always is): identity
      for (int i = tid; i < R; i += PL_THREADS) pl.list[i] = i;
      return;
    }
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
    // sizes of the backward row lists: this sample row feeds plane rows start and start + 1
    if (count_rows && start >= 0) {
      const int b = roi_image(rois, n, B);
      if (b < B) {
        atomicAdd(pl.rowcnt + b * H + start, 1);
        atomicAdd(pl.rowcnt + b * H + start + 1, 1);
      }
    }
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
    const int my_x = __shfl_sync(0xffffffffu, t.off, 16 + (lane & 7));
    if (lane < 8) {
      bc->cw0[lane] = my_cw0; bc->cw1[lane] = my_cw1; bc->ms[lane] = my_ms; bc->mh[lane] = my_mh;
      bc->xoff[lane] = (my_x < 0 ? 0 : my_x) * 4;  // invalid samples: cell 0 with zero weights
      const int t = lane + 1;
      const int ex = (t == 8) ? ex8 : ex_next;
      bc->soff[2 * lane] = ((en0 >> t) & 1u) ? ex * 4 : W * 4;  // W * 4: the dump cell
      bc->soff[2 * lane + 1] = ((en1 >> t) & 1u) ? (ex + 1) * 4 : W * 4;
    }
    if (lane == 0) pl.jump[n] = (unsigned char)all_jump;
  } else if (lane == 0) {
    pl.jump[n] = 0;
  }
}

// Backward row lists.  One CTA per (image, plane row) bin: its first item = sum of the counts of
// the bins before it; its 8 warps take the image's RoIs in list order, 32 per warp and round
// (all the dependent list -> row table loads of a pass are in flight at once), find the sample
// rows that touch this plane row, and write the items (RoI, ph, row weight) in (RoI, ph) order --
// a fixed order, so the backward sums are bitwise reproducible.
constexpr int PLR_WARPS = PL_THREADS / 32;
constexpr int PLR_ROUNDS = 4;  // 32-RoI chunks per warp and pass: a pass covers 1024 RoIs

__global__ void __launch_bounds__(PL_THREADS)
    roi_align_plan_rows_kernel(PlanPtrs pl, int B, int H, int AH) {
  __shared__ int chunk_cnt[PLR_WARPS * PLR_ROUNDS];
  __shared__ int bin_base;
  const int lane = lane_id(), wid = warp_id();
  const int bin = blockIdx.x;
  if (wid == 0) {
    int before = 0;
    for (int k = lane; k < bin; k += 32) before += __ldg(pl.rowcnt + k);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) before += __shfl_xor_sync(0xffffffffu, before, d);
    if (lane == 0) {
      pl.rowptr[bin] = before;
      bin_base = before;
    }
  }
  const int b = bin / H, y = bin - b * H;
  const int lo = __ldg(pl.cum + b), hi = __ldg(pl.cum + b + 1);
  int written = 0;  // items of the passes before this one
  for (int base = lo; base < hi; base += 32 * PLR_WARPS * PLR_ROUNDS) {
    // chunk c = r * PLR_WARPS + wid of this pass: RoIs base + 32 c .. base + 32 c + 31
    int n[PLR_ROUNDS];
    int4 q0[PLR_ROUNDS], q1[PLR_ROUNDS];
#pragma unroll
    for (int r = 0; r < PLR_ROUNDS; ++r) {
      const int e = base + 32 * (r * PLR_WARPS + wid) + lane;
      n[r] = e < hi ? __ldg(pl.list + e) : -1;
    }
#pragma unroll
    for (int r = 0; r < PLR_ROUNDS; ++r) {
      q0[r] = make_int4(-1, -1, -1, -1);
      q1[r] = q0[r];
      if (n[r] >= 0) {
        const int4* yr = reinterpret_cast<const int4*>(pl.yrow + (size_t)n[r] * 16);
        q0[r] = __ldg(yr);
        q1[r] = __ldg(yr + 1);
      }
    }
    unsigned hit0[PLR_ROUNDS], hit1[PLR_ROUNDS];  // bit ph: plane row y is the sample row's first / second row
    int incl[PLR_ROUNDS];
#pragma unroll
    for (int r = 0; r < PLR_ROUNDS; ++r) {
      hit0[r] = hit1[r] = 0u;
      const int w[8] = {q0[r].x, q0[r].y, q0[r].z, q0[r].w, q1[r].x, q1[r].y, q1[r].z, q1[r].w};
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        const int ys = (int)(short)((k & 1) ? (w[k >> 1] >> 16) : (w[k >> 1] & 0xffff));
        if (k < AH && ys >= 0) {
          if (ys == y) hit0[r] |= 1u << k;
          if (ys + 1 == y) hit1[r] |= 1u << k;
        }
      }
      int v = __popc(hit0[r] | hit1[r]);
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += t;
      }
      incl[r] = v;
      if (lane == 31) chunk_cnt[r * PLR_WARPS + wid] = v;
    }
    __syncthreads();
    int pass_total = 0;
#pragma unroll
    for (int c = 0; c < PLR_WARPS * PLR_ROUNDS; ++c) pass_total += chunk_cnt[c];
#pragma unroll
    for (int r = 0; r < PLR_ROUNDS; ++r) {
      unsigned both = hit0[r] | hit1[r];
      if (both) {
        int pos = bin_base + written + incl[r] - __popc(both);
        for (int c = 0; c < r * PLR_WARPS + wid; ++c) pos += chunk_cnt[c];
        RowItem* o = pl.items + pos;
        const int key = (n[r] << 5) | ((int)__ldg(pl.jump + n[r]) << 4);
        while (both) {
          const int ph = __ffs(both) - 1;
          both &= both - 1u;
          const float4 t = __ldg(pl.tabs + (size_t)n[r] * 32 + ph);
          RowItem it;
          it.roi_ph = key | ph;
          it.weight = ((hit0[r] >> ph) & 1u) ? t.y : t.z;
          *o++ = it;
        }
      }
    }
    written += pass_total;
    __syncthreads();  // chunk_cnt is rewritten by the next pass
  }
}

// ===========================================================================
// plane-resident forward
// ===========================================================================
constexpr int PR_CH = 16;        // channels per slab == warps per CTA
constexpr int PR_THREADS = 512;  // 16 warps

constexpr int PR_TILE_BYTES = 2048;  // output staging tile of one warp: 16 channels x 4 rows x 8 floats

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

__host__ __device__ inline size_t pr_tiles_offset(int Pp) {
  return ((size_t)PR_CH * Pp * sizeof(float) + 1023) / 1024 * 1024;
}

__device__ __forceinline__ void st_global_v8(float* p, const float (&o)[8]) {
  asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "f"(o[0]),
               "f"(o[1]), "f"(o[2]), "f"(o[3]), "f"(o[4]), "f"(o[5]), "f"(o[6]), "f"(o[7])
               : "memory");
}

__device__ __forceinline__ void sts_v4(unsigned addr, float a, float b, float c, float d) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
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
//
// TMA_OUT (aligned_h == 8): the rows are not stored from registers (fully sector-efficient, but
// each 32-byte sector costs the LSU data pipe ~1.2 cycles) -- they go to a per-warp 2 KB tile
// [16 channels][4 rows x 8] with two conflict-free STS.128 and leave with one TMA tensor store
// per half RoI slab (rows 0-3, then rows 4-7; the output is (R, C, 2, 32) to the tensor map,
// box 32 x 1 x 16 x 1, 128-byte swizzle: 16-byte chunk k of tile row r sits at chunk k ^ (r & 7)).
// tile_row = shared address of this lane's channel row in the tile, cs = its swizzle key (c & 7).
template <int WP, bool TMA_OUT>
__device__ __forceinline__ void pr_fwd_roi_w8(unsigned plane_addr, int Wp_rt,
                                              const float4* __restrict__ wtab, int AH, int slot,
                                              float* __restrict__ out_c /* channel c of the RoI */,
                                              const void* omap, unsigned tile, unsigned tile_row, unsigned cs,
                                              int c0, int n) {
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
    if (TMA_OUT) {
      if ((j & 1) == 0) {  // the previous half's store must have read the tile
        bulk_wait_read_all();
        __syncwarp();
      }
      const unsigned k = (unsigned)((ph & 3) * 2);
      sts_v4(tile_row + ((k ^ cs) << 4), o[0], o[1], o[2], o[3]);
      sts_v4(tile_row + (((k | 1u) ^ cs) << 4), o[4], o[5], o[6], o[7]);
      if (j & 1) {
        fence_proxy_async_smem();
        __syncwarp();
        if (elect_one()) {
          tma_store_4d(omap, tile, 0, j >> 1, c0, n);
          bulk_commit_group();
        }
      }
    } else if (2 * j + slot < AH) {
      st_global_v8(out_c + ph * 8, o);
    }
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
template <int W8, bool TMA_OUT>
__global__ void __launch_bounds__(PR_THREADS, 1)
    roi_align_fwd_planes_kernel(const __grid_constant__ CUtensorMap omap, const float* __restrict__ features,
                                float* __restrict__ output, PlanPtrs pl, int B, int C, int H, int W, int R,
                                int AH, int AW, int Pp) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  // [planes][output tiles (TMA_OUT), 1024-byte aligned][PRShared]
  float* planes = reinterpret_cast<float*>(smem_raw);
  const size_t tiles_off = pr_tiles_offset(Pp);
  PRShared& sh = *reinterpret_cast<PRShared*>(
      smem_raw + (TMA_OUT ? tiles_off + PR_TILE_BYTES * (PR_THREADS / 32) : (size_t)PR_CH * Pp * sizeof(float)));
  const int tid = threadIdx.x, lane = lane_id(), wid = warp_id();
  const int nslabs = C / PR_CH;
  const int P = H * W, S = AH * AW;
  const int Wp = pr_row_stride(W);
  // channel of the lane pair p = lane >> 1: bits (p0, p1, p2, p3) -> (c0, c2, c1, c3), so that the
  // four channels of a quarter warp differ in c0 and c2 (the output tile's swizzle needs that)
  const int pr = lane >> 1, slot = lane & 1;
  const int c = (pr & 9) | ((pr & 2) << 1) | ((pr & 4) >> 1);
  const unsigned tile = TMA_OUT ? smem_u32(smem_raw + tiles_off + (size_t)wid * PR_TILE_BYTES) : 0u;
  const unsigned tile_row = tile + 128u * (unsigned)c, cs = (unsigned)(c & 7);
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
          pr_fwd_roi_w8<(W8 > 1 ? W8 : 0), TMA_OUT>(plane_addr, Wp, wtab, AH, slot, out_roi + (size_t)c * S,
                                                     &omap, tile, tile_row, cs, c0, n);
        __syncwarp();
      }
      n = n2;
      t = t2;
    }
    __syncthreads();
    u += (r_hi - r_lo);
  }
  if (TMA_OUT) bulk_wait_all();  // the tiles must outlive their stores
}

// ===========================================================================
// host side
// ===========================================================================
int roi_align_check_common(const void* a, const void* b, const void* c, int batch, int channels,
                        int height, int width, int num_rois, int ah, int aw) {
  if (!a || !b || !c) return TLOD_ERR_NULL_POINTER;
  if (batch <= 0 || channels <= 0 || height < 2 || width < 2 || num_rois < 0 || ah < 2 || aw < 2)
    return TLOD_ERR_BAD_SHAPE;
  if ((long long)batch * channels * height * width >= (1LL << 31)) return TLOD_ERR_INT32_OVERFLOW;
  return TLOD_OK;
}

static size_t pr_smem_bytes(int h, int w, bool tma_out) {
  const int Pp = pr_plane_stride(h, w);
  if (tma_out) return pr_tiles_offset(Pp) + (size_t)PR_TILE_BYTES * (PR_THREADS / 32) + sizeof(PRShared);
  return (size_t)PR_CH * Pp * sizeof(float) + sizeof(PRShared);
}

// (R, C, 8, 8) fp32 output seen as (R, C, 2, 32): box = rows 0-3 or 4-7 of 16 consecutive channels
static bool make_out_tmap(CUtensorMap* map, float* output, int R, int C) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) return false;
  const cuuint64_t dims[4] = {32, 2, (cuuint64_t)C, (cuuint64_t)R};
  const cuuint64_t strides[3] = {128, 256, (cuuint64_t)C * 256};
  const cuuint32_t box[4] = {32, 1, (cuuint32_t)PR_CH, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, output, dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

int roi_align_generic_launch(bool backward, const float* src, const float* rois, float* dst, int batch,
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
  // backward row lists (aligned_w == 8 only, like the row-resident backward itself)
  const bool rows = aligned_w == 8 && plan_has_row_lists(batch, height);
  const int bins = batch * height;
  if (rows) {
    cudaError_t e = cudaMemsetAsync(pl.rowcnt, 0, (size_t)bins * sizeof(int), st);
    if (e != cudaSuccess) return (int)e;
  }
  {
    LaunchScope scope("roi_align_plan_kernel", st);
    roi_align_plan_kernel<<<grid, PL_THREADS, 0, st>>>(rois, pl, batch, height, width, num_rois,
                                                       aligned_h, aligned_w, spatial_scale, rows ? 1 : 0);
  }
  int rc = last_launch_status();
  if (rc != TLOD_OK || !rows) return rc;
  {
    LaunchScope scope("roi_align_plan_rows_kernel", st);
    roi_align_plan_rows_kernel<<<bins, PL_THREADS, 0, st>>>(pl, batch, height, aligned_h);
  }
  return last_launch_status();
}

extern "C" int tlod_roi_align_forward(const float* features, const float* rois, float* output,
                                      int batch, int channels, int height, int width, int num_rois,
                                      int aligned_h, int aligned_w, float spatial_scale,
                                      const void* plan, size_t plan_bytes, void* stream) {
  int rc = roi_align_check_common(features, rois, output, batch, channels, height, width, num_rois, aligned_h,
                        aligned_w);
  if (rc != TLOD_OK) return rc;
  if (num_rois == 0) return TLOD_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const bool planned = plan != nullptr && plan_supported(batch, height, width, aligned_h, aligned_w);
  if (plan && (plan_bytes < tlod_roi_align_plan_bytes(batch, num_rois) || ((uintptr_t)plan & 255)))
    return TLOD_ERR_WORKSPACE;
  if (planned && aligned_h == 8 && aligned_w == 8) {
    rc = roi_align_fwd8_launch(false, features, output, batch, channels, height, width, num_rois, plan, st);
    if (rc != TLOD_ERR_UNSUPPORTED) return rc;
  }
  if (planned && channels % PR_CH == 0 &&
      pr_smem_bytes(height, width, false) <= (size_t)device_info().max_smem_optin) {
    const int Pp = pr_plane_stride(height, width);
    const long long units = (long long)(channels / PR_CH) * num_rois;
    int grid = device_info().sm_count;
    if ((long long)grid > units) grid = (int)units;
    const bool w8 = aligned_w == 8 && ((uintptr_t)output & 31) == 0;
    // 8x8 outputs leave through per-warp staging tiles and TMA stores when the tiles fit beside the planes
    CUtensorMap omap;
    memset(&omap, 0, sizeof(omap));
    const bool tma_out = w8 && aligned_h == 8 && ((uintptr_t)output & 127) == 0 &&
                         pr_smem_bytes(height, width, true) <= (size_t)device_info().max_smem_optin &&
                         make_out_tmap(&omap, output, num_rois, channels);
    const size_t smem = pr_smem_bytes(height, width, tma_out);
    const bool w76 = pr_row_stride(width) == 76;
    auto kern = !w8 ? roi_align_fwd_planes_kernel<0, false>
                : tma_out ? (w76 ? roi_align_fwd_planes_kernel<76, true> : roi_align_fwd_planes_kernel<1, true>)
                          : (w76 ? roi_align_fwd_planes_kernel<76, false> : roi_align_fwd_planes_kernel<1, false>);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    const PlanPtrs pl = plan_ptrs(const_cast<void*>(plan), batch, num_rois);
    {
      LaunchScope scope("roi_align_fwd_planes_kernel", st);
      kern<<<grid, PR_THREADS, smem, st>>>(omap, features, output, pl, batch, channels, height, width,
                                           num_rois, aligned_h, aligned_w, Pp);
    }
    return last_launch_status();
  }
  return roi_align_generic_launch(false, features, rois, output, batch, channels, height, width, num_rois,
                        aligned_h, aligned_w, spatial_scale, st);
}
