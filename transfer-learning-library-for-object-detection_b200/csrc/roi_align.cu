// RoIAlign forward / backward for sm_100a.
//
// Reference semantics: lib/model/roi_align/src/roi_align_kernel.cu:15-70 (fwd),
// :94-143 (bwd); glue lib/model/roi_align/src/roi_align_cuda.c.
//
// Two families of kernels:
//
//  * "plane-resident" kernels (the product path for every BASELINE shape): a CTA
//    keeps the 16 feature planes of one (image, 16-channel slab) in shared
//    memory (16 x H*W fp32 = 178-182 KB for the 37x75 / 38x75 maps) and
//    streams the image's RoIs through them.  HBM traffic is then the
//    algorithmic minimum: each plane is read (fwd) or written (bwd) once, the
//    (R, C, AH, AW) tensor is streamed once with full-sector accesses.
//    Lane mapping: lane = 2*channel + slot.  The plane stride is padded to
//    2 (mod 4) floats so that the 16 channels land on the 16 even (or odd)
//    banks; the two slots of a channel always touch cells of opposite column
//    parity, so every shared-memory access of the gather/scatter is
//    bank-conflict free regardless of the RoI geometry.
//
//  * generic kernels for shapes the plane-resident layout cannot hold
//    (channels % 16 != 0, planes too large for shared memory, aligned size
//    > 16): one CTA per (RoI, channel block), geometry hoisted to shared
//    memory, coalesced output, fp32 atomics in the backward.
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
// plane-resident forward
// ===========================================================================
constexpr int PR_CH = 16;        // channels per slab == warps per CTA
constexpr int PR_THREADS = 512;  // 16 warps
constexpr int PR_LIST = 1024;    // RoI indices staged per refill
constexpr int PR_MAXB = 1024;    // images per call on this path
constexpr int PR_MAXA = 16;      // aligned_h / aligned_w limit on this path

struct PRShared {
  int cnt[PR_MAXB + 2];  // RoIs per image (+1 slot: RoIs with an invalid image index)
  int cum[PR_MAXB + 2];  // exclusive prefix of cnt
  int list[PR_LIST];     // RoI indices of the current (image, rank range)
  int warp_sums[PR_THREADS / 32];
  int cur[4];            // broadcast slots: image, slab, rank_lo, rank_hi
  AxisTab rows[PR_THREADS / 32][PR_MAXA];
  AxisTab cols[PR_THREADS / 32][PR_MAXA];
};

__host__ __device__ inline int pr_plane_stride(int hw) {
  int p = hw;
  while ((p & 3) != 2) ++p;  // 2 (mod 4): 16 channels -> 16 distinct same-parity banks
  return p;
}

// Exclusive prefix over cnt[0..nb) by warp 0; cum[nb] = total.
__device__ inline void pr_prefix(PRShared& sh, int nb) {
  if (warp_id() == 0) {
    int carry = 0;
    for (int base = 0; base < nb; base += 32) {
      int i = base + lane_id();
      int v = (i < nb) ? sh.cnt[i] : 0;
      int incl = v;
      for (int d = 1; d < 32; d <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane_id() >= d) incl += t;
      }
      if (i < nb) sh.cum[i] = carry + incl - v;
      carry += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (lane_id() == 0) sh.cum[nb] = carry;
  }
}

// Fill sh.list with the indices of the RoIs of image `img` whose rank (position
// among that image's RoIs in index order) lies in [lo, hi), hi - lo <= PR_LIST.
__device__ inline void pr_build_list(PRShared& sh, const float* __restrict__ rois, int R, int B,
                                     int img, int lo, int hi) {
  int running = 0;
  for (int base = 0; base < R && running < hi; base += PR_THREADS) {
    const int i = base + threadIdx.x;
    bool m = false;
    if (i < R) {
      int b = (int)__ldg(rois + (size_t)i * 5);
      if (b < 0 || b >= B) b = B;
      m = (b == img);
    }
    const unsigned bal = __ballot_sync(0xffffffffu, m);
    if (lane_id() == 0) sh.warp_sums[warp_id()] = __popc(bal);
    __syncthreads();
    int before = running;
    int total = 0;
#pragma unroll
    for (int w = 0; w < PR_THREADS / 32; ++w) {
      const int s = sh.warp_sums[w];
      if (w < warp_id()) before += s;
      total += s;
    }
    if (m) {
      const int rank = before + __popc(bal & ((1u << lane_id()) - 1u));
      if (rank >= lo && rank < hi) sh.list[rank - lo] = i;
    }
    running += total;
    __syncthreads();
  }
}

__device__ __forceinline__ void st_global_v4(float* p, float a, float b, float c, float d) {
  asm volatile("st.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d)
               : "memory");
}

// One warp, one RoI, 16 channels, AW == 8 and AH even: lane (c, slot) produces the
// full 8-wide output row ph = 2*j + slot of channel c (32 contiguous bytes).
__device__ __forceinline__ void pr_fwd_roi_fast8(const float* __restrict__ planes, int Pp, int W,
                                                 const AxisTab* rows, const AxisTab* cols, int AH,
                                                 float* __restrict__ out_roi /* (16, AH, 8) */) {
  const int lane = lane_id();
  const int c = lane >> 1, slot = lane & 1;
  const float* plane = planes + c * Pp;
  int cx[8];
  float cw0[8], cw1[8];
#pragma unroll
  for (int pw = 0; pw < 8; ++pw) {
    const AxisTab t = cols[pw];
    cx[pw] = t.off;
    cw0[pw] = t.w0;
    cw1[pw] = t.w1;
  }
  for (int j = 0; j < AH / 2; ++j) {
    const int ph = 2 * j + slot;
    const AxisTab row = rows[ph];
    const int roff = row.off < 0 ? 0 : row.off;
    const int other = __shfl_xor_sync(0xffffffffu, roff, 1);
    // slot 1 reads (x+1, x) instead of (x, x+1) when both rows start on the same
    // parity, so that the two half-warps always hit opposite bank parities.
    const int flip = slot & (((roff ^ other) & 1) ^ 1);
    float o[8];
#pragma unroll
    for (int pw = 0; pw < 8; ++pw) {
      const int x = cx[pw] < 0 ? 0 : cx[pw];
      const float* p = plane + roff + x;
      const float a = p[flip];
      const float b = p[flip ^ 1];
      const float cc = p[W + flip];
      const float d = p[W + (flip ^ 1)];
      const float wa = flip ? cw1[pw] : cw0[pw];
      const float wb = flip ? cw0[pw] : cw1[pw];
      float v = a * (row.w0 * wa);
      v = fmaf(b, row.w0 * wb, v);
      v = fmaf(cc, row.w1 * wa, v);
      v = fmaf(d, row.w1 * wb, v);
      o[pw] = (row.off < 0 || cx[pw] < 0) ? 0.f : v;
    }
    float* dst = out_roi + ((size_t)c * AH + ph) * 8;
    st_global_v4(dst, o[0], o[1], o[2], o[3]);
    st_global_v4(dst + 4, o[4], o[5], o[6], o[7]);
  }
}

// Any AH, AW <= 16: lane (c, slot) produces samples i = 2k + slot of channel c.
__device__ __forceinline__ void pr_fwd_roi_any(const float* __restrict__ planes, int Pp, int W,
                                               const AxisTab* rows, const AxisTab* cols, int AH,
                                               int AW, float* __restrict__ out_roi) {
  const int lane = lane_id();
  const int c = lane >> 1, slot = lane & 1;
  const float* plane = planes + c * Pp;
  const int S = AH * AW;
  for (int k = 0; 2 * k < S; ++k) {
    const int i = min(2 * k + slot, S - 1);
    const int ph = i / AW, pw = i - ph * AW;
    const AxisTab row = rows[ph], col = cols[pw];
    const bool ok = row.off >= 0 && col.off >= 0;
    const int off = ok ? row.off + col.off : 0;
    const int other = __shfl_xor_sync(0xffffffffu, off, 1);
    const int flip = slot & (((off ^ other) & 1) ^ 1);
    const float* p = plane + off;
    const float a = p[flip];
    const float b = p[flip ^ 1];
    const float cc = p[W + flip];
    const float d = p[W + (flip ^ 1)];
    const float wa = flip ? col.w1 : col.w0;
    const float wb = flip ? col.w0 : col.w1;
    float v = a * (row.w0 * wa);
    v = fmaf(b, row.w0 * wb, v);
    v = fmaf(cc, row.w1 * wa, v);
    v = fmaf(d, row.w1 * wb, v);
    if (2 * k + slot < S) out_roi[(size_t)c * S + i] = ok ? v : 0.f;
  }
}

template <bool FAST8>
__global__ void __launch_bounds__(PR_THREADS, 1)
    roi_align_fwd_planes_kernel(const float* __restrict__ features, const float* __restrict__ rois,
                                float* __restrict__ output, int B, int C, int H, int W, int R,
                                int AH, int AW, float scale, int Pp) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* planes = reinterpret_cast<float*>(smem_raw);
  PRShared& sh = *reinterpret_cast<PRShared*>(smem_raw + (size_t)PR_CH * Pp * sizeof(float));
  const int tid = threadIdx.x, lane = lane_id(), wid = warp_id();
  const int nslabs = C / PR_CH;
  const int P = H * W, S = AH * AW;

  // ---- RoIs per image (slot B collects RoIs with an invalid image index) ----
  for (int i = tid; i <= B; i += PR_THREADS) sh.cnt[i] = 0;
  __syncthreads();
  for (int i = tid; i < R; i += PR_THREADS) {
    int b = (int)__ldg(rois + (size_t)i * 5);
    if (b < 0 || b >= B) b = B;
    atomicAdd(&sh.cnt[b], 1);
  }
  __syncthreads();
  pr_prefix(sh, B + 1);
  __syncthreads();

  // ---- this CTA's contiguous range of (image, slab, RoI-rank) units ----
  const long long U = (long long)nslabs * R;
  long long u = U * blockIdx.x / gridDim.x;
  const long long u_end = U * (blockIdx.x + 1) / gridDim.x;
  int staged_img = -1, staged_slab = -1;

  while (u < u_end) {
    if (tid == 0) {
      // image b with nslabs*cum[b] <= u < nslabs*cum[b+1]
      int lo = 0, hi = B;  // invariant: answer in [lo, hi]
      while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if ((long long)nslabs * sh.cum[mid] <= u) lo = mid; else hi = mid - 1;
      }
      // skip images without RoIs (cum[b] == cum[b+1])
      while (lo < B && sh.cnt[lo] == 0) ++lo;
      const long long rem = u - (long long)nslabs * sh.cum[lo];
      const int slab = (int)(rem / sh.cnt[lo]);
      const int r_lo = (int)(rem - (long long)slab * sh.cnt[lo]);
      long long room = u_end - u;
      int r_hi = sh.cnt[lo];
      if (room < (long long)(r_hi - r_lo)) r_hi = r_lo + (int)room;
      sh.cur[0] = lo; sh.cur[1] = slab; sh.cur[2] = r_lo; sh.cur[3] = r_hi;
    }
    __syncthreads();
    const int img = sh.cur[0], slab = sh.cur[1], r_lo = sh.cur[2], r_hi = sh.cur[3];
    const int c0 = slab * PR_CH;

    // ---- stage the 16 planes of (img, slab): warp w copies channel w ----
    if (img < B && (img != staged_img || slab != staged_slab)) {
      const float* g = features + ((size_t)img * C + c0 + wid) * P;
      float* s = planes + wid * Pp;
      int i = lane;
      for (; i + 7 * 32 < P; i += 8 * 32) {
        float v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = __ldg(g + i + k * 32);
#pragma unroll
        for (int k = 0; k < 8; ++k) s[i + k * 32] = v[k];
      }
      for (; i < P; i += 32) s[i] = __ldg(g + i);
      staged_img = img;
      staged_slab = slab;
    }

    for (int base = r_lo; base < r_hi; base += PR_LIST) {
      const int top = min(r_hi, base + PR_LIST);
      pr_build_list(sh, rois, R, B, img, base, top);  // ends with __syncthreads()
      for (int e = wid; e < top - base; e += PR_THREADS / 32) {
        const int n = sh.list[e];
        float* out_roi = output + ((size_t)n * C + c0) * S;
        if (img == B) {  // invalid image index: zeros
          for (int i = lane; i < PR_CH * S; i += 32) out_roi[i] = 0.f;
          continue;
        }
        const float* r = rois + (size_t)n * 5;
        if (lane < AH)
          sh.rows[wid][lane] = make_tab(align_axis(__ldg(r + 2), __ldg(r + 4), scale, AH, H, lane), W);
        else if (lane >= 16 && lane < 16 + AW)
          sh.cols[wid][lane - 16] =
              make_tab(align_axis(__ldg(r + 1), __ldg(r + 3), scale, AW, W, lane - 16), 1);
        __syncwarp();
        if (FAST8)
          pr_fwd_roi_fast8(planes, Pp, W, sh.rows[wid], sh.cols[wid], AH, out_roi);
        else
          pr_fwd_roi_any(planes, Pp, W, sh.rows[wid], sh.cols[wid], AH, AW, out_roi);
        __syncwarp();
      }
      __syncthreads();
    }
    u += (r_hi - r_lo);
  }
}

// ===========================================================================
// plane-resident backward (no atomics)
// ===========================================================================
// One CTA owns the gradient planes of (image, 16-channel slab, row band) exclusively: they
// are accumulated in shared memory and written to HBM once with plain coalesced stores, so
// bottom_grad needs neither a memset nor a single atomic (the reference issues
// 4 * R * C * AH * AW global fp32 REDs, roi_align_kernel.cu:131-134).
//
// Inside the CTA the 16 warps never touch the same cell: warp w owns the plane rows
// y == w (mod 16) of the band.  A RoI contributes 2*AH "row entries" (ph, dy) -> row
// hs[ph] + dy; each entry is scattered by the one warp that owns that row, lane =
// 2*channel + slot, slot = half of the 8 samples of the row.  As in the forward the plane
// stride is 2 (mod 4) and the two slots update cells of opposite column parity, so every
// read-modify-write instruction is bank-conflict free.  The order of additions into a cell
// is fixed (RoI index, then ph, dy, pw), so the result is bitwise reproducible.
//
// The (16, AH, 8) gradient tiles of the next PB_NB RoIs are fetched with cp.async into a
// second shared-memory stage while the current batch is scattered.
constexpr int PB_NB = 4;      // RoIs per batch (stage)
constexpr int PB_LIST = 512;  // RoI indices per refill

struct PBShared {
  int list[PB_LIST];
  int warp_sums[PR_THREADS / 32];
  AxisTab rows[2][PB_NB][PR_MAXA];  // off = first row index (not scaled) or -1
  AxisTab cols[2][PB_NB][PR_MAXA];
};

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// RoIs of image `img` whose sampled rows can intersect [y_lo, y_hi), ranks [lo, hi) -> list
__device__ inline int pb_build_list(PBShared& sh, const float* __restrict__ rois, int R, int img,
                                    int AH, int H, float scale, int y_lo, int y_hi, int lo, int hi) {
  int running = 0;
  for (int base = 0; base < R; base += PR_THREADS) {
    const int i = base + threadIdx.x;
    bool m = false;
    if (i < R) {
      const float* r = rois + (size_t)i * 5;
      if ((int)__ldg(r) == img) {
        const AlignAxis a0 = align_axis(__ldg(r + 2), __ldg(r + 4), scale, AH, H, 0);
        const AlignAxis a1 = align_axis(__ldg(r + 2), __ldg(r + 4), scale, AH, H, AH - 1);
        m = (a0.start < y_hi) && (a1.start + 1 >= y_lo);
      }
    }
    const unsigned bal = __ballot_sync(0xffffffffu, m);
    if (lane_id() == 0) sh.warp_sums[warp_id()] = __popc(bal);
    __syncthreads();
    int before = running, total = 0;
#pragma unroll
    for (int w = 0; w < PR_THREADS / 32; ++w) {
      const int v = sh.warp_sums[w];
      if (w < warp_id()) before += v;
      total += v;
    }
    if (m) {
      const int rank = before + __popc(bal & ((1u << lane_id()) - 1u));
      if (rank >= lo && rank < hi) sh.list[rank - lo] = i;
    }
    running += total;
    __syncthreads();
  }
  return running;  // number of matching RoIs in the whole array
}

__global__ void __launch_bounds__(PR_THREADS, 1)
    roi_align_bwd_planes_kernel(const float* __restrict__ top_grad, const float* __restrict__ rois,
                                float* __restrict__ bottom_grad, int B, int C, int H, int W, int R,
                                int AH, float scale, int nsplit, int band_rows, int Ppb, int TS) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* planes = reinterpret_cast<float*>(smem_raw);
  float* tiles = planes + (size_t)PR_CH * Ppb;                       // [2][PB_NB][16 * TS]
  PBShared& sh = *reinterpret_cast<PBShared*>(tiles + 2 * PB_NB * PR_CH * TS);
  const int tid = threadIdx.x, lane = lane_id(), wid = warp_id();
  const int nslabs = C / PR_CH;
  const int band = blockIdx.x % nsplit;
  const int pair = blockIdx.x / nsplit;
  const int img = pair / nslabs, c0 = (pair % nslabs) * PR_CH;
  const int y_lo = band * band_rows, y_hi = min(H, y_lo + band_rows);
  const int S = AH * 8, S4 = S / 4;
  const int tile_floats = PR_CH * TS;

  for (int i = tid; i < PR_CH * Ppb; i += PR_THREADS) planes[i] = 0.f;

  const int c = lane >> 1, slot = lane & 1;
  int done = 0, total = 1;
  while (done < total) {
    total = pb_build_list(sh, rois, R, img, AH, H, scale, y_lo, y_hi, done, done + PB_LIST);
    const int cnt = min(PB_LIST, total - done);  // list entries valid this round
    const int nbatch = (cnt + PB_NB - 1) / PB_NB;

    auto issue = [&](int k) {  // gradient tiles of batch k -> stage k & 1
      const int nb = min(PB_NB, cnt - k * PB_NB);
      float* stage = tiles + (size_t)(k & 1) * PB_NB * tile_floats;
      for (int q = tid; q < nb * PR_CH * S4; q += PR_THREADS) {
        const int j = q / (PR_CH * S4);
        const int rem = q - j * (PR_CH * S4);
        const int ch = rem / S4, f = rem - ch * S4;
        const int n = sh.list[k * PB_NB + j];
        cp_async16(stage + j * tile_floats + ch * TS + f * 4,
                   top_grad + ((size_t)n * C + c0 + ch) * S + f * 4);
      }
      cp_async_commit();
    };

    if (nbatch > 0) issue(0);
    for (int k = 0; k < nbatch; ++k) {
      const int nb = min(PB_NB, cnt - k * PB_NB);
      if (k + 1 < nbatch) issue(k + 1);
      // sampling tables of batch k
      if (tid < nb * 32) {
        const int j = tid >> 5, e = tid & 31;
        const float* r = rois + (size_t)sh.list[k * PB_NB + j] * 5;
        if (e < AH) {
          const AlignAxis a = align_axis(__ldg(r + 2), __ldg(r + 4), scale, AH, H, e);
          AxisTab t = make_tab(a, 1);
          sh.rows[k & 1][j][e] = t;
        } else if (e >= 16 && e < 24) {
          sh.cols[k & 1][j][e - 16] = make_tab(align_axis(__ldg(r + 1), __ldg(r + 3), scale, 8, W, e - 16), 1);
        }
      }
      if (k + 1 < nbatch) cp_async_wait<1>(); else cp_async_wait<0>();
      __syncthreads();

      const float* stage = tiles + (size_t)(k & 1) * PB_NB * tile_floats;
      for (int j = 0; j < nb; ++j) {
        const AxisTab* rows = sh.rows[k & 1][j];
        const AxisTab* cols = sh.cols[k & 1][j];
        // which of the 2*AH row entries (ph = lane >> 1, dy = lane & 1) does this warp own?
        bool own = false;
        if ((lane >> 1) < AH) {
          const int r0 = rows[lane >> 1].off;
          const int y = r0 + (lane & 1);
          own = r0 >= 0 && y >= y_lo && y < y_hi && ((y - y_lo) & 15) == wid;
        }
        unsigned mine = __ballot_sync(0xffffffffu, own);
        if (!mine) continue;
        int cx[4], flip[4];
        float cw0[4], cw1[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const AxisTab t = cols[4 * slot + q];
          const int xp = cols[4 * (slot ^ 1) + q].off;
          cx[q] = t.off;
          cw0[q] = t.w0;
          cw1[q] = t.w1;
          // slot 1 swaps its (x, x+1) order when both slots start on the same parity
          flip[q] = slot & (((t.off ^ xp) & 1) ^ 1);
        }
        const float* g_roi = stage + j * tile_floats + c * TS + 4 * slot;
        while (mine) {
          const int e = __ffs(mine) - 1;
          mine &= mine - 1u;
          const int ph = e >> 1, dy = e & 1;
          const AxisTab rt = rows[ph];
          const float rw = dy ? rt.w1 : rt.w0;
          float* prow = planes + c * Ppb + (rt.off + dy - y_lo) * W;
          const float4 g = *reinterpret_cast<const float4*>(g_roi + ph * 8);
          const float gv[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float mv = rw * gv[q];
            const int x = cx[q], f = flip[q];
            if (x >= 0) prow[x + f] += mv * (f ? cw1[q] : cw0[q]);
            __syncwarp();
            if (x >= 0) prow[x + (f ^ 1)] += mv * (f ? cw0[q] : cw1[q]);
            __syncwarp();
          }
        }
      }
      __syncthreads();  // stage k & 1 and tables k & 1 are free again
    }
    done += cnt;
    if (cnt == 0) break;
  }
  __syncthreads();
  // ---- write the band of the 16 planes: warp w stores channel w ----
  {
    const int n_cells = (y_hi - y_lo) * W;
    float* g = bottom_grad + ((size_t)img * C + c0 + wid) * H * W + (size_t)y_lo * W;
    const float* sp = planes + wid * Ppb;
    for (int i = lane; i < n_cells; i += 32) g[i] = sp[i];
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

static size_t pr_smem_bytes(int hw) {
  return (size_t)PR_CH * pr_plane_stride(hw) * sizeof(float) + sizeof(PRShared);
}

static bool pr_applicable(int batch, int channels, int height, int width, int ah, int aw) {
  if (channels % PR_CH != 0 || batch > PR_MAXB || ah > PR_MAXA || aw > PR_MAXA) return false;
  return pr_smem_bytes(height * width) <= (size_t)device_info().max_smem_optin;
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

extern "C" int tlod_roi_align_forward(const float* features, const float* rois, float* output,
                                      int batch, int channels, int height, int width, int num_rois,
                                      int aligned_h, int aligned_w, float spatial_scale,
                                      void* stream) {
  int rc = check_common(features, rois, output, batch, channels, height, width, num_rois, aligned_h,
                        aligned_w);
  if (rc != TLOD_OK) return rc;
  if (num_rois == 0) return TLOD_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const bool aligned16 = ((uintptr_t)output & 15) == 0;
  if (pr_applicable(batch, channels, height, width, aligned_h, aligned_w)) {
    const int Pp = pr_plane_stride(height * width);
    const size_t smem = pr_smem_bytes(height * width);
    const long long units = (long long)(channels / PR_CH) * num_rois;
    int grid = device_info().sm_count;
    if ((long long)grid > units) grid = (int)units;
    const bool fast8 = aligned_w == 8 && (aligned_h % 2) == 0 && aligned16;
    auto kern = fast8 ? roi_align_fwd_planes_kernel<true> : roi_align_fwd_planes_kernel<false>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    {
      LaunchScope scope("roi_align_fwd_planes_kernel", st);
      kern<<<grid, PR_THREADS, smem, st>>>(features, rois, output, batch, channels, height, width,
                                           num_rois, aligned_h, aligned_w, spatial_scale, Pp);
    }
    return last_launch_status();
  }
  return generic_launch(false, features, rois, output, batch, channels, height, width, num_rois,
                        aligned_h, aligned_w, spatial_scale, st);
}

extern "C" int tlod_roi_align_backward(const float* top_grad, const float* rois, float* bottom_grad,
                                       int batch, int channels, int height, int width,
                                       int num_rois, int aligned_h, int aligned_w,
                                       float spatial_scale, void* stream) {
  int rc = check_common(top_grad, rois, bottom_grad, batch, channels, height, width, num_rois,
                        aligned_h, aligned_w);
  if (rc != TLOD_OK) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t grad_bytes = (size_t)batch * channels * height * width * sizeof(float);
  if (num_rois == 0) return (int)cudaMemsetAsync(bottom_grad, 0, grad_bytes, st);

  // plane-resident path: AW == 8 (one float4 per slot), tiles fetched in 16-byte pieces
  if (channels % PR_CH == 0 && aligned_w == 8 && aligned_h <= PR_MAXA &&
      ((uintptr_t)top_grad & 15) == 0) {
    const int pairs = batch * (channels / PR_CH);
    const int sms = device_info().sm_count;
    int nsplit = (2 * sms + pairs - 1) / pairs;  // row bands per plane: fill the machine twice over
    if (nsplit > 4) nsplit = 4;
    if (nsplit > height / 8) nsplit = height / 8 > 0 ? height / 8 : 1;
    if (nsplit < 1) nsplit = 1;
    const int S = aligned_h * 8;
    const int TS = (S + 31) / 32 * 32 + 8;  // channel stride of a tile: 8 (mod 32) floats
    for (; nsplit <= 64; ++nsplit) {  // more bands if the planes do not fit shared memory
      const int band_rows = (height + nsplit - 1) / nsplit;
      const int Ppb = pr_plane_stride(band_rows * width);
      const size_t smem = ((size_t)PR_CH * Ppb + (size_t)2 * PB_NB * PR_CH * TS) * sizeof(float) +
                          sizeof(PBShared);
      if (smem > (size_t)device_info().max_smem_optin) {
        if (band_rows <= 1) break;
        continue;
      }
      if ((long long)pairs * nsplit > 2147483647LL) break;
      cudaError_t e = cudaFuncSetAttribute(roi_align_bwd_planes_kernel,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return (int)e;
      {
        LaunchScope scope("roi_align_bwd_planes_kernel", st);
        roi_align_bwd_planes_kernel<<<pairs * nsplit, PR_THREADS, smem, st>>>(
            top_grad, rois, bottom_grad, batch, channels, height, width, num_rois, aligned_h,
            spatial_scale, nsplit, band_rows, Ppb, TS);
      }
      return last_launch_status();
    }
  }
  cudaError_t e = cudaMemsetAsync(bottom_grad, 0, grad_bytes, st);
  if (e != cudaSuccess) return (int)e;
  return generic_launch(true, top_grad, rois, bottom_grad, batch, channels, height, width, num_rois,
                        aligned_h, aligned_w, spatial_scale, st);
}
