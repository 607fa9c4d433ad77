// RoIPool (Caffe max pooling over RoI bins) forward / backward for sm_100a.
//
// Reference semantics: lib/model/roi_pooling/src/roi_pooling_kernel.cu:24-93
// (forward, argmax saved) and :128-203 (backward, a gather over all RoIs).
// The backward here is a scatter by argmax that applies exactly the
// reference's admission tests (same image, same channel, cell inside the
// rounded RoI, bin inside the cell's feasible set), so it differs from the
// reference only in fp32 summation order.
#include <float.h>
#include <stdlib.h>

#include "async_copy.cuh"
#include "common.cuh"

namespace tlod {

struct PoolRoi {
  int batch, sw, sh, ew, eh;
  float bin_h, bin_w;
};

// roi_pooling_kernel.cu:44-55
__device__ __forceinline__ PoolRoi pool_roi(const float* __restrict__ r, float scale, int PH,
                                            int PW) {
  PoolRoi g;
  g.batch = (int)__ldg(r);
  g.sw = (int)roundf(__fmul_rn(__ldg(r + 1), scale));
  g.sh = (int)roundf(__fmul_rn(__ldg(r + 2), scale));
  g.ew = (int)roundf(__fmul_rn(__ldg(r + 3), scale));
  g.eh = (int)roundf(__fmul_rn(__ldg(r + 4), scale));
  const int rw = max(g.ew - g.sw + 1, 1);
  const int rh = max(g.eh - g.sh + 1, 1);
  g.bin_h = __fdiv_rn((float)rh, (float)PH);
  g.bin_w = __fdiv_rn((float)rw, (float)PW);
  return g;
}

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return min(max(v, lo), hi); }

// PHT, PWT > 0: compile-time pooled size (7 x 7: the index divisions become multiplications).
template <int PHT, int PWT>
__global__ void __launch_bounds__(256)
    roi_pool_fwd_kernel(const float* __restrict__ features, const float* __restrict__ rois,
                        float* __restrict__ output, int* __restrict__ argmax, int B, int C, int H,
                        int W, int PH_rt, int PW_rt, float scale, int chans_per_block) {
  const int PH = PHT > 0 ? PHT : PH_rt, PW = PWT > 0 ? PWT : PW_rt;
  const int n = blockIdx.x;
  const int c0 = blockIdx.y * chans_per_block;
  const PoolRoi g = pool_roi(rois + (size_t)n * 5, scale, PH, PW);
  const int S = PH * PW;
  const int cb = min(chans_per_block, C - c0);
  const bool image_ok = g.batch >= 0 && g.batch < B;
  const size_t roi_base = ((size_t)n * C + c0) * S;
  for (int o = threadIdx.x; o < cb * S; o += blockDim.x) {
    const int c = o / S;
    const int i = o - c * S;
    const int ph = i / PW;
    const int pw = i - ph * PW;
    int hs = (int)floorf(__fmul_rn((float)ph, g.bin_h));
    int he = (int)ceilf(__fmul_rn((float)(ph + 1), g.bin_h));
    int ws = (int)floorf(__fmul_rn((float)pw, g.bin_w));
    int we = (int)ceilf(__fmul_rn((float)(pw + 1), g.bin_w));
    hs = clampi(hs + g.sh, 0, H);
    he = clampi(he + g.sh, 0, H);
    ws = clampi(ws + g.sw, 0, W);
    we = clampi(we + g.sw, 0, W);
    const bool empty = (he <= hs) || (we <= ws) || !image_ok;
    float best = empty ? 0.f : -FLT_MAX;
    int besti = -1;
    if (!empty) {
      const int plane = (g.batch * C + c0 + c) * H * W;
      const float* p = features + plane;
      for (int h = hs; h < he; ++h)
        for (int w = ws; w < we; ++w) {
          const float v = __ldg(p + h * W + w);
          if (v > best) {
            best = v;
            besti = plane + h * W + w;
          }
        }
    }
    output[roi_base + o] = best;
    if (argmax) argmax[roi_base + o] = besti;
  }
}

// ---------------------------------------------------------------------------
// Plane-resident forward, 7 x 7 bins.  The generic kernel above is bound by instruction issue
// (~280 thread instructions per output: per-output index arithmetic, window loops whose trip
// counts differ from lane to lane).  Here a CTA holds RPL_CH whole feature planes of one image in
// shared memory (one bulk copy: the planes are contiguous in NCHW) and its warps walk the
// image's RoIs.  Lane = (channel, one of four bin rows): all lanes of a warp work on the same
// RoI, so the column windows [ws, we) of the seven bin columns are warp-uniform -- the window
// loops have uniform trip counts, the bounds are computed once per RoI, and the strict-`>`
// row-major scan order of the reference (roi_pooling_kernel.cu:75-88) is kept per bin.  Values
// and argmax of a RoI's RPL_CH channels are contiguous in the output: they are staged in shared
// memory and leave as two bulk stores.
// ---------------------------------------------------------------------------
constexpr int RPL_CH = 8;                       // planes per CTA
constexpr int RPL_TILE = RPL_CH * 49 * 4;       // bytes of one RoI's values (or argmax) for the slab
constexpr int RPL_SLOT = 2 * RPL_TILE;          // per warp: values, argmax
constexpr int RPL_MAX_WARPS = 20;

__device__ __forceinline__ float rpl_max3(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));  // FMNMX3; like fmaxf, NaN operands are ignored
  return r;
}
__device__ __forceinline__ void rpl_bulk_store(void* gdst, unsigned smem_src, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_src), "r"(bytes)
               : "memory");
}
__host__ __device__ inline unsigned rpl_plane_bytes(int HW) { return ((unsigned)(RPL_CH * HW * 4) + 127u) & ~127u; }

// Column offsets of the SLOTS cells scanned per bin column, starting at ws[pw] + k0.  Slots past the
// bin's last column re-read that column (a repeated value never changes a first-maximum search).
// FAST: every bin has at least SLOTS - 1 columns, so only the last slot needs the clamp and the others
// are ws[pw] + k -- one address per bin column and row, immediates for the rest.
template <int SLOTS, bool FAST>
__device__ __forceinline__ void rpl_offsets(int (&off)[7][SLOTS], const int (&ws)[7], const int (&we)[7], int k0) {
#pragma unroll
  for (int pw = 0; pw < 7; ++pw)
#pragma unroll
    for (int k = 0; k < SLOTS; ++k)
      off[pw][k] = (FAST && k < SLOTS - 1) ? ws[pw] + k : max(min(ws[pw] + k0 + k, we[pw] - 1), 0);
}
// One row of the lane's bin row: the running maximum of each bin column absorbs SLOTS cells with
// FMNMX3 / FMNMX -- one or two ALU operations instead of a compare + two selects per cell (the ALU
// pipe, which issues every other cycle, bounds this kernel).  Where the maximum grew, `code` (the
// row, and the first slot for wide bins) is recorded: under the strict `>` of the reference the first
// row (and chunk) that reaches the final maximum holds the argmax; its column is found afterwards.
template <int SLOTS>
__device__ __forceinline__ void rpl_scan_row(const float* __restrict__ row, const int (&off)[7][SLOTS], int code,
                                             float (&best)[7], int (&bcode)[7]) {
  float v[7][SLOTS];
#pragma unroll
  for (int pw = 0; pw < 7; ++pw)
#pragma unroll
    for (int k = 0; k < SLOTS; ++k) v[pw][k] = row[off[pw][k]];
#pragma unroll
  for (int pw = 0; pw < 7; ++pw) {
    float nb = best[pw];
#pragma unroll
    for (int k = 0; k + 1 < SLOTS; k += 2) nb = rpl_max3(nb, v[pw][k], v[pw][k + 1]);
    if (SLOTS & 1) nb = fmaxf(nb, v[pw][SLOTS - 1]);
    if (nb > best[pw]) bcode[pw] = code;
    best[pw] = nb;
  }
}
// First of the row's SLOTS cells that holds the maximum: its column offset (a repeated last column
// comes after the original, so the match is a real cell of the bin).
template <int SLOTS>
__device__ __forceinline__ int rpl_find_col(const float* __restrict__ row, const int (&off)[SLOTS], float best) {
  float v[SLOTS];
#pragma unroll
  for (int k = 0; k < SLOTS; ++k) v[k] = row[off[k]];
  int col = off[SLOTS - 1];
#pragma unroll
  for (int k = SLOTS - 2; k >= 0; --k)
    if (v[k] == best) col = off[k];
  return col;
}

struct RplLane {
  const float* plane;      // this lane's channel plane in shared memory
  int gplane;              // (image * C + channel) * H * W: argmax indexes the whole feature tensor
  int sub, H, W;           // which of the four bin rows of a pass this lane scans
  unsigned out_v, out_a;   // shared-window address of this lane's channel in the warp's staging slot
};

// Both passes (bin rows 0-3, 4-6) of one RoI.  SLOTS > 0: every bin is at most SLOTS columns wide and
// at least SLOTS - 1 (RoIs inside the map): offsets once per RoI.  SLOTS == 0: wide or clipped bins,
// four clamped columns at a time in row-major order.
template <int SLOTS>
__device__ __forceinline__ void rpl_pool_roi(const RplLane& L, const PoolRoi& g, bool image_ok, const int (&ws)[7],
                                             const int (&we)[7], int maxw, bool wait_store) {
  constexpr int S = SLOTS > 0 ? SLOTS : 4;
  int off[7][S];
  if (SLOTS > 0) rpl_offsets<S, true>(off, ws, we, 0);
#pragma unroll 1
  for (int it = 0; it < 2; ++it) {
    const int ph = it * 4 + L.sub;
    const int hs = clampi((int)floorf(__fmul_rn((float)ph, g.bin_h)) + g.sh, 0, L.H);
    const int he = clampi((int)ceilf(__fmul_rn((float)(ph + 1), g.bin_h)) + g.sh, 0, L.H);
    const int nrows = (ph < 7 && image_ok && he > hs) ? he - hs : 0;
    const int maxrows = __reduce_max_sync(0xffffffffu, nrows);
    const float* __restrict__ first = L.plane + (nrows > 0 ? hs : 0) * L.W;
    float best[7];
    int bcode[7];
#pragma unroll
    for (int pw = 0; pw < 7; ++pw) { best[pw] = -FLT_MAX; bcode[pw] = -1; }
    // rows past this lane's last one re-read it (warp-uniform trip count), harmless like repeated columns
#pragma unroll 1
    for (int r = 0; r < maxrows; ++r) {
      const int rr = max(min(r, nrows - 1), 0);
      if (SLOTS > 0) {
        rpl_scan_row<S>(first + rr * L.W, off, rr, best, bcode);
      } else {
#pragma unroll 1
        for (int k0 = 0; k0 < maxw; k0 += 4) {
          rpl_offsets<S, false>(off, ws, we, k0);
          rpl_scan_row<S>(first + rr * L.W, off, (rr << 8) | k0, best, bcode);
        }
      }
    }
    if (it == 0 && wait_store) {  // the staging slot is about to be rewritten
      if (lane_id() == 0) bulk_wait_read_all();
      __syncwarp();
    }
    if (ph < 7) {
#pragma unroll
      for (int pw = 0; pw < 7; ++pw) {
        const bool empty = nrows == 0 || we[pw] <= ws[pw];
        const int code = max(bcode[pw], 0);
        const int rr = SLOTS > 0 ? code : code >> 8;
        if (SLOTS == 0) {
#pragma unroll
          for (int k = 0; k < S; ++k) off[pw][k] = max(min(ws[pw] + (code & 255) + k, we[pw] - 1), 0);
        }
        const float* __restrict__ row = first + rr * L.W;
        const int col = rpl_find_col<S>(row, off[pw], best[pw]);
        // the stored value is re-read (the running maximum may carry +0 where the cell holds -0);
        // nothing above -FLT_MAX: value -FLT_MAX, argmax -1 like the reference
        const float val = empty ? 0.f : (bcode[pw] < 0 ? best[pw] : row[col]);
        const int idx = (empty || bcode[pw] < 0) ? -1 : L.gplane + (int)(row - L.plane) + col;
        const unsigned o = (unsigned)(ph * 7 + pw) * 4u;
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(L.out_v + o), "f"(val) : "memory");
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(L.out_a + o), "r"(idx) : "memory");
      }
    }
  }
}

__global__ void __launch_bounds__(32 * RPL_MAX_WARPS)
    roi_pool_fwd_planes_kernel(const float* __restrict__ features, const float* __restrict__ rois,
                               float* __restrict__ output, int* __restrict__ argmax, int B, int C, int H, int W,
                               int R, float scale, int nchunks) {
  extern __shared__ __align__(128) unsigned char rpl_smem[];
  const int HW = H * W;
  const int nwarps = blockDim.x >> 5, wid = warp_id(), lane = lane_id();
  const int slabs = C / RPL_CH;
  const int chunk = blockIdx.x % nchunks;
  const int slab = (blockIdx.x / nchunks) % slabs;
  const int b = blockIdx.x / (nchunks * slabs);
  const int c0 = slab * RPL_CH;
  const unsigned planes = smem_u32(rpl_smem);
  const unsigned slots = planes + rpl_plane_bytes(HW);
  const unsigned bar = slots + (unsigned)nwarps * RPL_SLOT;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const unsigned bytes = (unsigned)(RPL_CH * HW * 4);
    mbar_arrive_expect_tx(bar, bytes);
    bulk_load(planes, features + ((size_t)b * C + c0) * HW, bytes, bar);
  }
  const int stride = nchunks * nwarps;
  int n = chunk + nchunks * wid;
  // the RoI list is read one RoI ahead (every lane the same five floats)
  float raw[5] = {-1.f, 0.f, 0.f, 0.f, 0.f};
  if (n < R) {
#pragma unroll
    for (int k = 0; k < 5; ++k) raw[k] = __ldg(rois + (size_t)n * 5 + k);
  }
  __syncthreads();
  mbar_wait(bar, 0);

  RplLane L;
  const int ch = lane & 7;
  L.sub = lane >> 3;
  L.H = H;
  L.W = W;
  L.plane = reinterpret_cast<const float*>(rpl_smem) + ch * HW;
  L.gplane = (b * C + c0 + ch) * HW;
  const unsigned slot_v = slots + (unsigned)wid * RPL_SLOT, slot_a = slot_v + RPL_TILE;
  L.out_v = slot_v + (unsigned)(ch * 49) * 4u;
  L.out_a = slot_a + (unsigned)(ch * 49) * 4u;
  bool stored = false;  // a bulk store of this warp's slot may still be reading it

  for (; n < R; n += stride) {
    float cur[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) cur[k] = raw[k];
    if (n + stride < R) {
#pragma unroll
      for (int k = 0; k < 5; ++k) raw[k] = __ldg(rois + (size_t)(n + stride) * 5 + k);
    }
    const int batch = (int)cur[0];
    const bool image_ok = batch >= 0 && batch < B;
    if ((image_ok ? batch : 0) != b) continue;  // RoIs without an image: all bins empty, image 0 writes them
    PoolRoi g;  // pool_roi(), from the prefetched floats
    g.batch = batch;
    g.sw = (int)roundf(__fmul_rn(cur[1], scale));
    g.sh = (int)roundf(__fmul_rn(cur[2], scale));
    g.ew = (int)roundf(__fmul_rn(cur[3], scale));
    g.eh = (int)roundf(__fmul_rn(cur[4], scale));
    g.bin_h = __fdiv_rn((float)max(g.eh - g.sh + 1, 1), 7.f);
    g.bin_w = __fdiv_rn((float)max(g.ew - g.sw + 1, 1), 7.f);
    int ws[7], we[7], maxw = 0, minw = 1 << 30;
#pragma unroll
    for (int pw = 0; pw < 7; ++pw) {
      ws[pw] = clampi((int)floorf(__fmul_rn((float)pw, g.bin_w)) + g.sw, 0, W);
      we[pw] = clampi((int)ceilf(__fmul_rn((float)(pw + 1), g.bin_w)) + g.sw, 0, W);
      maxw = max(maxw, we[pw] - ws[pw]);
      minw = min(minw, we[pw] - ws[pw]);
    }
    if (minw >= maxw - 1 && maxw == 2) rpl_pool_roi<2>(L, g, image_ok, ws, we, maxw, stored);
    else if (minw >= maxw - 1 && maxw == 3) rpl_pool_roi<3>(L, g, image_ok, ws, we, maxw, stored);
    else if (minw >= maxw - 1 && maxw == 4) rpl_pool_roi<4>(L, g, image_ok, ws, we, maxw, stored);
    else rpl_pool_roi<0>(L, g, image_ok, ws, we, maxw, stored);
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
      const size_t off = ((size_t)n * C + c0) * 49;
      rpl_bulk_store(output + off, slot_v, RPL_TILE);
      if (argmax) rpl_bulk_store(argmax + off, slot_a, RPL_TILE);
      bulk_commit_group();
    }
    stored = true;
  }
  if (lane == 0 && stored) bulk_wait_read_all();
}

struct RplLaunch {
  int warps, nchunks, grid;
  size_t smem;
};
// false: the shape is not served by the plane-resident kernel
static bool rpl_plan(int batch, int channels, int height, int width, int num_rois, RplLaunch* L) {
  if (channels % RPL_CH != 0 || getenv("TLOD_DISABLE_POOL_PLANES")) return false;
  const size_t smem_max = (size_t)device_info().max_smem_optin;
  const size_t planes = rpl_plane_bytes(height * width);
  if ((size_t)height * width * RPL_CH * 4 >= (1u << 20)) return false;  // one mbarrier transaction
  // one CTA per SM with as many warps as the staging slots allow (measured: the kernel is bound by
  // instruction issue, 20 warps of one CTA beat two CTAs of 7 that overlap staging and compute)
  int warps = planes + 16 < smem_max ? (int)((smem_max - 16 - planes) / RPL_SLOT) : 0;
  if (const char* e = getenv("TLOD_POOL_WARPS")) {
    const int w = atoi(e);
    if (w > 0 && planes + (size_t)w * RPL_SLOT + 16 <= smem_max) warps = w;
  }
  if (warps > RPL_MAX_WARPS) warps = RPL_MAX_WARPS;
  if (warps < 6) return false;
  const long long base = (long long)batch * (channels / RPL_CH);
  // RoI chunks only where (image, slab) pairs alone leave SMs idle (every CTA stages its planes
  // again), and never fewer than two RoIs per warp
  const int sms = device_info().sm_count;
  int nchunks = (int)((sms + base / 2) / base);
  const long long per_image = (num_rois + batch - 1) / batch;
  const long long most = per_image / (2LL * warps);
  if (nchunks > most) nchunks = (int)most;
  if (nchunks < 1) nchunks = 1;
  if (const char* e = getenv("TLOD_POOL_CHUNKS")) nchunks = atoi(e) > 0 ? atoi(e) : nchunks;
  if (base * nchunks > 2147483647LL) return false;
  L->warps = warps;
  L->nchunks = nchunks;
  L->grid = (int)(base * nchunks);
  L->smem = planes + (size_t)warps * RPL_SLOT + 16;
  return true;
}

template <int PHT, int PWT>
__global__ void __launch_bounds__(256)
    roi_pool_bwd_kernel(const float* __restrict__ top_grad, const int* __restrict__ argmax,
                        const float* __restrict__ rois, float* __restrict__ bottom_grad, int B, int C,
                        int H, int W, int PH_rt, int PW_rt, float scale, int chans_per_block) {
  const int PH = PHT > 0 ? PHT : PH_rt, PW = PWT > 0 ? PWT : PW_rt;
  const int n = blockIdx.x;
  const int c0 = blockIdx.y * chans_per_block;
  const PoolRoi g = pool_roi(rois + (size_t)n * 5, scale, PH, PW);
  if (g.batch < 0 || g.batch >= B) return;
  const int S = PH * PW;
  const int cb = min(chans_per_block, C - c0);
  const size_t roi_base = ((size_t)n * C + c0) * S;
  const int HW = H * W;
  const int total = B * C * HW;
  for (int o = threadIdx.x; o < cb * S; o += blockDim.x) {
    const int idx = __ldg(argmax + roi_base + o);
    if (idx < 0 || idx >= total) continue;
    const int c = c0 + o / S;
    const int i = o % S;
    const int ph = i / PW, pw = i - ph * PW;
    // decode the cell: the reference's thread for cell `idx` only looks at RoIs of its own image
    // (:150) and at argmax entries of its own channel (:189) -- i.e. idx must lie in the plane of
    // (this RoI's image, this channel); then one division gives the cell
    const int rem = idx - (g.batch * C + c) * HW;
    if (rem < 0 || rem >= HW) continue;
    const int h = rem / W;
    const int w = rem - h * W;
    if (!(w >= g.sw && w <= g.ew && h >= g.sh && h <= g.eh)) continue;  // :157-159
    int phs = (int)floorf(__fdiv_rn((float)(h - g.sh), g.bin_h));       // :178-186
    int phe = (int)ceilf(__fdiv_rn((float)(h - g.sh + 1), g.bin_h));
    int pws = (int)floorf(__fdiv_rn((float)(w - g.sw), g.bin_w));
    int pwe = (int)ceilf(__fdiv_rn((float)(w - g.sw + 1), g.bin_w));
    phs = clampi(phs, 0, PH);
    phe = clampi(phe, 0, PH);
    pws = clampi(pws, 0, PW);
    pwe = clampi(pwe, 0, PW);
    if (ph < phs || ph >= phe || pw < pws || pw >= pwe) continue;
    atomicAdd(bottom_grad + idx, __ldg(top_grad + roi_base + o));
  }
}

// Backward with the admission tests of roi_pooling_kernel.cu:157-186 tabulated per CTA: the bins
// that may own map row h (column w) depend on the RoI only, so they are computed once per CTA
// (the four exact divisions per row / column) instead of once per output (four per output).
// tab[h] = phstart | phend << 8 (an empty range for rows outside the RoI), tab[H + w] likewise.
constexpr int RPB_TAB = 1024;  // H + W served by the table kernel
constexpr int RPB_ITER = 8;    // outputs per thread: pool_grid() gives a CTA at most 2048
template <int PHT, int PWT>
__global__ void __launch_bounds__(256)
    roi_pool_bwd_tab_kernel(const float* __restrict__ top_grad, const int* __restrict__ argmax,
                            const float* __restrict__ rois, float* __restrict__ bottom_grad, int B, int C,
                            int H, int W, int PH_rt, int PW_rt, float scale, int chans_per_block, float inv_w) {
  const int PH = PHT > 0 ? PHT : PH_rt, PW = PWT > 0 ? PWT : PW_rt;
  __shared__ unsigned short tab[RPB_TAB];
  const int n = blockIdx.x;
  const int c0 = blockIdx.y * chans_per_block;
  const PoolRoi g = pool_roi(rois + (size_t)n * 5, scale, PH, PW);
  if (g.batch < 0 || g.batch >= B) return;
  const int S = PH * PW;
  const int cb = min(chans_per_block, C - c0);
  const size_t roi_base = ((size_t)n * C + c0) * S;
  // all of this thread's argmax / gradient pairs are requested before the table is built: one
  // memory latency per thread instead of two per output (pool_grid: at most RPB_ITER passes)
  int idx[RPB_ITER];
  float tg[RPB_ITER];
#pragma unroll
  for (int j = 0; j < RPB_ITER; ++j) {
    const int o = threadIdx.x + j * 256;
    idx[j] = -1;
    tg[j] = 0.f;
    if (o < cb * S) {
      idx[j] = __ldg(argmax + roi_base + o);
      tg[j] = __ldg(top_grad + roi_base + o);
    }
  }
  for (int t = threadIdx.x; t < H + W; t += blockDim.x) {
    const bool row = t < H;
    const int p = row ? t : t - H;
    const int lo = row ? g.sh : g.sw, hi = row ? g.eh : g.ew, P = row ? PH : PW;
    const float bin = row ? g.bin_h : g.bin_w;
    int s = (int)floorf(__fdiv_rn((float)(p - lo), bin));
    int e = (int)ceilf(__fdiv_rn((float)(p - lo + 1), bin));
    s = clampi(s, 0, P);
    e = clampi(e, 0, P);
    tab[t] = (p >= lo && p <= hi) ? (unsigned short)(s | (e << 8)) : (unsigned short)1;
  }
  __syncthreads();
  const int HW = H * W;
  const int plane0 = (g.batch * C + c0) * HW;
#pragma unroll
  for (int j = 0; j < RPB_ITER; ++j) {
    const int o = threadIdx.x + j * 256;
    const int c = o / S;
    const int i = o - c * S;
    const int ph = i / PW, pw = i - ph * PW;
    // the reference's thread for cell `idx` looks only at RoIs of its own image (:150) and at
    // argmax entries of its own channel (:189): idx must lie in the plane of (image, channel)
    const int rem = idx[j] - plane0 - c * HW;
    if (idx[j] < 0 || rem < 0 || rem >= HW) continue;
    int h = __float2int_rz(__fmul_rn((float)rem, inv_w));  // rem < 2^24: off by at most one
    int w = rem - h * W;
    if (w < 0) { --h; w += W; }
    if (w >= W) { ++h; w -= W; }
    const unsigned th = tab[h], tw = tab[H + w];
    if (ph < (int)(th & 255u) || ph >= (int)(th >> 8) || pw < (int)(tw & 255u) || pw >= (int)(tw >> 8)) continue;
    atomicAdd(bottom_grad + idx[j], tg[j]);
  }
}

static int pool_check(const void* a, const void* b, const void* c, int batch, int channels,
                      int height, int width, int num_rois, int ph, int pw) {
  if (!a || !b || !c) return TLOD_ERR_NULL_POINTER;
  if (batch <= 0 || channels <= 0 || height <= 0 || width <= 0 || num_rois < 0 || ph <= 0 || pw <= 0)
    return TLOD_ERR_BAD_SHAPE;
  if ((long long)batch * channels * height * width >= (1LL << 31)) return TLOD_ERR_INT32_OVERFLOW;
  return TLOD_OK;
}

static dim3 pool_grid(int num_rois, int channels, int S, int* cpb_out) {
  int cpb = S <= 2048 ? 2048 / S : 1;  // <= 8 full passes of 256 threads
  if (cpb > channels) cpb = channels;
  while ((channels + cpb - 1) / cpb > 65535) ++cpb;
  *cpb_out = cpb;
  return dim3(num_rois, (channels + cpb - 1) / cpb);
}

}  // namespace tlod

using namespace tlod;

extern "C" int tlod_roi_pool_forward(const float* features, const float* rois, float* output,
                                     int* argmax, int batch, int channels, int height, int width,
                                     int num_rois, int pooled_h, int pooled_w, float spatial_scale,
                                     void* stream) {
  return tlod::roi_pool_forward_launch(features, rois, output, argmax, batch, channels, height, width, num_rois,
                                       pooled_h, pooled_w, spatial_scale, true, (cudaStream_t)stream);
}

// batch_known = false: `batch` is only an upper bound for rois[:, 0] (the reference's forward
// launcher is not told the batch size), so nothing may be read per image: generic kernel only.
int tlod::roi_pool_forward_launch(const float* features, const float* rois, float* output, int* argmax, int batch,
                                  int channels, int height, int width, int num_rois, int pooled_h, int pooled_w,
                                  float spatial_scale, bool batch_known, cudaStream_t stream) {
  int rc = pool_check(features, rois, output, batch, channels, height, width, num_rois, pooled_h,
                      pooled_w);
  if (rc != TLOD_OK) return rc;
  if (num_rois == 0) return TLOD_OK;
  RplLaunch L;
  if (batch_known && pooled_h == 7 && pooled_w == 7 && ((uintptr_t)features & 15) == 0 && ((uintptr_t)output & 15) == 0 &&
      ((uintptr_t)argmax & 15) == 0 && rpl_plan(batch, channels, height, width, num_rois, &L)) {
    cudaError_t e = cudaFuncSetAttribute(roi_pool_fwd_planes_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)L.smem);
    if (e != cudaSuccess) return (int)e;
    {
      LaunchScope scope("roi_pool_fwd_planes_kernel", (cudaStream_t)stream);
      roi_pool_fwd_planes_kernel<<<L.grid, 32 * L.warps, L.smem, (cudaStream_t)stream>>>(
          features, rois, output, argmax, batch, channels, height, width, num_rois, spatial_scale, L.nchunks);
    }
    return last_launch_status();
  }
  int cpb;
  dim3 grid = pool_grid(num_rois, channels, pooled_h * pooled_w, &cpb);
  {
    LaunchScope scope("roi_pool_fwd_kernel", (cudaStream_t)stream);
    auto kern = (pooled_h == 7 && pooled_w == 7) ? roi_pool_fwd_kernel<7, 7> : roi_pool_fwd_kernel<0, 0>;
    kern<<<grid, 256, 0, (cudaStream_t)stream>>>(features, rois, output, argmax, batch, channels, height, width,
                                                 pooled_h, pooled_w, spatial_scale, cpb);
  }
  return last_launch_status();
}

extern "C" int tlod_roi_pool_backward(const float* top_grad, const int* argmax, const float* rois,
                                      float* bottom_grad, int batch, int channels, int height,
                                      int width, int num_rois, int pooled_h, int pooled_w,
                                      float spatial_scale, void* stream) {
  int rc = pool_check(top_grad, rois, bottom_grad, batch, channels, height, width, num_rois,
                      pooled_h, pooled_w);
  if (rc != TLOD_OK) return rc;
  if (!argmax) return TLOD_ERR_NULL_POINTER;
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(bottom_grad, 0,
                                  (size_t)batch * channels * height * width * sizeof(float), st);
  if (e != cudaSuccess) return (int)e;
  if (num_rois == 0) return TLOD_OK;
  int cpb;
  dim3 grid = pool_grid(num_rois, channels, pooled_h * pooled_w, &cpb);
  if (height + width <= RPB_TAB && pooled_h <= 255 && pooled_w <= 255 && height * width < (1 << 24) &&
      cpb * pooled_h * pooled_w <= 256 * RPB_ITER &&
      !getenv("TLOD_DISABLE_POOL_TAB")) {
    LaunchScope scope("roi_pool_bwd_tab_kernel", st);
    auto kern = (pooled_h == 7 && pooled_w == 7) ? roi_pool_bwd_tab_kernel<7, 7> : roi_pool_bwd_tab_kernel<0, 0>;
    kern<<<grid, 256, 0, st>>>(top_grad, argmax, rois, bottom_grad, batch, channels, height, width, pooled_h,
                               pooled_w, spatial_scale, cpb, 1.0f / (float)width);
    return last_launch_status();
  }
  {
    LaunchScope scope("roi_pool_bwd_kernel", st);
    auto kern = (pooled_h == 7 && pooled_w == 7) ? roi_pool_bwd_kernel<7, 7> : roi_pool_bwd_kernel<0, 0>;
    kern<<<grid, 256, 0, st>>>(top_grad, argmax, rois, bottom_grad, batch, channels, height, width, pooled_h,
                               pooled_w, spatial_scale, cpb);
  }
  return last_launch_status();
}
