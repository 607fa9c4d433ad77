// RoIAlign forward for 8 x 8 samples per RoI on sm_100a: RoIAlign(8, 8), and RoIAlignAvg(7, 7)
// with its 2x2 / stride-1 average fused in (the (R, C, 8, 8) intermediate never exists).
//
// Reference semantics: lib/model/roi_align/src/roi_align_kernel.cu:15-70 (one bilinear sample per
// output cell) and lib/model/roi_align/modules/roi_align.py:26-29 (`avg_pool2d(x, 2, 1)`).
//
// Same residency as roi_align.cu's plane kernel -- a CTA keeps the 16 feature planes of one
// (image, 16-channel slab) in shared memory and streams the image's RoIs through them, a warp per
// RoI -- but the work of a warp is cut differently, because the kernel is bound by shared-memory
// wavefronts (16 bytes of gather per sample), not by HBM:
//
//  * lane = (channel, half): half 0 owns samples pw 0..3 of every sample row, half 1 owns
//    pw 4..7.  All lanes walk the sample rows ph = 0..7 together, so the control flow over rows
//    is warp-uniform and the horizontally interpolated plane rows can be REUSED: sample row ph
//    needs plane rows y and y + 1; if the previous sample row used the same pair nothing is
//    loaded, if it used (y - 1, y) only row y + 1 is.  A RoI of 8 sample rows touches D <= 16
//    distinct plane rows (D = roi height in cells + 2 for small RoIs): 8 D shared-memory loads per
//    lane instead of 128.
//  * bank conflicts: the planes are stored unpadded with a plane stride of 2 (mod 4) floats, so
//    the 16 channels sit on 16 banks of one parity.  Both halves read the same plane row at the
//    same time; half 0 reads its cell pair as (x, x + 1), half 1 reads (x', x' + 1) in the order
//    that puts it on the other parity.  Every gather is conflict free for any RoI geometry.
//  * staging: with an unpadded layout the 16 planes of a slab are one contiguous, 64-byte
//    aligned block in HBM (16 * H * W * 4 bytes): when H * W = 2 (mod 4) (the 38 x 75 conv4 maps)
//    it arrives with a few cp.async.bulk copies on an mbarrier, no register round trip.
//  * output: RoIAlign(8, 8) rows leave through a per-warp 2 KB tile and one TMA tensor store
//    per half RoI slab (as in roi_align.cu).  RoIAlignAvg: the 2x2 sums are formed in registers
//    (one shuffle per sample row carries sample 4 to the lane that owns column 3), the 16 x 49
//    floats of a (RoI, slab) are contiguous in the (R, C, 7, 7) output, so they leave with ONE
//    3136-byte cp.async.bulk store.
#include <cstdlib>
#include <cstring>

#include "async_copy.cuh"
#include "roi_align_plan.cuh"

namespace tlod {

constexpr int F8_CH = 16;            // channels per slab
constexpr int F8_MAX_WARPS = 16;
constexpr int F8_TILE8 = 2048;       // RoIAlign(8, 8): 16 channels x 4 rows x 8 floats (one half RoI)
constexpr int F8_TILE7 = F8_CH * 49 * 4;  // RoIAlignAvg: 16 channels x 7 x 7 floats
constexpr int F8_WTAB = 256;         // per-warp tables: 8 row + 8 column entries

struct F8Layout {
  int warps;
  int Pp;  // plane stride in floats
  size_t tiles, wtab, ctl, total;
};

__host__ __device__ inline int f8_plane_stride(int P) {
  while ((P & 3) != 2) ++P;
  return P;
}

static F8Layout f8_layout(int H, int W, bool fused, size_t smem_max) {
  F8Layout L;
  L.Pp = f8_plane_stride(H * W);
  const size_t planes = (size_t)F8_CH * L.Pp * sizeof(float);
  const size_t tile = fused ? F8_TILE7 : F8_TILE8;
  L.tiles = fused ? (planes + 127) / 128 * 128 : (planes + 1023) / 1024 * 1024;
  L.warps = 0;
  for (int w = F8_MAX_WARPS; w >= 8; --w) {
    const size_t total = L.tiles + (size_t)w * (tile + F8_WTAB) + 64;
    if (total <= smem_max) {
      L.warps = w;
      break;
    }
  }
  L.wtab = L.tiles + (size_t)L.warps * tile;
  L.ctl = L.wtab + (size_t)L.warps * F8_WTAB;
  L.total = L.ctl + 64;
  return L;
}

__device__ __forceinline__ float f8_lds(unsigned addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void f8_sts(unsigned addr, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ void f8_sts_v4(unsigned addr, float a, float b, float c, float d) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
// the pooled tensor is written once and never read here: keep it from displacing the planes in L2
__device__ __forceinline__ unsigned long long f8_evict_first_policy() {
  unsigned long long p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void f8_bulk_store(void* gdst, unsigned smem_src, unsigned bytes, unsigned long long pol) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(gdst),
               "r"(smem_src), "r"(bytes), "l"(pol)
               : "memory");
}
__device__ __forceinline__ void f8_tma_store_4d(const void* tmap, unsigned smem_src, int c0, int c1, int c2, int c3,
                                                unsigned long long pol) {
  asm volatile(
      "cp.async.bulk.tensor.4d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%2, %3, %4, %5}], [%1], %6;" ::"l"(
          (unsigned long long)tmap),
      "r"(smem_src), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "l"(pol)
      : "memory");
}

// One horizontally interpolated plane row: this lane's four samples at byte offset `off` of its plane.
__device__ __forceinline__ void f8_hrow(float (&T)[4], const unsigned (&ca)[4], const unsigned (&cb)[4],
                                        const float (&wp)[4], const float (&wq)[4], unsigned off) {
  float a[4], b[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    a[k] = f8_lds(ca[k] + off);
    b[k] = f8_lds(cb[k] + off);
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) T[k] = fmaf(b[k], wq[k], a[k] * wp[k]);
}

// FUSED: write avg_pool2d(samples, 2, 1) as (R, C, 7, 7); else the samples as (R, C, 8, 8).
// WC > 0: the map width as a compile-time constant (75: the 600 x 1200 / stride-16 maps).
template <bool FUSED, int WC>
__global__ void __launch_bounds__(32 * F8_MAX_WARPS, 1)
    roi_align_fwd8_kernel(const __grid_constant__ CUtensorMap omap, const float* __restrict__ features,
                          float* __restrict__ output, PlanPtrs pl, int B, int C, int H, int W_rt, int R, int Pp,
                          int bulk_stage, unsigned tiles_off, unsigned wtab_off, unsigned ctl_off) {
  extern __shared__ __align__(1024) unsigned char smem_f8[];
  constexpr int TILE = FUSED ? F8_TILE7 : F8_TILE8;
  const int W = WC > 0 ? WC : W_rt;
  const int P = H * W;
  const int tid = threadIdx.x, lane = lane_id(), wid = warp_id();
  const int nwarps = blockDim.x >> 5;
  float* planes = reinterpret_cast<float*>(smem_f8);
  float4* wtab = reinterpret_cast<float4*>(smem_f8 + wtab_off) + wid * 16;
  int* cur_base = reinterpret_cast<int*>(smem_f8 + ctl_off);  // 2 x {image, slab, rank_lo, rank_hi}
  const unsigned bar = smem_u32(smem_f8 + ctl_off + 32);
  const int nslabs = C / F8_CH;
  const int S_out = FUSED ? 49 : 64;

  // lane pair p = lane >> 1 -> channel: pair bits (p0, p1, p2, p3) -> channel bits (c1, c2, c0, c3), so
  // that the four pairs of a quarter warp differ in c1 and c2 and the two halves of a pair fill
  // bit 0 of the 16-byte group index of the swizzled 8x8 output tile
  const int pr = lane >> 1, half = lane & 1;
  const int c = ((pr & 3) << 1) | ((pr >> 2) & 1) | (pr & 8);
  const unsigned tile = smem_u32(smem_f8 + tiles_off + (size_t)wid * TILE);
  const unsigned tile_row = tile + (FUSED ? 196u : 128u) * (unsigned)c;
  const unsigned cs = (unsigned)(c & 7);
  const unsigned plane_addr = smem_u32(planes + (size_t)c * Pp);
  const unsigned row_bytes = 4u * (unsigned)W;
  const unsigned long long pol = f8_evict_first_policy();
  const int* __restrict__ cum = pl.cum;

  if (tid == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  unsigned stage_parity = 0;

  // this CTA's contiguous range of (image, slab, RoI rank) units
  const long long U = (long long)nslabs * R;
  long long u = U * blockIdx.x / gridDim.x;
  const long long u_end = U * (blockIdx.x + 1) / gridDim.x;
  int staged_img = -1, staged_slab = -1;

  for (int iter = 0; u < u_end; ++iter) {
    int* cur = cur_base + 4 * (iter & 1);  // double buffered: thread 0 may run one unit ahead of a reader
    if (tid == 0) {
      int lo = 0, hi = B;  // image b with nslabs * cum[b] <= u < nslabs * cum[b + 1]
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
      cur[0] = lo; cur[1] = slab; cur[2] = r_lo; cur[3] = r_hi;
    }
    __syncthreads();  // also: every warp is done with the planes staged before
    const int img = cur[0], slab = cur[1], r_lo = cur[2], r_hi = cur[3];
    const int c0 = slab * F8_CH;

    if (img < B && (img != staged_img || slab != staged_slab)) {
      const float* g = features + ((size_t)img * C + c0) * P;
      if (bulk_stage) {
        // the 16 unpadded planes are one contiguous block of 64 * P bytes
        if (tid == 0) {
          const unsigned total = 64u * (unsigned)P;
          mbar_arrive_expect_tx(bar, total);
          const unsigned dst = smem_u32(planes);
          for (unsigned done = 0; done < total; done += 32768u) {
            const unsigned nb = total - done < 32768u ? total - done : 32768u;
            bulk_load(dst + done, reinterpret_cast<const unsigned char*>(g) + done, nb, bar);
          }
        }
        mbar_wait(bar, stage_parity);
        stage_parity ^= 1u;
      } else {
        for (int ch = wid; ch < F8_CH; ch += nwarps) {
          const float* gp = g + (size_t)ch * P;
          float* sp = planes + (size_t)ch * Pp;
          for (int i = lane; i < P; i += 32 * 16) {
            float v[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) v[k] = (i + k * 32 < P) ? __ldg(gp + i + k * 32) : 0.f;
#pragma unroll
            for (int k = 0; k < 16; ++k)
              if (i + k * 32 < P) sp[i + k * 32] = v[k];
          }
        }
        __syncthreads();
      }
      staged_img = img;
      staged_slab = slab;
    }

    // pull the slab this CTA stages next into L2 while this one is being gathered from
    if (tid == 32 && bulk_stage) {
      const long long un = u + (r_hi - r_lo);
      if (un < u_end) {
        const int cnt = __ldg(cum + img + 1) - __ldg(cum + img);
        int ni = img, ns = slab + 1;  // the next unit is the next slab of this image, or the next image's first
        if (r_hi < cnt) ns = slab;    // (same slab: nothing to fetch)
        if (ns >= nslabs) {
          ns = 0;
          do { ++ni; } while (ni < B && __ldg(cum + ni + 1) == __ldg(cum + ni));
        }
        if (ni < B && (ni != img || ns != slab)) {
          const unsigned char* g = reinterpret_cast<const unsigned char*>(features + ((size_t)ni * C + ns * F8_CH) * P);
          const unsigned total = 64u * (unsigned)P;
          for (unsigned done = 0; done < total; done += 32768u) {
            const unsigned nb = total - done < 32768u ? total - done : 32768u;
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(g + done), "r"(nb) : "memory");
          }
        }
      }
    }
    const int base = __ldg(cum + img);
    int e = r_lo + wid;
    int n = (e < r_hi) ? __ldg(pl.list + base + e) : 0;
    float4 t = (e < r_hi && img < B) ? __ldg(pl.tabs + (size_t)n * 32 + lane) : make_float4(0, 0, 0, 0);
    for (; e < r_hi; e += nwarps) {
      const int e2 = e + nwarps;  // prefetch the next RoI of this warp
      const int n2 = (e2 < r_hi) ? __ldg(pl.list + base + e2) : 0;
      float4 t2 = make_float4(0, 0, 0, 0);
      if (e2 < r_hi && img < B) t2 = __ldg(pl.tabs + (size_t)n2 * 32 + lane);

      float* out_roi = output + ((size_t)n * C + c0) * S_out;
      if (img == B) {  // invalid image index: zeros
        for (int i = lane; i < F8_CH * S_out; i += 32) out_roi[i] = 0.f;
      } else {
        // rows 0..7: {byte offset of the first sampled row, w0, w1, first row | -1}; columns 8..15
        if (lane < 8) {
          const int start = __float_as_int(t.w);
          t.x = __int_as_float(start < 0 ? 0 : start * (int)row_bytes);
          if (FUSED) { t.y *= 0.25f; t.z *= 0.25f; }  // exact: the 2x2 average's 1/4 rides on the row weights
          wtab[lane] = t;
        } else if (lane >= 16 && lane < 24) {
          wtab[lane - 8] = t;
        }
        __syncwarp();

        // all table reads of the RoI are issued up front: the row loop below then never waits on one
        float4 rt[8];
#pragma unroll
        for (int ph = 0; ph < 8; ++ph) rt[ph] = wtab[ph];
        unsigned ca[4], cb[4];
        float wp[4], wq[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float4 m0 = wtab[8 + k], m1 = wtab[12 + k];
          int x0 = __float_as_int(m0.x), x1 = __float_as_int(m1.x);
          x0 = x0 < 0 ? 0 : x0;
          x1 = x1 < 0 ? 0 : x1;
          const int flip = half & (((x0 ^ x1) & 1) ^ 1);  // half 1 goes to the parity half 0 does not use
          const int x = half ? x1 : x0;
          const float w0 = half ? m1.y : m0.y, w1 = half ? m1.z : m0.z;
          ca[k] = plane_addr + 4u * (unsigned)(x + flip);
          cb[k] = plane_addr + 4u * (unsigned)(x + (flip ^ 1));
          wp[k] = flip ? w1 : w0;
          wq[k] = flip ? w0 : w1;
        }

        float Tlo[4] = {0.f, 0.f, 0.f, 0.f}, Thi[4] = {0.f, 0.f, 0.f, 0.f}, hp[4] = {0.f, 0.f, 0.f, 0.f};
        int prev = -100;
#pragma unroll
        for (int ph = 0; ph < 8; ++ph) {
          const float4 r = rt[ph];
          const int st = __float_as_int(r.w);
          float o[4] = {0.f, 0.f, 0.f, 0.f};
          if (st >= 0) {  // warp-uniform
            const unsigned ro = (unsigned)__float_as_int(r.x);
            if (st != prev) {
              if (st == prev + 1) {
#pragma unroll
                for (int k = 0; k < 4; ++k) Tlo[k] = Thi[k];
              } else {
                f8_hrow(Tlo, ca, cb, wp, wq, ro);
              }
              f8_hrow(Thi, ca, cb, wp, wq, ro + row_bytes);
              prev = st;
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) o[k] = fmaf(Thi[k], r.z, Tlo[k] * r.y);
          }
          if (!FUSED) {
            if ((ph & 3) == 0) {  // the previous half's store must have read the tile
              bulk_wait_read_all();
              __syncwarp();
            }
            const unsigned k16 = (unsigned)((ph & 3) * 2 + half);
            f8_sts_v4(tile_row + ((k16 ^ cs) << 4), o[0], o[1], o[2], o[3]);
            if ((ph & 3) == 3) {
              fence_proxy_async_smem();
              __syncwarp();
              if (elect_one()) {
                f8_tma_store_4d(&omap, tile, 0, ph >> 2, c0, n, pol);
                bulk_commit_group();
              }
            }
          } else {
            // 2x2 sums: half 0 owns pooled columns 0..3 (column 3 needs sample 4), half 1 owns 4..6
            const float nb = __shfl_xor_sync(0xffffffffu, o[0], 1);
            float h[4];
            h[0] = o[0] + o[1]; h[1] = o[1] + o[2]; h[2] = o[2] + o[3]; h[3] = o[3] + nb;
            if (ph > 0) {
              if (ph == 1) {  // the previous RoI's store must have read the tile
                bulk_wait_read_all();
                __syncwarp();
              }
              const unsigned a = tile_row + 4u * (unsigned)((ph - 1) * 7 + 4 * half);
              f8_sts(a, hp[0] + h[0]);
              f8_sts(a + 4u, hp[1] + h[1]);
              f8_sts(a + 8u, hp[2] + h[2]);
              if (!half) f8_sts(a + 12u, hp[3] + h[3]);
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) hp[k] = h[k];
          }
        }
        if (FUSED) {
          fence_proxy_async_smem();
          __syncwarp();
          if (elect_one()) {
            f8_bulk_store(out_roi, tile, F8_TILE7, pol);
            bulk_commit_group();
          }
        }
        __syncwarp();  // wtab is rewritten for the next RoI
      }
      n = n2;
      t = t2;
    }
    u += (r_hi - r_lo);
  }
  bulk_wait_all();  // the tiles must outlive their stores
}

// (R, C, 8, 8) fp32 output seen as (R, C, 2, 32): box = rows 0-3 or 4-7 of 16 consecutive channels
static bool f8_make_out_tmap(CUtensorMap* map, float* output, int R, int C) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) return false;
  const cuuint64_t dims[4] = {32, 2, (cuuint64_t)C, (cuuint64_t)R};
  const cuuint64_t strides[3] = {128, 256, (cuuint64_t)C * 256};
  const cuuint32_t box[4] = {32, 1, (cuuint32_t)F8_CH, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, output, dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// Launches the 8x8-sample forward if this shape is served by it; returns TLOD_ERR_UNSUPPORTED
// (nothing launched) otherwise.
int roi_align_fwd8_launch(bool fused, const float* features, float* output, int batch, int channels, int height,
                          int width, int num_rois, const void* plan, cudaStream_t st) {
  if (!plan || channels % F8_CH != 0 || !plan_supported(batch, height, width, 8, 8)) return TLOD_ERR_UNSUPPORTED;
  if (getenv("TLOD_DISABLE_FWD8")) return TLOD_ERR_UNSUPPORTED;  // A/B timing knob (tools/prof_roi_align.py)
  if (((uintptr_t)features & 15) || ((uintptr_t)output & (fused ? 15 : 127))) return TLOD_ERR_UNSUPPORTED;
  const F8Layout L = f8_layout(height, width, fused, (size_t)device_info().max_smem_optin);
  if (L.warps < 8) return TLOD_ERR_UNSUPPORTED;
  CUtensorMap omap;
  memset(&omap, 0, sizeof(omap));
  if (!fused && !f8_make_out_tmap(&omap, output, num_rois, channels)) return TLOD_ERR_UNSUPPORTED;
  const int P = height * width;
  const int bulk = (L.Pp == P) ? 1 : 0;  // unpadded planes: the slab is one contiguous 64-byte aligned block
  const long long units = (long long)(channels / F8_CH) * num_rois;
  int grid = device_info().sm_count;
  if ((long long)grid > units) grid = (int)units;
  auto kern = fused ? (width == 75 ? roi_align_fwd8_kernel<true, 75> : roi_align_fwd8_kernel<true, 0>)
                    : (width == 75 ? roi_align_fwd8_kernel<false, 75> : roi_align_fwd8_kernel<false, 0>);
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total);
  if (e != cudaSuccess) return (int)e;
  const PlanPtrs pl = plan_ptrs(const_cast<void*>(plan), batch, num_rois);
  {
    LaunchScope scope(fused ? "roi_align_avg_fwd8_kernel" : "roi_align_fwd8_kernel", st);
    kern<<<grid, 32 * L.warps, L.total, st>>>(omap, features, output, pl, batch, channels, height, width, num_rois,
                                              L.Pp, bulk, (unsigned)L.tiles, (unsigned)L.wtab, (unsigned)L.ctl);
  }
  return last_launch_status();
}

}  // namespace tlod

using namespace tlod;

extern "C" size_t tlod_roi_align_avg_scratch_bytes(int channels, int num_rois, int pooled_h, int pooled_w) {
  if (channels <= 0 || num_rois < 0 || pooled_h < 1 || pooled_w < 1) return 0;
  return (size_t)num_rois * channels * (pooled_h + 1) * (pooled_w + 1) * sizeof(float);
}

extern "C" int tlod_roi_align_avg_forward(const float* features, const float* rois, float* output, int batch,
                                          int channels, int height, int width, int num_rois, int pooled_h,
                                          int pooled_w, float spatial_scale, const void* plan, size_t plan_bytes,
                                          void* scratch, size_t scratch_bytes, void* stream) {
  const int ah = pooled_h + 1, aw = pooled_w + 1;
  int rc = roi_align_check_common(features, rois, output, batch, channels, height, width, num_rois, ah, aw);
  if (rc != TLOD_OK) return rc;
  if (pooled_h < 1 || pooled_w < 1) return TLOD_ERR_BAD_SHAPE;
  if (num_rois == 0) return TLOD_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (plan && (plan_bytes < tlod_roi_align_plan_bytes(batch, num_rois) || ((uintptr_t)plan & 255)))
    return TLOD_ERR_WORKSPACE;
  if (ah == 8 && aw == 8) {
    rc = roi_align_fwd8_launch(true, features, output, batch, channels, height, width, num_rois, plan, st);
    if (rc != TLOD_ERR_UNSUPPORTED) return rc;
  }
  // composed: samples into the scratch tensor, then the 2x2 average
  if (!scratch || scratch_bytes < tlod_roi_align_avg_scratch_bytes(channels, num_rois, pooled_h, pooled_w))
    return TLOD_ERR_WORKSPACE;
  rc = tlod_roi_align_forward(features, rois, (float*)scratch, batch, channels, height, width, num_rois, ah, aw,
                              spatial_scale, plan, plan_bytes, stream);
  if (rc != TLOD_OK) return rc;
  return tlod_avgpool2x2_forward((const float*)scratch, output, (long long)num_rois * channels, ah, aw, stream);
}

extern "C" int tlod_roi_align_avg_backward(const float* top_grad, const float* rois, float* bottom_grad, int batch,
                                           int channels, int height, int width, int num_rois, int pooled_h,
                                           int pooled_w, float spatial_scale, const void* plan, size_t plan_bytes,
                                           void* scratch, size_t scratch_bytes, void* stream) {
  const int ah = pooled_h + 1, aw = pooled_w + 1;
  int rc = roi_align_check_common(top_grad, rois, bottom_grad, batch, channels, height, width, num_rois, ah, aw);
  if (rc != TLOD_OK) return rc;
  if (pooled_h < 1 || pooled_w < 1) return TLOD_ERR_BAD_SHAPE;
  if (num_rois > 0) {
    if (!scratch || scratch_bytes < tlod_roi_align_avg_scratch_bytes(channels, num_rois, pooled_h, pooled_w))
      return TLOD_ERR_WORKSPACE;
    rc = tlod_avgpool2x2_backward(top_grad, (float*)scratch, (long long)num_rois * channels, ah, aw, stream);
    if (rc != TLOD_OK) return rc;
  }
  return tlod_roi_align_backward(num_rois > 0 ? (const float*)scratch : top_grad, rois, bottom_grad, batch, channels,
                                 height, width, num_rois, ah, aw, spatial_scale, plan, plan_bytes, stream);
}
