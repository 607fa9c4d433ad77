// RoICrop on an axis-aligned 14 x 14 grid fused with its max_pool2d(2, 2): the 'crop' pooling mode
// of the detector (lib/model/faster_rcnn/faster_rcnn.py:73-80: _affine_grid_gen -> RoICropFunction
// -> F.max_pool2d(., 2, 2)), BASELINE cfg3 (ii).  The (R, C, 14, 14) sample tensor -- 1.64 GB at cfg3,
// written by the crop kernel and read back by the pooling kernel in the reference -- never exists.
//
// Sampling semantics: lib/model/roi_crop/src/roi_crop_cuda_kernel.cu:12-23 (getTopLeft: coordinate
// (x + 1) * (size - 1) / 2, top-left cell = floor, weight of it = 1 - frac) and :45-108 (four
// corners, corners outside the map contribute zero).  The grid of _affine_grid_gen has no rotation
// (theta01 = theta10 = 0, lib/model/utils/net_utils.py:142-164), so it is the outer product of a
// row-coordinate vector grid_y (R, 14) and a column-coordinate vector grid_x (R, 14); the entry
// points take those two vectors.
//
// Forward = the plane-resident gather of roi_align_fwd8.cu with 14 samples per axis: a CTA keeps
// the 16 planes of (image, 16-channel slab) in shared memory, a warp takes a RoI, lane =
// (channel, half); half 0 owns samples 0..6 of every sample row, half 1 samples 8..13 and 7 (in
// that order, so that both halves pool their local pairs (0,1), (2,3), (4,5) and half 0 pools
// its sample 6 with half 1's sample 7, one shuffle per row).  Sample rows are walked in order
// with the horizontally interpolated plane rows reused between consecutive sample rows (bins are
// roi / 13 cells apart: most sample rows share a plane row with their predecessor).  The pooled
// 16 x 49 values and the 16 x 49 one-byte argmax codes (which of the four samples won: 2 * dy +
// dx, first maximum in row-major order like max_pool2d) leave with one cp.async.bulk store each.
//
// Backward: the gradient of a pooled cell goes to the four corners of its winning sample with
// that sample's bilinear weights.  The winner differs per channel, so the scatter is not
// separable (no row-resident scheme): one thread per pooled cell, four fp32 REDs into the
// zero-filled gradient map -- a quarter of the reference path's atomics (which scatters all
// 196 samples of a tile, three quarters of them with the pooling's zero gradient).
#include <cstring>

#include "async_copy.cuh"
#include "common.cuh"

namespace tlod {

constexpr int CP_NS = 14;   // samples per axis

constexpr int CP_CH = 16;
constexpr int CP_TILE_V = CP_CH * 49 * 4;  // pooled values of a (RoI, slab)
constexpr int CP_TILE_A = CP_CH * 49;      // argmax codes
constexpr int CP_WTAB = 512;               // 32 float4: rows 0..13, columns 16..29
constexpr int CP_PER_WARP = CP_TILE_V + CP_TILE_A + CP_WTAB;

__host__ __device__ inline int cp_plane_stride(int P) {
  while ((P & 3) != 2) ++P;
  return P;
}

// One axis of a sampling site as a table entry {first cell | -1, weight of it, weight of the next
// cell, first cell | -1}: both cells are always inside the map, weights of outside cells are zero.
__device__ __forceinline__ float4 cp_axis_entry(float coord_norm, int size) {
  const float coord = __fdiv_rn(__fmul_rn(__fadd_rn(coord_norm, 1.0f), (float)(size - 1)), 2.0f);
  const float fl = floorf(coord);
  const int p = (int)fl;
  const float w_tl = __fsub_rn(1.0f, __fsub_rn(coord, fl));  // weight of cell p
  const float w_br = __fsub_rn(1.0f, w_tl);                  // weight of cell p + 1
  int start = -1;
  float w0 = 0.f, w1 = 0.f;
  if (p >= 0 && p <= size - 2) { start = p; w0 = w_tl; w1 = w_br; }
  else if (p == -1) { start = 0; w0 = w_br; }                      // only cell 0 = p + 1 is inside
  else if (p == size - 1 && size >= 2) { start = size - 2; w1 = w_tl; }  // only cell size - 1 = p
  return make_float4(__int_as_float(start), w0, w1, __int_as_float(start));
}

// tabs[n][32]: rows 0..13 from grid_y, columns 16..29 from grid_x
__global__ void __launch_bounds__(256)
    roi_crop_pool_plan_kernel(const float* __restrict__ grid_y, const float* __restrict__ grid_x,
                              float4* __restrict__ tabs, int R, int H, int W) {
  const int n = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (n >= R) return;
  float4 t = make_float4(__int_as_float(-1), 0.f, 0.f, __int_as_float(-1));
  if (lane < CP_NS) t = cp_axis_entry(__ldg(grid_y + (size_t)n * CP_NS + lane), H);
  else if (lane >= 16 && lane < 16 + CP_NS) t = cp_axis_entry(__ldg(grid_x + (size_t)n * CP_NS + lane - 16), W);
  tabs[(size_t)n * 32 + lane] = t;
}

__device__ __forceinline__ float cp_lds(unsigned addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void cp_bulk_store(void* gdst, unsigned smem_src, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_src), "r"(bytes)
               : "memory");
}

__device__ __forceinline__ void cp_hrow(float (&T)[7], const unsigned (&ca)[7], const unsigned (&cb)[7],
                                        const float (&wp)[7], const float (&wq)[7], unsigned off) {
  float a[7], b[7];
#pragma unroll
  for (int k = 0; k < 7; ++k) {
    a[k] = cp_lds(ca[k] + off);
    b[k] = cp_lds(cb[k] + off);
  }
#pragma unroll
  for (int k = 0; k < 7; ++k) T[k] = fmaf(b[k], wq[k], a[k] * wp[k]);
}

// RoIs are grouped by image by construction: image of RoI n = n / per_image.
__global__ void __launch_bounds__(512, 1)
    roi_crop_pool_fwd_kernel(const float* __restrict__ features, const float4* __restrict__ tabs,
                             float* __restrict__ output, unsigned char* __restrict__ argmax, int B, int C, int H,
                             int W, int per_image, int Pp, int bulk_stage, unsigned warp_off) {
  extern __shared__ __align__(128) unsigned char smem_cp[];
  const int tid = threadIdx.x, lane = lane_id(), wid = warp_id();
  const int nwarps = blockDim.x >> 5;
  const int P = H * W;
  float* planes = reinterpret_cast<float*>(smem_cp);
  unsigned char* mine = smem_cp + warp_off + (size_t)wid * CP_PER_WARP;
  const unsigned tile_v = smem_u32(mine), tile_a = tile_v + CP_TILE_V;
  float4* wtab = reinterpret_cast<float4*>(mine + CP_TILE_V + CP_TILE_A);
  int* cur_base = reinterpret_cast<int*>(smem_cp + warp_off + (size_t)nwarps * CP_PER_WARP);
  const unsigned bar = smem_u32(cur_base) + 32u;
  const int nslabs = C / CP_CH;
  const int c = lane >> 1, half = lane & 1;
  const unsigned plane_addr = smem_u32(planes + (size_t)c * Pp);
  const unsigned row_bytes = 4u * (unsigned)W;
  const unsigned tv_row = tile_v + 196u * (unsigned)c, ta_row = tile_a + 49u * (unsigned)c;

  if (tid == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  unsigned stage_parity = 0;

  // contiguous range of (image, slab, RoI rank) units
  const long long U = (long long)nslabs * per_image * B;
  long long u = U * blockIdx.x / gridDim.x;
  const long long u_end = U * (blockIdx.x + 1) / gridDim.x;
  int staged = -1;
  for (int iter = 0; u < u_end; ++iter) {
    const long long per_img_units = (long long)nslabs * per_image;
    const int img = (int)(u / per_img_units);
    const long long rem = u - (long long)img * per_img_units;
    const int slab = (int)(rem / per_image);
    const int r_lo = (int)(rem - (long long)slab * per_image);
    int r_hi = per_image;
    if (u_end - u < (long long)(r_hi - r_lo)) r_hi = r_lo + (int)(u_end - u);
    const int c0 = slab * CP_CH;
    __syncthreads();  // every warp is done with the planes staged before
    if (img * nslabs + slab != staged) {
      const float* g = features + ((size_t)img * C + c0) * P;
      if (bulk_stage) {
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
        for (int ch = wid; ch < CP_CH; ch += nwarps) {
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
      staged = img * nslabs + slab;
    }

    for (int e = r_lo + wid; e < r_hi; e += nwarps) {
      const int n = img * per_image + e;
      float4 t = __ldg(tabs + (size_t)n * 32 + lane);
      if (lane < 16) {
        const int start = __float_as_int(t.w);
        t.x = __int_as_float(start < 0 ? 0 : start * (int)row_bytes);
      }
      __syncwarp();
      wtab[lane] = t;
      __syncwarp();

      // local sample k of this lane: half 0 -> sample k; half 1 -> samples 8..13, then 7
      unsigned ca[7], cb[7];
      float wp[7], wq[7];
#pragma unroll
      for (int k = 0; k < 7; ++k) {
        const float4 m0 = wtab[16 + k], m1 = wtab[16 + (k < 6 ? 8 + k : 7)];
        int x0 = __float_as_int(m0.x), x1 = __float_as_int(m1.x);
        x0 = x0 < 0 ? 0 : x0;
        x1 = x1 < 0 ? 0 : x1;
        const int flip = half & (((x0 ^ x1) & 1) ^ 1);
        const int x = half ? x1 : x0;
        const float w0 = half ? m1.y : m0.y, w1 = half ? m1.z : m0.z;
        ca[k] = plane_addr + 4u * (unsigned)(x + flip);
        cb[k] = plane_addr + 4u * (unsigned)(x + (flip ^ 1));
        wp[k] = flip ? w1 : w0;
        wq[k] = flip ? w0 : w1;
      }

      float Tlo[7], Thi[7], best[4];
      int code[4];
#pragma unroll
      for (int k = 0; k < 7; ++k) { Tlo[k] = 0.f; Thi[k] = 0.f; }
      int prev = -100;
      bulk_wait_read_all();  // the previous RoI's stores have read the tiles
      __syncwarp();
#pragma unroll 2
      for (int ph = 0; ph < CP_NS; ++ph) {
        const float4 r = wtab[ph];
        const int st = __float_as_int(r.w);
        float o[7];
#pragma unroll
        for (int k = 0; k < 7; ++k) o[k] = 0.f;
        if (st >= 0) {  // warp-uniform
          const unsigned ro = (unsigned)__float_as_int(r.x);
          if (st != prev) {
            if (st == prev + 1) {
#pragma unroll
              for (int k = 0; k < 7; ++k) Tlo[k] = Thi[k];
            } else {
              cp_hrow(Tlo, ca, cb, wp, wq, ro);
            }
            cp_hrow(Thi, ca, cb, wp, wq, ro + row_bytes);
            prev = st;
          }
#pragma unroll
          for (int k = 0; k < 7; ++k) o[k] = fmaf(Thi[k], r.z, Tlo[k] * r.y);
        }
        const float nb = __shfl_xor_sync(0xffffffffu, o[6], 1);  // half 0 receives sample 7
        const float sa[4] = {o[0], o[2], o[4], o[6]}, sb[4] = {o[1], o[3], o[5], nb};
        if ((ph & 1) == 0) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            best[j] = sa[j];
            code[j] = 0;
            if (sb[j] > best[j]) { best[j] = sb[j]; code[j] = 1; }
          }
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (sa[j] > best[j]) { best[j] = sa[j]; code[j] = 2; }
            if (sb[j] > best[j]) { best[j] = sb[j]; code[j] = 3; }
          }
          const unsigned i7 = (unsigned)((ph >> 1) * 7 + 4 * half);
          const unsigned av = tv_row + 4u * i7, aa = ta_row + i7;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (j < 3 || !half) {
              asm volatile("st.shared.f32 [%0], %1;" ::"r"(av + 4u * j), "f"(best[j]) : "memory");
              asm volatile("st.shared.u8 [%0], %1;" ::"r"(aa + (unsigned)j), "r"(code[j]) : "memory");
            }
          }
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (elect_one()) {
        cp_bulk_store(output + ((size_t)n * C + c0) * 49, tile_v, CP_TILE_V);
        if (argmax) cp_bulk_store(argmax + ((size_t)n * C + c0) * 49, tile_a, CP_TILE_A);
        bulk_commit_group();
      }
    }
    u += (r_hi - r_lo);
  }
  bulk_wait_all();
}

// one thread per pooled cell: 4 REDs
__global__ void __launch_bounds__(256)
    roi_crop_pool_bwd_kernel(const float* __restrict__ grad_out, const unsigned char* __restrict__ argmax,
                             const float4* __restrict__ tabs, float* __restrict__ grad_features, int C, int H,
                             int W, int per_image, int chans_per_block) {
  __shared__ float4 tab[32];
  const int n = blockIdx.x;
  const int c0 = blockIdx.y * chans_per_block;
  if (threadIdx.x < 32) tab[threadIdx.x] = __ldg(tabs + (size_t)n * 32 + threadIdx.x);
  __syncthreads();
  const int cb = min(chans_per_block, C - c0);
  const int img = n / per_image;
  const size_t tile = ((size_t)n * C + c0) * 49;
  float* gimg = grad_features + ((size_t)img * C + c0) * H * W;
  for (int o = threadIdx.x; o < cb * 49; o += blockDim.x) {
    const int c = o / 49, i = o - c * 49;
    const int pi = i / 7, pj = i - pi * 7;
    const int code = argmax[tile + o];
    const float g = __ldg(grad_out + tile + o);
    const float4 row = tab[2 * pi + (code >> 1)], col = tab[16 + 2 * pj + (code & 1)];
    const int ys = __float_as_int(row.x), xs = __float_as_int(col.x);
    if (ys < 0 || xs < 0) continue;
    float* p = gimg + (size_t)c * H * W + ys * W + xs;
    const float g0 = g * row.y, g1 = g * row.z;
    if (col.y != 0.f && row.y != 0.f) atomicAdd(p, g0 * col.y);
    if (col.z != 0.f && row.y != 0.f) atomicAdd(p + 1, g0 * col.z);
    if (col.y != 0.f && row.z != 0.f) atomicAdd(p + W, g1 * col.y);
    if (col.z != 0.f && row.z != 0.f) atomicAdd(p + W + 1, g1 * col.z);
  }
}

struct CpLayout {
  int warps, Pp;
  size_t warp_off, total;
};
static CpLayout cp_layout(int H, int W, size_t smem_max) {
  CpLayout L;
  L.Pp = cp_plane_stride(H * W);
  L.warp_off = ((size_t)CP_CH * L.Pp * 4 + 127) / 128 * 128;
  L.warps = 0;
  for (int w = 16; w >= 6; --w)
    if (L.warp_off + (size_t)w * CP_PER_WARP + 64 <= smem_max) {
      L.warps = w;
      break;
    }
  L.total = L.warp_off + (size_t)L.warps * CP_PER_WARP + 64;
  return L;
}

}  // namespace tlod

using namespace tlod;

extern "C" size_t tlod_roi_crop_pool_workspace_bytes(int out_batch) {
  return out_batch < 0 ? 0 : (size_t)out_batch * 32 * sizeof(float4) + 256;
}

static int cp_check(const void* a, const void* gy, const void* gx, const void* o, int in_batch, int channels,
                    int height, int width, int out_batch, int grid_h, int grid_w, const void* ws, size_t ws_bytes) {
  if (!a || !gy || !gx || !o) return TLOD_ERR_NULL_POINTER;
  if (in_batch <= 0 || channels <= 0 || height < 2 || width < 2 || out_batch < 0) return TLOD_ERR_BAD_SHAPE;
  if (out_batch % in_batch != 0) return TLOD_ERR_BAD_SHAPE;
  if (grid_h != CP_NS || grid_w != CP_NS) return TLOD_ERR_UNSUPPORTED;
  if ((long long)in_batch * channels * height * width >= (1LL << 31)) return TLOD_ERR_INT32_OVERFLOW;
  if (!ws || ws_bytes < tlod_roi_crop_pool_workspace_bytes(out_batch) || ((uintptr_t)ws & 15)) return TLOD_ERR_WORKSPACE;
  return TLOD_OK;
}

extern "C" int tlod_roi_crop_pool_forward(const float* features, const float* grid_y, const float* grid_x,
                                          float* output, unsigned char* argmax, int in_batch, int channels,
                                          int height, int width, int out_batch, int grid_h, int grid_w,
                                          void* workspace, size_t workspace_bytes, void* stream) {
  int rc = cp_check(features, grid_y, grid_x, output, in_batch, channels, height, width, out_batch, grid_h, grid_w,
                    workspace, workspace_bytes);
  if (rc != TLOD_OK) return rc;
  if (out_batch == 0) return TLOD_OK;
  if (channels % CP_CH != 0 || ((uintptr_t)features & 15) || ((uintptr_t)output & 15) || ((uintptr_t)argmax & 15))
    return TLOD_ERR_UNSUPPORTED;
  const CpLayout L = cp_layout(height, width, (size_t)device_info().max_smem_optin);
  if (L.warps < 6) return TLOD_ERR_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  float4* tabs = (float4*)workspace;
  {
    LaunchScope scope("roi_crop_pool_plan_kernel", st);
    roi_crop_pool_plan_kernel<<<(out_batch + 7) / 8, 256, 0, st>>>(grid_y, grid_x, tabs, out_batch, height, width);
  }
  rc = last_launch_status();
  if (rc != TLOD_OK) return rc;
  const int per_image = out_batch / in_batch;
  const long long units = (long long)(channels / CP_CH) * out_batch;
  int grid = device_info().sm_count;
  if ((long long)grid > units) grid = (int)units;
  cudaError_t e = cudaFuncSetAttribute(roi_crop_pool_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)L.total);
  if (e != cudaSuccess) return (int)e;
  const int bulk = (L.Pp == height * width) ? 1 : 0;
  {
    LaunchScope scope("roi_crop_pool_fwd_kernel", st);
    roi_crop_pool_fwd_kernel<<<grid, 32 * L.warps, L.total, st>>>(features, tabs, output, argmax, in_batch, channels,
                                                                  height, width, per_image, L.Pp, bulk,
                                                                  (unsigned)L.warp_off);
  }
  return last_launch_status();
}

extern "C" int tlod_roi_crop_pool_backward(const float* grad_output, const unsigned char* argmax, const float* grid_y,
                                           const float* grid_x, float* grad_features, int in_batch, int channels,
                                           int height, int width, int out_batch, int grid_h, int grid_w,
                                           void* workspace, size_t workspace_bytes, void* stream) {
  int rc = cp_check(grad_output, grid_y, grid_x, grad_features, in_batch, channels, height, width, out_batch, grid_h,
                    grid_w, workspace, workspace_bytes);
  if (rc != TLOD_OK) return rc;
  if (!argmax) return TLOD_ERR_NULL_POINTER;
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(grad_features, 0, (size_t)in_batch * channels * height * width * sizeof(float), st);
  if (e != cudaSuccess) return (int)e;
  if (out_batch == 0) return TLOD_OK;
  float4* tabs = (float4*)workspace;
  {
    LaunchScope scope("roi_crop_pool_plan_kernel", st);
    roi_crop_pool_plan_kernel<<<(out_batch + 7) / 8, 256, 0, st>>>(grid_y, grid_x, tabs, out_batch, height, width);
  }
  rc = last_launch_status();
  if (rc != TLOD_OK) return rc;
  int cpb = 64;
  if (cpb > channels) cpb = channels;
  while ((channels + cpb - 1) / cpb > 65535) ++cpb;
  dim3 g(out_batch, (channels + cpb - 1) / cpb);
  {
    LaunchScope scope("roi_crop_pool_bwd_kernel", st);
    roi_crop_pool_bwd_kernel<<<g, 256, 0, st>>>(grad_output, argmax, tabs, grad_features, channels, height, width,
                                                out_batch / in_batch, cpb);
  }
  return last_launch_status();
}
