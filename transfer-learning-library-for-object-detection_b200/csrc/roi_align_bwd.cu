// RoIAlign backward for sm_100a: row-resident, no global atomics, no memset.
//
// Reference semantics: lib/model/roi_align/src/roi_align_kernel.cu:94-143 (4 fp32 atomicAdd
// per output element, 537 M at BASELINE cfg3); glue roi_align_cuda.c:42-76.
//
// One warp owns one row of the gradient map for 32 (or 64) channels: (image b, plane row y,
// channel group g).  Lane = channel (or the channel pair c, c + 32), so a cell is only ever
// touched by one thread: the scatter is plain load-add-store on shared memory in a fixed order
// (bitwise reproducible), and the finished row is written to HBM once, coalesced.
//
// The plan (roi_align.cu) holds, per (image, plane row), the list of gradient rows (RoI n,
// output row ph, row weight) that feed it.  The warp walks its list:
//   * one elected lane fetches the 32-byte gradient rows of the warp's channels (256 B apart in
//     HBM) with one TMA tile copy (tensor map over (R, C, AH, 8), box 8 x 1 x 32|64 x 1, 32-byte
//     swizzle) into a private ring of stages, and the RoI's column chain (BwdCols, 96 or 224 B)
//     with a bulk copy on the same mbarrier when the RoI changes;
//   * every lane reads its channel's row with two conflict-free LDS.128 and adds weight * row
//     into 8 registers -- all output rows of a RoI have the same column structure, so the rows
//     that feed this plane row are summed first;
//   * when the RoI changes, the column chain turns the 8 sums into <= 16 distinct cells, which
//     are added to the row: 16 loads, 16 adds, 16 stores, no predicates (sites that are not
//     emitted point at a dump cell behind the row).
// Warps are independent: no synchronisation between rows, and the grid (B * H * C/32 or C/64
// CTAs) is balanced by the hardware scheduler.  On small grids a CTA has K warps that split the
// row's list and sum their copies of the row at the end.
#include "async_copy.cuh"
#include "roi_align_plan.cuh"

namespace tlod {

constexpr int RW_STAGES = 2;  // ring depth per warp (2..4 measure within 3 %; fewer = more warps per SM)
constexpr int RW_META_BYTES = 256;  // BwdCols (224), padded: keeps the tiles' swizzle phase
// CPL = channels per lane (1 or 2): a warp owns 32 * CPL channels of its row.  With CPL = 2 a
// cell holds the pair (channel lane, channel lane + 32) as a float2: one 64-bit shared-memory
// access updates both, and the per-item bookkeeping and the column state are paid once per 64
// channels -- but a warp needs twice the shared memory.  Used whenever channels % 64 == 0.
__host__ __device__ constexpr int rw_tile_bytes(int cpl) { return 32 * 32 * cpl; }  // 32-byte rows
__host__ __device__ constexpr int rw_stage_bytes(int cpl) { return rw_tile_bytes(cpl) + RW_META_BYTES; }
// bytes of shared memory one warp owns: [stages][row: 32 lanes x Ws cells x CPL floats][full barriers]
__host__ __device__ inline unsigned rw_warp_bytes(int Ws, int cpl) {
  return (unsigned)(RW_STAGES * rw_stage_bytes(cpl) + 32 * Ws * 4 * cpl + RW_STAGES * 8 + 255) / 256u * 256u;
}

__device__ __forceinline__ float lds_f32(unsigned addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts_f32(unsigned addr, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ float2 lds_v2(unsigned addr) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts_v2(unsigned addr, float a, float b) {
  asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ float4 lds_v4(unsigned addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
// one cell of the lane's row: CPL floats
template <int CPL>
struct Cell {
  float v[CPL];
};
template <int CPL>
__device__ __forceinline__ Cell<CPL> cell_load(unsigned addr) {
  Cell<CPL> c;
  if (CPL == 1) {
    c.v[0] = lds_f32(addr);
  } else {
    const float2 t = lds_v2(addr);
    c.v[0] = t.x;
    c.v[CPL - 1] = t.y;
  }
  return c;
}
template <int CPL>
__device__ __forceinline__ void cell_store(unsigned addr, const Cell<CPL>& c) {
  if (CPL == 1) sts_f32(addr, c.v[0]);
  else sts_v2(addr, c.v[0], c.v[CPL - 1]);
}

// column state of the current RoI, for this lane's row
struct RowCols {
  float cw0[8], cw1[8], ms[8], mh[8];
  unsigned sa[16];  // shared-window address of site j in this lane's row (or of its dump cell);
                    // all-jump RoIs: sa[t] = first cell of sample t, t < 8
};

__device__ __forceinline__ void rw_unpack(float (&d)[8], const float4 a, const float4 b) {
  d[0] = a.x; d[1] = a.y; d[2] = a.z; d[3] = a.w; d[4] = b.x; d[5] = b.y; d[6] = b.z; d[7] = b.w;
}
// plan offsets are bytes of a one-float cell: scaled by CPL here
template <int CPL>
__device__ __forceinline__ void rw_unpack_addr(unsigned* d, const float4 a, unsigned row_addr) {
  d[0] = row_addr + (unsigned)__float_as_int(a.x) * CPL; d[1] = row_addr + (unsigned)__float_as_int(a.y) * CPL;
  d[2] = row_addr + (unsigned)__float_as_int(a.z) * CPL; d[3] = row_addr + (unsigned)__float_as_int(a.w) * CPL;
}

// all lanes read the same bytes (broadcast).  JUMP: the 96-byte head of BwdCols; else all of it.
template <bool JUMP, int CPL>
__device__ __forceinline__ void rw_load_cols(RowCols& s, unsigned meta, unsigned row_addr) {
  rw_unpack(s.cw0, lds_v4(meta), lds_v4(meta + 16));
  rw_unpack(s.cw1, lds_v4(meta + 32), lds_v4(meta + 48));
  if (JUMP) {
    rw_unpack_addr<CPL>(s.sa, lds_v4(meta + 64), row_addr);
    rw_unpack_addr<CPL>(s.sa + 4, lds_v4(meta + 80), row_addr);
  } else {
    rw_unpack(s.ms, lds_v4(meta + 96), lds_v4(meta + 112));
    rw_unpack(s.mh, lds_v4(meta + 128), lds_v4(meta + 144));
#pragma unroll
    for (int i = 0; i < 4; ++i) rw_unpack_addr<CPL>(s.sa + 4 * i, lds_v4(meta + 160 + 16 * i), row_addr);
  }
}

// The RoI's summed gradient rows m -> values of the 16 column sites -> added to the row.
// All-jump RoIs: sample t owns cells (x_t, x_t + 1) alone.
template <int CPL>
__device__ __forceinline__ void rw_flush_jump(const RowCols& s, const float (&m)[CPL][8]) {
  Cell<CPL> o0[8], o1[8];
#pragma unroll
  for (int t = 0; t < 8; ++t) {
    o0[t] = cell_load<CPL>(s.sa[t]);
    o1[t] = cell_load<CPL>(s.sa[t] + 4u * CPL);
  }
#pragma unroll
  for (int t = 0; t < 8; ++t) {
#pragma unroll
    for (int k = 0; k < CPL; ++k) {
      o0[t].v[k] = fmaf(m[k][t], s.cw0[t], o0[t].v[k]);
      o1[t].v[k] = fmaf(m[k][t], s.cw1[t], o1[t].v[k]);
    }
    cell_store<CPL>(s.sa[t], o0[t]);
    cell_store<CPL>(s.sa[t] + 4u * CPL, o1[t]);
  }
}
// General RoIs: the chain of BwdCols.  The loads of the old values are issued first, so that the
// serial chain runs under their latency; all 16 cells are distinct (or dump cells), so the order
// is free.
template <int CPL>
__device__ __forceinline__ void rw_flush_chain(const RowCols& s, const float (&m)[CPL][8]) {
  Cell<CPL> o[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) o[j] = cell_load<CPL>(s.sa[j]);
#pragma unroll
  for (int k = 0; k < CPL; ++k) {
    float a0 = 0.f, a1 = 0.f;
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      if (t > 0) { o[2 * (t - 1)].v[k] += a0; o[2 * (t - 1) + 1].v[k] += a1; }
      const float na0 = fmaf(s.ms[t], a0, fmaf(s.mh[t], a1, m[k][t] * s.cw0[t]));
      a1 = fmaf(s.ms[t], a1, m[k][t] * s.cw1[t]);
      a0 = na0;
    }
    o[14].v[k] += a0;
    o[15].v[k] += a1;
  }
#pragma unroll
  for (int j = 0; j < 16; ++j) cell_store<CPL>(s.sa[j], o[j]);
}

// blockDim.x = 32 * K: on small grids the K warps of a CTA split the row's item list (each with
// its own copy of the row, summed when the row is written), so that the machine is filled.
template <int K, int CPL>
__global__ void __launch_bounds__(32 * K)
    roi_align_bwd_rows_kernel(const __grid_constant__ CUtensorMap tmap, float* __restrict__ bottom_grad,
                              PlanPtrs pl, int B, int C, int H, int W, int Ws) {
  constexpr int TILE = rw_tile_bytes(CPL), STAGE = rw_stage_bytes(CPL), CH = 32 * CPL;
  extern __shared__ __align__(1024) unsigned char smem_rw[];
  const int lane = lane_id(), wid = K > 1 ? warp_id() : 0;
  unsigned char* mine = smem_rw + (size_t)wid * rw_warp_bytes(Ws, CPL);
  const unsigned stages = smem_u32(mine);
  float* row = reinterpret_cast<float*>(mine + RW_STAGES * STAGE);
  const unsigned bars = smem_u32(row + 32 * Ws * CPL);
  const int y = blockIdx.x % H;
  const int rest = blockIdx.x / H;
  const int groups = C / CH;
  const int grp = rest % groups, img = rest / groups;
  const int c0 = grp * CH;
  const unsigned row_addr = smem_u32(row + lane * Ws * CPL);

  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < RW_STAGES; ++s) mbar_init(bars + 8u * s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = lane; i < 32 * Ws * CPL; i += 32) row[i] = 0.f;
  __syncwarp();

  const int bin = img * H + y;
  const int cnt_all = __ldg(pl.rowcnt + bin);
  const int seg_lo = (int)((long long)cnt_all * wid / K), seg_hi = (int)((long long)cnt_all * (wid + 1) / K);
  const int cnt = seg_hi - seg_lo;  // this warp's share of the list
  const RowItem* __restrict__ items = pl.items + __ldg(pl.rowptr + bin) + seg_lo;

  // lane's 32-byte row inside a tile, 16-byte halves swapped by the 32-byte swizzle (the second
  // channel's row, 32 rows further, has the same swizzle phase)
  const unsigned g_lo = (unsigned)(lane * 32 + (((lane >> 2) & 1) << 4));
  unsigned g_hi;  // = g_lo ^ 16, kept in a register (the compiler would recompute it from %tid per item)
  asm volatile("xor.b32 %0, %1, 16;" : "=r"(g_hi) : "r"(g_lo));

  // The issue stream reads the list sequentially: 32 entries per coalesced load, one chunk ahead,
  // handed out by shuffles (no global-memory latency on the per-item path).
  const int2* __restrict__ items2 = reinterpret_cast<const int2*>(items);
  int2 chunk = lane < cnt ? __ldg(items2 + lane) : make_int2(0, 0);
  int2 chunk_next = 32 + lane < cnt ? __ldg(items2 + 32 + lane) : make_int2(0, 0);
  int issue_prev = -1;  // key of the item issued last (the BwdCols travel with the first item of a RoI)
  int en[RW_STAGES];    // key = (RoI << 1) | all_jump of the item in stage s
  float ew[RW_STAGES];  // its row weight
  // all lanes run this (warp-uniform); one elected lane issues the asynchronous copies
  auto issue = [&](int j, int s) {
    if ((j & 31) == 0 && j > 0) {  // the issue stream enters the next chunk
      chunk = chunk_next;
      chunk_next = j + 32 + lane < cnt ? __ldg(items2 + j + 32 + lane) : make_int2(0, 0);
    }
    const int x = __shfl_sync(0xffffffffu, chunk.x, j & 31);
    const int n = x >> 4, ph = x & 15;  // n = key
    en[s] = n;
    ew[s] = __int_as_float(__shfl_sync(0xffffffffu, chunk.y, j & 31));
    const unsigned bar = bars + 8u * s, stg = stages + (unsigned)(s * STAGE);
    const unsigned cols = n == issue_prev ? 0u : ((n & 1) ? BWDCOLS_JUMP_BYTES : (unsigned)sizeof(BwdCols));
    if (elect_one()) {
      mbar_arrive_expect_tx(bar, TILE + cols);
      tma_load_4d(stg, &tmap, bar, 0, ph, c0, n >> 1);
      if (cols) bulk_load(stg + TILE, pl.bwdx + (n >> 1), cols, bar);
    }
    issue_prev = n;
  };

#pragma unroll
  for (int s = 0; s < RW_STAGES; ++s)
    if (s < cnt) issue(s, s);

  // column state of the current RoI; before the first one every site is the dump cell
  RowCols st;
#pragma unroll
  for (int t = 0; t < 8; ++t) { st.cw0[t] = 0.f; st.cw1[t] = 0.f; st.ms[t] = 0.f; st.mh[t] = 0.f; }
#pragma unroll
  for (int j = 0; j < 16; ++j) st.sa[j] = row_addr + 4u * CPL * (unsigned)W;
  int cur = -1;  // never an item's key; odd: the first flush takes the all-jump path, into the dump cell
  float m[CPL][8];
#pragma unroll
  for (int k = 0; k < CPL; ++k)
#pragma unroll
    for (int t = 0; t < 8; ++t) m[k][t] = 0.f;

  for (int base = 0; base < cnt; base += RW_STAGES) {
    const unsigned parity = (unsigned)(base / RW_STAGES) & 1u;
#pragma unroll
    for (int s = 0; s < RW_STAGES; ++s) {
      const int i = base + s;
      if (i < cnt) {
        const unsigned stg = stages + (unsigned)(s * STAGE);
        mbar_wait(bars + 8u * s, parity);
        const int n = en[s];
        const float w = ew[s];
        float4 ga[CPL], gb[CPL];
#pragma unroll
        for (int k = 0; k < CPL; ++k) {
          ga[k] = lds_v4(stg + g_lo + 1024u * k);
          gb[k] = lds_v4(stg + g_hi + 1024u * k);
        }
        if (n != cur) {
          if (cur & 1) rw_flush_jump<CPL>(st, m); else rw_flush_chain<CPL>(st, m);
#pragma unroll
          for (int k = 0; k < CPL; ++k)
#pragma unroll
            for (int t = 0; t < 8; ++t) m[k][t] = 0.f;
          if (n & 1) rw_load_cols<true, CPL>(st, stg + TILE, row_addr);
          else rw_load_cols<false, CPL>(st, stg + TILE, row_addr);
          cur = n;
        }
#pragma unroll
        for (int k = 0; k < CPL; ++k) {
          m[k][0] = fmaf(w, ga[k].x, m[k][0]); m[k][1] = fmaf(w, ga[k].y, m[k][1]);
          m[k][2] = fmaf(w, ga[k].z, m[k][2]); m[k][3] = fmaf(w, ga[k].w, m[k][3]);
          m[k][4] = fmaf(w, gb[k].x, m[k][4]); m[k][5] = fmaf(w, gb[k].y, m[k][5]);
          m[k][6] = fmaf(w, gb[k].z, m[k][6]); m[k][7] = fmaf(w, gb[k].w, m[k][7]);
        }
        // the stage has been read (the sums above depend on it): refill it
        __syncwarp();
        if (i + RW_STAGES < cnt) issue(i + RW_STAGES, s);
      }
    }
  }
  if (cur & 1) rw_flush_jump<CPL>(st, m); else rw_flush_chain<CPL>(st, m);
  if (K > 1) __syncthreads(); else __syncwarp();

  // ---- write the row: lanes along the cells; warp k takes lane-rows k, k + K, ... ----
  float* out = bottom_grad + (((size_t)img * C + c0) * H + y) * W;
  const size_t plane = (size_t)H * W;
  if (CPL == 1) {
    const float* row0 = reinterpret_cast<const float*>(smem_rw + RW_STAGES * STAGE);
    const unsigned wstride = rw_warp_bytes(Ws, CPL) / 4u;
    for (int ch = wid; ch < 32; ch += K)
      for (int x = lane; x < W; x += 32) {
        float v = row0[ch * Ws + x];
#pragma unroll
        for (int k = 1; k < K; ++k) v += row0[k * wstride + ch * Ws + x];
        out[ch * plane + x] = v;
      }
  } else {
    const unsigned row_base = smem_u32(smem_rw + RW_STAGES * STAGE);
    const unsigned wbytes = rw_warp_bytes(Ws, CPL);
    for (int ch = wid; ch < 32; ch += K)
      for (int x = lane; x < W; x += 32) {
        float2 v = lds_v2(row_base + 8u * (unsigned)(ch * Ws + x));
#pragma unroll
        for (int k = 1; k < K; ++k) {
          const float2 t = lds_v2(row_base + (unsigned)k * wbytes + 8u * (unsigned)(ch * Ws + x));
          v.x += t.x;
          v.y += t.y;
        }
        out[ch * plane + x] = v.x;
        out[(ch + 32) * plane + x] = v.y;
      }
  }
}

// (R, C, AH, 8) fp32 gradient tensor; box = one 8-wide row of `box_ch` consecutive channels
static bool make_grad_tmap(CUtensorMap* map, const float* top_grad, int R, int C, int AH, int box_ch) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) return false;
  const cuuint64_t dims[4] = {8, (cuuint64_t)AH, (cuuint64_t)C, (cuuint64_t)R};
  const cuuint64_t strides[3] = {32, (cuuint64_t)AH * 32, (cuuint64_t)C * AH * 32};
  const cuuint32_t box[4] = {8, 1, (cuuint32_t)box_ch, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(top_grad), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace tlod

using namespace tlod;

extern "C" int tlod_roi_align_backward(const float* top_grad, const float* rois, float* bottom_grad,
                                       int batch, int channels, int height, int width,
                                       int num_rois, int aligned_h, int aligned_w,
                                       float spatial_scale, const void* plan, size_t plan_bytes,
                                       void* stream) {
  int rc = roi_align_check_common(top_grad, rois, bottom_grad, batch, channels, height, width, num_rois,
                                  aligned_h, aligned_w);
  if (rc != TLOD_OK) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t grad_bytes = (size_t)batch * channels * height * width * sizeof(float);
  if (num_rois == 0) return (int)cudaMemsetAsync(bottom_grad, 0, grad_bytes, st);
  if (plan && (plan_bytes < tlod_roi_align_plan_bytes(batch, num_rois) || ((uintptr_t)plan & 255)))
    return TLOD_ERR_WORKSPACE;
  const bool planned = plan != nullptr && plan_supported(batch, height, width, aligned_h, aligned_w);

  // row-resident path: AW == 8 (one 32-byte gradient row per item), C % 32 == 0, row lists in the plan
  if (planned && plan_has_row_lists(batch, height) && channels % 32 == 0 && aligned_w == 8 &&
      ((uintptr_t)top_grad & 15) == 0) {
    const int Ws = (width + 1) | 1;  // + dump cell; odd stride: lane = channel is conflict free
    const int sms = device_info().sm_count;
    const size_t smem_max = (size_t)device_info().max_smem_optin;
    // Two channels per lane (half the instructions per channel, twice the shared memory per warp)
    // whenever channels % 64 == 0; K warps per row on grids that would leave SMs idle.
    int cpl = 1, K = 1;
    if (channels % 64 == 0 && rw_warp_bytes(Ws, 2) <= smem_max) {
      cpl = 2;
      const long long grid2 = (long long)batch * height * (channels / 64);
      while (K < 4 && grid2 * K < 6LL * sms && (size_t)(K + 1) * rw_warp_bytes(Ws, 2) <= smem_max) ++K;
    } else {
      const long long grid1 = (long long)batch * height * (channels / 32);
      while (K < 4 && grid1 * K < 12LL * sms && (size_t)(K + 1) * rw_warp_bytes(Ws, 1) <= smem_max) ++K;
    }
    const long long grid = (long long)batch * height * (channels / (32 * cpl));
    const size_t smem = (size_t)K * rw_warp_bytes(Ws, cpl);
    CUtensorMap tmap;
    if (grid <= 2147483647LL && smem <= smem_max &&
        make_grad_tmap(&tmap, top_grad, num_rois, channels, aligned_h, 32 * cpl)) {
      auto kern = cpl == 2 ? (K == 1 ? roi_align_bwd_rows_kernel<1, 2> : K == 2 ? roi_align_bwd_rows_kernel<2, 2>
                              : K == 3 ? roi_align_bwd_rows_kernel<3, 2> : roi_align_bwd_rows_kernel<4, 2>)
                           : (K == 1 ? roi_align_bwd_rows_kernel<1, 1> : K == 2 ? roi_align_bwd_rows_kernel<2, 1>
                              : K == 3 ? roi_align_bwd_rows_kernel<3, 1> : roi_align_bwd_rows_kernel<4, 1>);
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return (int)e;
      const PlanPtrs pl = plan_ptrs(const_cast<void*>(plan), batch, num_rois);
      {
        LaunchScope scope("roi_align_bwd_rows_kernel", st);
        kern<<<(int)grid, 32 * K, smem, st>>>(tmap, bottom_grad, pl, batch, channels, height, width, Ws);
      }
      return last_launch_status();
    }
  }
  cudaError_t e = cudaMemsetAsync(bottom_grad, 0, grad_bytes, st);
  if (e != cudaSuccess) return (int)e;
  return roi_align_generic_launch(true, top_grad, rois, bottom_grad, batch, channels, height, width, num_rois,
                                  aligned_h, aligned_w, spatial_scale, st);
}
