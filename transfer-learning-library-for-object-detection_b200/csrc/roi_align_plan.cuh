// RoIAlign plan: layout of the caller-provided workspace shared by roi_align.cu (plan kernels,
// forward) and roi_align_bwd.cu (backward).  See tlod_roi_align_plan in include/tlod_b200.h.
#pragma once
#include "common.cuh"

namespace tlod {

constexpr int PL_MAXB = 1024;     // images per call on the planned paths
constexpr int PL_MAXA = 16;       // aligned_h / aligned_w limit on the planned paths
constexpr int PL_MAXBINS = 8192;  // (image, plane row) bins of the backward row lists

// Column chain of one RoI for the backward pass (aligned_w == 8).  Walking the 8 samples of
// a row left to right, (a0, a1) accumulate the values of cells (cur, cur + 1):
//     a0' = g*cw0[t] + ms[t]*a0 + mh[t]*a1        a1' = g*cw1[t] + ms[t]*a1
// (same cell: ms=1; moved right by one: mh=1; jumped: both 0).  Before sample t (t = 1..7)
// and after the last one (t = 8) the accumulators that fall out of the window are final:
// a0 -> cell ex[t], a1 -> cell ex[t] + 1 ("sites" 2(t-1) and 2(t-1)+1).  All cells emitted for
// one row are distinct, so their read-modify-writes are independent.  A site that is not
// emitted points at the dump cell behind the row (column `width`), so the scatter needs no
// predicates.
struct __align__(16) BwdCols {
  float cw0[8], cw1[8];
  int xoff[8];   // byte offset of sample t's first cell in the row (0 with zero weights for an invalid
                 // sample); all-jump RoIs: sites 2t, 2t+1 = xoff[t], xoff[t] + 4
  float ms[8], mh[8];
  int soff[16];  // site 2(t-1)+k (t = 1..8, k = 0/1): byte offset (ex[t] + k) * 4 in the row; width * 4 = dump
};
constexpr unsigned BWDCOLS_JUMP_BYTES = 96;  // what an all-jump RoI needs: cw0, cw1, xoff
static_assert(sizeof(BwdCols) == 224, "BwdCols layout");

// One entry of a backward row list: gradient row (RoI, ph) adds weight * (its column scatter)
// to the plane row the list belongs to.
struct __align__(8) RowItem {
  int roi_ph;  // (roi << 5) | (all_jump << 4) | ph; all_jump: every sample of the RoI's rows is valid
               // and starts a new pair of cells, so the column chain is the identity
  float weight;
};

struct PlanLayout {
  size_t cum, list, yrow, jump, tabs, bwdx, rowcnt, rowptr, items, total;
};
__host__ __device__ inline size_t pl_align(size_t v) { return (v + 255) / 256 * 256; }
__host__ __device__ inline PlanLayout plan_layout(int B, int R) {
  PlanLayout L;
  size_t off = 0;
  L.cum = off;    off = pl_align(off + (size_t)(B + 2) * 4);
  L.list = off;   off = pl_align(off + (size_t)R * 4);
  L.yrow = off;   off = pl_align(off + (size_t)R * 32);
  L.jump = off;   off = pl_align(off + (size_t)R);
  L.tabs = off;   off = pl_align(off + (size_t)R * 512);
  L.bwdx = off;   off = pl_align(off + (size_t)R * sizeof(BwdCols));
  L.rowcnt = off; off = pl_align(off + (size_t)PL_MAXBINS * 4);
  L.rowptr = off; off = pl_align(off + (size_t)PL_MAXBINS * 4);
  L.items = off;  off = pl_align(off + (size_t)R * 2 * PL_MAXA * sizeof(RowItem));
  L.total = off;
  return L;
}

struct PlanPtrs {
  int* cum;        // [B + 2] exclusive prefix of RoIs per image; slot B = invalid image index
  int* list;       // [R] RoI indices sorted by image, stable
  short* yrow;     // [R][16] first sampled row of each output row, -1 if none
  unsigned char* jump;  // [R] all-jump flag of the RoI's column chain
  float4* tabs;    // [R][32] rows 0..15: {row*W | -1, w0, w1, row | -1}; cols 16..31: {col | -1, w0, w1, col | -1}
  BwdCols* bwdx;   // [R]
  int* rowcnt;     // [B * H] items per (image, plane row)
  int* rowptr;     // [B * H] first item of (image, plane row) in `items`
  RowItem* items;  // row lists, each in (image-sorted RoI order, ph) order
};
__host__ __device__ inline PlanPtrs plan_ptrs(void* base, int B, int R) {
  const PlanLayout L = plan_layout(B, R);
  unsigned char* p = (unsigned char*)base;
  PlanPtrs q;
  q.cum = (int*)(p + L.cum);
  q.list = (int*)(p + L.list);
  q.yrow = (short*)(p + L.yrow);
  q.jump = p + L.jump;
  q.tabs = (float4*)(p + L.tabs);
  q.bwdx = (BwdCols*)(p + L.bwdx);
  q.rowcnt = (int*)(p + L.rowcnt);
  q.rowptr = (int*)(p + L.rowptr);
  q.items = (RowItem*)(p + L.items);
  return q;
}

inline bool plan_supported(int batch, int height, int width, int ah, int aw) {
  return batch <= PL_MAXB && ah <= PL_MAXA && aw <= PL_MAXA && height < 32768 &&
         (long long)height * width < (1LL << 30);
}
inline bool plan_has_row_lists(int batch, int height) {
  return (long long)batch * height <= PL_MAXBINS;
}

// shared host helpers (roi_align.cu)
int roi_align_check_common(const void* a, const void* b, const void* c, int batch, int channels, int height,
                           int width, int num_rois, int ah, int aw);
int roi_align_generic_launch(bool backward, const float* src, const float* rois, float* dst, int batch,
                             int channels, int height, int width, int num_rois, int ah, int aw, float scale,
                             cudaStream_t st);

// roi_align_fwd8.cu: the 8 x 8-sample forward (fused: RoIAlignAvg(7, 7) output).  Returns
// TLOD_ERR_UNSUPPORTED, with nothing launched, for shapes it does not serve.
int roi_align_fwd8_launch(bool fused, const float* features, float* output, int batch, int channels, int height,
                          int width, int num_rois, const void* plan, cudaStream_t st);

}  // namespace tlod
