// The reference's tensor-free launcher ABI, served by this library's kernels: a maintainer can
// link libtlod_b200.so where lib/model/*/src/*_kernel.cu.o used to be linked and keep the cffi
// glue (roi_align_cuda.c, roi_pooling_cuda.c, nms_cuda.c) unchanged.
//
//   ROIAlignForwardLaucher / ROIAlignBackwardLaucher   lib/model/roi_align/src/roi_align_kernel.h:13-27
//   ROIPoolForwardLaucher / ROIPoolBackwardLaucher     lib/model/roi_pooling/src/roi_pooling_kernel.h:8-18
//   nms_cuda_compute                                   lib/model/nms/src/nms_cuda_kernel.h:5-6
//
// Same argument lists, names (including the reference's "Laucher" spelling) and return values
// (1 = launched; the reference exits the process on a launch error, these return 0 instead).
// The reference ABI has no workspace argument, so scratch (the RoIAlign plan, the NMS bitmask) is
// taken from the stream-ordered allocator (cudaMallocAsync / cudaFreeAsync: no synchronisation) --
// the reference's own nms_cuda_compute cudaMallocs and cudaFrees per call as well.
#include <climits>

#include "common.cuh"

namespace {

struct Scratch {
  void* p = nullptr;
  cudaStream_t st;
  Scratch(size_t bytes, cudaStream_t s) : st(s) {
    if (cudaMallocAsync(&p, bytes < 256 ? 256 : bytes, s) != cudaSuccess) p = nullptr;
  }
  ~Scratch() {
    if (p) cudaFreeAsync(p, st);
  }
};

// the forward launcher is not told the batch size (the reference kernel trusts rois[:, 0]): the
// largest batch the 32-bit indexing admits, so that every index the reference would accept is valid
int max_batch(int channels, int height, int width) {
  const long long plane = (long long)channels * height * width;
  long long b = plane > 0 ? ((1LL << 31) - 1) / plane : 1;
  if (b > 1024) b = 1024;  // the planned kernels' limit; plenty for any detector batch
  return b < 1 ? 1 : (int)b;
}

}  // namespace

extern "C" int ROIAlignForwardLaucher(const float* bottom_data, const float spatial_scale, const int num_rois,
                                      const int height, const int width, const int channels,
                                      const int aligned_height, const int aligned_width, const float* bottom_rois,
                                      float* top_data, cudaStream_t stream) {
  const int batch = max_batch(channels, height, width);
  Scratch plan(tlod_roi_align_plan_bytes(batch, num_rois), stream);
  const void* pl = nullptr;
  size_t pl_bytes = 0;
  if (plan.p && tlod_roi_align_plan(bottom_rois, batch, height, width, num_rois, aligned_height, aligned_width,
                                    spatial_scale, plan.p, tlod_roi_align_plan_bytes(batch, num_rois),
                                    stream) == TLOD_OK) {
    pl = plan.p;
    pl_bytes = tlod_roi_align_plan_bytes(batch, num_rois);
  }
  return tlod_roi_align_forward(bottom_data, bottom_rois, top_data, batch, channels, height, width, num_rois,
                                aligned_height, aligned_width, spatial_scale, pl, pl_bytes, stream) == TLOD_OK;
}

// bottom_diff: the reference accumulates into a buffer its caller zeroed (functions/roi_align.py:42);
// here it is overwritten with the same values.
extern "C" int ROIAlignBackwardLaucher(const float* top_diff, const float spatial_scale, const int batch_size,
                                       const int num_rois, const int height, const int width, const int channels,
                                       const int aligned_height, const int aligned_width,
                                       const float* bottom_rois, float* bottom_diff, cudaStream_t stream) {
  Scratch plan(tlod_roi_align_plan_bytes(batch_size, num_rois), stream);
  const void* pl = nullptr;
  size_t pl_bytes = 0;
  if (plan.p && tlod_roi_align_plan(bottom_rois, batch_size, height, width, num_rois, aligned_height,
                                    aligned_width, spatial_scale, plan.p,
                                    tlod_roi_align_plan_bytes(batch_size, num_rois), stream) == TLOD_OK) {
    pl = plan.p;
    pl_bytes = tlod_roi_align_plan_bytes(batch_size, num_rois);
  }
  return tlod_roi_align_backward(top_diff, bottom_rois, bottom_diff, batch_size, channels, height, width, num_rois,
                                 aligned_height, aligned_width, spatial_scale, pl, pl_bytes, stream) == TLOD_OK;
}

extern "C" int ROIPoolForwardLaucher(const float* bottom_data, const float spatial_scale, const int num_rois,
                                     const int height, const int width, const int channels,
                                     const int pooled_height, const int pooled_width, const float* bottom_rois,
                                     float* top_data, int* argmax_data, cudaStream_t stream) {
  return tlod::roi_pool_forward_launch(bottom_data, bottom_rois, top_data, argmax_data,
                                       max_batch(channels, height, width), channels, height, width, num_rois,
                                       pooled_height, pooled_width, spatial_scale, false, stream) == TLOD_OK;
}

extern "C" int ROIPoolBackwardLaucher(const float* top_diff, const float spatial_scale, const int batch_size,
                                      const int num_rois, const int height, const int width, const int channels,
                                      const int pooled_height, const int pooled_width, const float* bottom_rois,
                                      float* bottom_diff, const int* argmax_data, cudaStream_t stream) {
  return tlod_roi_pool_backward(top_diff, argmax_data, bottom_rois, bottom_diff, batch_size, channels, height, width,
                                num_rois, pooled_height, pooled_width, spatial_scale, stream) == TLOD_OK;
}

// keep_out (boxes_num ints) and num_out (1 int) are DEVICE pointers (nms_cuda.c:12-17 passes
// THCudaIntTensor data); like the reference the call works on the legacy default stream and has
// completed when it returns (nms_gpu.py:11 reads num_out[0] right after).
extern "C" void nms_cuda_compute(int* keep_out, int* num_out, float* boxes_host, int boxes_num, int boxes_dim,
                                 float nms_overlap_thresh) {
  cudaStream_t st = 0;
  if (boxes_num <= 0) {
    cudaMemsetAsync(num_out, 0, sizeof(int), st);
    cudaStreamSynchronize(st);
    return;
  }
  const size_t bytes = tlod_nms_workspace_bytes(boxes_num);
  {
    Scratch ws(bytes, st);
    if (ws.p)
      tlod_nms(boxes_host, boxes_num, boxes_dim, nms_overlap_thresh, 0, keep_out, num_out, ws.p, bytes, st);
  }
  cudaStreamSynchronize(st);
}
