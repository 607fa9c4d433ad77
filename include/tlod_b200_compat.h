/*
 * tlod_b200_compat.h -- the reference's own tensor-free launcher ABI, exported by
 * libtlod_b200.so next to the tlod_* entry points of tlod_b200.h.
 *
 * These are, symbol for symbol and argument for argument, what the reference's cffi glue
 * (lib/model/roi_align/src/roi_align_cuda.c:31,66, lib/model/roi_pooling/src/roi_pooling_cuda.c,
 * lib/model/nms/src/nms_cuda.c:17) calls; the declarations below restate
 *   lib/model/roi_align/src/roi_align_kernel.h:13-27
 *   lib/model/roi_pooling/src/roi_pooling_kernel.h:8-18
 *   lib/model/nms/src/nms_cuda_kernel.h:5-6
 * (the "Laucher" spelling is the reference's).  Linking this library in place of the reference's
 * *_kernel.cu.o objects keeps the glue and every Python call site unchanged.
 *
 * Differences a caller can observe: a failed launch returns 0 instead of calling exit(-1)
 * (roi_align_kernel.cu:84-88); the backward launchers overwrite bottom_diff (the reference adds
 * into a buffer its caller has just zeroed -- same values); scratch comes from the stream-ordered
 * allocator instead of cudaMalloc / cudaFree.
 */
#ifndef TLOD_B200_COMPAT_H
#define TLOD_B200_COMPAT_H

#include <cuda_runtime_api.h>

#ifdef __cplusplus
extern "C" {
#endif

int ROIAlignForwardLaucher(const float* bottom_data, const float spatial_scale, const int num_rois,
                           const int height, const int width, const int channels,
                           const int aligned_height, const int aligned_width,
                           const float* bottom_rois, float* top_data, cudaStream_t stream);
int ROIAlignBackwardLaucher(const float* top_diff, const float spatial_scale, const int batch_size,
                            const int num_rois, const int height, const int width,
                            const int channels, const int aligned_height, const int aligned_width,
                            const float* bottom_rois, float* bottom_diff, cudaStream_t stream);
int ROIPoolForwardLaucher(const float* bottom_data, const float spatial_scale, const int num_rois,
                          const int height, const int width, const int channels,
                          const int pooled_height, const int pooled_width, const float* bottom_rois,
                          float* top_data, int* argmax_data, cudaStream_t stream);
int ROIPoolBackwardLaucher(const float* top_diff, const float spatial_scale, const int batch_size,
                           const int num_rois, const int height, const int width, const int channels,
                           const int pooled_height, const int pooled_width, const float* bottom_rois,
                           float* bottom_diff, const int* argmax_data, cudaStream_t stream);
/* keep_out (boxes_num ints) and num_out (one int) are device pointers; boxes_host, despite its
 * name in the reference, is the DEVICE pointer of the score-sorted (boxes_num, boxes_dim) boxes
 * (nms_cuda.c:12-17).  Blocking, on the legacy default stream, like the reference. */
void nms_cuda_compute(int* keep_out, int* num_out, float* boxes_host, int boxes_num, int boxes_dim,
                      float nms_overlap_thresh);

#ifdef __cplusplus
}
#endif

#endif /* TLOD_B200_COMPAT_H */
