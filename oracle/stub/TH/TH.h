/* Minimal stand-in for torch-0.4's <TH/TH.h>, just enough for the reference's
 * CPU RoI sources (lib/model/roi_align/src/roi_align.c, roi_pooling.c) to
 * compile unmodified into oracle/_ref/libref_cpu.so.  Only the tensor-free
 * functions (ROIAlignForwardCpu ...) are ever called; the TH accessors are
 * declared here and defined as aborting stubs in stub/th_stubs.c. */
#ifndef TLOD_ORACLE_TH_STUB_H
#define TLOD_ORACLE_TH_STUB_H
typedef struct THFloatTensor THFloatTensor;
typedef struct THFloatStorage THFloatStorage;
float* THFloatTensor_data(THFloatTensor* t);
long THFloatTensor_size(THFloatTensor* t, int dim);
THFloatStorage* THFloatTensor_storage(THFloatTensor* t);
void THFloatStorage_fill(THFloatStorage* s, float v);
#endif
