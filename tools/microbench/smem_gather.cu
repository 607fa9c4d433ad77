// Microbenchmark: shared-memory gather throughput for the lane = 2*channel + slot patterns
// used by the plane-resident RoIAlign kernels.  Prints cycles per warp-wide LDS.
#include <cstdio>
#include <cuda_runtime.h>

__global__ void bench(const int* __restrict__ offs, float* out, long long* cyc, int iters, int step) {
  extern __shared__ float sm[];
  for (int i = threadIdx.x; i < 45600 + 4096; i += blockDim.x) sm[i] = (float)i;
  __syncthreads();
  int off = offs[threadIdx.x & 31];
  float acc = 0.f;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 16; ++u) acc += sm[off + u * step];
    off = (off + 7) % 64 + offs[threadIdx.x & 31];
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

int main() {
  const int P = 2850;
  struct Pat { const char* name; int off[32]; int step; } pats[8];
  int np = 0;
  auto add = [&](const char* n, int step, auto f) { pats[np].name = n; pats[np].step = step; for (int l = 0; l < 32; ++l) pats[np].off[l] = f(l); ++np; };
  add("contiguous lane", 32, [](int l) { return l; });
  add("16ch x (x, x+1)", 2, [&](int l) { return (l >> 1) * P + (l & 1); });
  add("16ch x rows 750/901 (opposite parity)", 2, [&](int l) { return (l >> 1) * P + ((l & 1) ? 901 : 750); });
  add("16ch x rows 750/900 (same parity: 2-way)", 2, [&](int l) { return (l >> 1) * P + ((l & 1) ? 900 : 750); });
  add("32 lanes x stride 2850 (2-way)", 1, [&](int l) { return l * 1425; });
  add("16ch x (x, x+17)", 2, [&](int l) { return (l >> 1) * P + (l & 1) * 17; });
  add("all lanes same address (broadcast)", 1, [](int l) { return 5; });
  add("16ch stride 2851 (odd) x (x,x+1)", 2, [&](int l) { return (l >> 1) * 2851 + (l & 1); });
  int* d_off; float* d_out; long long* d_cyc;
  cudaMalloc(&d_off, 32 * 4); cudaMalloc(&d_out, 148 * 1024 * 4); cudaMalloc(&d_cyc, 148 * 8);
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int threads : {32, 512, 1024}) {
    for (int p = 0; p < np; ++p) {
      cudaMemcpy(d_off, pats[p].off, 128, cudaMemcpyHostToDevice);
      const int iters = 2000;
      bench<<<1, threads, 200 * 1024>>>(d_off, d_out, d_cyc, iters, pats[p].step);
      long long c; cudaMemcpy(&c, d_cyc, 8, cudaMemcpyDeviceToHost);
      double per_lds_sm = (double)c / (iters * 16.0 * (threads / 32));
      printf("threads %4d  %-45s cycles/LDS/SM %.3f\n", threads, pats[p].name, per_lds_sm);
    }
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("%s\n", cudaGetErrorString(e));
  return 0;
}
