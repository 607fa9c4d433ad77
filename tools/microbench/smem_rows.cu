// How does the cost of a bank-conflict-free 32-bit shared-memory gather depend on the number
// of distinct 128-byte rows the 32 lanes touch?  (B200, sm_100a)
#include <cstdio>
#include <cuda_runtime.h>

template <int VEC>
__global__ void bench(const int* __restrict__ offs, float* out, long long* cyc, int iters) {
  extern __shared__ float sm[];
  for (int i = threadIdx.x; i < 48 * 1024; i += blockDim.x) sm[i] = (float)i;
  __syncthreads();
  const int off = offs[threadIdx.x & 31];
  float a0 = 0, a1 = 0, a2 = 0, a3 = 0;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      const float* p = sm + off + ((it + u) & 7) * 512;
      if (VEC == 1) { a0 += p[0]; }
      if (VEC == 2) { float2 v = *reinterpret_cast<const float2*>(p); a0 += v.x; a1 += v.y; }
      if (VEC == 4) { float4 v = *reinterpret_cast<const float4*>(p); a0 += v.x; a1 += v.y; a2 += v.z; a3 += v.w; }
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
  out[threadIdx.x] = a0 + a1 + a2 + a3;
}

int main() {
  int* d_off; float* d_out; long long* d_cyc;
  cudaMalloc(&d_off, 128); cudaMalloc(&d_out, 4096); cudaMalloc(&d_cyc, 8);
  const int smem = 200 * 1024, threads = 512, iters = 4000;
  cudaFuncSetAttribute(bench<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(bench<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(bench<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  auto run = [&](const char* name, int vec, auto f) {
    int h[32];
    for (int l = 0; l < 32; ++l) h[l] = f(l);
    cudaMemcpy(d_off, h, 128, cudaMemcpyHostToDevice);
    if (vec == 1) bench<1><<<1, threads, smem>>>(d_off, d_out, d_cyc, iters);
    if (vec == 2) bench<2><<<1, threads, smem>>>(d_off, d_out, d_cyc, iters);
    if (vec == 4) bench<4><<<1, threads, smem>>>(d_off, d_out, d_cyc, iters);
    long long c; cudaMemcpy(&c, d_cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-58s vec%d  cycles/instr/SM %.3f   bytes/cycle %.1f\n", name, vec, (double)c / (iters * 16.0 * (threads / 32)),
           32.0 * 4 * vec / ((double)c / (iters * 16.0 * (threads / 32))));
  };
  run("32-bit stride 1 (contiguous)", 1, [](int l) { return l; });
  run("32-bit stride 2 words", 1, [](int l) { return l * 2; });
  run("32-bit stride 4 words", 1, [](int l) { return l * 4; });
  run("32-bit stride 8 words", 1, [](int l) { return l * 8; });
  run("32-bit stride 16 words", 1, [](int l) { return l * 16; });
  run("32-bit stride 32 words", 1, [](int l) { return l * 32; });
  run("32-bit stride 64 words", 1, [](int l) { return l * 64; });
  run("32-bit word = 64*l + (l&1)", 1, [](int l) { return 64 * l + (l & 1); });
  run("32-bit word = 32*l + (l&1)", 1, [](int l) { return 32 * l + (l & 1); });
  run("32-bit word = 32*l + (l&3)", 1, [](int l) { return 32 * l + (l & 3); });
  run("old fwd pattern: c*2850 + (slot?901:750)", 1, [](int l) { return (l >> 1) * 2850 + ((l & 1) ? 901 : 750); });
  run("old fwd pattern: c*2850 + (slot?1051:750)", 1, [](int l) { return (l >> 1) * 2850 + ((l & 1) ? 1051 : 750); });
  run("new fwd pattern: c*2852 + 751 + dx", 1, [](int l) { return (l >> 1) * 2852 + 751 + (l & 1); });
  run("new fwd pattern: c*2852 + 750 + dx", 1, [](int l) { return (l >> 1) * 2852 + 750 + (l & 1); });
  run("c*2850 + 751 + dx", 1, [](int l) { return (l >> 1) * 2850 + 751 + (l & 1); });
  run("c*2850 + 750 + dx", 1, [](int l) { return (l >> 1) * 2850 + 750 + (l & 1); });
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
