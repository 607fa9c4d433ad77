// Which per-warp store shape streams a (R, C, 8, 8) fp32 tensor to HBM fastest?  (B200)
// Every variant writes the same 512 MB; only the lane -> address mapping differs.
#include <cstdio>
#include <cuda_runtime.h>

// unit = one "RoI slab": 16 channels x 64 floats = 4 KB contiguous
template <int MODE>
__global__ void store_k(float* __restrict__ out, long long units) {
  const int lane = threadIdx.x & 31;
  const long long warp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  const float v = (float)lane;
  for (long long u = warp; u < units; u += nwarps) {
    float* base = out + u * 1024;
    if (MODE == 0) {  // contiguous: 8 x (32 lanes x 16 B)
#pragma unroll
      for (int i = 0; i < 8; ++i) *reinterpret_cast<float4*>(base + i * 128 + lane * 4) = make_float4(v, v, v, v);
    } else if (MODE == 1) {  // lane=(c,slot): 8 rows x [16 lines x 32 B]   (current forward kernel)
      const int c = lane >> 1, s = lane & 1;
#pragma unroll
      for (int ph = 0; ph < 8; ++ph) *reinterpret_cast<float4*>(base + c * 64 + ph * 8 + s * 4) = make_float4(v, v, v, v);
    } else if (MODE == 2) {  // lane=(c,row parity): 2 x STG.128 per row, 32 half-sectors per instr (first kernel)
      const int c = lane >> 1, s = lane & 1;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float* p = base + c * 64 + (2 * j + s) * 8;
        *reinterpret_cast<float4*>(p) = make_float4(v, v, v, v);
        *reinterpret_cast<float4*>(p + 4) = make_float4(v, v, v, v);
      }
    } else if (MODE == 3) {  // lane=(c,row parity): one 256-bit store per row: 16 lines x 64 B per instr
      const int c = lane >> 1, s = lane & 1;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float* p = base + c * 64 + (2 * j + s) * 8;
        asm volatile("st.global.v8.f32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"l"(p), "f"(v) : "memory");
      }
    } else if (MODE == 4) {  // contiguous 256-bit
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float* p = base + i * 256 + lane * 8;
        asm volatile("st.global.v8.f32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"l"(p), "f"(v) : "memory");
      }
    } else if (MODE == 5) {  // scalar, lane = channel-major (16 lines x 8 B per instr): generic path shape
      const int c = lane >> 1, s = lane & 1;
#pragma unroll 8
      for (int i = 0; i < 32; ++i) base[c * 64 + 2 * i + s] = v;
    }
  }
}

int main() {
  const long long units = 131072;  // 512 MB
  float* d; cudaMalloc(&d, units * 4096);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  const char* names[6] = {"STG.128 contiguous 512 B/instr", "STG.128 16 lines x 32 B (lane=(c,slot))", "STG.128 32 half-sectors x 16 B",
                          "STG.256 16 lines x 64 B", "STG.256 contiguous 1 KB/instr", "STG.32  16 lines x 8 B"};
  for (int threads : {512, 1024}) for (int mode = 0; mode < 6; ++mode) {
    auto launch = [&]() {
      const int grid = 148;
      switch (mode) {
        case 0: store_k<0><<<grid, threads>>>(d, units); break;
        case 1: store_k<1><<<grid, threads>>>(d, units); break;
        case 2: store_k<2><<<grid, threads>>>(d, units); break;
        case 3: store_k<3><<<grid, threads>>>(d, units); break;
        case 4: store_k<4><<<grid, threads>>>(d, units); break;
        case 5: store_k<5><<<grid, threads>>>(d, units); break;
      }
    };
    launch(); launch();
    cudaEventRecord(a);
    for (int i = 0; i < 5; ++i) launch();
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); ms /= 5;
    printf("threads %4d  %-45s %.1f us  %.0f GB/s\n", threads, names[mode], ms * 1e3, units * 4096.0 / ms / 1e6);
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
}
