// Gradient reversal and the domain-classifier loss reductions for sm_100a.
//
// Reference semantics:
//   GRLayer.backward        lib/DAF/DA.py:19-30   (grad.neg() * alpha, two launches)
//   weighted GRL            lib/MAF/DA.py:34-53
//   image / instance / consistency losses   lib/DAF/faster_rcnn.py:181-220
//     img = F.nll_loss(F.log_softmax(score, 1), label)        (mean over B*H*W)
//     ins = nn.BCELoss()(sigmoid_out, label)                  (mean over R, log clamped at -100)
//     cst = MSELoss(size_average=False)(sigmoid_out, mean(softmax(score,1)[:, d]))   (sum)
// The reference spends ~15 launches and four D2H copies per head on these
// (LabelResizeLayer.py:28-29, :51-52); here it is one launch and no copy: the
// label map is the domain scalar, so it is never materialised.
#include "common.cuh"

namespace tlod {

__global__ void __launch_bounds__(256)
    grl_kernel(const float* __restrict__ g, float* __restrict__ out, float neg_alpha, long long n) {
  const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i + 3 < n && (((uintptr_t)(g + i) | (uintptr_t)(out + i)) & 15) == 0) {
    float4 v = __ldg(reinterpret_cast<const float4*>(g + i));
    v.x *= neg_alpha; v.y *= neg_alpha; v.z *= neg_alpha; v.w *= neg_alpha;
    *reinterpret_cast<float4*>(out + i) = v;
  } else {
    for (long long j = i; j < n && j < i + 4; ++j) out[j] = __ldg(g + j) * neg_alpha;
  }
}

__global__ void __launch_bounds__(256)
    grl_weighted_kernel(const float* __restrict__ g, const float* __restrict__ w,
                        float* __restrict__ out, float neg_alpha, int rows, int cols) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (long long)rows * cols) return;
  const int r = (int)(e / cols);
  out[e] = __ldg(g + e) * (neg_alpha * __ldg(w + r));
}

constexpr int DA_THREADS = 1024;

__device__ inline double block_sum(double v, double* scratch) {
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  __syncthreads();  // protect scratch reuse
  if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  for (int w = 0; w < DA_THREADS / 32; ++w) t += scratch[w];
  return t;
}

// single CTA: the tensors are a few thousand elements (B*2*H*W ~ 22 K, R ~ 10^3)
__global__ void __launch_bounds__(DA_THREADS)
    da_loss_fwd_kernel(const float* __restrict__ score, const float* __restrict__ prob,
                       const float* __restrict__ label, int d, float* __restrict__ out, int B,
                       int HW, int R) {
  __shared__ double scratch[DA_THREADS / 32];
  double nll = 0.0, psum = 0.0;
  const long long cells = (long long)B * HW;
  for (long long e = threadIdx.x; e < cells; e += DA_THREADS) {
    const long long b = e / HW, c = e - b * HW;
    const float s0 = __ldg(score + (b * 2) * HW + c), s1 = __ldg(score + (b * 2 + 1) * HW + c);
    const float m = fmaxf(s0, s1);
    const float lse = m + logf(expf(s0 - m) + expf(s1 - m));
    const float sd = d ? s1 : s0;
    nll += (double)(lse - sd);
    psum += (double)expf(sd - lse);
  }
  nll = block_sum(nll, scratch);
  psum = block_sum(psum, scratch);
  const float cons = cells > 0 ? (float)(psum / (double)cells) : 0.f;
  double bce = 0.0, mse = 0.0;
  for (int r = threadIdx.x; r < R; r += DA_THREADS) {
    const float p = __ldg(prob + r);
    const float y = label ? __ldg(label + r) : (float)d;
    const float lp = fmaxf(logf(p), -100.f), l1p = fmaxf(log1pf(-p), -100.f);
    bce -= (double)(y * lp + (1.f - y) * l1p);
    const float df = p - cons;
    mse += (double)(df * df);
  }
  bce = block_sum(bce, scratch);
  mse = block_sum(mse, scratch);
  if (threadIdx.x == 0) {
    out[0] = cells > 0 ? (float)(nll / (double)cells) : 0.f;
    out[1] = R > 0 ? (float)(bce / (double)R) : 0.f;
    out[2] = (float)mse;
    out[3] = cons;
  }
}

__global__ void __launch_bounds__(256)
    da_loss_bwd_kernel(const float* __restrict__ score, const float* __restrict__ prob,
                       const float* __restrict__ label, int d, const float* __restrict__ losses,
                       const float* __restrict__ upstream, float w_img, float w_ins, float w_cst,
                       float* __restrict__ g_score, float* __restrict__ g_prob, int B, int HW, int R) {
  if (upstream) {
    w_img *= __ldg(upstream);
    w_ins *= __ldg(upstream + 1);
    w_cst *= __ldg(upstream + 2);
  }
  const long long cells = (long long)B * HW;
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < cells) {
    const long long b = e / HW, c = e - b * HW;
    const float s0 = __ldg(score + (b * 2) * HW + c), s1 = __ldg(score + (b * 2 + 1) * HW + c);
    const float m = fmaxf(s0, s1);
    const float e0 = expf(s0 - m), e1 = expf(s1 - m);
    const float inv = 1.f / (e0 + e1);
    const float k = w_img / (float)cells;
    g_score[(b * 2) * HW + c] = k * (e0 * inv - (d == 0 ? 1.f : 0.f));
    g_score[(b * 2 + 1) * HW + c] = k * (e1 * inv - (d == 1 ? 1.f : 0.f));
  }
  if (e < R) {
    const float p = __ldg(prob + e);
    const float y = label ? __ldg(label + e) : (float)d;
    const float cons = __ldg(losses + 3);
    // torch binary_cross_entropy backward: (p - y) / max((1 - p) * p, 1e-12) / R
    float g = w_ins * (p - y) / fmaxf((1.f - p) * p, 1e-12f) / (float)R;
    g += w_cst * 2.f * (p - cons);
    g_prob[e] = g;
  }
}

}  // namespace tlod

using namespace tlod;

extern "C" int tlod_grl_backward(const float* grad, float* out, float alpha, long long n,
                                 void* stream) {
  if (!grad || !out) return TLOD_ERR_NULL_POINTER;
  if (n < 0) return TLOD_ERR_BAD_SHAPE;
  if (n == 0) return TLOD_OK;
  const long long threads = (n + 3) / 4;
  {
    LaunchScope scope("grl_kernel", (cudaStream_t)stream);
    grl_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(grad, out, -alpha, n);
  }
  return last_launch_status();
}

extern "C" int tlod_grl_backward_weighted(const float* grad, const float* row_weight, float* out,
                                          float alpha, int rows, int cols, void* stream) {
  if (!grad || !row_weight || !out) return TLOD_ERR_NULL_POINTER;
  if (rows < 0 || cols < 0) return TLOD_ERR_BAD_SHAPE;
  const long long n = (long long)rows * cols;
  if (n == 0) return TLOD_OK;
  {
    LaunchScope scope("grl_weighted_kernel", (cudaStream_t)stream);
    grl_weighted_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        grad, row_weight, out, -alpha, rows, cols);
  }
  return last_launch_status();
}

extern "C" size_t tlod_da_loss_workspace_bytes(void) { return 16; }

extern "C" int tlod_da_loss_forward(const float* img_score, const float* ins_prob,
                                    const float* ins_label, int domain_label, float* losses_out,
                                    int batch, int height, int width, int num_ins, void* workspace,
                                    size_t workspace_bytes, void* stream) {
  (void)workspace;
  (void)workspace_bytes;
  if (!losses_out || (batch > 0 && !img_score) || (num_ins > 0 && !ins_prob)) return TLOD_ERR_NULL_POINTER;
  if (batch < 0 || height < 0 || width < 0 || num_ins < 0 || (domain_label != 0 && domain_label != 1))
    return TLOD_ERR_BAD_SHAPE;
  {
    LaunchScope scope("da_loss_fwd_kernel", (cudaStream_t)stream);
    da_loss_fwd_kernel<<<1, DA_THREADS, 0, (cudaStream_t)stream>>>(
        img_score, ins_prob, ins_label, domain_label, losses_out, batch, height * width, num_ins);
  }
  return last_launch_status();
}

extern "C" int tlod_da_loss_backward(const float* img_score, const float* ins_prob,
                                     const float* ins_label, int domain_label,
                                     const float* losses_out, const float* upstream, float w_img,
                                     float w_ins, float w_cst, float* grad_img_score,
                                     float* grad_ins_prob, int batch, int height, int width,
                                     int num_ins, void* stream) {
  if (!losses_out || (batch > 0 && (!img_score || !grad_img_score)) ||
      (num_ins > 0 && (!ins_prob || !grad_ins_prob)))
    return TLOD_ERR_NULL_POINTER;
  if (batch < 0 || height < 0 || width < 0 || num_ins < 0 || (domain_label != 0 && domain_label != 1))
    return TLOD_ERR_BAD_SHAPE;
  const long long cells = (long long)batch * height * width;
  const long long n = cells > num_ins ? cells : num_ins;
  if (n == 0) return TLOD_OK;
  {
    LaunchScope scope("da_loss_bwd_kernel", (cudaStream_t)stream);
    da_loss_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        img_score, ins_prob, ins_label, domain_label, losses_out, upstream, w_img, w_ins, w_cst,
        grad_img_score, grad_ins_prob, batch, height * width, num_ins);
  }
  return last_launch_status();
}
