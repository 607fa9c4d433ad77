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
// pre_img: NULL, or {mean NLL, mean softmax probability of channel d} of the image head already
// reduced by da_image_loss_fwd_kernel (maps too large for one CTA)
__global__ void __launch_bounds__(DA_THREADS)
    da_loss_fwd_kernel(const float* __restrict__ score, const float* __restrict__ prob,
                       const float* __restrict__ label, int d, float* __restrict__ out, int B,
                       int HW, int R, const float* __restrict__ pre_img) {
  __shared__ double scratch[DA_THREADS / 32];
  double nll = 0.0, psum = 0.0;
  const long long cells = (long long)B * HW;
  for (long long e = threadIdx.x; e < cells && !pre_img; e += DA_THREADS) {
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
  if (pre_img) {
    nll = (double)__ldg(pre_img) * (double)cells;
    psum = (double)__ldg(pre_img + 1) * (double)cells;
  }
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

// ---------------------------------------------------------------------------------------------
// Image-level losses of several feature levels in one launch (MAF / PT-MAF: conv3, conv4, conv5
// heads, lib/MAF/faster_rcnn.py:188-205; ATF: the same with ignore_index = -1,
// lib/ATF/faster_rcnn.py:303-321).  Multi-CTA: per-CTA partial sums in fp64, one atomicAdd(double)
// per CTA and quantity, the last CTA of a level writes its results.
// ---------------------------------------------------------------------------------------------
constexpr int DAL_THREADS = 256;
constexpr int DAL_MAX_CTAS = 64;  // per level

struct DaLevels {
  const float* score[TLOD_DA_MAX_LEVELS];
  const long long* label[TLOD_DA_MAX_LEVELS];  // (B, H, W) int64 or NULL
  float* grad[TLOD_DA_MAX_LEVELS];
  long long cells[TLOD_DA_MAX_LEVELS];
  int hw[TLOD_DA_MAX_LEVELS];
  int first_block[TLOD_DA_MAX_LEVELS + 1];
  float weight[TLOD_DA_MAX_LEVELS];
  int levels;
};

__device__ __forceinline__ int dal_level(const DaLevels& L) {
  int l = 0;
  while (l + 1 < L.levels && (int)blockIdx.x >= L.first_block[l + 1]) ++l;
  return l;
}

__device__ __forceinline__ double dal_block_sum(double v, double* scratch) {
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  for (int w = 0; w < DAL_THREADS / 32; ++w) t += scratch[w];
  return t;
}

// ws: per level {nll sum, prob sum, count} as doubles + an arrival counter (zeroed before the launch)
__global__ void __launch_bounds__(DAL_THREADS)
    da_image_loss_fwd_kernel(DaLevels L, int d, int ignore_index, double* __restrict__ ws,
                             unsigned* __restrict__ arrived, float* __restrict__ out) {
  __shared__ double scratch[DAL_THREADS / 32];
  __shared__ bool last;
  const int l = dal_level(L);
  const int nb = L.first_block[l + 1] - L.first_block[l];
  const int hw = L.hw[l];
  const float* __restrict__ score = L.score[l];
  const long long* __restrict__ label = L.label[l];
  double nll = 0.0, psum = 0.0, cnt = 0.0;
  for (long long e = (long long)(blockIdx.x - L.first_block[l]) * DAL_THREADS + threadIdx.x; e < L.cells[l];
       e += (long long)nb * DAL_THREADS) {
    const long long b = e / hw, c = e - b * hw;
    long long y = d;
    if (label) {
      y = __ldg(label + e);
      if (y == ignore_index) continue;
    }
    const float s0 = __ldg(score + (b * 2) * hw + c), s1 = __ldg(score + (b * 2 + 1) * hw + c);
    const float m = fmaxf(s0, s1);
    const float lse = m + logf(expf(s0 - m) + expf(s1 - m));
    nll += (double)(lse - (y ? s1 : s0));
    psum += (double)expf((d ? s1 : s0) - lse);
    cnt += 1.0;
  }
  nll = dal_block_sum(nll, scratch);
  psum = dal_block_sum(psum, scratch);
  cnt = dal_block_sum(cnt, scratch);
  if (threadIdx.x == 0) {
    atomicAdd(ws + l * 3 + 0, nll);
    atomicAdd(ws + l * 3 + 1, psum);
    atomicAdd(ws + l * 3 + 2, cnt);
    __threadfence();
    last = atomicAdd(arrived + l, 1u) == (unsigned)(nb - 1);
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {
    __threadfence();
    const double n = atomicAdd(ws + l * 3 + 2, 0.0);
    out[l * 4 + 0] = (float)(atomicAdd(ws + l * 3 + 0, 0.0) / n);  // 0 / 0 = NaN, like torch with nothing counted
    out[l * 4 + 1] = n > 0.0 ? (float)(atomicAdd(ws + l * 3 + 1, 0.0) / n) : 0.f;
    out[l * 4 + 2] = (float)n;
    out[l * 4 + 3] = 0.f;
  }
}

__global__ void __launch_bounds__(DAL_THREADS)
    da_image_loss_bwd_kernel(DaLevels L, int d, int ignore_index, const float* __restrict__ out,
                             const float* __restrict__ upstream) {
  const int l = dal_level(L);
  const long long e = (long long)(blockIdx.x - L.first_block[l]) * DAL_THREADS + threadIdx.x;
  if (e >= L.cells[l]) return;
  const int hw = L.hw[l];
  const long long b = e / hw, c = e - b * hw;
  float* __restrict__ g = L.grad[l];
  long long y = d;
  if (L.label[l]) y = __ldg(L.label[l] + e);
  float g0 = 0.f, g1 = 0.f;
  if (!(L.label[l] && y == ignore_index)) {
    const float* __restrict__ score = L.score[l];
    const float s0 = __ldg(score + (b * 2) * hw + c), s1 = __ldg(score + (b * 2 + 1) * hw + c);
    const float m = fmaxf(s0, s1);
    const float e0 = expf(s0 - m), e1 = expf(s1 - m);
    const float inv = 1.f / (e0 + e1);
    float k = L.weight[l] / __ldg(out + l * 4 + 2);
    if (upstream) k *= __ldg(upstream + l);
    g0 = k * (e0 * inv - (y == 0 ? 1.f : 0.f));
    g1 = k * (e1 * inv - (y == 1 ? 1.f : 0.f));
  }
  g[(b * 2) * hw + c] = g0;
  g[(b * 2 + 1) * hw + c] = g1;
}

static int dal_fill(DaLevels& L, int levels, const float* const* scores, const long long* const* labels,
                    const int* batch, const int* height, const int* width, bool backward) {
  if (levels < 1 || levels > TLOD_DA_MAX_LEVELS) return TLOD_ERR_BAD_SHAPE;
  if (!scores || !batch || !height || !width) return TLOD_ERR_NULL_POINTER;
  L.levels = levels;
  int blocks = 0;
  for (int l = 0; l < levels; ++l) {
    if (batch[l] < 0 || height[l] < 0 || width[l] < 0) return TLOD_ERR_BAD_SHAPE;
    const long long hw = (long long)height[l] * width[l];
    const long long cells = (long long)batch[l] * hw;
    if (cells * 2 >= (1LL << 31)) return TLOD_ERR_INT32_OVERFLOW;
    if (cells > 0 && !scores[l]) return TLOD_ERR_NULL_POINTER;
    L.score[l] = scores[l];
    L.label[l] = labels ? labels[l] : nullptr;
    L.grad[l] = nullptr;
    L.cells[l] = cells;
    L.hw[l] = (int)(hw > 0 ? hw : 1);
    L.weight[l] = 1.f;
    L.first_block[l] = blocks;
    long long nb = (cells + DAL_THREADS - 1) / DAL_THREADS;
    if (!backward) nb = (cells + DAL_THREADS * 8 - 1) / (DAL_THREADS * 8);  // ~8 cells per thread
    if (nb < 1) nb = 1;
    if (!backward && nb > DAL_MAX_CTAS) nb = DAL_MAX_CTAS;
    blocks += (int)nb;
  }
  L.first_block[levels] = blocks;
  return TLOD_OK;
}

}  // namespace tlod

using namespace tlod;

extern "C" size_t tlod_da_image_loss_workspace_bytes(void) {
  return TLOD_DA_MAX_LEVELS * (3 * sizeof(double) + sizeof(unsigned));
}

extern "C" int tlod_da_image_loss_forward(int levels, const float* const* h_scores,
                                          const long long* const* h_labels, const int* h_batch,
                                          const int* h_height, const int* h_width, int domain_label,
                                          int ignore_index, float* out, void* workspace,
                                          size_t workspace_bytes, void* stream) {
  DaLevels L;
  int rc = dal_fill(L, levels, h_scores, h_labels, h_batch, h_height, h_width, false);
  if (rc != TLOD_OK) return rc;
  if (!out) return TLOD_ERR_NULL_POINTER;
  if (domain_label != 0 && domain_label != 1) return TLOD_ERR_BAD_SHAPE;
  if (!workspace || workspace_bytes < tlod_da_image_loss_workspace_bytes() || ((uintptr_t)workspace & 7))
    return TLOD_ERR_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(workspace, 0, tlod_da_image_loss_workspace_bytes(), st);
  if (e != cudaSuccess) return (int)e;
  double* ws = (double*)workspace;
  unsigned* arrived = (unsigned*)(ws + TLOD_DA_MAX_LEVELS * 3);
  {
    LaunchScope scope("da_image_loss_fwd_kernel", st);
    da_image_loss_fwd_kernel<<<L.first_block[levels], DAL_THREADS, 0, st>>>(L, domain_label, ignore_index, ws,
                                                                            arrived, out);
  }
  return last_launch_status();
}

extern "C" int tlod_da_image_loss_backward(int levels, const float* const* h_scores,
                                           const long long* const* h_labels, const int* h_batch,
                                           const int* h_height, const int* h_width, int domain_label,
                                           int ignore_index, const float* out, const float* upstream,
                                           const float* h_weights, float* const* h_grad_scores,
                                           void* stream) {
  DaLevels L;
  int rc = dal_fill(L, levels, h_scores, h_labels, h_batch, h_height, h_width, true);
  if (rc != TLOD_OK) return rc;
  if (!out || !h_grad_scores) return TLOD_ERR_NULL_POINTER;
  if (domain_label != 0 && domain_label != 1) return TLOD_ERR_BAD_SHAPE;
  bool any = false;
  for (int l = 0; l < levels; ++l) {
    if (L.cells[l] > 0 && !h_grad_scores[l]) return TLOD_ERR_NULL_POINTER;
    L.grad[l] = h_grad_scores[l];
    if (h_weights) L.weight[l] = h_weights[l];
    any = any || L.cells[l] > 0;
  }
  if (!any) return TLOD_OK;
  {
    LaunchScope scope("da_image_loss_bwd_kernel", (cudaStream_t)stream);
    da_image_loss_bwd_kernel<<<L.first_block[levels], DAL_THREADS, 0, (cudaStream_t)stream>>>(
        L, domain_label, ignore_index, out, upstream);
  }
  return last_launch_status();
}


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

// image-head partials of da_image_loss_fwd_kernel, then its 4 output floats
extern "C" size_t tlod_da_loss_workspace_bytes(void) { return tlod_da_image_loss_workspace_bytes() + 16; }

// maps above this many cells reduce the image head over many CTAs first (needs the workspace)
static const long long DA_SINGLE_CTA_CELLS = 1 << 16;

extern "C" int tlod_da_loss_forward(const float* img_score, const float* ins_prob,
                                    const float* ins_label, int domain_label, float* losses_out,
                                    int batch, int height, int width, int num_ins, void* workspace,
                                    size_t workspace_bytes, void* stream) {
  if (!losses_out || (batch > 0 && !img_score) || (num_ins > 0 && !ins_prob)) return TLOD_ERR_NULL_POINTER;
  if (batch < 0 || height < 0 || width < 0 || num_ins < 0 || (domain_label != 0 && domain_label != 1))
    return TLOD_ERR_BAD_SHAPE;
  const float* pre_img = nullptr;
  if ((long long)batch * height * width > DA_SINGLE_CTA_CELLS && workspace &&
      workspace_bytes >= tlod_da_loss_workspace_bytes()) {
    float* img_out = (float*)((char*)workspace + tlod_da_image_loss_workspace_bytes());
    int rc = tlod_da_image_loss_forward(1, &img_score, nullptr, &batch, &height, &width, domain_label,
                                        -100, img_out, workspace, tlod_da_image_loss_workspace_bytes(),
                                        stream);
    if (rc != TLOD_OK) return rc;
    pre_img = img_out;
  }
  {
    LaunchScope scope("da_loss_fwd_kernel", (cudaStream_t)stream);
    da_loss_fwd_kernel<<<1, DA_THREADS, 0, (cudaStream_t)stream>>>(
        img_score, ins_prob, ins_label, domain_label, losses_out, batch, height * width, num_ins,
        pre_img);
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
