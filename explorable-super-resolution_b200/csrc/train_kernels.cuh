// Training-step kernels outside the networks: fused multi-tensor Adam, L1 loss (pixel / VGG-feature criterion) and the relativistic
// average BCE of the GAN terms.  All HBM-bound (Adam: 16 B read + 12 B written per parameter; L1: 8 B read per element forward,
// 8 B read + 4 B written backward) or launch-latency bound (the [B,1] logits of the GAN terms).
// Reference: models/SRRaGAN_model.py:182,188,403,499 (torch.optim.Adam x2), :98,129,434,448-451 (nn.L1Loss), :353-354,475-476 with
// models/modules/loss.py:212-246 (GANLoss 'vanilla' = BCEWithLogitsLoss on D(real) - mean(D(fake)) and vice versa).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace esr {

struct AdamTensor {
  float* p; const float* g; float* m; float* v; unsigned long long n;
};

// torch.optim.Adam (amsgrad off, maximize off), fp32, the exact operation order of torch/optim/adam.py::_single_tensor_adam:
//   g' = g * grad_scale (+ wd * p);  m += (g' - m) * (1 - b1);  v = v * b2 + (1 - b2) * g' * g';
//   p -= step_size * m / (sqrt(v) / bc2_sqrt + eps),   step_size = lr / (1 - b1^t),  bc2_sqrt = sqrt(1 - b2^t)
// grid = (blocks per tensor, tensors); one launch per optimizer step.
__global__ void adam_multi_kernel(const AdamTensor* __restrict__ tab, float b1, float b2, float eps, float wd, float step_size, float bc2_sqrt,
                                  float grad_scale) {
  const AdamTensor t = tab[blockIdx.y];
  const size_t n = (size_t)t.n;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const bool vec = ((((uintptr_t)t.p | (uintptr_t)t.g | (uintptr_t)t.m | (uintptr_t)t.v) & 15) == 0);
  const size_t n4 = vec ? n / 4 : 0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 p = reinterpret_cast<float4*>(t.p)[i];
    const float4 g4 = __ldg(reinterpret_cast<const float4*>(t.g) + i);
    float4 m = reinterpret_cast<float4*>(t.m)[i], v = reinterpret_cast<float4*>(t.v)[i];
    float* pp = &p.x; float* mm = &m.x; float* vv = &v.x; const float* gg = &g4.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float g = gg[k] * grad_scale;
      if (wd != 0.f) g = fmaf(wd, pp[k], g);
      mm[k] = mm[k] + (g - mm[k]) * (1.f - b1);
      vv[k] = vv[k] * b2 + (1.f - b2) * g * g;
      pp[k] = pp[k] - step_size * (mm[k] / (sqrtf(vv[k]) / bc2_sqrt + eps));
    }
    reinterpret_cast<float4*>(t.p)[i] = p;
    reinterpret_cast<float4*>(t.m)[i] = m;
    reinterpret_cast<float4*>(t.v)[i] = v;
  }
  for (size_t i = n4 * 4 + blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += stride) {
    float g = t.g[i] * grad_scale;
    if (wd != 0.f) g = fmaf(wd, t.p[i], g);
    const float m = t.m[i] + (g - t.m[i]) * (1.f - b1);
    const float v = t.v[i] * b2 + (1.f - b2) * g * g;
    t.m[i] = m; t.v[i] = v;
    t.p[i] = t.p[i] - step_size * (m / (sqrtf(v) / bc2_sqrt + eps));
  }
}

// ---- L1 (mean absolute difference) ---------------------------------------------------------------------------------------------
constexpr int kL1Blocks = 592;   // 148 SMs x 4

__device__ __forceinline__ float block_sum_256(float s, float* sm) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) sm[warp] = s;
  __syncthreads();
  float t = 0.f;
  if (threadIdx.x < 8) t = sm[threadIdx.x];
  if (threadIdx.x < 32) {
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  }
  return t;   // valid in thread 0
}

// partial[b] = sum over the block's grid-stride share of |a - b|
__global__ void l1_partial_kernel(const float* __restrict__ a, const float* __restrict__ b, size_t n, float* __restrict__ partial) {
  __shared__ float sm[8];
  float s = 0.f;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const bool vec = ((((uintptr_t)a | (uintptr_t)b) & 15) == 0);
  const size_t n4 = vec ? n / 4 : 0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 x = __ldg(reinterpret_cast<const float4*>(a) + i), y = __ldg(reinterpret_cast<const float4*>(b) + i);
    s += fabsf(x.x - y.x) + fabsf(x.y - y.y) + fabsf(x.z - y.z) + fabsf(x.w - y.w);
  }
  for (size_t i = n4 * 4 + blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += stride) s += fabsf(__ldg(a + i) - __ldg(b + i));
  const float t = block_sum_256(s, sm);
  if (threadIdx.x == 0) partial[blockIdx.x] = t;
}
// out[0] = scale * sum(partial)  (double combine, deterministic)
__global__ void sum_partials_kernel(const float* __restrict__ partial, int count, double scale, float* __restrict__ out) {
  __shared__ double sm[32];
  double s = 0.0;
  for (int i = threadIdx.x; i < count; i += blockDim.x) s += (double)partial[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sm[w];
    out[0] = (float)(t * scale);
  }
}
// ga = sign(a - b) * gout[0] * scale  (torch's l1_loss backward: sign(0) = 0); gb (optional) = -ga
__global__ void l1_grad_kernel(const float* __restrict__ a, const float* __restrict__ b, size_t n, const float* __restrict__ gout, float scale,
                               float* __restrict__ ga, float* __restrict__ gb) {
  const float g = __ldg(gout) * scale;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += stride) {
    const float d = __ldg(a + i) - __ldg(b + i);
    const float s = d > 0.f ? g : (d < 0.f ? -g : 0.f);
    if (ga) ga[i] = s;
    if (gb) gb[i] = -s;
  }
}

// ---- relativistic average BCE on logits -------------------------------------------------------------------------------------------
//   la = mean_i bce(a_i - mean(b), ta),  lb = mean_i bce(b_i - mean(a), tb),  bce(x, t) = max(x, 0) - x t + log1p(exp(-|x|))
// One block.  sums[0..1] = sum of a / sum of b over the GLOBAL batch of n_global samples (the caller all-reduces them when the batch
// is sharded over ranks; NULL = compute them here from the local values).  out[0..3] = local sum of bce_a, of bce_b, S_a = sum_i e_a[i],
// S_b = sum_i e_b[i] with e_a[i] = sigmoid(a_i - mean_b) - ta (the derivative of bce), kept in ea / eb for the backward.
__global__ void bce_rel_fwd_kernel(const float* __restrict__ a, const float* __restrict__ b, int n, const float* __restrict__ sums, float n_global,
                                   float ta, float tb, float* __restrict__ out, float* __restrict__ ea, float* __restrict__ eb) {
  __shared__ float sm[8];
  __shared__ float means[2];
  if (sums) {
    if (threadIdx.x == 0) { means[0] = sums[0] / n_global; means[1] = sums[1] / n_global; }
    __syncthreads();
  } else {
    float sa = 0.f, sb = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) { sa += a[i]; sb += b[i]; }
    const float ta_ = block_sum_256(sa, sm);
    __syncthreads();
    const float tb_ = block_sum_256(sb, sm);
    if (threadIdx.x == 0) { means[0] = ta_ / (float)n; means[1] = tb_ / (float)n; }
    __syncthreads();
  }
  const float ma = means[0], mb = means[1];
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float xa = a[i] - mb, xb = b[i] - ma;
    acc[0] += fmaxf(xa, 0.f) - xa * ta + log1pf(expf(-fabsf(xa)));
    acc[1] += fmaxf(xb, 0.f) - xb * tb + log1pf(expf(-fabsf(xb)));
    const float da = 1.f / (1.f + expf(-xa)) - ta, db = 1.f / (1.f + expf(-xb)) - tb;
    ea[i] = da; eb[i] = db;
    acc[2] += da; acc[3] += db;
  }
  for (int k = 0; k < 4; ++k) {
    __syncthreads();
    const float t = block_sum_256(acc[k], sm);
    if (threadIdx.x == 0) out[k] = t;
  }
}
// gradients of (ga_up * la + gb_up * lb) with respect to the logits, la / lb being the LOCAL means (rank-mean losses whose parameter
// gradients are averaged over ranks afterwards reproduce the global-batch gradient when the shards are equal):
//   d/da_i = ga_up * ea[i] / n - gb_up * S_b / (n_global * n)      d/db_i = gb_up * eb[i] / n - ga_up * S_a / (n_global * n)
// S = [S_a, S_b] over the GLOBAL batch.  detach_a / detach_b: that input is a constant (the generator step detaches the real logits).
__global__ void bce_rel_bwd_kernel(const float* __restrict__ ea, const float* __restrict__ eb, int n, const float* __restrict__ S, float n_global,
                                   const float* __restrict__ ga_up, const float* __restrict__ gb_up, float* __restrict__ ga, float* __restrict__ gb) {
  const float gu = ga_up ? ga_up[0] : 0.f, gv = gb_up ? gb_up[0] : 0.f;
  const float inv = 1.f / (float)n, cross = 1.f / (n_global * (float)n);
  for (int i = threadIdx.x + blockIdx.x * blockDim.x; i < n; i += gridDim.x * blockDim.x) {
    if (ga) ga[i] = gu * ea[i] * inv - gv * S[1] * cross;
    if (gb) gb[i] = gv * eb[i] * inv - gu * S[0] * cross;
  }
}

}  // namespace esr
