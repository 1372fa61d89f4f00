// Discriminator_VGG_128 helper kernels (models/modules/architecture.py:446-508): BatchNorm2d with batch statistics fused
// with LeakyReLU on planar-8 tensors, space-to-depth addressing for the 4x4 stride-2 convolutions (which run as 3x3
// convolutions over the 2x2 space-to-depth image), and the two fully connected layers.  All HBM-bound, read-once/write-once.
#pragma once
#include "aux_kernels.cuh"

namespace esr {

constexpr int kBnMaxChunks = 64;   // partial sums per plane (workspace = planes * kBnMaxChunks * 16 floats)

// offset (in floats) of channel group g of pixel (img, y, x) in a gradient tensor.
//   layout 0: fp32 planes [n][P][h][w][8]
//   layout 1: fp32 planes of the space-to-depth image [n][4P][h/2][w/2][8], plane (py*2+px)*P + g, py = y&1, px = x&1
//             (what the transposed 3x3 conv of a 4x4 stride-2 conv produces)
//   layout 2: NCHW fp32 [n][C][h][w] (returns the offset of channel g*8; channel stride = h*w)
__device__ __forceinline__ size_t grad_offset(int layout, int img, int g, int y, int x, int P, int h, int w, int C) {
  if (layout == 0) return ((((size_t)img * P + g) * h + y) * w + x) * 8;
  if (layout == 1) {
    const int h2 = h >> 1, w2 = w >> 1;
    const int pl = ((y & 1) * 2 + (x & 1)) * P + g;
    return ((((size_t)img * 4 * P + pl) * h2 + (y >> 1)) * w2 + (x >> 1)) * 8;
  }
  return (((size_t)img * C + g * 8) * h + y) * w + x;
}

__device__ __forceinline__ void load8(const float* p, float* v) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}

__device__ __forceinline__ void load_grad8(const float* g, int layout, int img, int gi, int y, int x, int P, int h, int w, int C, float* v) {
  const size_t o = grad_offset(layout, img, gi, y, x, P, h, w, C);
  if (layout == 2) {
    const size_t hw = (size_t)h * w;
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = (gi * 8 + k < C) ? __ldg(g + o + k * hw) : 0.f;
  } else {
    load8(g + o, v);
  }
}

// block-level sum of 16 per-thread values -> part[0..15] (thread 0..15 write)
__device__ __forceinline__ void block_sum16(float* acc, float* part) {
  __shared__ float sm[8][16];
#pragma unroll
  for (int k = 0; k < 16; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < 16; ++k) sm[warp][k] = acc[k];
  }
  __syncthreads();
  if (threadIdx.x < 16) {
    float s = 0.f;
    for (int wv = 0; wv < (int)(blockDim.x >> 5); ++wv) s += sm[wv][threadIdx.x];
    part[threadIdx.x] = s;
  }
}

// ---- forward ---------------------------------------------------------------------------------------------------------------------
// per-(plane, chunk) partial sums of y and y^2 over (n, h, w).  grid (chunks, P), 256 threads.
__global__ void bn_partial_kernel(const float* __restrict__ y, int n, int P, int hw, float* __restrict__ part, int chunks) {
  const int g = blockIdx.y, chunk = blockIdx.x;
  const size_t M = (size_t)n * hw;
  const size_t per = (M + chunks - 1) / chunks;
  const size_t m0 = (size_t)chunk * per, m1 = m0 + per < M ? m0 + per : M;
  float acc[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) acc[k] = 0.f;
  for (size_t m = m0 + threadIdx.x; m < m1; m += blockDim.x) {
    const size_t img = m / hw, pix = m - img * hw;
    float v[8];
    load8(y + ((img * P + g) * hw + pix) * 8, v);
#pragma unroll
    for (int k = 0; k < 8; ++k) { acc[k] += v[k]; acc[8 + k] = fmaf(v[k], v[k], acc[8 + k]); }
  }
  block_sum16(acc, part + ((size_t)g * chunks + chunk) * 16);
}

// one thread per channel: batch mean / biased variance (double combine of the partials), the affine the apply kernel uses
// (scale = gamma * invstd, shift = beta - mean * scale) and the running-statistics update of nn.BatchNorm2d
// (momentum 0.1, unbiased variance).  train == 0: statistics come from running_mean / running_var.
__global__ void bn_finalize_kernel(const float* __restrict__ part, int chunks, int C, double M, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float eps, float momentum, int train, float* running_mean,
                                   float* running_var, float* __restrict__ save_mean, float* __restrict__ save_invstd,
                                   float* __restrict__ scale, float* __restrict__ shift) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double mean, var;
  if (train) {
    double s = 0.0, q = 0.0;
    const float* pp = part + (size_t)(c >> 3) * chunks * 16 + (c & 7);
    for (int k = 0; k < chunks; ++k) { s += pp[k * 16]; q += pp[k * 16 + 8]; }
    mean = s / M;
    var = q / M - mean * mean;
    if (var < 0.0) var = 0.0;
    if (running_mean) running_mean[c] = (float)((1.0 - momentum) * running_mean[c] + momentum * mean);
    if (running_var) running_var[c] = (float)((1.0 - momentum) * running_var[c] + momentum * var * (M > 1.0 ? M / (M - 1.0) : 1.0));
  } else {
    mean = running_mean[c];
    var = running_var[c];
  }
  const double invstd = 1.0 / sqrt(var + (double)eps);
  const double sc = (double)gamma[c] * invstd;
  save_mean[c] = (float)mean;
  save_invstd[c] = (float)invstd;
  scale[c] = (float)sc;
  shift[c] = (float)((double)beta[c] - mean * sc);
}

// out = LeakyReLU(y * scale[c] + shift[c]) -> 16-bit planes (plain, or space-to-depth for a following 4x4 stride-2 conv)
// and / or NCHW fp32 (the classifier's input).  One thread per (img, plane, y, x).
__global__ void bn_lrelu_apply_kernel(const float* __restrict__ y32, const float* __restrict__ scale, const float* __restrict__ shift,
                                      float slope, int n, int P, int h, int w, int C, int dtype, uint16_t* __restrict__ dst16, int s2d,
                                      float* __restrict__ dst_nchw, int split) {
  const size_t total = (size_t)n * P * h * w;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    size_t r = idx;
    const int x = r % w; r /= w;
    const int yy = r % h; r /= h;
    const int g = r % P; r /= P;
    const int img = (int)r;
    float v[8], sc[8], sh[8];
    load8(y32 + idx * 8, v);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int c = g * 8 + k;
      sc[k] = c < C ? __ldg(scale + c) : 0.f;
      sh[k] = c < C ? __ldg(shift + c) : 0.f;
      float t = fmaf(v[k], sc[k], sh[k]);
      v[k] = t > 0.f ? t : t * slope;
    }
    if (dst16) {   // (split precision: hi and lo halves per image, so the image stride doubles; both layouts hold P*h*w*8 per half)
      const size_t o = grad_offset(s2d ? 1 : 0, split ? 2 * img : img, g, yy, x, P, h, w, C);
      store16x8(dst16 + o, v, dtype, split ? (size_t)P * h * w * 8 : 0);
    }
    if (dst_nchw) {
      const size_t hw = (size_t)h * w;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int c = g * 8 + k;
        if (c < C) dst_nchw[((size_t)img * C + c) * hw + (size_t)yy * w + x] = v[k];
      }
    }
  }
}

// 16-bit planes -> space-to-depth 16-bit planes (for a 4x4 stride-2 conv whose input was produced by another path)
__global__ void s2d_planes16_kernel(const uint4* __restrict__ src, int n, int P, int h, int w, uint4* __restrict__ dst) {
  const size_t total = (size_t)n * P * h * w;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    size_t r = idx;
    const int x = r % w; r /= w;
    const int yy = r % h; r /= h;
    const int g = r % P; r /= P;
    dst[grad_offset(1, (int)r, g, yy, x, P, h, w, 0) >> 3] = __ldg(src + idx);
  }
}

// ---- backward --------------------------------------------------------------------------------------------------------------------
// gb = g * lrelu'(y*scale+shift); partial sums of gb and gb * xhat (xhat = (y - mean) * invstd) per (plane, chunk)
__global__ void bn_bwd_partial_kernel(const float* __restrict__ g, int layout, const float* __restrict__ y32, const float* __restrict__ scale,
                                      const float* __restrict__ shift, const float* __restrict__ mean, const float* __restrict__ invstd,
                                      float slope, int n, int P, int h, int w, int C, float* __restrict__ part, int chunks) {
  const int gi = blockIdx.y, chunk = blockIdx.x;
  const size_t hw = (size_t)h * w, M = (size_t)n * hw;
  const size_t per = (M + chunks - 1) / chunks;
  const size_t m0 = (size_t)chunk * per, m1 = m0 + per < M ? m0 + per : M;
  float sc[8], sh[8], mu[8], is[8], acc[16];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int c = gi * 8 + k;
    sc[k] = c < C ? scale[c] : 0.f; sh[k] = c < C ? shift[c] : 0.f; mu[k] = c < C ? mean[c] : 0.f; is[k] = c < C ? invstd[c] : 0.f;
    acc[k] = 0.f; acc[8 + k] = 0.f;
  }
  for (size_t m = m0 + threadIdx.x; m < m1; m += blockDim.x) {
    const int img = (int)(m / hw);
    const int pix = (int)(m - (size_t)img * hw);
    const int yy = pix / w, x = pix - yy * w;
    float v[8], gv[8];
    load8(y32 + (((size_t)img * P + gi) * hw + pix) * 8, v);
    load_grad8(g, layout, img, gi, yy, x, P, h, w, C, gv);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float gb = fmaf(v[k], sc[k], sh[k]) > 0.f ? gv[k] : gv[k] * slope;
      acc[k] += gb;
      acc[8 + k] = fmaf(gb, (v[k] - mu[k]) * is[k], acc[8 + k]);
    }
  }
  block_sum16(acc, part + ((size_t)gi * chunks + chunk) * 16);
}

// dbeta = sum gb, dgamma = sum gb*xhat (gscale undoes a loss scale); c1 = dbeta / M, c2 = dgamma / M for the apply pass
// (zeros when the layer normalised with running statistics)
__global__ void bn_bwd_finalize_kernel(const float* __restrict__ part, int chunks, int C, double M, int train, float gscale, int accumulate,
                                       float* dgamma, float* dbeta, float* __restrict__ c1, float* __restrict__ c2) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double s = 0.0, q = 0.0;
  const float* pp = part + (size_t)(c >> 3) * chunks * 16 + (c & 7);
  for (int k = 0; k < chunks; ++k) { s += pp[k * 16]; q += pp[k * 16 + 8]; }
  if (dbeta) dbeta[c] = (accumulate ? dbeta[c] : 0.f) + (float)(s * gscale);
  if (dgamma) dgamma[c] = (accumulate ? dgamma[c] : 0.f) + (float)(q * gscale);
  c1[c] = train ? (float)(s / M) : 0.f;
  c2[c] = train ? (float)(q / M) : 0.f;
}

// gy = scale * (gb - c1 - xhat * c2) -> 16-bit planes: the gradient of the conv's output, operand of its dgrad / wgrad launches.
// Layers without normalisation pass scale = 1, shift = 0, mean = 0, invstd = 1, c1 = c2 = 0.
__global__ void bn_bwd_apply_kernel(const float* __restrict__ g, int layout, const float* __restrict__ y32, const float* __restrict__ scale,
                                    const float* __restrict__ shift, const float* __restrict__ mean, const float* __restrict__ invstd,
                                    const float* __restrict__ c1, const float* __restrict__ c2, float slope, int n, int P, int h, int w, int C,
                                    int dtype, uint16_t* __restrict__ gy16, int split) {
  const size_t total = (size_t)n * P * h * w;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    size_t r = idx;
    const int x = r % w; r /= w;
    const int yy = r % h; r /= h;
    const int gi = r % P; r /= P;
    const int img = (int)r;
    float v[8], gv[8], o[8];
    load8(y32 + idx * 8, v);
    load_grad8(g, layout, img, gi, yy, x, P, h, w, C, gv);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int c = gi * 8 + k;
      if (c < C) {
        const float sc = __ldg(scale + c);
        const float gb = fmaf(v[k], sc, __ldg(shift + c)) > 0.f ? gv[k] : gv[k] * slope;
        const float xh = (v[k] - __ldg(mean + c)) * __ldg(invstd + c);
        o[k] = sc * (gb - __ldg(c1 + c) - xh * __ldg(c2 + c));
      } else {
        o[k] = 0.f;
      }
    }
    const size_t per_img = (size_t)P * h * w;
    const size_t o16 = split ? (size_t)img * 2 * per_img + (idx - (size_t)img * per_img) : idx;
    store16x8(gy16 + o16 * 8, o, dtype, split ? per_img * 8 : 0);
  }
}

// ---- second-order pass of the gradient penalty (WGAN-GP, models/modules/loss.py:260-279) ------------------------------------------
// L_gp = mean_b (||g_b|| - 1)^2 with g = d(sum D(x))/dx.  Its parameter gradient is  grad_theta [ D'(x; theta)[v] ],  v = dL_gp/dg: the
// directional derivative of the critic along v (a forward-mode tangent through the layers) differentiated by an ordinary backward pass
// over the (primal, tangent) pair of streams.  Convolutions are linear (tangent = the same conv without bias), LeakyReLU has a zero
// second derivative (tangent and both adjoints take the primal's mask); the only second-order arithmetic is BatchNorm's, whose
// batch statistics couple all pixels of a channel:
//   primal   yh = (y - mu) r,  z = gamma yh + beta                       r = 1 / sqrt(var + eps)
//   tangent  w  = gamma r A,   A = t - mean(t) - yh c,   c = mean(yh t)   (the BatchNorm Jacobian applied to the conv's tangent t)
// bn_tangent_apply_kernel: LeakyReLU'(z) * w  -> 16-bit planes (plain / space-to-depth) and / or NCHW fp32.  c1 = mean(t), c2 = c
// come from bn_bwd_partial_kernel (slope 1) + bn_bwd_finalize_kernel.  Layers without normalisation pass scale 1, c1 = c2 = 0.
__global__ void bn_tangent_apply_kernel(const float* __restrict__ t32, const float* __restrict__ y32, const float* __restrict__ scale,
                                        const float* __restrict__ shift, const float* __restrict__ mean, const float* __restrict__ invstd,
                                        const float* __restrict__ c1, const float* __restrict__ c2, float slope, int n, int P, int h, int w, int C,
                                        int dtype, uint16_t* __restrict__ dst16, int s2d, float* __restrict__ dst_nchw, int split) {
  const size_t total = (size_t)n * P * h * w;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    size_t r = idx;
    const int x = r % w; r /= w;
    const int yy = r % h; r /= h;
    const int g = r % P; r /= P;
    const int img = (int)r;
    float v[8], t[8], o[8];
    load8(y32 + idx * 8, v);
    load8(t32 + idx * 8, t);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int c = g * 8 + k;
      if (c < C) {
        const float sc = __ldg(scale + c);
        const float xh = (v[k] - __ldg(mean + c)) * __ldg(invstd + c);
        const float wv = sc * (t[k] - __ldg(c1 + c) - xh * __ldg(c2 + c));
        o[k] = fmaf(v[k], sc, __ldg(shift + c)) > 0.f ? wv : wv * slope;
      } else {
        o[k] = 0.f;
      }
    }
    if (dst16) {
      const size_t off = grad_offset(s2d ? 1 : 0, split ? 2 * img : img, g, yy, x, P, h, w, C);
      store16x8(dst16 + off, o, dtype, split ? (size_t)P * h * w * 8 : 0);
    }
    if (dst_nchw) {
      const size_t hw = (size_t)h * w;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int c = g * 8 + k;
        if (c < C) dst_nchw[((size_t)img * C + c) * hw + (size_t)yy * w + x] = o[k];
      }
    }
  }
}

// Backward of the pair (z, w) above.  Incoming adjoints: zb (of the primal activation) and wb (of the tangent activation), both in
// one of the three gradient layouts; either may be NULL (= zero).  With m = LeakyReLU'(z): p = m zb, q = m wb.  Per channel
//   sums  S0 = sum p, S1 = sum p yh, S2 = sum q, S3 = sum q yh, S4 = sum q A
//   tb  = gamma r (q - S2/N - yh S3/N)                                   adjoint of the conv's tangent t
//   yb  = gamma r (p - S0/N - yh S1/N) - gamma r^2 [ (S4/N) yh + c (q - S2/N - yh S3/N) + (S3/N) A ]      adjoint of the conv's output y
//   dgamma = S1 + r S4,  dbeta = S0
// partial kernel: grid (chunks, P) -> part[(g*chunks + chunk)*40 + 8*j + k]
__global__ void bn_dbl_partial_kernel(const float* __restrict__ zb, const float* __restrict__ wb, int layout, const float* __restrict__ y32,
                                      const float* __restrict__ t32, const float* __restrict__ scale, const float* __restrict__ shift,
                                      const float* __restrict__ mean, const float* __restrict__ invstd, const float* __restrict__ c1,
                                      const float* __restrict__ c2, float slope, int n, int P, int h, int w, int C, float* __restrict__ part,
                                      int chunks) {
  const int gi = blockIdx.y, chunk = blockIdx.x;
  const size_t hw = (size_t)h * w, M = (size_t)n * hw;
  const size_t per = (M + chunks - 1) / chunks;
  const size_t m0 = (size_t)chunk * per, m1 = m0 + per < M ? m0 + per : M;
  float sc[8], sh[8], mu[8], is[8], tm[8], tc[8], acc[40];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int c = gi * 8 + k;
    const bool on = c < C;
    sc[k] = on ? scale[c] : 0.f; sh[k] = on ? shift[c] : 0.f; mu[k] = on ? mean[c] : 0.f; is[k] = on ? invstd[c] : 0.f;
    tm[k] = on ? c1[c] : 0.f; tc[k] = on ? c2[c] : 0.f;
  }
#pragma unroll
  for (int k = 0; k < 40; ++k) acc[k] = 0.f;
  for (size_t m = m0 + threadIdx.x; m < m1; m += blockDim.x) {
    const int img = (int)(m / hw);
    const int pix = (int)(m - (size_t)img * hw);
    const int yy = pix / w, x = pix - yy * w;
    float v[8], t[8], pz[8], qw[8];
    const size_t o = (((size_t)img * P + gi) * hw + pix) * 8;
    load8(y32 + o, v);
    load8(t32 + o, t);
    if (zb) load_grad8(zb, layout, img, gi, yy, x, P, h, w, C, pz);
    if (wb) load_grad8(wb, layout, img, gi, yy, x, P, h, w, C, qw);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float msk = fmaf(v[k], sc[k], sh[k]) > 0.f ? 1.f : slope;
      const float xh = (v[k] - mu[k]) * is[k];
      const float p = zb ? pz[k] * msk : 0.f, q = wb ? qw[k] * msk : 0.f;
      const float A = t[k] - tm[k] - xh * tc[k];
      acc[k] += p; acc[8 + k] = fmaf(p, xh, acc[8 + k]);
      acc[16 + k] += q; acc[24 + k] = fmaf(q, xh, acc[24 + k]); acc[32 + k] = fmaf(q, A, acc[32 + k]);
    }
  }
  // block-level sums of the 40 accumulators
  __shared__ float smem[8][40];
#pragma unroll
  for (int k = 0; k < 40; ++k) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], off);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < 40; ++k) smem[warp][k] = acc[k];
  }
  __syncthreads();
  if (threadIdx.x < 40) {
    float sum = 0.f;
    for (int wv = 0; wv < (int)(blockDim.x >> 5); ++wv) sum += smem[wv][threadIdx.x];
    part[((size_t)gi * chunks + chunk) * 40 + threadIdx.x] = sum;
  }
}

// coef[j*C + c] = S_j / N for j = 0..4; dgamma / dbeta (+= when accumulate), scaled by gscale
__global__ void bn_dbl_finalize_kernel(const float* __restrict__ part, int chunks, int C, double M, const float* __restrict__ invstd, int has_bn,
                                       float gscale, int accumulate, float* dgamma, float* dbeta, float* __restrict__ coef) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double S[5] = {0, 0, 0, 0, 0};
  const float* pp = part + (size_t)(c >> 3) * chunks * 40 + (c & 7);
  for (int k = 0; k < chunks; ++k)
    for (int j = 0; j < 5; ++j) S[j] += pp[(size_t)k * 40 + 8 * j];
  if (has_bn) {
    if (dbeta) dbeta[c] = (accumulate ? dbeta[c] : 0.f) + (float)(S[0] * gscale);
    if (dgamma) dgamma[c] = (accumulate ? dgamma[c] : 0.f) + (float)((S[1] + (double)invstd[c] * S[4]) * gscale);
  }
  for (int j = 0; j < 5; ++j) coef[j * C + c] = has_bn ? (float)(S[j] / M) : 0.f;
}

// tb16 / yb16 (16-bit planes, operands of the dgrad and wgrad launches of the layer's conv)
__global__ void bn_dbl_apply_kernel(const float* __restrict__ zb, const float* __restrict__ wb, int layout, const float* __restrict__ y32,
                                    const float* __restrict__ t32, const float* __restrict__ scale, const float* __restrict__ shift,
                                    const float* __restrict__ mean, const float* __restrict__ invstd, const float* __restrict__ c1,
                                    const float* __restrict__ c2, const float* __restrict__ coef, int has_bn, float slope, int n, int P, int h, int w,
                                    int C, int dtype, uint16_t* __restrict__ tb16, uint16_t* __restrict__ yb16, int split) {
  const size_t total = (size_t)n * P * h * w;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    size_t r = idx;
    const int x = r % w; r /= w;
    const int yy = r % h; r /= h;
    const int gi = r % P; r /= P;
    const int img = (int)r;
    float v[8], t[8], pz[8], qw[8], ot[8], oy[8];
    load8(y32 + idx * 8, v);
    load8(t32 + idx * 8, t);
    if (zb) load_grad8(zb, layout, img, gi, yy, x, P, h, w, C, pz);
    if (wb) load_grad8(wb, layout, img, gi, yy, x, P, h, w, C, qw);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int c = gi * 8 + k;
      if (c < C) {
        const float sc = __ldg(scale + c), is = __ldg(invstd + c);
        const float msk = fmaf(v[k], sc, __ldg(shift + c)) > 0.f ? 1.f : slope;
        const float p = zb ? pz[k] * msk : 0.f, q = wb ? qw[k] * msk : 0.f;
        if (has_bn) {
          const float xh = (v[k] - __ldg(mean + c)) * is;
          const float cc = __ldg(c2 + c);
          const float A = t[k] - __ldg(c1 + c) - xh * cc;
          const float s0 = __ldg(coef + c), s1 = __ldg(coef + C + c), s2 = __ldg(coef + 2 * C + c), s3 = __ldg(coef + 3 * C + c),
                      s4 = __ldg(coef + 4 * C + c);
          const float qc = q - s2 - xh * s3;
          ot[k] = sc * qc;                                                    // sc = gamma r
          oy[k] = sc * (p - s0 - xh * s1) - sc * is * (s4 * xh + cc * qc + s3 * A);
        } else {      // no normalisation (scale = 1) or running statistics (a fixed affine map: scale = gamma * invstd)
          ot[k] = sc * q;
          oy[k] = sc * p;
        }
      } else {
        ot[k] = 0.f; oy[k] = 0.f;
      }
    }
    const size_t per_img = (size_t)P * h * w;
    const size_t o16 = split ? (size_t)img * 2 * per_img + (idx - (size_t)img * per_img) : idx;
    if (tb16) store16x8(tb16 + o16 * 8, ot, dtype, split ? per_img * 8 : 0);
    if (yb16) store16x8(yb16 + o16 * 8, oy, dtype, split ? per_img * 8 : 0);
  }
}

// ---- fully connected layers (fp32; 8192 -> 100 -> 1 at the reference's sizes: a few MFLOP, weight-read bound) ---------------------
// out[b][j] = act(bias[j] + sum_k x[b][k] * W[j][k]); one block per output feature, batch in groups of 8
__global__ void linear_fwd_kernel(const float* __restrict__ x, const float* __restrict__ W, const float* __restrict__ bias, int B, int K, int J,
                                  int lrelu, float slope, float* __restrict__ out) {
  __shared__ float sm[8][8];
  const int j = blockIdx.x;
  const float* wr = W + (size_t)j * K;
  for (int b0 = 0; b0 < B; b0 += 8) {
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
      const float wv = __ldg(wr + k);
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (b0 + i < B) acc[i] = fmaf(wv, __ldg(x + (size_t)(b0 + i) * K + k), acc[i]);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], o);
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();   // sm reuse across batch groups
    if (lane == 0) {
#pragma unroll
      for (int i = 0; i < 8; ++i) sm[warp][i] = acc[i];
    }
    __syncthreads();
    if (threadIdx.x < 8 && b0 + threadIdx.x < B) {
      float s = bias ? bias[j] : 0.f;
      for (int wv = 0; wv < (int)(blockDim.x >> 5); ++wv) s += sm[wv][threadIdx.x];
      if (lrelu) s = s > 0.f ? s : s * slope;
      out[(size_t)(b0 + threadIdx.x) * J + j] = s;
    }
  }
}

__device__ __forceinline__ float masked_grad(const float* g, const float* act, float slope, int b, int j, int J) {
  const float v = g[(size_t)b * J + j];
  return (act && !(act[(size_t)b * J + j] > 0.f)) ? v * slope : v;
}

// gx[b][k] = sum_j g'[b][j] * W[j][k], g' = g masked by the layer's own LeakyReLU (act = its output, NULL = linear)
__global__ void linear_bwd_input_kernel(const float* __restrict__ g, const float* __restrict__ act, float slope, const float* __restrict__ W,
                                        int B, int K, int J, float* __restrict__ gx) {
  const size_t total = (size_t)B * K;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int b = (int)(idx / K), k = (int)(idx - (size_t)b * K);
    float acc = 0.f;
    for (int j = 0; j < J; ++j) acc = fmaf(masked_grad(g, act, slope, b, j, J), __ldg(W + (size_t)j * K + k), acc);
    gx[idx] = acc;
  }
}

// dW[j][k] (+)= gscale * sum_b g'[b][j] * x[b][k];  db[j] (+)= gscale * sum_b g'[b][j]
__global__ void linear_bwd_weight_kernel(const float* __restrict__ g, const float* __restrict__ act, float slope, const float* __restrict__ x,
                                         int B, int K, int J, float gscale, int accumulate, float* dW, float* db) {
  const size_t total = (size_t)J * K;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int j = (int)(idx / K), k = (int)(idx - (size_t)j * K);
    float acc = 0.f, accb = 0.f;
    for (int b = 0; b < B; ++b) {
      const float gm = masked_grad(g, act, slope, b, j, J);
      acc = fmaf(gm, __ldg(x + (size_t)b * K + k), acc);
      accb += gm;
    }
    if (dW) dW[idx] = (accumulate ? dW[idx] : 0.f) + gscale * acc;
    if (db && k == 0) db[j] = (accumulate ? db[j] : 0.f) + gscale * accb;
  }
}

// ---- structure-tensor statistics of the latent-control loss (models/modules/loss.py:51-62,133-147: FilterLoss with
// 'structure_tensor' latent channels) -------------------------------------------------------------------------------------------------
// dx = x[i][j+1] - x[i][j], dy = x[i+1][j] - x[i][j] (the reference's 2x2 depth-wise filters, no padding); per image the
// sums over channels and the (h-1) x (w-1) valid positions of dx^2, dy^2, dx*dy.  HBM-bound: the image is read once.
constexpr int kStChunks = 32;

__global__ void structure_tensor_partial_kernel(const float* __restrict__ img, int c, int h, int w, float* __restrict__ part) {
  const int n = blockIdx.y, chunk = blockIdx.x;
  const size_t per_c = (size_t)(h - 1) * (w - 1), total = per_c * c;
  const size_t per = (total + kStChunks - 1) / kStChunks;
  const size_t t0 = (size_t)chunk * per, t1 = t0 + per < total ? t0 + per : total;
  const float* base = img + (size_t)n * c * h * w;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f;
  for (size_t t = t0 + threadIdx.x; t < t1; t += blockDim.x) {
    const size_t ch = t / per_c, r = t - ch * per_c;
    const int i = (int)(r / (w - 1)), j = (int)(r - (size_t)i * (w - 1));
    const float* p = base + (ch * h + i) * w + j;
    const float v = __ldg(p), dx = __ldg(p + 1) - v, dy = __ldg(p + w) - v;
    a0 = fmaf(dx, dx, a0); a1 = fmaf(dy, dy, a1); a2 = fmaf(dx, dy, a2);
  }
  __shared__ float sm[8][3];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a0 += __shfl_xor_sync(0xffffffffu, a0, o); a1 += __shfl_xor_sync(0xffffffffu, a1, o); a2 += __shfl_xor_sync(0xffffffffu, a2, o);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { sm[warp][0] = a0; sm[warp][1] = a1; sm[warp][2] = a2; }
  __syncthreads();
  if (threadIdx.x < 3) {
    float s = 0.f;
    for (int wv = 0; wv < (int)(blockDim.x >> 5); ++wv) s += sm[wv][threadIdx.x];
    part[((size_t)n * kStChunks + chunk) * 3 + threadIdx.x] = s;
  }
}

// out[n][k] = mean over channels and valid positions (double combine of the partials)
__global__ void structure_tensor_finalize_kernel(const float* __restrict__ part, int n, double inv_count, float* __restrict__ out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * 3) return;
  const int img = t / 3, k = t - img * 3;
  double s = 0.0;
  for (int ch = 0; ch < kStChunks; ++ch) s += part[((size_t)img * kStChunks + ch) * 3 + k];
  out[t] = (float)(s * inv_count);
}

// gradient of sum_k g[n][k] * mean_k with respect to the image: with px = (2 g0 dx + g2 dy)/N, py = (2 g1 dy + g2 dx)/N on the
// valid positions, grad x[i][j] = -px(i,j) - py(i,j) + px(i,j-1) + py(i-1,j).  One thread per pixel (gather form, no atomics).
__global__ void structure_tensor_bwd_kernel(const float* __restrict__ img, const float* __restrict__ g, int n, int c, int h, int w, float inv_count,
                                            float* __restrict__ grad) {
  const size_t total = (size_t)n * c * h * w;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    size_t r = idx;
    const int j = r % w; r /= w;
    const int i = r % h; r /= h;
    const int im = (int)(r / c);
    const float g0 = 2.f * __ldg(g + im * 3) * inv_count, g1 = 2.f * __ldg(g + im * 3 + 1) * inv_count, g2 = __ldg(g + im * 3 + 2) * inv_count;
    const float* p = img + idx;
    const float v = __ldg(p);
    float acc = 0.f;
    if (i < h - 1 && j < w - 1) {
      const float dx = __ldg(p + 1) - v, dy = __ldg(p + w) - v;
      acc -= fmaf(g0, dx, g2 * dy) + fmaf(g1, dy, g2 * dx);
    }
    if (j > 0 && i < h - 1) {     // position (i, j-1): this pixel is its right neighbour
      const float u = __ldg(p - 1), dx = v - u, dy = __ldg(p + w - 1) - u;
      acc += fmaf(g0, dx, g2 * dy);
    }
    if (i > 0 && j < w - 1) {     // position (i-1, j): this pixel is its lower neighbour
      const float u = __ldg(p - w), dx = __ldg(p - w + 1) - u, dy = v - u;
      acc += fmaf(g1, dy, g2 * dx);
    }
    grad[idx] = acc;
  }
}

}  // namespace esr
