// Kernel-density soft histogram / dictionary distance of the GUI's imprinting tools (Z_optimization.py:24-230, SoftHistogramLoss):
//   E[p][b] = exp( -(1/D) sum_d (c(x[d][p] - bins[d][b]) + eps)^2 / T ),   c(u) = min(|u|, |u - vmax|, |u + vmax|)   (cyclic distance)
//   histogram mode : hist[b] = (1/P) sum_p E[p][b]            (the reference: exp(hist.mean(0)).mean(0), :196-210)
//   dictionary mode: out[p]  = -log( (1/B) sum_b E[p][b] )    (:206-207)
// The reference materialises the [D, P, B] distance tensor in double precision; here it never exists: O(D P B) flops on the fly,
// fp64 like the reference, O(D (P + B)) bytes.  D = 1 (grey-level histogram, 256 bins) ... 36 (6x6 patches), B up to a few thousand.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace esr {

constexpr int kHistMaxD = 64;
constexpr int kHistTileB = 128;     // bins per block (one per thread)
constexpr int kHistTileP = 32;      // samples staged per shared-memory tile

__device__ __forceinline__ double cyc_dist(double u, double vmax, double& sgn) {
  double c = fabs(u);
  sgn = u > 0 ? 1.0 : (u < 0 ? -1.0 : 0.0);
  const double a = fabs(u - vmax), b = fabs(u + vmax);
  if (a < c) { c = a; sgn = (u - vmax) > 0 ? 1.0 : ((u - vmax) < 0 ? -1.0 : 0.0); }
  if (b < c) { c = b; sgn = (u + vmax) > 0 ? 1.0 : ((u + vmax) < 0 ? -1.0 : 0.0); }
  return c;
}

// hist[b] += (1/P) sum over the block's samples.  grid = (ceil(B / 128), sample chunks); hist zero-initialised by the caller.
__global__ void soft_hist_fwd_kernel(const double* __restrict__ x, int D, int P, const double* __restrict__ bins, int B, double vmax, double eps,
                                     double inv_DT, int p_per_block, double* __restrict__ hist) {
  __shared__ double xs[kHistMaxD][kHistTileP];
  const int b = blockIdx.x * kHistTileB + threadIdx.x;
  const int p0 = blockIdx.y * p_per_block, p1 = min(p0 + p_per_block, P);
  double acc = 0.0;
  for (int pt = p0; pt < p1; pt += kHistTileP) {
    const int np = min(kHistTileP, p1 - pt);
    __syncthreads();
    for (int e = threadIdx.x; e < D * np; e += blockDim.x) {
      const int d = e / np, pp = e - d * np;
      xs[d][pp] = x[(size_t)d * P + pt + pp];
    }
    __syncthreads();
    if (b < B) {
      for (int pp = 0; pp < np; ++pp) {
        double s = 0.0, sg;
        for (int d = 0; d < D; ++d) {
          const double c = cyc_dist(xs[d][pp] - bins[(size_t)d * B + b], vmax, sg) + eps;
          s += c * c;
        }
        acc += exp(-s * inv_DT);
      }
    }
  }
  if (b < B) atomicAdd(hist + b, acc / (double)P);
}

// dictionary mode forward: out[p] = -log(mean_b E[p][b]); also keeps sumE[p] for the backward.  One thread per sample.
__global__ void soft_dict_fwd_kernel(const double* __restrict__ x, int D, int P, const double* __restrict__ bins, int B, double vmax, double eps,
                                     double inv_DT, double* __restrict__ out, double* __restrict__ sumE) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  double xv[kHistMaxD];
  for (int d = 0; d < D; ++d) xv[d] = x[(size_t)d * P + p];
  double acc = 0.0;
  for (int b = 0; b < B; ++b) {
    double s = 0.0, sg;
    for (int d = 0; d < D; ++d) {
      const double c = cyc_dist(xv[d] - __ldg(bins + (size_t)d * B + b), vmax, sg) + eps;
      s += c * c;
    }
    acc += exp(-s * inv_DT);
  }
  sumE[p] = acc;
  out[p] = -log(acc / (double)B);
}

// gradient with respect to x (one thread per sample):
//   histogram : dx[d][p] = sum_b (g_hist[b] / P) E[p][b] (-2 inv_DT) (c + eps) sgn
//   dictionary: dx[d][p] = (-g_out[p] / sumE[p]) sum_b E[p][b] (-2 inv_DT) (c + eps) sgn
__global__ void soft_hist_bwd_kernel(const double* __restrict__ x, int D, int P, const double* __restrict__ bins, int B, double vmax, double eps,
                                     double inv_DT, const double* __restrict__ g_hist, const double* __restrict__ g_out, const double* __restrict__ sumE,
                                     double* __restrict__ dx) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  double xv[kHistMaxD], gx[kHistMaxD];
  for (int d = 0; d < D; ++d) { xv[d] = x[(size_t)d * P + p]; gx[d] = 0.0; }
  const double lead = g_out ? (-g_out[p] / sumE[p]) : (1.0 / (double)P);
  for (int b = 0; b < B; ++b) {
    double s = 0.0, sg;
    for (int d = 0; d < D; ++d) {
      const double c = cyc_dist(xv[d] - __ldg(bins + (size_t)d * B + b), vmax, sg) + eps;
      s += c * c;
    }
    const double wgt = (g_hist ? g_hist[b] : 1.0) * lead * exp(-s * inv_DT) * (-2.0 * inv_DT);
    for (int d = 0; d < D; ++d) {
      const double c = cyc_dist(xv[d] - __ldg(bins + (size_t)d * B + b), vmax, sg);
      gx[d] += wgt * (c + eps) * sg;
    }
  }
  for (int d = 0; d < D; ++d) dx[(size_t)d * P + p] = gx[d];
}

}  // namespace esr
