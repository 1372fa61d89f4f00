// HBM-bound helper kernels: boundary layout conversion, nearest x2, and the CEM filters.
#pragma once
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace esr {

__device__ __forceinline__ uint16_t to16(float v, int dtype) {
  if (dtype == 0) return __half_as_ushort(__float2half_rn(v));
  return __bfloat16_as_ushort(__float2bfloat16_rn(v));
}
__device__ __forceinline__ float from16(uint16_t v, int dtype) {
  if (dtype == 0) return __half2float(__ushort_as_half(v));
  return __bfloat162float(__ushort_as_bfloat16(v));
}
__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// NCHW fp32 [n][c][h][w] -> planes [n][planes_total][h+2pad][w+2pad][8] (replicate padding), channels
// beyond c are written as zeros.  One thread per (n, plane, y, x): 8 strided-by-HW reads (coalesced
// over x), one 16 B (and optionally one 32 B) store.
__global__ void pack_nchw_kernel(const float* __restrict__ src, int n, int c, int h, int w, int pad, int dtype,
                                 uint16_t* __restrict__ dst16, float* __restrict__ dst32, int planes_total,
                                 int plane_off, int planes) {
  const int ho = h + 2 * pad, wo = w + 2 * pad;
  const size_t total = (size_t)n * planes * ho * wo;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    size_t r = idx;
    const int x = r % wo; r /= wo;
    const int y = r % ho; r /= ho;
    const int g = r % planes; r /= planes;
    const int img = (int)r;
    const int sy = clampi(y - pad, 0, h - 1), sx = clampi(x - pad, 0, w - 1);
    float v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int ch = g * 8 + k;
      v[k] = ch < c ? __ldg(src + (((size_t)img * c + ch) * h + sy) * w + sx) : 0.f;
    }
    const size_t o = ((((size_t)img * planes_total + plane_off + g) * ho + y) * wo + x) * 8;
    if (dst16) {
      uint4 pk;
      pk.x = (uint32_t)to16(v[0], dtype) | ((uint32_t)to16(v[1], dtype) << 16);
      pk.y = (uint32_t)to16(v[2], dtype) | ((uint32_t)to16(v[3], dtype) << 16);
      pk.z = (uint32_t)to16(v[4], dtype) | ((uint32_t)to16(v[5], dtype) << 16);
      pk.w = (uint32_t)to16(v[6], dtype) | ((uint32_t)to16(v[7], dtype) << 16);
      *reinterpret_cast<uint4*>(dst16 + o) = pk;
    }
    if (dst32) {
      float4* op = reinterpret_cast<float4*>(dst32 + o);
      op[0] = make_float4(v[0], v[1], v[2], v[3]);
      op[1] = make_float4(v[4], v[5], v[6], v[7]);
    }
  }
}

template <bool kIs16>
__global__ void unpack_planes_kernel(const void* __restrict__ src, int dtype, int n, int c, int h, int w,
                                     int planes_total, int plane_off, float* __restrict__ dst) {
  const size_t total = (size_t)n * c * h * w;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    size_t r = idx;
    const int x = r % w; r /= w;
    const int y = r % h; r /= h;
    const int ch = r % c; r /= c;
    const int img = (int)r;
    const size_t o = ((((size_t)img * planes_total + plane_off + (ch >> 3)) * h + y) * w + x) * 8 + (ch & 7);
    if (kIs16) dst[idx] = from16(reinterpret_cast<const uint16_t*>(src)[o], dtype);
    else dst[idx] = reinterpret_cast<const float*>(src)[o];
  }
}

// Bilinear (align_corners=False) 1/s down-scaling of the HR latent map, applied to the replicate-padded
// map (architecture.py:284 after CEMnet.py:290-292).  NCHW fp32 in ([n*c][hh][wh], unpadded) -> NCHW fp32 out
// ([n*c][(hh+2*pad)/s][(wh+2*pad)/s]).
__global__ void latent_downscale_kernel(const float* __restrict__ src, size_t nc, int hh, int wh, int s, int pad,
                                        float* __restrict__ dst) {
  const int hp = hh + 2 * pad, wp = wh + 2 * pad;
  const int ho = hp / s, wo = wp / s;
  const size_t total = nc * ho * wo;
  const float fs = 1.0f / (1.0f / (float)s);
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    size_t r = idx;
    const int j = r % wo; r /= wo;
    const int i = r % ho; r /= ho;
    const float sy = fmaxf(((float)i + 0.5f) * fs - 0.5f, 0.f), sx = fmaxf(((float)j + 0.5f) * fs - 0.5f, 0.f);
    const int y0 = (int)sy, x0 = (int)sx;
    const float ly = sy - (float)y0, lx = sx - (float)x0;
    const int y1 = min(y0 + 1, hp - 1), x1 = min(x0 + 1, wp - 1);
    const int ya = clampi(y0 - pad, 0, hh - 1), yb = clampi(y1 - pad, 0, hh - 1);
    const int xa = clampi(x0 - pad, 0, wh - 1), xb = clampi(x1 - pad, 0, wh - 1);
    const float* p = src + r * (size_t)hh * wh;
    const float v00 = __ldg(p + (size_t)ya * wh + xa), v01 = __ldg(p + (size_t)ya * wh + xb);
    const float v10 = __ldg(p + (size_t)yb * wh + xa), v11 = __ldg(p + (size_t)yb * wh + xb);
    dst[idx] = (1.f - ly) * ((1.f - lx) * v00 + lx * v01) + ly * ((1.f - lx) * v10 + lx * v11);
  }
}

// nearest x2 on 16-bit planes: one thread per source pixel-plane, four 16 B stores.
__global__ void upsample2x_kernel(const uint4* __restrict__ src, size_t nplanes, int h, int w, uint4* __restrict__ dst) {
  const size_t total = nplanes * h * w;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    size_t r = idx;
    const int x = r % w; r /= w;
    const int y = r % h; r /= h;
    const uint4 v = __ldg(src + idx);
    uint4* o = dst + (r * (2 * (size_t)h) + 2 * y) * (2 * (size_t)w) + 2 * x;
    o[0] = v; o[1] = v;
    o[2 * (size_t)w] = v; o[2 * (size_t)w + 1] = v;
  }
}

// ------------------------------------------------------------------------------------------------
// CEM filters (CEM/CEMnet.py:254-311).  All images NCHW fp32; one block per (tile, n*c).
// Filters arrive as `rank` separable terms: K[a][b] = sum_r kv[r][a] * kh[r][b]  (rank 1 for the
// reference's bicubic kernels; more terms cover estimated, non-separable kernels exactly).
// ------------------------------------------------------------------------------------------------
constexpr int kCemTJ = 32, kCemTI = 8, kCemThreads = 256;

// DownscaleOP: out[i][j] = sum_{a,b} K[a][b] * G[clamp(s*i+phase+a-r)][clamp(s*j+phase+b-r)],  r = len/2
// optional fused residual: out = sub_from - Down(G)
__global__ void __launch_bounds__(kCemThreads)
cem_down_kernel(const float* __restrict__ g, int hh, int wh, int s, int phase, const float* __restrict__ kv,
                const float* __restrict__ kh, int len, int rank, const float* __restrict__ sub_from,
                float* __restrict__ out) {
  extern __shared__ float sm[];
  const int hl = hh / s, wl = wh / s;
  const int rows = (kCemTI - 1) * s + len, cols = (kCemTJ - 1) * s + len;
  float* tile = sm;                 // [rows][cols]
  float* hbuf = sm + rows * cols;   // [rows][TJ]
  const int nc = blockIdx.z;
  const int i0 = blockIdx.y * kCemTI, j0 = blockIdx.x * kCemTJ;
  const int r = len / 2;
  const float* gp = g + (size_t)nc * hh * wh;
  const int ybase = s * i0 + phase - r, xbase = s * j0 + phase - r;
  for (int e = threadIdx.x; e < rows * cols; e += kCemThreads) {
    const int rr = e / cols, cc = e - rr * cols;
    tile[e] = __ldg(gp + (size_t)clampi(ybase + rr, 0, hh - 1) * wh + clampi(xbase + cc, 0, wh - 1));
  }
  __syncthreads();
  const int tj = threadIdx.x % kCemTJ, ti = threadIdx.x / kCemTJ;  // 32 x 8
  float acc = 0.f;
  for (int t = 0; t < rank; ++t) {
    const float* khr = kh + t * len;
    const float* kvr = kv + t * len;
    for (int e = threadIdx.x; e < rows * kCemTJ; e += kCemThreads) {
      const int rr = e / kCemTJ, jj = e - rr * kCemTJ;
      const float* tp = tile + rr * cols + jj * s;
      float a = 0.f;
      for (int b = 0; b < len; ++b) a = fmaf(__ldg(khr + b), tp[b], a);
      hbuf[e] = a;
    }
    __syncthreads();
    for (int a = 0; a < len; ++a) acc = fmaf(__ldg(kvr + a), hbuf[(ti * s + a) * kCemTJ + tj], acc);
    __syncthreads();
  }
  const int i = i0 + ti, j = j0 + tj;
  if (i < hl && j < wl) {
    const size_t o = (size_t)nc * hl * wl + (size_t)i * wl + j;
    out[o] = sub_from ? (__ldg(sub_from + o) - acc) : acc;
  }
}

// Conv_LR_with_Inv_hTh_OP: same-size correlation with replicate padding len/2.
__global__ void __launch_bounds__(kCemThreads)
cem_inv_kernel(const float* __restrict__ e_in, int hl, int wl, const float* __restrict__ kv,
               const float* __restrict__ kh, int len, int rank, float* __restrict__ out) {
  extern __shared__ float sm[];
  const int rows = kCemTI + len - 1, cols = kCemTJ + len - 1;
  float* tile = sm;
  float* hbuf = sm + rows * cols;
  const int nc = blockIdx.z;
  const int i0 = blockIdx.y * kCemTI, j0 = blockIdx.x * kCemTJ;
  const int r = len / 2;
  const float* ep = e_in + (size_t)nc * hl * wl;
  for (int e = threadIdx.x; e < rows * cols; e += kCemThreads) {
    const int rr = e / cols, cc = e - rr * cols;
    tile[e] = __ldg(ep + (size_t)clampi(i0 - r + rr, 0, hl - 1) * wl + clampi(j0 - r + cc, 0, wl - 1));
  }
  __syncthreads();
  const int tj = threadIdx.x % kCemTJ, ti = threadIdx.x / kCemTJ;
  float acc = 0.f;
  for (int t = 0; t < rank; ++t) {
    const float* khr = kh + t * len;
    const float* kvr = kv + t * len;
    for (int e = threadIdx.x; e < rows * kCemTJ; e += kCemThreads) {
      const int rr = e / kCemTJ, jj = e - rr * kCemTJ;
      const float* tp = tile + rr * cols + jj;
      float a = 0.f;
      for (int b = 0; b < len; ++b) a = fmaf(__ldg(khr + b), tp[b], a);
      hbuf[e] = a;
    }
    __syncthreads();
    for (int a = 0; a < len; ++a) acc = fmaf(__ldg(kvr + a), hbuf[(ti + a) * kCemTJ + tj], acc);
    __syncthreads();
  }
  const int i = i0 + ti, j = j0 + tj;
  if (i < hl && j < wl) out[(size_t)nc * hl * wl + (size_t)i * wl + j] = acc;
}

// Upscale_OP + add:  out[Y][X] = g[Y][X] + sum_{a,b} K[a][b] * S[clamp(Y+a-r)][clamp(X+b-r)]
//   S = zero-stuffed F: S[s*i+phase][s*j+phase] = F[i][j]; replicate padding acts on S (CEMnet.py:268-272).
// Output is cropped by `crop` pixels per side.  Block = 32 x 32 output pixels of one (n, c) image.
constexpr int kUpT = 32;
__global__ void __launch_bounds__(kCemThreads)
cem_up_add_kernel(const float* __restrict__ f, const float* __restrict__ g, int hl, int wl, int s, int phase,
                  const float* __restrict__ kv, const float* __restrict__ kh, int len, int rank, int crop,
                  float* __restrict__ out) {
  extern __shared__ float sm[];
  const int hh = hl * s, wh = wl * s;
  const int ho = hh - 2 * crop, wo = wh - 2 * crop;
  const int r = len / 2;
  const int nc = blockIdx.z;
  const int Y0 = blockIdx.y * kUpT + crop, X0 = blockIdx.x * kUpT + crop;  // full-res coordinates
  // LR rows / cols that can contribute to this tile
  const int ylo = clampi(Y0 - r, 0, hh - 1), yhi = clampi(Y0 + kUpT - 1 + r, 0, hh - 1);
  const int xlo = clampi(X0 - r, 0, wh - 1), xhi = clampi(X0 + kUpT - 1 + r, 0, wh - 1);
  const int ilo = max((ylo - phase + s - 1) / s, 0), ihi = min((yhi - phase) / s, hl - 1);
  const int jlo = max((xlo - phase + s - 1) / s, 0), jhi = min((xhi - phase) / s, wl - 1);
  const int ni = max(ihi - ilo + 1, 0), nj = max(jhi - jlo + 1, 0);
  const int maxn = (kUpT + len) / s + 2;
  float* ft = sm;                    // [maxn][maxn] LR window
  float* hb = sm + maxn * maxn;      // [maxn][kUpT]  horizontally filtered rows
  const float* fp = f + (size_t)nc * hl * wl;
  for (int e = threadIdx.x; e < ni * nj; e += kCemThreads) {
    const int ii = e / nj, jj = e - ii * nj;
    ft[ii * maxn + jj] = __ldg(fp + (size_t)(ilo + ii) * wl + (jlo + jj));
  }
  __syncthreads();
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  const int tx = threadIdx.x % kUpT, ty0 = threadIdx.x / kUpT;  // 32 x 8, each thread does 4 rows
  for (int t = 0; t < rank; ++t) {
    const float* khr = kh + t * len;
    const float* kvr = kv + t * len;
    // horizontal: hb[ii][X] = sum_b kh[b] * Srow_ii[clamp(X+b-r)]
    for (int e = threadIdx.x; e < ni * kUpT; e += kCemThreads) {
      const int ii = e / kUpT, xx = e - ii * kUpT;
      const int X = X0 + xx;
      float a = 0.f;
      for (int b = 0; b < len; ++b) {
        const int xs = clampi(X + b - r, 0, wh - 1) - phase;
        if (xs >= 0 && xs % s == 0) {
          const int j = xs / s - jlo;
          if (j >= 0 && j < nj) a = fmaf(__ldg(khr + b), ft[ii * maxn + j], a);
        }
      }
      hb[ii * kUpT + xx] = a;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int Y = Y0 + ty0 + 8 * k;
      float a2 = 0.f;
      for (int a = 0; a < len; ++a) {
        const int ys = clampi(Y + a - r, 0, hh - 1) - phase;
        if (ys >= 0 && ys % s == 0) {
          const int i = ys / s - ilo;
          if (i >= 0 && i < ni) a2 = fmaf(__ldg(kvr + a), hb[i * kUpT + tx], a2);
        }
      }
      acc[k] += a2;
    }
    __syncthreads();
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int Y = Y0 + ty0 + 8 * k, X = X0 + tx;
    const int yo = Y - crop, xo = X - crop;
    if (yo < ho && xo < wo) {
      float v = acc[k];
      if (g) v += __ldg(g + (size_t)nc * hh * wh + (size_t)Y * wh + X);
      out[(size_t)nc * ho * wo + (size_t)yo * wo + xo] = v;
    }
  }
}

}  // namespace esr
