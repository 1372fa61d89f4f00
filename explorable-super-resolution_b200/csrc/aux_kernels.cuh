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

// Split precision (esr_dtype ESR_BF16X3): a 16-bit tensor holds bf16 "hi" planes and, `lo` ELEMENTS further, the bf16 residuals
// v - float(hi).  lo == 0 means a plain tensor of `dtype`.  store16x8 / load16x8 move 8 channels of one pixel-plane.
__device__ __forceinline__ void store16x8(uint16_t* dst, const float* v, int dtype, size_t lo) {
  uint16_t h[8];
  const int dt = lo ? 1 : dtype;
#pragma unroll
  for (int k = 0; k < 8; ++k) h[k] = to16(v[k], dt);
  *reinterpret_cast<uint4*>(dst) = make_uint4((uint32_t)h[0] | ((uint32_t)h[1] << 16), (uint32_t)h[2] | ((uint32_t)h[3] << 16),
                                              (uint32_t)h[4] | ((uint32_t)h[5] << 16), (uint32_t)h[6] | ((uint32_t)h[7] << 16));
  if (lo) {
    uint16_t l[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) l[k] = to16(v[k] - from16(h[k], 1), 1);
    *reinterpret_cast<uint4*>(dst + lo) = make_uint4((uint32_t)l[0] | ((uint32_t)l[1] << 16), (uint32_t)l[2] | ((uint32_t)l[3] << 16),
                                                     (uint32_t)l[4] | ((uint32_t)l[5] << 16), (uint32_t)l[6] | ((uint32_t)l[7] << 16));
  }
}
__device__ __forceinline__ void load16x8(const uint16_t* src, int dtype, size_t lo, float* v) {
  const uint4 q = __ldg(reinterpret_cast<const uint4*>(src));
  const uint32_t w[4] = {q.x, q.y, q.z, q.w};
  const int dt = lo ? 1 : dtype;
#pragma unroll
  for (int k = 0; k < 8; ++k) v[k] = from16((uint16_t)((w[k >> 1] >> ((k & 1) * 16)) & 0xFFFFu), dt);
  if (lo) {
    const uint4 ql = __ldg(reinterpret_cast<const uint4*>(src + lo));
    const uint32_t wl[4] = {ql.x, ql.y, ql.z, ql.w};
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] += from16((uint16_t)((wl[k >> 1] >> ((k & 1) * 16)) & 0xFFFFu), 1);
  }
}

// NCHW fp32 [n][c][h][w] -> planes [n][planes_total][h+2pad][w+2pad][8] (replicate padding), channels
// beyond c are written as zeros.  One thread per (n, plane, y, x): 8 strided-by-HW reads (coalesced
// over x), one 16 B (and optionally one 32 B) store.
__global__ void pack_nchw_kernel(const float* __restrict__ src, int n, int c, int h, int w, int pad, int dtype,
                                 uint16_t* __restrict__ dst16, float* __restrict__ dst32, int planes_total,
                                 int plane_off, int planes, size_t lo16) {
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
    if (dst16) store16x8(dst16 + o, v, dtype, lo16);
    if (dst32) {
      float4* op = reinterpret_cast<float4*>(dst32 + o);
      op[0] = make_float4(v[0], v[1], v[2], v[3]);
      op[1] = make_float4(v[4], v[5], v[6], v[7]);
    }
  }
}

template <bool kIs16>
__global__ void unpack_planes_kernel(const void* __restrict__ src, int dtype, int n, int c, int h, int w,
                                     int planes_total, int plane_off, float* __restrict__ dst, size_t lo16) {
  const size_t total = (size_t)n * c * h * w;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    size_t r = idx;
    const int x = r % w; r /= w;
    const int y = r % h; r /= h;
    const int ch = r % c; r /= c;
    const int img = (int)r;
    const size_t o = ((((size_t)img * planes_total + plane_off + (ch >> 3)) * h + y) * w + x) * 8 + (ch & 7);
    if (kIs16) {
      const uint16_t* sp = reinterpret_cast<const uint16_t*>(src);
      dst[idx] = lo16 ? from16(sp[o], 1) + from16(sp[o + lo16], 1) : from16(sp[o], dtype);
    } else {
      dst[idx] = reinterpret_cast<const float*>(src)[o];
    }
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

// adjoint of nn.PixelShuffle(2) (models/modules/block.py:287) on 16-bit planes: dst channel 4c + 2dy + dx at (y, x) = src channel c at
// (2y + dy, 2x + dx).  One thread per (image, source plane, y, x): four 16 B loads (the 2x2 positions of 8 channels), four 16 B stores
// (destination planes 4P .. 4P+3).  src [N][src_pt][2h][2w][8] from plane src_po, dst [N][dst_pt][h][w][8] from plane dst_po.
__global__ void pixel_unshuffle2_kernel(const uint16_t* __restrict__ src, int n, int planes, int h, int w, int src_pt, int src_po,
                                        uint16_t* __restrict__ dst, int dst_pt, int dst_po) {
  const size_t total = (size_t)n * planes * h * w;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    size_t r = idx;
    const int x = r % w; r /= w;
    const int y = r % h; r /= h;
    const int P = r % planes; r /= planes;
    const size_t img = r;
    const uint16_t* sp = src + (((img * src_pt + src_po + P) * (2 * (size_t)h) + 2 * y) * (2 * (size_t)w) + 2 * x) * 8;
    union { uint4 q; uint16_t e[8]; } in[2][2], out;
    in[0][0].q = __ldg(reinterpret_cast<const uint4*>(sp));
    in[0][1].q = __ldg(reinterpret_cast<const uint4*>(sp + 8));
    in[1][0].q = __ldg(reinterpret_cast<const uint4*>(sp + 2 * (size_t)w * 8));
    in[1][1].q = __ldg(reinterpret_cast<const uint4*>(sp + 2 * (size_t)w * 8 + 8));
#pragma unroll
    for (int s4 = 0; s4 < 4; ++s4) {
#pragma unroll
      for (int j = 0; j < 8; ++j) out.e[j] = in[(j >> 1) & 1][j & 1].e[2 * s4 + (j >> 2)];
      *reinterpret_cast<uint4*>(dst + (((img * dst_pt + dst_po + 4 * P + s4) * (size_t)h + y) * w + x) * 8) = out.q;
    }
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
  // tiles whose filter footprint stays inside the image need no replicate-pad clamping (the common case)
  const bool interior_y = Y0 - r >= 0 && Y0 + kUpT - 1 + r <= hh - 1;
  const bool interior_x = X0 - r >= 0 && X0 + kUpT - 1 + r <= wh - 1;
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
      if (interior_x) {
        // polyphase: only every s-th tap meets a non-zero sample of the zero-stuffed row; no clamping inside the image
        const int b0 = (((phase + r - X) % s) + s) % s;
        int j = (X + b0 - r - phase) / s - jlo;
        for (int b = b0; b < len; b += s, ++j) a = fmaf(__ldg(khr + b), ft[ii * maxn + j], a);
      } else {
        for (int b = 0; b < len; ++b) {
          const int xs = clampi(X + b - r, 0, wh - 1) - phase;
          if (xs >= 0 && xs % s == 0) {
            const int j = xs / s - jlo;
            if (j >= 0 && j < nj) a = fmaf(__ldg(khr + b), ft[ii * maxn + j], a);
          }
        }
      }
      hb[ii * kUpT + xx] = a;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int Y = Y0 + ty0 + 8 * k;
      float a2 = 0.f;
      if (interior_y) {
        const int a0 = (((phase + r - Y) % s) + s) % s;
        int i = (Y + a0 - r - phase) / s - ilo;
        for (int a = a0; a < len; a += s, ++i) a2 = fmaf(__ldg(kvr + a), hb[i * kUpT + tx], a2);
      } else {
        for (int a = 0; a < len; ++a) {
          const int ys = clampi(Y + a - r, 0, hh - 1) - phase;
          if (ys >= 0 && ys % s == 0) {
            const int i = ys / s - ilo;
            if (i >= 0 && i < ni) a2 = fmaf(__ldg(kvr + a), hb[i * kUpT + tx], a2);
          }
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

// ------------------------------------------------------------------------------------------------
// Backward helpers (gradient of the same operators; used by the Z-optimisation / training paths)
// ------------------------------------------------------------------------------------------------

// Adjoint of nearest x2 (block.py:299-300): dst[y][x] = sum of the 2x2 block of src.  Optionally multiplies by the
// LeakyReLU derivative of the saved activation `act16` (given at the HIGH resolution, where it is stored 2x2
// replicated) before the 16-bit store.  src fp32 planes [nplanes][2h][2w][8] -> dst32 / dst16 planes [nplanes][h][w][8].
// Split precision: the 16-bit tensors (act16, dst16) have `planes` hi planes followed by as many lo planes per image.
__global__ void downsum2x_kernel(const float4* __restrict__ src, size_t nplanes, int h, int w, const uint4* __restrict__ act16,
                                 float slope, int dtype, float4* __restrict__ dst32, uint4* __restrict__ dst16, int planes, int split) {
  const size_t total = nplanes * h * w;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    size_t r = idx;
    const int x = r % w; r /= w;
    const int y = r % h; r /= h;
    const size_t hi = (r * (2 * (size_t)h) + 2 * y) * (2 * (size_t)w) + 2 * x;  // pixel index in the hi-res plane
    float v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        const float4* sp = src + (hi + (size_t)dy * 2 * w + dx) * 2;
        const float4 a = __ldg(sp), b = __ldg(sp + 1);
        v[0] += a.x; v[1] += a.y; v[2] += a.z; v[3] += a.w; v[4] += b.x; v[5] += b.y; v[6] += b.z; v[7] += b.w;
      }
    // plane index inside a 16-bit tensor: split tensors hold twice the planes per image
    const size_t r16 = split ? (r / planes) * (2 * (size_t)planes) + (r % planes) : r;
    if (act16) {
      const uint4 q = __ldg(act16 + (r16 * (2 * (size_t)h) + 2 * y) * (2 * (size_t)w) + 2 * x);
      const uint32_t wq[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float a = from16((uint16_t)((wq[k >> 1] >> ((k & 1) * 16)) & 0xFFFFu), dtype);
        if (!(a > 0.f)) v[k] *= slope;
      }
    }
    if (dst32) {
      dst32[idx * 2] = make_float4(v[0], v[1], v[2], v[3]);
      dst32[idx * 2 + 1] = make_float4(v[4], v[5], v[6], v[7]);
    }
    if (dst16)
      store16x8(reinterpret_cast<uint16_t*>(dst16) + ((r16 * h + y) * w + x) * 8, v, dtype, split ? (size_t)planes * h * w * 8 : 0);
  }
}

// out = a + b on fp32 planes; optional 16-bit copy
// (split precision: per_img = pixel-planes of one image in the fp32 tensors; the 16-bit output holds twice as many)
__global__ void planes_add_kernel(const float4* __restrict__ a, const float4* __restrict__ b, size_t n8, int dtype,
                                  float4* __restrict__ out32, uint4* __restrict__ out16, size_t per_img, int split) {
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < n8; idx += (size_t)gridDim.x * blockDim.x) {
    const float4 a0 = __ldg(a + 2 * idx), a1 = __ldg(a + 2 * idx + 1), b0 = __ldg(b + 2 * idx), b1 = __ldg(b + 2 * idx + 1);
    const float v[8] = {a0.x + b0.x, a0.y + b0.y, a0.z + b0.z, a0.w + b0.w, a1.x + b1.x, a1.y + b1.y, a1.z + b1.z, a1.w + b1.w};
    if (out32) {
      out32[2 * idx] = make_float4(v[0], v[1], v[2], v[3]);
      out32[2 * idx + 1] = make_float4(v[4], v[5], v[6], v[7]);
    }
    if (out16) {
      const size_t o = split ? (idx / per_img) * (2 * per_img) + (idx % per_img) : idx;
      store16x8(reinterpret_cast<uint16_t*>(out16) + o * 8, v, dtype, split ? per_img * 8 : 0);
    }
  }
}

// Adjoint of one 1-D pass of a clamp-addressed (replicate-padded) strided filter along one axis:
//   forward:  out[o] = sum_t k[t] * in[clamp(A*o + t + c, 0, n_in-1)]          (in: logical length n_in)
//   adjoint:  gin[m] = sum_t k[t] * sum_{o : clamp(A*o+t+c) == m} gout[o]
// Only the logical positions m = Am*mi + pm (mi < n_store) are produced (Am = s, pm = phase for the zero-stuffed
// input of Upscale_OP; Am = 1, pm = 0 otherwise).  `gout` may be a crop-adjoint view: logical index o maps to
// stored index o - crop, zero outside [0, n_out_store).  Layout: [imgs][outer][axis][inner] with inner = 1 for the
// x axis and inner = row length for the y axis.  Optional fused epilogue: dst = sub_from_view - result (used for
// g_G = g_out - Down^T(...)), where sub_from has the same crop convention.
__global__ void sep_adjoint_1d_kernel(const float* __restrict__ gout, int imgs, int outer, int inner, int n_out, int o_lo,
                                      int o_cnt, int n_in, int n_store, int A, int c, int Am, int pm,
                                      const float* __restrict__ k, int len, const float* __restrict__ sub_from, int sub_crop,
                                      float* __restrict__ gin) {
  const size_t total = (size_t)imgs * outer * n_store * inner;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    size_t r = idx;
    const int in_i = r % inner; r /= inner;
    const int mi = r % n_store; r /= n_store;
    const int ou = r % outer; r /= outer;
    const int img = (int)r;
    const int m = Am * mi + pm;
    // stored gout covers logical o in [o_lo, o_lo + o_cnt) (crop adjoint = zeros elsewhere)
    const float* gp = gout + (((size_t)img * outer + ou) * o_cnt) * inner + in_i;
    const int o_hi = o_lo + o_cnt;
    float acc = 0.f;
    // un-clamped hits: A*o + t + c == m.  Only every A-th tap can hit (t == (m - c) mod A): walk those, o falls by one per step -
    // no division inside the loop (the strided Down^T pass touches ceil(len / A) taps per output instead of len).
    {
      int t = (m - c) % A;
      if (t < 0) t += A;
      int o = (m - t - c) / A;            // exact
      for (; t < len && o >= o_lo; t += A, --o)
        if (o < o_hi) acc = fmaf(__ldg(k + t), __ldg(gp + (size_t)(o - o_lo) * inner), acc);
    }
    if (m == 0 || m == n_in - 1) {        // border positions also collect everything the replicate padding clamped onto them
      for (int t = 0; t < len; ++t) {
        const float kt = __ldg(k + t);
        float ssum = 0.f;
        if (m == 0) {            // everything that fell off the low end was clamped onto index 0
          for (int o = o_lo; o < o_hi && A * o + t + c < 0; ++o) ssum += __ldg(gp + (size_t)(o - o_lo) * inner);
        }
        if (m == n_in - 1) {     // ... and off the high end onto index n_in-1
          int o0 = (n_in - 1 - t - c) / A + 1;
          if (o0 < o_lo) o0 = o_lo;
          while (o0 > o_lo && A * (o0 - 1) + t + c > n_in - 1) --o0;
          for (int o = o0; o < o_hi; ++o)
            if (A * o + t + c > n_in - 1) ssum += __ldg(gp + (size_t)(o - o_lo) * inner);
        }
        acc = fmaf(kt, ssum, acc);
      }
    }
    if (sub_from) {  // x-axis pass only (inner == 1): dst = crop_adjoint(sub_from) - acc
      const int y = ou, x = mi;
      const int hs = outer - 2 * sub_crop, ws = n_store - 2 * sub_crop;
      float base = 0.f;
      if (y >= sub_crop && y < sub_crop + hs && x >= sub_crop && x < sub_crop + ws)
        base = __ldg(sub_from + ((size_t)img * hs + (y - sub_crop)) * ws + (x - sub_crop));
      acc = base - acc;
    }
    gin[idx] = acc;
  }
}

// gradient of the latent map: scatter-add (atomics) of
//   (a) the HR-resolution latent gradient given on the replicate-padded domain (planes32 [n][1][hp][wp][8], first c ch), and
//   (b) the LR-resolution latent gradient through the adjoint of the bilinear 1/s resize of the padded map,
// onto the un-padded map dst [n][c][hh][wh] (NCHW fp32, zero-initialised by the caller).
__global__ void latent_grad_hr_kernel(const float* __restrict__ gz_hr, int n, int c, int hh, int wh, int pad, float* __restrict__ dst) {
  const int hp = hh + 2 * pad, wp = wh + 2 * pad;
  const size_t total = (size_t)n * hp * wp;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    size_t r = idx;
    const int x = r % wp; r /= wp;
    const int y = r % hp; r /= hp;
    const int img = (int)r;
    const int sy = clampi(y - pad, 0, hh - 1), sx = clampi(x - pad, 0, wh - 1);
    for (int ch = 0; ch < c; ++ch)
      atomicAdd(dst + (((size_t)img * c + ch) * hh + sy) * wh + sx, gz_hr[idx * 8 + ch]);
  }
}
__global__ void latent_grad_lr_kernel(const float* __restrict__ gz_lr, int n, int c, int hh, int wh, int s, int pad,
                                      float* __restrict__ dst) {
  const int hp = hh + 2 * pad, wp = wh + 2 * pad;
  const int ho = hp / s, wo = wp / s;
  const size_t total = (size_t)n * ho * wo;
  const float fs = 1.0f / (1.0f / (float)s);
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    size_t r = idx;
    const int j = r % wo; r /= wo;
    const int i = r % ho; r /= ho;
    const int img = (int)r;
    const float sy = fmaxf(((float)i + 0.5f) * fs - 0.5f, 0.f), sx = fmaxf(((float)j + 0.5f) * fs - 0.5f, 0.f);
    const int y0 = (int)sy, x0 = (int)sx;
    const float ly = sy - (float)y0, lx = sx - (float)x0;
    const int y1 = min(y0 + 1, hp - 1), x1 = min(x0 + 1, wp - 1);
    const int ya = clampi(y0 - pad, 0, hh - 1), yb = clampi(y1 - pad, 0, hh - 1);
    const int xa = clampi(x0 - pad, 0, wh - 1), xb = clampi(x1 - pad, 0, wh - 1);
    for (int ch = 0; ch < c; ++ch) {
      const float g = gz_lr[idx * 8 + ch];
      float* p = dst + ((size_t)img * c + ch) * hh * wh;
      atomicAdd(p + (size_t)ya * wh + xa, (1.f - ly) * (1.f - lx) * g);
      atomicAdd(p + (size_t)ya * wh + xb, (1.f - ly) * lx * g);
      atomicAdd(p + (size_t)yb * wh + xa, ly * (1.f - lx) * g);
      atomicAdd(p + (size_t)yb * wh + xb, ly * lx * g);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// VGG feature extractor helpers (models/modules/architecture.py:658-724)
// ------------------------------------------------------------------------------------------------

// NCHW fp32 -> 16-bit planes with a per-channel affine map v*scale[c] + shift[c]: the ImageNet input normalisation
// (x - mean) / std of VGGFeatureExtractor.forward (:719-720).  One thread per (n, plane, y, x).
__global__ void pack_nchw_affine_kernel(const float* __restrict__ src, int n, int c, int h, int w, const float* __restrict__ scale,
                                        const float* __restrict__ shift, int dtype, uint16_t* __restrict__ dst16, int planes_total,
                                        int plane_off, int planes, size_t lo16) {
  const size_t total = (size_t)n * planes * h * w;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    size_t r = idx;
    const int x = r % w; r /= w;
    const int y = r % h; r /= h;
    const int g = r % planes; r /= planes;
    const int img = (int)r;
    float v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int ch = g * 8 + k;
      v[k] = ch < c ? fmaf(__ldg(src + (((size_t)img * c + ch) * h + y) * w + x), __ldg(scale + ch), __ldg(shift + ch)) : 0.f;
    }
    store16x8(dst16 + ((((size_t)img * planes_total + plane_off + g) * h + y) * w + x) * 8, v, dtype, lo16);
  }
}

// 2x2 / stride 2 max pooling of 16-bit planes (nn.MaxPool2d(2, 2) inside torchvision's vgg19.features).
// split precision: `planes` hi planes + as many lo planes per image (nplanes counts the hi planes of all images); a window's
// maximum is decided on hi + lo and both halves of the winner are carried over
__global__ void maxpool2x2_split_kernel(const uint16_t* __restrict__ src, size_t nplanes, int planes, int h, int w, uint16_t* __restrict__ dst) {
  const int ho = h / 2, wo = w / 2;
  const size_t total = nplanes * ho * wo;
  const size_t lo_in = (size_t)planes * h * w * 8, lo_out = (size_t)planes * ho * wo * 8;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    size_t r = idx;
    const int x = r % wo; r /= wo;
    const int y = r % ho; r /= ho;
    const size_t r16 = (r / planes) * (2 * (size_t)planes) + (r % planes);
    const uint16_t* sp = src + ((r16 * h + 2 * y) * w + 2 * x) * 8;
    float best[8], v[8];
    load16x8(sp, 1, lo_in, best);
    const size_t offs[3] = {8, (size_t)w * 8, (size_t)w * 8 + 8};
#pragma unroll
    for (int t = 0; t < 3; ++t) {
      load16x8(sp + offs[t], 1, lo_in, v);
#pragma unroll
      for (int k = 0; k < 8; ++k) best[k] = v[k] > best[k] ? v[k] : best[k];
    }
    store16x8(dst + ((r16 * ho + y) * wo + x) * 8, best, 1, lo_out);
  }
}

__global__ void maxpool2x2_kernel(const uint4* __restrict__ src, size_t nplanes, int h, int w, int dtype, uint4* __restrict__ dst) {
  const int ho = h / 2, wo = w / 2;
  const size_t total = nplanes * ho * wo;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    size_t r = idx;
    const int x = r % wo; r /= wo;
    const int y = r % ho; r /= ho;
    const uint4* sp = src + (r * h + 2 * y) * w + 2 * x;
    const uint4 q[4] = {__ldg(sp), __ldg(sp + 1), __ldg(sp + w), __ldg(sp + w + 1)};
    uint32_t pk[4] = {0, 0, 0, 0};
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      uint16_t best_bits = 0;
      float best = 0.f;
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const uint32_t wq = (k >> 1) == 0 ? q[t].x : ((k >> 1) == 1 ? q[t].y : ((k >> 1) == 2 ? q[t].z : q[t].w));
        const uint16_t bits = (uint16_t)((wq >> ((k & 1) * 16)) & 0xFFFFu);
        const float v = from16(bits, dtype);
        if (t == 0 || v > best) { best = v; best_bits = bits; }
      }
      pk[k >> 1] |= (uint32_t)best_bits << ((k & 1) * 16);
    }
    dst[idx] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
  }
}

// Backward of the pooling (+ the ReLU in front of it): the gradient of a pooled pixel goes to the FIRST maximum of its window
// in row-major order (what torch's max_pool2d backward does) if that activation is positive, zeros elsewhere.
// gout: 16-bit planes [nplanes][h/2][w/2], act: the pooling's input (post-ReLU) [nplanes][h][w]; gin like act.
__global__ void maxpool2x2_bwd_kernel(const uint4* __restrict__ gout, const uint4* __restrict__ act, size_t nplanes, int h, int w, int dtype,
                                      uint4* __restrict__ gin) {
  const int ho = h / 2, wo = w / 2;
  const size_t total = nplanes * ho * wo;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    size_t r = idx;
    const int x = r % wo; r /= wo;
    const int y = r % ho; r /= ho;
    const size_t base = (r * h + 2 * y) * w + 2 * x;
    const uint4 q[4] = {__ldg(act + base), __ldg(act + base + 1), __ldg(act + base + w), __ldg(act + base + w + 1)};
    const uint4 g = __ldg(gout + idx);
    uint32_t o[4][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}, {0, 0, 0, 0}, {0, 0, 0, 0}};
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      int arg = 0;
      float best = 0.f;
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const uint32_t wq = (k >> 1) == 0 ? q[t].x : ((k >> 1) == 1 ? q[t].y : ((k >> 1) == 2 ? q[t].z : q[t].w));
        const float v = from16((uint16_t)((wq >> ((k & 1) * 16)) & 0xFFFFu), dtype);
        if (t == 0 || v > best) { best = v; arg = t; }
      }
      const uint32_t gw = (k >> 1) == 0 ? g.x : ((k >> 1) == 1 ? g.y : ((k >> 1) == 2 ? g.z : g.w));
      const uint32_t gb = best > 0.f ? ((gw >> ((k & 1) * 16)) & 0xFFFFu) : 0u;
#pragma unroll
      for (int t = 0; t < 4; ++t)
        if (t == arg) o[t][k >> 1] |= gb << ((k & 1) * 16);
    }
    gin[base] = make_uint4(o[0][0], o[0][1], o[0][2], o[0][3]);
    gin[base + 1] = make_uint4(o[1][0], o[1][1], o[1][2], o[1][3]);
    gin[base + w] = make_uint4(o[2][0], o[2][1], o[2][2], o[2][3]);
    gin[base + w + 1] = make_uint4(o[3][0], o[3][1], o[3][2], o[3][3]);
  }
}

// split-precision variant of the pooling backward: argmax on hi + lo of the saved activation, the gradient's hi and lo halves
// both go to the winner
__global__ void maxpool2x2_bwd_split_kernel(const uint16_t* __restrict__ gout, const uint16_t* __restrict__ act, size_t nplanes, int planes,
                                            int h, int w, uint16_t* __restrict__ gin) {
  const int ho = h / 2, wo = w / 2;
  const size_t total = nplanes * ho * wo;
  const size_t lo_in = (size_t)planes * h * w * 8, lo_out = (size_t)planes * ho * wo * 8;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    size_t r = idx;
    const int x = r % wo; r /= wo;
    const int y = r % ho; r /= ho;
    const size_t r16 = (r / planes) * (2 * (size_t)planes) + (r % planes);
    const size_t base = ((r16 * h + 2 * y) * w + 2 * x) * 8;
    const size_t offs[4] = {0, 8, (size_t)w * 8, (size_t)w * 8 + 8};
    float a[4][8], g[8];
#pragma unroll
    for (int t = 0; t < 4; ++t) load16x8(act + base + offs[t], 1, lo_in, a[t]);
    load16x8(gout + ((r16 * ho + y) * wo + x) * 8, 1, lo_out, g);
    float o[4][8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      int arg = 0;
      float best = a[0][k];
#pragma unroll
      for (int t = 1; t < 4; ++t)
        if (a[t][k] > best) { best = a[t][k]; arg = t; }
#pragma unroll
      for (int t = 0; t < 4; ++t) o[t][k] = (t == arg && best > 0.f) ? g[k] : 0.f;
    }
#pragma unroll
    for (int t = 0; t < 4; ++t) store16x8(gin + base + offs[t], o[t], 1, lo_in);
  }
}

}  // namespace esr
