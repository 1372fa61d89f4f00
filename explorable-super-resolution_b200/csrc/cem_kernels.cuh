// CEM projection, HBM-bound fast path for rank-1 (separable) filters: out = G + Up(Inv(x - Down(G)))   (CEM/CEMnet.py:303-310 by
// linearity; the reference runs five dense depth-wise convs on zero-stuffed / un-decimated HR tensors plus three pad copies).
//
//   cem_down_fast   : e = x - Down(G)      reads G once (float4, halo 1.3x through L2), writes the LR residual
//   cem_inv_fast    : f = Inv(e)           27-tap separable on the LR grid (16x fewer pixels than HR)
//   cem_up_add_fast : out = G + Up(f)      polyphase (every s-th tap meets a sample of the zero-stuffed image), float4 G read and
//                                          float4 store, 128 x 32 output tiles, optional crop (eval mode's HR_unpadder)
// Algorithmic bytes per HR pixel (fp32, x4): 12 (G) + 12 (G again, for the add) + 12 (out) + 5 x 0.75 (LR tensors) = 39.75 B for the
// three-kernel form against 24.75 B for a (not tile-able: 68-pixel receptive radius) single pass.
// Interior tiles take index-arithmetic-free paths; tiles touching the image border reproduce the replicate padding of the
// reference's three pads by clamped addressing (same code path as the general kernels in aux_kernels.cuh).
#pragma once
#include "aux_kernels.cuh"

namespace esr {

constexpr int kDnTJ = 32, kDnTI = 16;      // LR output tile of the down kernel
constexpr int kInvTJ = 64, kInvTI = 32;    // LR tile of the inverse filter
constexpr int kUpTX = 128, kUpTY = 32;     // HR output tile of the up kernel
constexpr int kCemMaxTaps = 64;

// e[i][j] = sub_from[i][j] - sum_{a,b} kv[a] kh[b] G[clamp(s i + phase + a - r)][clamp(s j + phase + b - r)]     (sub_from optional)
// S and LEN are compile-time (bicubic: (2,9) (3,11) (4,17) (8,33)): the taps live in registers and both passes are register-blocked -
// a thread computes 4 horizontally adjacent / 2 vertically adjacent outputs from one sliding window, 3-5x fewer shared-memory loads
// than one output per thread (the un-blocked kernel was LSU-bound at 1.2 TB/s).
template <int S, int LEN>
__global__ void __launch_bounds__(256)
cem_down_fast_kernel(const float* __restrict__ g, int hh, int wh, int phase, const float* __restrict__ kv, const float* __restrict__ kh,
                     const float* __restrict__ sub_from, float* __restrict__ out) {
  extern __shared__ float sm[];
  constexpr int rows = (kDnTI - 1) * S + LEN, cols = (kDnTJ - 1) * S + LEN;
  constexpr int pitch = ((cols + 3 + 3) & ~3) | 1;  // room for the alignment shift; odd: the strided row walk is bank-conflict free
  constexpr int hp = kDnTJ + 1;
  const int hl = hh / S, wl = wh / S;
  float* tile = sm;                                 // [rows][pitch]
  float* hbuf = sm + rows * pitch;                  // [rows][hp]
  const int nc = blockIdx.z;
  const int i0 = blockIdx.y * kDnTI, j0 = blockIdx.x * kDnTJ;
  constexpr int r = LEN / 2;
  const float* gp = g + (size_t)nc * hh * wh;
  const int ybase = S * i0 + phase - r, xbase = S * j0 + phase - r;
  const int xa = xbase & ~3;                        // 16-byte aligned start of the window (two's complement floor for negative xbase)
  const int shift = xbase - xa;                     // 0..3
  const int nvec = (shift + cols + 3) >> 2;
  const bool interior = ybase >= 0 && ybase + rows <= hh && xa >= 0 && xa + 4 * nvec <= wh && (wh & 3) == 0 && (((uintptr_t)gp) & 15) == 0;
  if (interior) {
    const int total = rows * nvec;
    for (int e0 = threadIdx.x; e0 < total; e0 += 256 * 4) {      // four 16-byte loads in flight per thread
      float4 q[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int e = e0 + k * 256;
        if (e < total) {
          const int rr = e / nvec, v = e - rr * nvec;
          q[k] = __ldg(reinterpret_cast<const float4*>(gp + (size_t)(ybase + rr) * wh + xa) + v);
        }
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int e = e0 + k * 256;
        if (e < total) {
          const int rr = e / nvec, v = e - rr * nvec;
          float* d = tile + rr * pitch + 4 * v;
          d[0] = q[k].x; d[1] = q[k].y; d[2] = q[k].z; d[3] = q[k].w;
        }
      }
    }
  } else {
    for (int e = threadIdx.x; e < rows * cols; e += 256) {
      const int rr = e / cols, cc = e - rr * cols;
      tile[rr * pitch + shift + cc] = __ldg(gp + (size_t)clampi(ybase + rr, 0, hh - 1) * wh + clampi(xbase + cc, 0, wh - 1));
    }
  }
  float th[LEN], tv[LEN];
#pragma unroll
  for (int b = 0; b < LEN; ++b) { th[b] = __ldg(kh + b); tv[b] = __ldg(kv + b); }
  __syncthreads();
  // horizontal pass: a thread produces 4 adjacent outputs of one row from a window of 3 S + LEN inputs
  constexpr int HB = 4, quads = kDnTJ / HB;
  for (int e = threadIdx.x; e < rows * quads; e += 256) {
    const int qd = e / rows, rr = e - qd * rows;
    const float* tp = tile + rr * pitch + shift + qd * HB * S;
    float win[(HB - 1) * S + LEN];
#pragma unroll
    for (int k = 0; k < (HB - 1) * S + LEN; ++k) win[k] = tp[k];
    float acc[HB];
#pragma unroll
    for (int o = 0; o < HB; ++o) {
      acc[o] = 0.f;
#pragma unroll
      for (int b = 0; b < LEN; ++b) acc[o] = fmaf(th[b], win[o * S + b], acc[o]);
      hbuf[rr * hp + qd * HB + o] = acc[o];
    }
  }
  __syncthreads();
  // vertical pass: a thread produces 2 vertically adjacent outputs of one column
  constexpr int VB = 2;
  const int tj = threadIdx.x & (kDnTJ - 1);
  for (int tp2 = threadIdx.x / kDnTJ; tp2 < kDnTI / VB; tp2 += 256 / kDnTJ) {
    const int ti = tp2 * VB;
    const float* hq = hbuf + (ti * S) * hp + tj;
    float win[(VB - 1) * S + LEN];
#pragma unroll
    for (int k = 0; k < (VB - 1) * S + LEN; ++k) win[k] = hq[k * hp];
#pragma unroll
    for (int o = 0; o < VB; ++o) {
      float acc = 0.f;
#pragma unroll
      for (int a = 0; a < LEN; ++a) acc = fmaf(tv[a], win[o * S + a], acc);
      const int i = i0 + ti + o, j = j0 + tj;
      if (i < hl && j < wl) {
        const size_t oo = (size_t)nc * hl * wl + (size_t)i * wl + j;
        out[oo] = sub_from ? (__ldg(sub_from + oo) - acc) : acc;
      }
    }
  }
}

// f = replicate-padded correlation of e with kv (x) kh on the LR grid; LEN compile-time, 8 outputs per thread in both passes
template <int LEN>
__global__ void __launch_bounds__(256)
cem_inv_fast_kernel(const float* __restrict__ e_in, int hl, int wl, const float* __restrict__ kv, const float* __restrict__ kh,
                    float* __restrict__ out) {
  extern __shared__ float sm[];
  constexpr int rows = kInvTI + LEN - 1, cols = kInvTJ + LEN - 1;
  constexpr int pitch = cols | 1;
  constexpr int hp = kInvTJ + 1;
  float* tile = sm;                        // [rows][pitch]
  float* hbuf = sm + rows * pitch;         // [rows][hp]
  const int nc = blockIdx.z;
  const int i0 = blockIdx.y * kInvTI, j0 = blockIdx.x * kInvTJ;
  constexpr int r = LEN / 2;
  const float* ep = e_in + (size_t)nc * hl * wl;
  for (int e = threadIdx.x; e < rows * cols; e += 256) {
    const int rr = e / cols, cc = e - rr * cols;
    tile[rr * pitch + cc] = __ldg(ep + (size_t)clampi(i0 - r + rr, 0, hl - 1) * wl + clampi(j0 - r + cc, 0, wl - 1));
  }
  float th[LEN], tv[LEN];
#pragma unroll
  for (int b = 0; b < LEN; ++b) { th[b] = __ldg(kh + b); tv[b] = __ldg(kv + b); }
  __syncthreads();
  constexpr int B8 = 8;
  for (int e = threadIdx.x; e < rows * (kInvTJ / B8); e += 256) {       // consecutive threads walk down the rows (odd pitch)
    const int oc = e / rows, rr = e - oc * rows;
    const float* tp = tile + rr * pitch + oc * B8;
    float win[B8 + LEN - 1];
#pragma unroll
    for (int k = 0; k < B8 + LEN - 1; ++k) win[k] = tp[k];
#pragma unroll
    for (int o = 0; o < B8; ++o) {
      float a = 0.f;
#pragma unroll
      for (int b = 0; b < LEN; ++b) a = fmaf(th[b], win[o + b], a);
      hbuf[rr * hp + oc * B8 + o] = a;
    }
  }
  __syncthreads();
  const int tj = threadIdx.x & (kInvTJ - 1);
  for (int tg = threadIdx.x / kInvTJ; tg < kInvTI / B8; tg += 256 / kInvTJ) {
    const int ti = tg * B8;
    const float* hq = hbuf + ti * hp + tj;
    float win[B8 + LEN - 1];
#pragma unroll
    for (int k = 0; k < B8 + LEN - 1; ++k) win[k] = hq[k * hp];
#pragma unroll
    for (int o = 0; o < B8; ++o) {
      float acc = 0.f;
#pragma unroll
      for (int a = 0; a < LEN; ++a) acc = fmaf(tv[a], win[o + a], acc);
      const int i = i0 + ti + o, j = j0 + tj;
      if (i < hl && j < wl) out[(size_t)nc * hl * wl + (size_t)i * wl + j] = acc;
    }
  }
}

// out[Y - crop][X - crop] = G[Y][X] + sum_{a,b} kv[a] kh[b] S[clamp(Y + a - r)][clamp(X + b - r)],  S = F zero-stuffed at `phase`
// (S[S_ i + phase][S_ j + phase] = F[i][j]).  Polyphase: of a sample's `len` taps only every S_-th one meets a non-zero entry of S.
// Replicate padding of S (CEMnet.py:268-272 pads the zero-stuffed image): every tap that leaves the image reads S's border row /
// column, which is F's first / last row / column when the border index is on the sampling lattice and zero otherwise - so a border
// sample is the in-range polyphase walk plus (sum of the taps that left) x (that border entry).  No per-tap clamping anywhere.
// S_ is compile-time (2, 3, 4, 8): the lattice arithmetic is shifts and masks.
template <int S_>
__device__ __forceinline__ float up_axis(const float* __restrict__ taps, int len, int r, int phase, int pos, int n_hr, int n_lr, int lo,
                                         const float* __restrict__ samples, int stride) {
  // taps: shared memory; samples[(j - lo) * stride] = the LR entries of this line (window starts at LR index `lo`)
  const int b_lo = max(0, r - pos), b_hi = min(len - 1, n_hr - 1 - pos + r);
  int m = (phase + r - pos - b_lo) % S_;
  if (m < 0) m += S_;
  int b = b_lo + m;
  int j = (pos + b - r - phase) / S_ - lo;           // exact: pos + b - r - phase is a non-negative multiple of S_
  float a = 0.f;
  for (; b <= b_hi; b += S_, ++j) a = fmaf(taps[b], samples[j * stride], a);
  if (b_lo > 0 && phase == 0) {                       // taps that fell off the low end read S[0] = F[0]
    float t = 0.f;
    for (int k = 0; k < b_lo; ++k) t += taps[k];
    a = fmaf(t, samples[(0 - lo) * stride], a);
  }
  if (b_hi < len - 1 && (n_hr - 1 - phase) % S_ == 0) {   // ... off the high end: S[n_hr - 1] = F[n_lr - 1]
    float t = 0.f;
    for (int k = b_hi + 1; k < len; ++k) t += taps[k];
    a = fmaf(t, samples[(n_lr - 1 - lo) * stride], a);
  }
  return a;
}

// the same walk for four adjacent columns at once (vertical pass: the taps and LR rows depend on the output row only)
template <int S_>
__device__ __forceinline__ float4 up_axis4(const float* __restrict__ taps, int len, int r, int phase, int pos, int n_hr, int n_lr, int lo,
                                           const float* __restrict__ samples, int stride) {
  const int b_lo = max(0, r - pos), b_hi = min(len - 1, n_hr - 1 - pos + r);
  int m = (phase + r - pos - b_lo) % S_;
  if (m < 0) m += S_;
  int b = b_lo + m;
  int j = (pos + b - r - phase) / S_ - lo;
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
  for (; b <= b_hi; b += S_, ++j) {
    const float t = taps[b];
    const float4 v = *reinterpret_cast<const float4*>(samples + j * stride);
    a.x = fmaf(t, v.x, a.x); a.y = fmaf(t, v.y, a.y); a.z = fmaf(t, v.z, a.z); a.w = fmaf(t, v.w, a.w);
  }
  if (b_lo > 0 && phase == 0) {
    float t = 0.f;
    for (int k = 0; k < b_lo; ++k) t += taps[k];
    const float4 v = *reinterpret_cast<const float4*>(samples + (0 - lo) * stride);
    a.x = fmaf(t, v.x, a.x); a.y = fmaf(t, v.y, a.y); a.z = fmaf(t, v.z, a.z); a.w = fmaf(t, v.w, a.w);
  }
  if (b_hi < len - 1 && (n_hr - 1 - phase) % S_ == 0) {
    float t = 0.f;
    for (int k = b_hi + 1; k < len; ++k) t += taps[k];
    const float4 v = *reinterpret_cast<const float4*>(samples + (n_lr - 1 - lo) * stride);
    a.x = fmaf(t, v.x, a.x); a.y = fmaf(t, v.y, a.y); a.z = fmaf(t, v.z, a.z); a.w = fmaf(t, v.w, a.w);
  }
  return a;
}

template <int S_>
__global__ void __launch_bounds__(256)
cem_up_add_fast_kernel(const float* __restrict__ f, const float* __restrict__ g, int hl, int wl, int phase, const float* __restrict__ kv,
                       const float* __restrict__ kh, int len, int crop, float* __restrict__ out) {
  extern __shared__ float sm[];
  __shared__ float tv[kCemMaxTaps], th[kCemMaxTaps];
  const int hh = hl * S_, wh = wl * S_;
  const int ho = hh - 2 * crop, wo = wh - 2 * crop;
  const int r = len / 2;
  const int nc = blockIdx.z;
  const int Y0 = blockIdx.y * kUpTY + crop, X0 = blockIdx.x * kUpTX + crop;     // tile origin in the un-cropped HR domain
  const int maxni = (kUpTY + len) / S_ + 2, maxnj = (kUpTX + len) / S_ + 2;
  float* ft = sm;                           // [maxni][maxnj]   LR window
  float* hb = sm + maxni * maxnj;           // [maxni][kUpTX]   horizontally filtered rows
  if (threadIdx.x < len) { tv[threadIdx.x] = __ldg(kv + threadIdx.x); th[threadIdx.x] = __ldg(kh + threadIdx.x); }
  // LR rows / columns that can contribute to this tile (clamped to the image: border entries included)
  const int ylo = clampi(Y0 - r, 0, hh - 1), yhi = clampi(Y0 + kUpTY - 1 + r, 0, hh - 1);
  const int xlo = clampi(X0 - r, 0, wh - 1), xhi = clampi(X0 + kUpTX - 1 + r, 0, wh - 1);
  const int ilo = max((ylo - phase + S_ - 1) / S_, 0), ihi = min((yhi - phase) / S_, hl - 1);
  const int jlo = max((xlo - phase + S_ - 1) / S_, 0), jhi = min((xhi - phase) / S_, wl - 1);
  const int ni = max(ihi - ilo + 1, 0), nj = max(jhi - jlo + 1, 0);
  // the thread's four G vectors go in flight first: their latency hides behind the filter passes
  const int xg = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int X = X0 + 4 * xg;
  const bool vec = (crop & 3) == 0 && (wo & 3) == 0 && (wh & 3) == 0 && X + 3 < wh && (X - crop) + 3 < wo &&
                   ((((uintptr_t)out) | (g ? (uintptr_t)g : 0)) & 15) == 0;
  float4 gq[kUpTY / 8];
#pragma unroll
  for (int k = 0; k < kUpTY / 8; ++k) {
    const int Y = Y0 + ty + 8 * k;
    gq[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (vec && g && Y < hh && Y - crop < ho) gq[k] = __ldg(reinterpret_cast<const float4*>(g + (size_t)nc * hh * wh + (size_t)Y * wh + X));
  }
  const float* fp = f + (size_t)nc * hl * wl;
  for (int ii = threadIdx.x / 64; ii < ni; ii += 4)
    for (int jj = threadIdx.x & 63; jj < nj; jj += 64) ft[ii * maxnj + jj] = __ldg(fp + (size_t)(ilo + ii) * wl + (jlo + jj));
  __syncthreads();
  // horizontal pass: one column per thread pair
  {
    const int xx = threadIdx.x & (kUpTX - 1), half = threadIdx.x / kUpTX;
    const int Xc = X0 + xx;
    if (Xc < wh)
      for (int ii = half; ii < ni; ii += 256 / kUpTX) hb[ii * kUpTX + xx] = up_axis<S_>(th, len, r, phase, Xc, wh, wl, jlo, ft + ii * maxnj, 1);
  }
  __syncthreads();
  // vertical pass + G + store: a thread owns 4 consecutive X of rows ty, ty+8, ty+16, ty+24
#pragma unroll
  for (int k = 0; k < kUpTY / 8; ++k) {
    const int Y = Y0 + ty + 8 * k;
    const int yo = Y - crop;
    if (yo >= ho || Y >= hh) continue;
    const float4 a4 = up_axis4<S_>(tv, len, r, phase, Y, hh, hl, ilo, hb + 4 * xg, kUpTX);
    const float acc[4] = {a4.x, a4.y, a4.z, a4.w};
    const float* gp = g ? g + (size_t)nc * hh * wh + (size_t)Y * wh + X : nullptr;
    float* op = out + (size_t)nc * ho * wo + (size_t)yo * wo + (X - crop);
    if (vec) {
      *reinterpret_cast<float4*>(op) = make_float4(acc[0] + gq[k].x, acc[1] + gq[k].y, acc[2] + gq[k].z, acc[3] + gq[k].w);
    } else {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (X + q < wh && (X + q - crop) < wo) op[q] = acc[q] + (gp ? __ldg(gp + q) : 0.f);
      }
    }
  }
}

}  // namespace esr
