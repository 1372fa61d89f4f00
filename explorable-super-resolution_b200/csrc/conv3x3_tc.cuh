// 3x3 / stride 1 / zero-pad 1 convolution as a tcgen05 implicit GEMM (sm_100a).
//
//   GEMM view:  D[pixel, cout] = sum_{tap, cin} A[pixel + tap, cin] * W[tap, cin, cout]
//
// Data layout in HBM: planar-8 ("planes")  [N][C/8][H][W][8] 16-bit.  One pixel of one plane is exactly
// one 16-byte row of a no-swizzle K-major UMMA core matrix, so
//   * a TMA box (8ch*P px as ONE 512-byte inner dimension, R+2 rows, kcp planes) lands in shared memory as kcp
//     halo planes whose pixels sit at a 16-byte pitch (a 16-byte inner box would make TMA fetch a whole 32-byte
//     sector per pixel: measured 2x over-fetch on the L2->SM path);
//   * a 128-row MMA operand tile is 128 consecutive pixels of the flattened (row pitch P) halo plane, and
//     the nine filter taps are nine *address offsets* ((dy*P + dx) * 16 B) into the same halo tile — the
//     activation tile is fetched from L2 once per CTA tile, not once per tap;
//   * the dense-block concatenation (block.py:234) is a plane offset, never a copy.
// Columns tx >= P-2 of each flattened row are junk (they wrap into the next row) and are discarded by the
// epilogue: M efficiency (P-2)/P.
//
// CTA = 12 warps: warp 0 = TMA producer, warp 1 = MMA issuer (+TMEM owner), warps 4..11 = epilogue; setmaxnreg moves
// the registers the first warpgroup does not need (56/thread) to the epilogue warpgroups (224/thread).
// Persistent over (tile, n-block) work items; TMEM accumulators are double buffered so the epilogue of
// item i overlaps the MMAs of item i+1.
#pragma once
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include "ptx.cuh"

namespace esr {

constexpr int kConvThreads = 384;  // warpgroup 0: TMA warp, MMA warp (+2 idle); warpgroups 1,2: 8 epilogue warps
constexpr int kMaxStages = 8;
constexpr uint32_t kSmemHeader = 1024;  // barriers + tmem pointer
constexpr uint32_t kASlack = 128;       // junk rows of the last M tile may read a few pixels past the last plane

struct ConvParams {
  // geometry
  int n, h, w;
  int P, TW, R, MT;
  int tiles_x, tiles_y, num_tiles;
  int kcp, nchunks, in_plane_off;
  int nb_n;       // MMA N (couts per n-block)
  int n_blocks;
  uint32_t plane_stride;  // (R+2)*P*16
  uint32_t a_bytes;       // kcp*plane_stride (TMA transaction bytes)
  uint32_t b_bytes;       // 9*kcp*nb_n*16
  uint32_t a_alloc;       // a_bytes + slack, multiple of 128
  uint32_t stage_bytes;
  int stages;
  int w_resident;         // 1: all K chunks of the weights are loaded once per CTA and stay in shared memory
  uint32_t w_bytes;       // nchunks*b_bytes (resident mode)
  uint32_t idesc;
  uint32_t tmem_cols;
  // row-streaming kernel (conv3x3_rows.cuh): work = n*strips*h output rows of 128 pixels, split into `ranges`
  // contiguous ranges (one per CTA and n-block); accumulators live in a ring of `slots` TMEM slots of nb_n columns
  const uint8_t* in; int in_pt;
  int strips, ranges, slots, cin_planes, issuers;
  long long units;
  uint32_t idesc_n[3];    // instruction descriptors for N = 1, 2, 3 x nb_n
  const uint8_t* wts;
  const float* bias;
  int cout;
  // epilogue
  int dtype;
  int lrelu;
  float slope, alpha;
  const void* res1; int res1_is16, res1_pt, res1_po; float beta1;
  const float* res2; int res2_pt, res2_po; float beta2;
  const float* res3; int res3_pt, res3_po; float beta3;
  // dgrad helpers: the first `lead_planes` output planes (latent channels) are accumulated (+=) into lead_acc and take
  // no other part in the epilogue; every other plane index below is relative to the first non-lead plane.
  int lead_planes; float* lead_acc; int lead_pt;
  // LeakyReLU derivative: planes >= tail_first are multiplied by (act > 0 ? 1 : mask_slope) before the 16-bit store;
  // the 16-bit store itself is limited to planes >= tail_first.
  const uint16_t* mask16; int mask_pt, mask_po; float mask_slope; int tail_first;
  uint16_t* out16; int out16_pt, out16_po, out16_up2, out16_ps;
  float* out32; int out32_pt, out32_po;
  float* out_nchw; int out_nchw_c;
  // split precision (esr_dtype ESR_BF16X3): every 16-bit tensor carries bf16 hi planes and, `*_lo` elements further, the bf16
  // residual v - hi ("lo") planes; the K loop runs over three plane segments (hi, lo, hi) against weights packed as
  // (w_hi, w_hi, w_lo), i.e. x*w ~= x_hi*w_hi + x_lo*w_hi + x_hi*w_lo with fp32 accumulation: 16 mantissa bits per operand on
  // the kind::f16 tensor pipe.  cps = K chunks per segment, seg_base = first input plane of each segment.
  int split, cps;
  int seg_base[3];
  size_t out16_lo, res1_lo;
};

// first input plane of K chunk c (K = planes per chunk of the calling kernel)
__device__ __forceinline__ int conv_chunk_plane(const ConvParams& p, int c, int K) {
  if (!p.split) return p.in_plane_off + c * K;
  const int s = c / p.cps;
  return p.seg_base[s] + (c - s * p.cps) * K;
}

__device__ __forceinline__ uint32_t pack2(float a, float b, int dtype) {
  if (dtype == 0) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  } else {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
}

// P (tile pitch), KCP (planes per K chunk) and NBN (MMA N) are compile-time so that every operand descriptor
// of the 9 x KCP/2 MMAs of a (chunk, M tile) is the chunk's base descriptor plus an immediate: the single
// issuing thread spends one 64-bit add per operand per MMA instead of rebuilding descriptors.
// kBwd selects the epilogue: false = forward (bias/LeakyReLU/two residuals, 32-channel units), true = adds the dgrad
// features (latent lead planes, LeakyReLU-derivative mask, third residual) with 16-channel units to stay in registers.
__device__ __forceinline__ void unpack8(const uint4& q, int dtype, float (&a)[8]) {
  const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (dtype == 0) {
      const __half2 h = *reinterpret_cast<const __half2*>(&w[k]);
      const float2 f = __half22float2(h);
      a[2 * k] = f.x; a[2 * k + 1] = f.y;
    } else {
      const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&w[k]);
      const float2 f = __bfloat1622float2(h);
      a[2 * k] = f.x; a[2 * k + 1] = f.y;
    }
  }
}

// 8 fp32 values -> 8 16-bit values (`hi`); split mode also returns the bf16 residuals v - float(hi) (`lo`)
__device__ __forceinline__ void pack8_hl(const float (&v)[8], int dtype, int split, uint4& hi, uint4& lo) {
  hi.x = pack2(v[0], v[1], dtype); hi.y = pack2(v[2], v[3], dtype); hi.z = pack2(v[4], v[5], dtype); hi.w = pack2(v[6], v[7], dtype);
  if (split) {
    float h[8];
    unpack8(hi, 1, h);
    lo.x = pack2(v[0] - h[0], v[1] - h[1], 1); lo.y = pack2(v[2] - h[2], v[3] - h[3], 1);
    lo.z = pack2(v[4] - h[4], v[5] - h[5], 1); lo.w = pack2(v[6] - h[6], v[7] - h[7], 1);
  }
}
// a[k] += lo part of a split 16-bit tensor (8 channels at `ptr`, the lo planes `lo_elems` further)
__device__ __forceinline__ void add_lo8(const uint16_t* ptr, size_t lo_elems, float (&a)[8]) {
  float l[8];
  unpack8(__ldg(reinterpret_cast<const uint4*>(ptr + lo_elems)), 1, l);
#pragma unroll
  for (int k = 0; k < 8; ++k) a[k] += l[k];
}

// Epilogue of one accumulator row of one thread (= one output pixel): TMEM -> registers -> bias / LeakyReLU /
// residuals -> stores.  `trow` addresses the thread's TMEM lane and the first of the NBN accumulator columns.
// Shared by the tile kernel below and the row-streaming kernel (conv3x3_rows.cuh).
template <int NBN, bool kBwd>
__device__ __forceinline__ void conv_epilogue_px(const ConvParams& p, uint32_t trow, int img, int y, int x, bool valid, int nblk) {
  constexpr int UW = (!kBwd && NBN % 32 == 0) ? 32 : 16;  // dgrad prefetches more (3 residuals + mask): smaller units
  const size_t hw = (size_t)p.h * p.w;
  const size_t pix = (size_t)y * p.w + x;
#pragma unroll
  for (int u = 0; u < NBN / UW; ++u) {
    uint32_t r[UW / 16][16];
#pragma unroll
    for (int k = 0; k < UW / 16; ++k) tmem_ld16(trow + u * UW + k * 16, r[k]);
    const int chu = nblk * NBN + u * UW;   // first conv output channel of this unit
    const int gu = chu >> 3;               // its plane index in the conv's output
    constexpr int G = UW / 8;              // groups of 8 channels in a unit
    uint4 q1[G], qm[G];                    // residual 1 as 16-bit planes, activation for the LeakyReLU mask
    float4 f1[G][2], f2[G][2], f3[G][2], bb[G][2];
    const bool on = valid;
    if (on) {
#pragma unroll
      for (int g = 0; g < G; ++g) {
        if (chu + g * 8 < p.cout) {
          if (!kBwd) {  // forward: bias prefetched with the residuals (dgrad has no bias and fewer spare registers)
            bb[g][0] = __ldg(reinterpret_cast<const float4*>(p.bias + chu + g * 8));
            bb[g][1] = __ldg(reinterpret_cast<const float4*>(p.bias + chu + g * 8 + 4));
          }
          const int gp = gu + g - (kBwd ? p.lead_planes : 0);
          if (kBwd && gp < 0) {
            const float4* rp = reinterpret_cast<const float4*>(p.lead_acc + (((size_t)img * p.lead_pt + gu + g) * hw + pix) * 8);
            f2[g][0] = rp[0]; f2[g][1] = rp[1];
            continue;
          }
          if (p.res1) {
            if (p.res1_is16) {
              q1[g] = __ldg(reinterpret_cast<const uint4*>(
                  reinterpret_cast<const uint16_t*>(p.res1) + (((size_t)img * p.res1_pt + p.res1_po + gp) * hw + pix) * 8));
            } else {
              const float4* rp = reinterpret_cast<const float4*>(
                  reinterpret_cast<const float*>(p.res1) + (((size_t)img * p.res1_pt + p.res1_po + gp) * hw + pix) * 8);
              f1[g][0] = __ldg(rp); f1[g][1] = __ldg(rp + 1);
            }
          }
          if (p.res2) {
            // plain loads: res2 may alias out32 (in-place accumulation of gradients)
            const float4* rp = reinterpret_cast<const float4*>(
                p.res2 + (((size_t)img * p.res2_pt + p.res2_po + gp) * hw + pix) * 8);
            if (kBwd) { f2[g][0] = rp[0]; f2[g][1] = rp[1]; }
            else { f2[g][0] = __ldg(rp); f2[g][1] = __ldg(rp + 1); }
          }
          if (kBwd && p.res3) {
            const float4* rp = reinterpret_cast<const float4*>(
                p.res3 + (((size_t)img * p.res3_pt + p.res3_po + gp) * hw + pix) * 8);
            f3[g][0] = __ldg(rp); f3[g][1] = __ldg(rp + 1);
          }
          if (kBwd && p.mask16 && gp >= p.tail_first) {
            qm[g] = __ldg(reinterpret_cast<const uint4*>(p.mask16 + (((size_t)img * p.mask_pt + p.mask_po + gp) * hw + pix) * 8));
          }
        }
      }
    }
    tc_wait_ld();
    if (on) {
#pragma unroll
      for (int g = 0; g < G; ++g) {
        const int ch0 = chu + g * 8;
        if (ch0 >= p.cout) continue;
        const int gp = gu + g - (kBwd ? p.lead_planes : 0);
        const uint32_t* rg = &r[g / 2][(g & 1) * 8];
        float v[8];
        {
          float4 b0, b1;
          if (kBwd) { b0 = make_float4(0.f, 0.f, 0.f, 0.f); b1 = b0; }   // transposed convs carry no bias
          else { b0 = bb[g][0]; b1 = bb[g][1]; }
          v[0] = __uint_as_float(rg[0]) + b0.x; v[1] = __uint_as_float(rg[1]) + b0.y;
          v[2] = __uint_as_float(rg[2]) + b0.z; v[3] = __uint_as_float(rg[3]) + b0.w;
          v[4] = __uint_as_float(rg[4]) + b1.x; v[5] = __uint_as_float(rg[5]) + b1.y;
          v[6] = __uint_as_float(rg[6]) + b1.z; v[7] = __uint_as_float(rg[7]) + b1.w;
        }
        if (kBwd && gp < 0) {  // latent planes: lead_acc += alpha * acc
          float4* op = reinterpret_cast<float4*>(p.lead_acc + (((size_t)img * p.lead_pt + gu + g) * hw + pix) * 8);
          op[0] = make_float4(fmaf(p.alpha, v[0], f2[g][0].x), fmaf(p.alpha, v[1], f2[g][0].y), fmaf(p.alpha, v[2], f2[g][0].z),
                              fmaf(p.alpha, v[3], f2[g][0].w));
          op[1] = make_float4(fmaf(p.alpha, v[4], f2[g][1].x), fmaf(p.alpha, v[5], f2[g][1].y), fmaf(p.alpha, v[6], f2[g][1].z),
                              fmaf(p.alpha, v[7], f2[g][1].w));
          continue;
        }
        if (p.lrelu) {
#pragma unroll
          for (int k = 0; k < 8; ++k) v[k] = v[k] > 0.f ? v[k] : v[k] * p.slope;
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] *= p.alpha;
        if (p.res1) {
          float a[8];
          if (p.res1_is16) {
            unpack8(q1[g], p.dtype, a);
            if (p.split)
              add_lo8(reinterpret_cast<const uint16_t*>(p.res1) + (((size_t)img * p.res1_pt + p.res1_po + gp) * hw + pix) * 8, p.res1_lo, a);
          } else {
            a[0] = f1[g][0].x; a[1] = f1[g][0].y; a[2] = f1[g][0].z; a[3] = f1[g][0].w;
            a[4] = f1[g][1].x; a[5] = f1[g][1].y; a[6] = f1[g][1].z; a[7] = f1[g][1].w;
          }
#pragma unroll
          for (int k = 0; k < 8; ++k) v[k] = fmaf(p.beta1, a[k], v[k]);
        }
        if (p.res2) {
          v[0] = fmaf(p.beta2, f2[g][0].x, v[0]); v[1] = fmaf(p.beta2, f2[g][0].y, v[1]);
          v[2] = fmaf(p.beta2, f2[g][0].z, v[2]); v[3] = fmaf(p.beta2, f2[g][0].w, v[3]);
          v[4] = fmaf(p.beta2, f2[g][1].x, v[4]); v[5] = fmaf(p.beta2, f2[g][1].y, v[5]);
          v[6] = fmaf(p.beta2, f2[g][1].z, v[6]); v[7] = fmaf(p.beta2, f2[g][1].w, v[7]);
        }
        if (kBwd && p.res3) {
          v[0] = fmaf(p.beta3, f3[g][0].x, v[0]); v[1] = fmaf(p.beta3, f3[g][0].y, v[1]);
          v[2] = fmaf(p.beta3, f3[g][0].z, v[2]); v[3] = fmaf(p.beta3, f3[g][0].w, v[3]);
          v[4] = fmaf(p.beta3, f3[g][1].x, v[4]); v[5] = fmaf(p.beta3, f3[g][1].y, v[5]);
          v[6] = fmaf(p.beta3, f3[g][1].z, v[6]); v[7] = fmaf(p.beta3, f3[g][1].w, v[7]);
        }
        if (p.out32) {
          float4* op = reinterpret_cast<float4*>(
              p.out32 + (((size_t)img * p.out32_pt + p.out32_po + gp) * hw + pix) * 8);
          op[0] = make_float4(v[0], v[1], v[2], v[3]);
          op[1] = make_float4(v[4], v[5], v[6], v[7]);
        }
        if (p.out_nchw) {
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int ch = gp * 8 + k;
            if (ch < p.out_nchw_c) p.out_nchw[((size_t)img * p.out_nchw_c + ch) * hw + pix] = v[k];
          }
        }
        if (p.out16 && (!kBwd || gp >= p.tail_first)) {
          if (kBwd && p.mask16) {
            float a[8];
            unpack8(qm[g], p.dtype, a);
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] = a[k] > 0.f ? v[k] : v[k] * p.mask_slope;
          }
          if (p.out16_ps == 0) {
            uint4 o, ol;
            pack8_hl(v, p.dtype, p.split, o, ol);
            if (!p.out16_up2) {
              uint16_t* op = p.out16 + (((size_t)img * p.out16_pt + p.out16_po + gp) * hw + pix) * 8;
              *reinterpret_cast<uint4*>(op) = o;
              if (p.split) *reinterpret_cast<uint4*>(op + p.out16_lo) = ol;
            } else {
              const size_t w2 = 2 * (size_t)p.w;
              uint16_t* basep = p.out16 + (((size_t)img * p.out16_pt + p.out16_po + gp) * (4 * hw) +
                                           (size_t)(2 * y) * w2 + 2 * x) * 8;
              uint4* o0 = reinterpret_cast<uint4*>(basep);
              uint4* o1 = reinterpret_cast<uint4*>(basep + w2 * 8);
              o0[0] = o; o0[1] = o; o1[0] = o; o1[1] = o;
              if (p.split) {
                uint4* l0 = reinterpret_cast<uint4*>(basep + p.out16_lo);
                uint4* l1 = reinterpret_cast<uint4*>(basep + p.out16_lo + w2 * 8);
                l0[0] = ol; l0[1] = ol; l1[0] = ol; l1[1] = ol;
              }
            }
          } else {
            // pixel shuffle (block.py:287): conv channel c*r*r + i*r + j -> channel c at (r*y+i, r*x+j).
            const int rs = p.out16_ps, r2 = rs * rs;
            const size_t wr = (size_t)rs * p.w;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const int ch = ch0 + k;
              if (ch >= p.cout) break;
              const int oc = ch / r2, ij = ch - oc * r2;
              const int i = ij / rs, j = ij - i * rs;
              uint16_t* op = p.out16 + (((size_t)img * p.out16_pt + p.out16_po + (oc >> 3)) * (r2 * hw) +
                                        (size_t)(rs * y + i) * wr + (rs * x + j)) * 8 + (oc & 7);
              const uint32_t pk = pack2(v[k], 0.f, p.dtype);
              *op = (uint16_t)(pk & 0xFFFFu);
              if (p.split) {
                const float hv = __bfloat162float(__ushort_as_bfloat16((uint16_t)(pk & 0xFFFFu)));
                op[p.out16_lo] = (uint16_t)(pack2(v[k] - hv, 0.f, 1) & 0xFFFFu);
              }
            }
          }
        }
      }
    }
  }
}

// Specialised epilogues of the row-streaming kernel for the launches that make up a generator step; everything the
// generic epilogue decides at run time is fixed here (cout % 32 == 0, forward only), which cuts the instruction count
// per output row ~3x (the generic epilogue is what bounds the row kernel otherwise).
//   kMode 1: out16 = lrelu(acc + bias) [+ beta1*res1(fp32)]  plain or nearest-x2 replicated store   (growth convs, HR convs,
//            LR_conv + ShortcutBlock add)
//   kMode 2: v = alpha*(acc + bias) + beta1*res1(16-bit) [+ beta2*res2(fp32)] -> out16 [+ out32]   (conv5 of a dense block)
//   kMode 4: out16 = acc * (act > 0 ? 1 : mask_slope)   no bias                (gradient slices of the dense-block backward)
//   kMode 5: v = acc + beta3*res3(fp32) [+ beta1*res1(fp32)] -> out32 and out16, no bias   (closing launch of the dense-block backward:
//            the block's input gradient plus what arrives at its output; 43 % of the backward's FLOPs)
template <int NBN, int kMode>
__device__ __forceinline__ void conv_epilogue_fast(const ConvParams& p, uint32_t trow, int img, int y, int x, bool valid, int nblk) {
  const size_t hw = (size_t)p.h * p.w;
  const size_t pix = (size_t)y * p.w + x;
#pragma unroll
  for (int u = 0; u < NBN / 32; ++u) {
    uint32_t r[2][16];
    tmem_ld16(trow + u * 32, r[0]);
    tmem_ld16(trow + u * 32 + 16, r[1]);
    const int chu = nblk * NBN + u * 32;
    const int gu = chu >> 3;
    uint4 q1[4];
    float4 f2[4][2], bb[4][2];
    const bool has2 = (kMode == 2 && p.res2 != nullptr) || (kMode == 1 && p.res1 != nullptr) || kMode == 5;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      if (kMode == 4 || kMode == 5) {
        bb[g][0] = make_float4(0.f, 0.f, 0.f, 0.f);
        bb[g][1] = bb[g][0];
      } else {
        bb[g][0] = __ldg(reinterpret_cast<const float4*>(p.bias + chu + g * 8));
        bb[g][1] = __ldg(reinterpret_cast<const float4*>(p.bias + chu + g * 8 + 4));
      }
    }
    if (kMode == 4 && valid) {   // saved activation of the slice: only its sign is used
      const uint16_t* mk = p.mask16 + (((size_t)img * p.mask_pt + p.mask_po + gu) * hw + pix) * 8;
#pragma unroll
      for (int g = 0; g < 4; ++g) q1[g] = __ldg(reinterpret_cast<const uint4*>(mk + (size_t)g * hw * 8));
    }
    if (kMode == 1 && has2 && valid) {   // fp32 residual (the ShortcutBlock's skip) rides in the f2 registers
      const float* r2 = reinterpret_cast<const float*>(p.res1) + (((size_t)img * p.res1_pt + p.res1_po + gu) * hw + pix) * 8;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        f2[g][0] = __ldg(reinterpret_cast<const float4*>(r2 + (size_t)g * hw * 8));
        f2[g][1] = __ldg(reinterpret_cast<const float4*>(r2 + (size_t)g * hw * 8) + 1);
      }
    }
    if (kMode == 5 && valid) {   // what arrives at the block's output (fp32 trunk gradient) rides in the f2 registers
      const float* r3 = p.res3 + (((size_t)img * p.res3_pt + p.res3_po + gu) * hw + pix) * 8;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        f2[g][0] = __ldg(reinterpret_cast<const float4*>(r3 + (size_t)g * hw * 8));
        f2[g][1] = __ldg(reinterpret_cast<const float4*>(r3 + (size_t)g * hw * 8) + 1);
      }
    }
    if (kMode == 2 && valid) {
      const uint16_t* r1 = reinterpret_cast<const uint16_t*>(p.res1) + (((size_t)img * p.res1_pt + p.res1_po + gu) * hw + pix) * 8;
#pragma unroll
      for (int g = 0; g < 4; ++g) q1[g] = __ldg(reinterpret_cast<const uint4*>(r1 + (size_t)g * hw * 8));
      if (has2) {
        const float* r2 = p.res2 + (((size_t)img * p.res2_pt + p.res2_po + gu) * hw + pix) * 8;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          f2[g][0] = __ldg(reinterpret_cast<const float4*>(r2 + (size_t)g * hw * 8));
          f2[g][1] = __ldg(reinterpret_cast<const float4*>(r2 + (size_t)g * hw * 8) + 1);
        }
      }
    }
    tc_wait_ld();
    if (valid) {
      const float slope = p.lrelu ? p.slope : 1.0f;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const uint32_t* rg = &r[g / 2][(g & 1) * 8];
        float v[8];
        v[0] = __uint_as_float(rg[0]) + bb[g][0].x; v[1] = __uint_as_float(rg[1]) + bb[g][0].y;
        v[2] = __uint_as_float(rg[2]) + bb[g][0].z; v[3] = __uint_as_float(rg[3]) + bb[g][0].w;
        v[4] = __uint_as_float(rg[4]) + bb[g][1].x; v[5] = __uint_as_float(rg[5]) + bb[g][1].y;
        v[6] = __uint_as_float(rg[6]) + bb[g][1].z; v[7] = __uint_as_float(rg[7]) + bb[g][1].w;
        if (kMode == 4) {
          float a[8];
          unpack8(q1[g], p.dtype, a);
#pragma unroll
          for (int k = 0; k < 8; ++k) v[k] = a[k] > 0.f ? v[k] : v[k] * p.mask_slope;
        } else if (kMode == 5) {
          v[0] = fmaf(p.beta3, f2[g][0].x, v[0]); v[1] = fmaf(p.beta3, f2[g][0].y, v[1]);
          v[2] = fmaf(p.beta3, f2[g][0].z, v[2]); v[3] = fmaf(p.beta3, f2[g][0].w, v[3]);
          v[4] = fmaf(p.beta3, f2[g][1].x, v[4]); v[5] = fmaf(p.beta3, f2[g][1].y, v[5]);
          v[6] = fmaf(p.beta3, f2[g][1].z, v[6]); v[7] = fmaf(p.beta3, f2[g][1].w, v[7]);
          if (p.res1) {     // the RRDB skip connection's gradient (first block of an RRDB), fp32
            const float4* r1 = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.res1) +
                                                               (((size_t)img * p.res1_pt + p.res1_po + gu + g) * hw + pix) * 8);
            const float4 a0 = __ldg(r1), a1 = __ldg(r1 + 1);
            v[0] = fmaf(p.beta1, a0.x, v[0]); v[1] = fmaf(p.beta1, a0.y, v[1]); v[2] = fmaf(p.beta1, a0.z, v[2]); v[3] = fmaf(p.beta1, a0.w, v[3]);
            v[4] = fmaf(p.beta1, a1.x, v[4]); v[5] = fmaf(p.beta1, a1.y, v[5]); v[6] = fmaf(p.beta1, a1.z, v[6]); v[7] = fmaf(p.beta1, a1.w, v[7]);
          }
          float4* op = reinterpret_cast<float4*>(p.out32 + (((size_t)img * p.out32_pt + p.out32_po + gu + g) * hw + pix) * 8);
          op[0] = make_float4(v[0], v[1], v[2], v[3]);
          op[1] = make_float4(v[4], v[5], v[6], v[7]);
        } else if (kMode == 1) {
#pragma unroll
          for (int k = 0; k < 8; ++k) v[k] = v[k] > 0.f ? v[k] : v[k] * slope;
          if (has2) {
            v[0] = fmaf(p.beta1, f2[g][0].x, v[0]); v[1] = fmaf(p.beta1, f2[g][0].y, v[1]);
            v[2] = fmaf(p.beta1, f2[g][0].z, v[2]); v[3] = fmaf(p.beta1, f2[g][0].w, v[3]);
            v[4] = fmaf(p.beta1, f2[g][1].x, v[4]); v[5] = fmaf(p.beta1, f2[g][1].y, v[5]);
            v[6] = fmaf(p.beta1, f2[g][1].z, v[6]); v[7] = fmaf(p.beta1, f2[g][1].w, v[7]);
          }
        } else {
          float a[8];
          unpack8(q1[g], p.dtype, a);
          if (p.split)
            add_lo8(reinterpret_cast<const uint16_t*>(p.res1) + (((size_t)img * p.res1_pt + p.res1_po + gu + g) * hw + pix) * 8, p.res1_lo, a);
#pragma unroll
          for (int k = 0; k < 8; ++k) v[k] = fmaf(p.beta1, a[k], v[k] * p.alpha);
          if (has2) {
            v[0] = fmaf(p.beta2, f2[g][0].x, v[0]); v[1] = fmaf(p.beta2, f2[g][0].y, v[1]);
            v[2] = fmaf(p.beta2, f2[g][0].z, v[2]); v[3] = fmaf(p.beta2, f2[g][0].w, v[3]);
            v[4] = fmaf(p.beta2, f2[g][1].x, v[4]); v[5] = fmaf(p.beta2, f2[g][1].y, v[5]);
            v[6] = fmaf(p.beta2, f2[g][1].z, v[6]); v[7] = fmaf(p.beta2, f2[g][1].w, v[7]);
          }
          if (p.out32) {
            float4* op = reinterpret_cast<float4*>(p.out32 + (((size_t)img * p.out32_pt + p.out32_po + gu + g) * hw + pix) * 8);
            op[0] = make_float4(v[0], v[1], v[2], v[3]);
            op[1] = make_float4(v[4], v[5], v[6], v[7]);
          }
        }
        uint4 o, ol;
        pack8_hl(v, p.dtype, p.split, o, ol);
        if (kMode == 1 && p.out16_up2) {
          const size_t w2 = 2 * (size_t)p.w;
          uint16_t* basep = p.out16 + (((size_t)img * p.out16_pt + p.out16_po + gu + g) * (4 * hw) + (size_t)(2 * y) * w2 + 2 * x) * 8;
          uint4* o0 = reinterpret_cast<uint4*>(basep);
          uint4* o1 = reinterpret_cast<uint4*>(basep + w2 * 8);
          o0[0] = o; o0[1] = o; o1[0] = o; o1[1] = o;
          if (p.split) {
            uint4* l0 = reinterpret_cast<uint4*>(basep + p.out16_lo);
            uint4* l1 = reinterpret_cast<uint4*>(basep + p.out16_lo + w2 * 8);
            l0[0] = ol; l0[1] = ol; l1[0] = ol; l1[1] = ol;
          }
        } else {
          uint16_t* op = p.out16 + (((size_t)img * p.out16_pt + p.out16_po + gu + g) * hw + pix) * 8;
          *reinterpret_cast<uint4*>(op) = o;
          if (p.split) *reinterpret_cast<uint4*>(op + p.out16_lo) = ol;
        }
      }
    }
  }
}

//   kMode 3: out_nchw[c] = lrelu?(acc + bias), c < out_nchw_c <= 8                                    (HR_conv1 -> image)
template <int NBN>
__device__ __forceinline__ void conv_epilogue_nchw(const ConvParams& p, uint32_t trow, int img, int y, int x, bool valid) {
  uint32_t r[16];
  tmem_ld16(trow, r);
  const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias));
  const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + 4));
  tc_wait_ld();
  if (valid) {
    const float slope = p.lrelu ? p.slope : 1.0f;
    const size_t hw = (size_t)p.h * p.w;
    float* op = p.out_nchw + (size_t)img * p.out_nchw_c * hw + (size_t)y * p.w + x;
    const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (k < p.out_nchw_c) {
        float v = __uint_as_float(r[k]) + bb[k];
        v = v > 0.f ? v : v * slope;
        op[(size_t)k * hw] = v * p.alpha;
      }
    }
  }
}

template <int P, int KCP, int NBN, bool kBwd>
__global__ void __launch_bounds__(kConvThreads, 1)
conv3x3_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ ConvParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 127u) & ~127u;
  // header: full[8] | empty[8] | tmem_full[2] | tmem_empty[2] | tmem_ptr
  const uint32_t bar_full = smem_base;
  const uint32_t bar_empty = smem_base + 8 * kMaxStages;
  const uint32_t bar_tfull = smem_base + 16 * kMaxStages;
  const uint32_t bar_tempty = bar_tfull + 16;
  const uint32_t tmem_slot = bar_tempty + 16;
  const uint32_t bar_w = tmem_slot + 8;
  const uint32_t wres = smem_base + kSmemHeader;                       // resident weights (if any)
  const uint32_t stage0 = wres + (p.w_resident ? p.w_bytes : 0u);
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_tfull + 8 * b, 1);
      mbar_init(bar_tempty + 8 * b, 8);
    }
    mbar_init(bar_w, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, p.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  const int total_items = p.num_tiles * p.n_blocks;
  const int tiles_per_img = p.tiles_x * p.tiles_y;
  // PDL: the next conv of the stream may begin its prologue (barriers, TMEM, weight loads) while this grid drains;
  // every access to activations below is preceded by pdl_wait().
  if (threadIdx.x == 0) pdl_launch_dependents();

  if (warp < 4) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (elect.sync, not lane == 0: the compiler then keeps the
    // copy operands in uniform registers instead of an ELECT + R2UR.BROADCAST sequence in front of every TMA instruction)
    if (elect_one()) {
      int s = 0;
      uint32_t ph = 0;
      if (p.w_resident && (int)blockIdx.x < total_items) {
        mbar_expect_tx(bar_w, p.w_bytes);
        for (int c = 0; c < p.nchunks; ++c) bulk_load(wres + c * p.b_bytes, p.wts + (size_t)c * p.b_bytes, p.b_bytes, bar_w);
      }
      pdl_wait();  // weights do not depend on the previous launch, activations do
      for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
        const int nblk = item / p.num_tiles;
        const int tile = item - nblk * p.num_tiles;
        const int img = tile / tiles_per_img;
        const int trem = tile - img * tiles_per_img;
        const int tyi = trem / p.tiles_x;
        const int txi = trem - tyi * p.tiles_x;
        const int x0 = txi * p.TW, y0 = tyi * p.R;
        const uint8_t* wsrc = p.wts + (size_t)nblk * p.nchunks * p.b_bytes;
        for (int c = 0; c < p.nchunks; ++c) {
          mbar_wait(bar_empty + 8 * s, ph ^ 1u, 1u);
          const uint32_t sa = stage0 + s * p.stage_bytes;
          mbar_expect_tx(bar_full + 8 * s, p.w_resident ? p.a_bytes : p.a_bytes + p.b_bytes);
          tma_load_4d(sa, &tmA, bar_full + 8 * s, (x0 - 1) * 8, y0 - 1, conv_chunk_plane(p, c, p.kcp), img);
          if (!p.w_resident) bulk_load(sa + p.a_alloc, wsrc + (size_t)c * p.b_bytes, p.b_bytes, bar_full + 8 * s);
          if (++s == p.stages) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    int s = 0;
    uint32_t ph = 0;
    int it = 0;
    // descriptor templates: LBO/SBO/version fields; the 14-bit start-address field is added per use
    const uint64_t adesc_t = make_smem_desc(0u, p.plane_stride, 128u);
    const uint64_t bdesc_t = make_smem_desc(0u, (uint32_t)NBN * 16u, 128u);
    const uint64_t a_kstep = (uint64_t)((2u * p.plane_stride) >> 4);  // two planes = one K=16 step
    if (p.w_resident && (int)blockIdx.x < total_items) mbar_wait(bar_w, 0u, 5u);
    for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++it) {
      const int buf = it & 1;
      mbar_wait(bar_tempty + 8 * buf, ((uint32_t)(it >> 1) & 1u) ^ 1u, 2u);
      tc_fence_after();
      for (int c = 0; c < p.nchunks; ++c) {
        mbar_wait(bar_full + 8 * s, ph, 3u);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sa = stage0 + s * p.stage_bytes;
          const uint64_t ad0 = adesc_t + (uint64_t)(sa >> 4);
          const uint64_t bd0 = bdesc_t + (uint64_t)((p.w_resident ? wres + c * p.b_bytes : sa + p.a_alloc) >> 4);
          const uint32_t acc_first = c != 0 ? 1u : 0u;
          for (int t = 0; t < p.MT; ++t) {
            const uint32_t d = tmem_base + (uint32_t)((buf * p.MT + t) * NBN);
            const uint64_t at = ad0 + (uint64_t)(t * 128);
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
#pragma unroll
              for (int j = 0; j < KCP / 2; ++j) {
                const uint64_t ad = at + (uint64_t)((tap / 3) * P + (tap % 3)) + (uint64_t)j * a_kstep;
                const uint64_t bd = bd0 + (uint64_t)((tap * KCP + 2 * j) * NBN);
                umma_f16(d, ad, bd, p.idesc, (tap | j) != 0 ? 1u : acc_first);
              }
            }
          }
          umma_commit(bar_empty + 8 * s);
          if (c == p.nchunks - 1) umma_commit(bar_tfull + 8 * buf);
        }
        __syncwarp();
        if (++s == p.stages) { s = 0; ph ^= 1u; }
      }
    }
  }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
    // ------------------------------------------------------------------ epilogue (warps 4..11)
    // Two warps per TMEM lane quarter; they alternate over the M tiles of an item.  Per (row, unit of UW
    // channels) every global load (residuals, bias) is issued before the first store so that a thread
    // keeps UW/8 .. 3*UW/8 128-bit loads in flight.
    const int wq = warp & 3;            // TMEM lane quarter this warp may touch
    const int eh = (warp - 4) >> 2;     // which half of the M tiles this warp takes
    int it = 0;
    pdl_wait();  // residual reads and all stores must not overtake the previous launch
    for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++it) {
      const int buf = it & 1;
      const int nblk = item / p.num_tiles;
      const int tile = item - nblk * p.num_tiles;
      const int img = tile / tiles_per_img;
      const int trem = tile - img * tiles_per_img;
      const int tyi = trem / p.tiles_x;
      const int txi = trem - tyi * p.tiles_x;
      const int x0 = txi * p.TW, y0 = tyi * p.R;
      mbar_wait(bar_tfull + 8 * buf, (uint32_t)(it >> 1) & 1u, 4u);
      tc_fence_after();
      for (int t = (p.MT > 1 ? eh : 0); t < (p.MT > 1 ? p.MT : 1 - eh); t += 2) {
        const int q = t * 128 + wq * 32 + lane;
        const int ty = q / P;
        const int tx = q - ty * P;
        const int y = y0 + ty, x = x0 + tx;
        const bool valid = (tx < p.TW) && (x < p.w) && (y < p.h);
        const uint32_t trow = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)((buf * p.MT + t) * NBN);
        conv_epilogue_px<NBN, kBwd>(p, trow, img, y, x, valid, nblk);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_tempty + 8 * buf);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// ---------------------------------------------------------------------------------------------
// weight packing: OIHW fp32 -> [n_block][chunk][tap][plane-in-chunk][cout-in-block][8 cin] 16-bit
// ---------------------------------------------------------------------------------------------
// one element of the packed image (shared by the per-conv kernel and the batched one)
// split mode (dtype 2): the chunk index runs over three segments of nchunks/3 chunks each; segments 0 and 1 hold bf16(w) (they meet
// the hi and lo activation planes), segment 2 holds the bf16 residual w - bf16(w) (it meets the hi planes again)
__device__ __forceinline__ uint16_t split_weight_bits(float val, int seg) {
  const float hi = __bfloat162float(__float2bfloat16_rn(val));
  return __bfloat16_as_ushort(__float2bfloat16_rn(seg == 2 ? val - hi : hi));
}

__device__ __forceinline__ uint16_t pack_weights_elem(const float* __restrict__ w, int cout, int cin, int lead, int kcp, int nb_n, int nchunks,
                                                      int dtype, int transpose_flip, size_t idx) {
  // logical conv: out channels = (transpose_flip ? cin : cout) of the source tensor
  const int lc_out = transpose_flip ? cin : cout;
  const int lc_in = transpose_flip ? cout : cin;
  const int lead_pad = (lead + 7) / 8 * 8;
  size_t r = idx;
  const int ci8 = r % 8; r /= 8;
  const int n = r % nb_n; r /= nb_n;
  const int j = r % kcp; r /= kcp;
  const int tap = r % 9; r /= 9;
  int c = r % nchunks; r /= nchunks;
  const int nb = (int)r;
  int seg = 0;
  if (dtype == 2) { const int cps = nchunks / 3; seg = c / cps; c -= seg * cps; }
  int o = nb * nb_n + n;
  int i = (c * kcp + j) * 8 + ci8;
  // channel position in plane space -> source channel: the `lead` latent channels sit in their own zero-padded
  // plane group in front (forward: on the input side; transpose/dgrad: on the output side)
  if (!transpose_flip) {
    if (i < lead_pad) i = i < lead ? i : -1;
    else i = i - lead_pad + lead;
  } else {
    if (o < lead_pad) o = o < lead ? o : -1;
    else o = o - lead_pad + lead;
  }
  float val = 0.f;
  if (o >= 0 && o < lc_out && i >= 0 && i < lc_in) {
    const int ky = tap / 3, kx = tap % 3;
    if (!transpose_flip) val = w[(((size_t)o * cin + i) * 3 + ky) * 3 + kx];
    else val = w[(((size_t)i * cin + o) * 3 + (2 - ky)) * 3 + (2 - kx)];
  }
  if (dtype == 2) return split_weight_bits(val, seg);
  return (uint16_t)(pack2(val, 0.f, dtype) & 0xFFFFu);
}

__global__ void pack_weights_kernel(const float* __restrict__ w, int cout, int cin, int lead, int kcp, int nb_n, int n_blocks,
                                    int nchunks, int dtype, int transpose_flip, uint16_t* __restrict__ dst,
                                    size_t total) {
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x)
    dst[idx] = pack_weights_elem(w, cout, cin, lead, kcp, nb_n, nchunks, dtype, transpose_flip, idx);
}

}  // namespace esr
