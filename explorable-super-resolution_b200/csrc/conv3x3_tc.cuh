// 3x3 / stride 1 / zero-pad 1 convolution as a tcgen05 implicit GEMM (sm_100a).
//
//   GEMM view:  D[pixel, cout] = sum_{tap, cin} A[pixel + tap, cin] * W[tap, cin, cout]
//
// Data layout in HBM: planar-8 ("planes")  [N][C/8][H][W][8] 16-bit.  One pixel of one plane is exactly
// one 16-byte row of a no-swizzle K-major UMMA core matrix, so
//   * a TMA box (8ch, P px, R+2 rows, kcp planes) lands in shared memory as kcp halo planes whose pixels
//     sit at a 16-byte pitch;
//   * a 128-row MMA operand tile is 128 consecutive pixels of the flattened (row pitch P) halo plane, and
//     the nine filter taps are nine *address offsets* ((dy*P + dx) * 16 B) into the same halo tile — the
//     activation tile is fetched from L2 once per CTA tile, not once per tap;
//   * the dense-block concatenation (block.py:234) is a plane offset, never a copy.
// Columns tx >= P-2 of each flattened row are junk (they wrap into the next row) and are discarded by the
// epilogue: M efficiency (P-2)/P.
//
// CTA = 6 warps: warp 0 = TMA producer, warp 1 = MMA issuer (+TMEM owner), warps 2..5 = epilogue.
// Persistent over (tile, n-block) work items; TMEM accumulators are double buffered so the epilogue of
// item i overlaps the MMAs of item i+1.
#pragma once
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include "ptx.cuh"

namespace esr {

constexpr int kConvThreads = 192;
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
  uint32_t idesc;
  uint32_t tmem_cols;
  const uint8_t* wts;
  const float* bias;
  int cout;
  // epilogue
  int dtype;
  int lrelu;
  float slope, alpha;
  const float* res1; int res1_pt, res1_po; float beta1;
  const float* res2; int res2_pt, res2_po; float beta2;
  uint16_t* out16; int out16_pt, out16_po, out16_up2, out16_ps;
  float* out32; int out32_pt, out32_po;
  float* out_nchw; int out_nchw_c;
};

__device__ __forceinline__ uint32_t pack2(float a, float b, int dtype) {
  if (dtype == 0) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  } else {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
}

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
  const uint32_t stage0 = smem_base + kSmemHeader;
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
      mbar_init(bar_tempty + 8 * b, 4);
    }
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

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
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
          mbar_expect_tx(bar_full + 8 * s, p.a_bytes + p.b_bytes);
          tma_load_5d(sa, &tmA, bar_full + 8 * s, 0, x0 - 1, y0 - 1, p.in_plane_off + c * p.kcp, img);
          bulk_load(sa + p.a_alloc, wsrc + (size_t)c * p.b_bytes, p.b_bytes, bar_full + 8 * s);
          if (++s == p.stages) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    int s = 0;
    uint32_t ph = 0;
    int it = 0;
    const uint32_t b_lbo = (uint32_t)p.nb_n * 16u;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++it) {
      const int buf = it & 1;
      mbar_wait(bar_tempty + 8 * buf, ((uint32_t)(it >> 1) & 1u) ^ 1u, 2u);
      tc_fence_after();
      for (int c = 0; c < p.nchunks; ++c) {
        mbar_wait(bar_full + 8 * s, ph, 3u);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t sa = stage0 + s * p.stage_bytes;
          const uint32_t sb = sa + p.a_alloc;
          for (int t = 0; t < p.MT; ++t) {
            const uint32_t d = tmem_base + (uint32_t)((buf * p.MT + t) * p.nb_n);
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
              const int dy = tap / 3, dx = tap - dy * 3;
              const uint32_t a0 = sa + (uint32_t)(t * 128 + dy * p.P + dx) * 16u;
              const uint32_t b0 = sb + (uint32_t)(tap * p.kcp * p.nb_n) * 16u;
              for (int j = 0; j < (p.kcp >> 1); ++j) {
                const uint64_t ad = make_smem_desc(a0 + (uint32_t)(2 * j) * p.plane_stride, p.plane_stride, 128u);
                const uint64_t bd = make_smem_desc(b0 + (uint32_t)(2 * j) * b_lbo, b_lbo, 128u);
                umma_f16(d, ad, bd, p.idesc, (c | tap | j) != 0 ? 1u : 0u);
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
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..5)
    const int wq = warp & 3;  // TMEM lane quarter this warp may touch
    int it = 0;
    const size_t hw = (size_t)p.h * p.w;
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
      for (int t = 0; t < p.MT; ++t) {
        const int q = t * 128 + wq * 32 + lane;
        const int ty = q / p.P;
        const int tx = q - ty * p.P;
        const int y = y0 + ty, x = x0 + tx;
        const bool valid = (tx < p.TW) && (x < p.w) && (y < p.h);
        const size_t pix = (size_t)y * p.w + x;
        const uint32_t trow = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)((buf * p.MT + t) * p.nb_n);
        for (int cb = 0; cb < p.nb_n; cb += 16) {
          uint32_t r[16];
          tmem_ld16(trow + cb, r);
          tc_wait_ld();
          if (valid) {
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
              const int ch0 = nblk * p.nb_n + cb + hh * 8;  // first conv output channel of this group of 8
              if (ch0 >= p.cout) continue;
              const int g = ch0 >> 3;
              float v[8];
              const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + ch0));
              const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + ch0 + 4));
              v[0] = __uint_as_float(r[hh * 8 + 0]) + b0.x;
              v[1] = __uint_as_float(r[hh * 8 + 1]) + b0.y;
              v[2] = __uint_as_float(r[hh * 8 + 2]) + b0.z;
              v[3] = __uint_as_float(r[hh * 8 + 3]) + b0.w;
              v[4] = __uint_as_float(r[hh * 8 + 4]) + b1.x;
              v[5] = __uint_as_float(r[hh * 8 + 5]) + b1.y;
              v[6] = __uint_as_float(r[hh * 8 + 6]) + b1.z;
              v[7] = __uint_as_float(r[hh * 8 + 7]) + b1.w;
              if (p.lrelu) {
#pragma unroll
                for (int k = 0; k < 8; ++k) v[k] = v[k] > 0.f ? v[k] : v[k] * p.slope;
              }
#pragma unroll
              for (int k = 0; k < 8; ++k) v[k] *= p.alpha;
              if (p.res1) {
                const float4* rp = reinterpret_cast<const float4*>(
                    p.res1 + (((size_t)img * p.res1_pt + p.res1_po + g) * hw + pix) * 8);
                const float4 r0 = __ldg(rp), r1 = __ldg(rp + 1);
                v[0] += p.beta1 * r0.x; v[1] += p.beta1 * r0.y; v[2] += p.beta1 * r0.z; v[3] += p.beta1 * r0.w;
                v[4] += p.beta1 * r1.x; v[5] += p.beta1 * r1.y; v[6] += p.beta1 * r1.z; v[7] += p.beta1 * r1.w;
              }
              if (p.res2) {
                const float4* rp = reinterpret_cast<const float4*>(
                    p.res2 + (((size_t)img * p.res2_pt + p.res2_po + g) * hw + pix) * 8);
                const float4 r0 = __ldg(rp), r1 = __ldg(rp + 1);
                v[0] += p.beta2 * r0.x; v[1] += p.beta2 * r0.y; v[2] += p.beta2 * r0.z; v[3] += p.beta2 * r0.w;
                v[4] += p.beta2 * r1.x; v[5] += p.beta2 * r1.y; v[6] += p.beta2 * r1.z; v[7] += p.beta2 * r1.w;
              }
              if (p.out32) {
                float4* op = reinterpret_cast<float4*>(
                    p.out32 + (((size_t)img * p.out32_pt + p.out32_po + g) * hw + pix) * 8);
                op[0] = make_float4(v[0], v[1], v[2], v[3]);
                op[1] = make_float4(v[4], v[5], v[6], v[7]);
              }
              if (p.out_nchw) {
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                  const int ch = ch0 + k;
                  if (ch < p.out_nchw_c) p.out_nchw[((size_t)img * p.out_nchw_c + ch) * hw + pix] = v[k];
                }
              }
              if (p.out16) {
                if (p.out16_ps == 0) {
                  uint4 o;
                  o.x = pack2(v[0], v[1], p.dtype);
                  o.y = pack2(v[2], v[3], p.dtype);
                  o.z = pack2(v[4], v[5], p.dtype);
                  o.w = pack2(v[6], v[7], p.dtype);
                  if (!p.out16_up2) {
                    uint4* op = reinterpret_cast<uint4*>(
                        p.out16 + (((size_t)img * p.out16_pt + p.out16_po + g) * hw + pix) * 8);
                    *op = o;
                  } else {
                    const size_t w2 = 2 * (size_t)p.w;
                    uint16_t* basep = p.out16 + (((size_t)img * p.out16_pt + p.out16_po + g) * (4 * hw) +
                                                 (size_t)(2 * y) * w2 + 2 * x) * 8;
                    uint4* o0 = reinterpret_cast<uint4*>(basep);
                    uint4* o1 = reinterpret_cast<uint4*>(basep + w2 * 8);
                    o0[0] = o; o0[1] = o; o1[0] = o; o1[1] = o;
                  }
                } else {
                  // pixel shuffle (block.py:287): conv channel c*r*r + i*r + j -> channel c at (r*y+i, r*x+j).
                  const int rr = p.out16_ps, r2 = rr * rr;
                  const size_t wr = (size_t)rr * p.w;
#pragma unroll
                  for (int k = 0; k < 8; ++k) {
                    const int ch = ch0 + k;
                    if (ch >= p.cout) break;
                    const int oc = ch / r2, ij = ch - oc * r2;
                    const int i = ij / rr, j = ij - i * rr;
                    uint16_t* op = p.out16 + (((size_t)img * p.out16_pt + p.out16_po + (oc >> 3)) * (r2 * hw) +
                                              (size_t)(rr * y + i) * wr + (rr * x + j)) * 8 + (oc & 7);
                    const uint32_t pk = pack2(v[k], 0.f, p.dtype);
                    *op = (uint16_t)(pk & 0xFFFFu);
                  }
                }
              }
            }
          }
        }
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
__global__ void pack_weights_kernel(const float* __restrict__ w, int cout, int cin, int lead, int kcp, int nb_n, int n_blocks,
                                    int nchunks, int dtype, int transpose_flip, uint16_t* __restrict__ dst,
                                    size_t total) {
  // logical conv: out channels = (transpose_flip ? cin : cout) of the source tensor
  const int lc_out = transpose_flip ? cin : cout;
  const int lc_in = transpose_flip ? cout : cin;
  const int lead_pad = (lead + 7) / 8 * 8;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    size_t r = idx;
    const int ci8 = r % 8; r /= 8;
    const int n = r % nb_n; r /= nb_n;
    const int j = r % kcp; r /= kcp;
    const int tap = r % 9; r /= 9;
    const int c = r % nchunks; r /= nchunks;
    const int nb = (int)r;
    const int o = nb * nb_n + n;
    int i = (c * kcp + j) * 8 + ci8;  // channel position in plane space -> source input channel
    if (i < lead_pad) i = i < lead ? i : -1;
    else i = i - lead_pad + lead;
    float val = 0.f;
    if (o < lc_out && i >= 0 && i < lc_in) {
      const int ky = tap / 3, kx = tap % 3;
      if (!transpose_flip) val = w[(((size_t)o * cin + i) * 3 + ky) * 3 + kx];
      else val = w[(((size_t)i * cin + o) * 3 + (2 - ky)) * 3 + (2 - kx)];
    }
    const uint32_t pk = pack2(val, 0.f, dtype);
    dst[idx] = (uint16_t)(pk & 0xFFFFu);
  }
}

}  // namespace esr
