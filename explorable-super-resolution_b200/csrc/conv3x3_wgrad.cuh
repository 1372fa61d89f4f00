// Weight gradient of the 3x3 / stride 1 / zero-pad 1 convolution on the tensor cores (sm_100a, tcgen05).
//
//   dW[co][ci][ky][kx] = sum_{n,y,x} gy[n][co][y][x] * xin[n][ci][y+ky-1][x+kx-1]
//
// GEMM view with K = pixels: both operands are read straight from the planar-8 activation layout, which is the
// "MN-major" no-swizzle UMMA layout when 16 consecutive pixels of a row are the K extent of one MMA: a core matrix is
// 8 pixels x 8 channels = 128 contiguous bytes, the next 8 channels (next plane) are one M/N group further (SBO), the
// next 8 pixels one K block further (LBO = 128 B).
//   * vertical taps live in M: the x ring buffer is laid out [row][plane][pixel], so the (row, plane) groups of input
//     rows r-1, r, r+1 are CONSECUTIVE M groups: an MMA of M = 128 covers 16 of the 3*Cin/8 groups
//       D[(ky, ci), .] += x[r + ky - 1][.. , ci] * gy[r][..]
//   * horizontal taps live in N: the producer loads the gy row three times, shifted by 1 - kx pixels, as
//     [kx][plane][pixel]; the (kx, plane) groups are consecutive N groups, so one MMA has N = 3 * Cout_block.
//   The sum over pixels is partitioned by OUTPUT row (y) and INPUT column (x): a CTA owns a 64-pixel strip of x
//   positions and a contiguous run of output rows, marches down it (every activation row is fetched once), and keeps
//   its partial dW in TMEM for the whole launch (ceil(3*Cin/128) chunks x 3*Cout_block fp32 columns).  At the end the
//   partials go to a workspace and a small kernel reduces them over CTAs into the OIHW fp32 gradient (the layout
//   torch.optim / the reference's optimizers read, models/SRRaGAN_model.py:163-188).
#pragma once
#include "conv3x3_rows.cuh"

namespace esr {

constexpr int kWgThreads = 256;       // warp 0 producer, warp 1 MMA issuer, warps 4..7 TMEM zero-fill + final read-out
constexpr int kWgPW = 64;             // pixels of a strip (4 K steps of 16 pixels per row)
constexpr int kWgMaxRing = 8;
constexpr int kWgGStages = 3;

struct WgradParams {
  int n, h, w;
  int strips, ranges, n_blocks;
  long long units;
  const uint8_t* x; int x_pt, x_po, cp;        // conv input (16-bit planes), cp planes used
  const uint8_t* gy; int gy_pt, gy_po, gyp;    // output gradient (16-bit planes), gyp planes exist
  int nbn, cpb;                                // channels / planes of one n-block
  int mt;                                      // M chunks of 16 (row, plane) groups
  int rb;                                      // ring rows (+2 duplicate slots)
  uint32_t slot_bytes, gstage_bytes;
  uint32_t idesc;
  float* part;                                 // [cta][mt][128][3*nbn] fp32 partial gradients
};

__device__ __forceinline__ bool wg_next_segment(const WgradParams& p, long long& u, long long u1, int& img, int& x0, int& ya,
                                                int& yb) {
  if (u >= u1) return false;
  const long long col = u / p.h;
  ya = (int)(u - col * p.h);
  const long long rem = u1 - u;
  yb = (rem < (long long)(p.h - ya)) ? ya + (int)rem : p.h;
  img = (int)(col / p.strips);
  x0 = (int)(col - (long long)img * p.strips) * kWgPW;
  u += yb - ya;
  return true;
}

__global__ void __launch_bounds__(kWgThreads, 1) conv3x3_wgrad_kernel(const __grid_constant__ WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 127u) & ~127u;
  // header: x row full[8] | x row empty[8] | gy full[4] | gy empty[4] | done | zeroed | tmem pointer
  const uint32_t bar_xfull = smem_base;
  const uint32_t bar_xempty = smem_base + 64;
  const uint32_t bar_gfull = smem_base + 128;
  const uint32_t bar_gempty = smem_base + 160;
  const uint32_t bar_done = smem_base + 192;
  const uint32_t bar_zero = smem_base + 200;
  const uint32_t tmem_slot = smem_base + 208;
  const uint32_t xring = smem_base + kSmemHeader;
  const uint32_t group_bytes = kWgPW * 16u;                               // one plane of one row
  const uint32_t gst0 = xring + (uint32_t)(p.rb + 2) * p.slot_bytes + 16u * group_bytes;   // + slack for the last M chunk
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int N3 = 3 * p.nbn;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.rb; ++s) {
      mbar_init(bar_xfull + 8 * s, 1);
      mbar_init(bar_xempty + 8 * s, 1);
    }
    for (int s = 0; s < kWgGStages; ++s) {
      mbar_init(bar_gfull + 8 * s, 1);
      mbar_init(bar_gempty + 8 * s, 1);
    }
    mbar_init(bar_done, 1);
    mbar_init(bar_zero, 4);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512u);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  const int nblk = (int)blockIdx.x % p.n_blocks;
  const int rid = (int)blockIdx.x / p.n_blocks;
  const long long u0 = p.units * rid / p.ranges, u1 = p.units * (rid + 1) / p.ranges;
  const size_t hw = (size_t)p.h * p.w;
  const size_t plane16 = hw * 16;

  if (warp == 0) {
    // ------------------------------------------------------------------ producer
    // x rows: one lane per plane (cp <= 32 planes); gy rows: one lane per (kx, plane) copy (3 * cpb <= 24)
    int xi = 0;            // x rows loaded so far (ring index = xi % rb)
    int gi = 0;            // gy rows loaded so far
    long long u = u0;
    int img, x0, ya, yb;
    while (wg_next_segment(p, u, u1, img, x0, ya, yb)) {
      const int valid = p.w - x0 < kWgPW ? p.w - x0 : kWgPW;   // x positions of this strip that exist
      const uint8_t* xcol = p.x + (((size_t)img * p.x_pt + p.x_po) * hw + x0) * 16;
      const uint8_t* gcol = p.gy + (((size_t)img * p.gy_pt + p.gy_po + nblk * p.cpb) * hw) * 16;
      int seg_x = 0, seg_g = 0;
      // gy copy handled by this lane: kx = lane / cpb, plane = lane % cpb; copy[k] = gy[row][x0 + k - kx + 1]
      const int kx = lane / p.cpb, gpl = lane - kx * p.cpb;
      const int gpl_src = nblk * p.cpb + gpl < p.gyp ? gpl : p.gyp - 1 - nblk * p.cpb;   // pad planes re-load the last real one
      int gs = x0 + 1 - kx, gd = 0;                       // source pixel, destination pixel
      if (gs < 0) { gd = -gs; gs = 0; }
      int gcnt = kWgPW - gd;
      if (gs + gcnt > p.w) gcnt = p.w - gs;
      if (gcnt < 0) gcnt = 0;
      for (int row = ya - 1; row <= yb; ++row) {
        // ---- x row `row` (zeros outside the image)
        {
          const int j = xi % p.rb;
          mbar_wait(bar_xempty + 8 * j, (uint32_t)(((xi / p.rb) & 1) ^ 1), 1u);
          const uint32_t dst = xring + (uint32_t)j * p.slot_bytes;
          const bool dup = j < 2;                         // ring slots 0,1 are mirrored behind the last slot
          const uint32_t dst2 = xring + (uint32_t)(p.rb + j) * p.slot_bytes;
          const bool real = row >= 0 && row < p.h;
          if (!real) {
            // zero row: all 32 lanes clear cp * kWgPW 16-byte pixels
            for (uint32_t o = lane * 16u; o < p.slot_bytes; o += 512u) {
              st_shared_zero16(dst + o);
              if (dup) st_shared_zero16(dst2 + o);
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_xfull + 8 * j);
          } else {
            if (valid < kWgPW && seg_x < p.rb + 2) {
              // pixels beyond the image edge are never written by this segment's copies: clear them once per slot
              if (lane < p.cp) {
                for (int k = valid; k < kWgPW; ++k) {
                  st_shared_zero16(dst + lane * group_bytes + k * 16);
                  if (dup) st_shared_zero16(dst2 + lane * group_bytes + k * 16);
                }
                fence_proxy_async();
              }
              __syncwarp();
            }
            if (lane == 0) mbar_expect_tx(bar_xfull + 8 * j, (uint32_t)p.cp * valid * 16u * (dup ? 2u : 1u));
            __syncwarp();
            if (lane < p.cp) {
              const uint8_t* src = xcol + (size_t)lane * plane16 + (size_t)row * p.w * 16;
              bulk_load(dst + lane * group_bytes, src, (uint32_t)valid * 16u, bar_xfull + 8 * j);
              if (dup) bulk_load(dst2 + lane * group_bytes, src, (uint32_t)valid * 16u, bar_xfull + 8 * j);
            }
          }
          ++xi;
          ++seg_x;
        }
        // ---- gy row `row` (output rows of the segment only)
        if (row >= ya && row < yb) {
          const int s = gi % kWgGStages;
          mbar_wait(bar_gempty + 8 * s, (uint32_t)(((gi / kWgGStages) & 1) ^ 1), 2u);
          const uint32_t dst = gst0 + (uint32_t)s * p.gstage_bytes + lane * group_bytes;
          const bool mine = lane < 3 * p.cpb;
          if (seg_g < kWgGStages) {
            if (mine) {
              for (int k = 0; k < gd; ++k) st_shared_zero16(dst + k * 16);
              for (int k = gd + gcnt; k < kWgPW; ++k) st_shared_zero16(dst + k * 16);
              fence_proxy_async();
            }
            __syncwarp();
          }
          // every lane's byte count differs by at most one pixel: lane 0 posts the total
          uint32_t tot = mine ? (uint32_t)gcnt * 16u : 0u;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
          if (lane == 0) mbar_expect_tx(bar_gfull + 8 * s, tot);
          __syncwarp();
          if (mine && gcnt > 0)
            bulk_load(dst + gd * 16, gcol + (size_t)gpl_src * plane16 + ((size_t)row * p.w + gs) * 16, (uint32_t)gcnt * 16u,
                      bar_gfull + 8 * s);
          ++gi;
          ++seg_g;
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (one thread)
    if (elect_one()) {
      if (u0 < u1) mbar_wait(bar_zero, 0u, 6u);
      tc_fence_after();
      const uint64_t adesc_t = make_smem_desc(0u, 128u, group_bytes);   // MN-major: LBO = K-block (8 px) stride, SBO = group stride
      const uint64_t bdesc_t = make_smem_desc(0u, 128u, group_bytes);
      int xi = 0, gi = 0;
      long long u = u0;
      int img, x0, ya, yb;
      while (wg_next_segment(p, u, u1, img, x0, ya, yb)) {
        const int valid = p.w - x0 < kWgPW ? p.w - x0 : kWgPW;
        const int ksteps = (valid + 15) >> 4;
        // rows ya-1, ya are the first two ring rows of this segment
        for (int r = ya; r < yb; ++r) {
          // window = x rows r-1, r, r+1 = ring indices xi, xi+1, xi+2
          if (r == ya) {
            mbar_wait(bar_xfull + 8 * (xi % p.rb), (uint32_t)((xi / p.rb) & 1), 3u);
            mbar_wait(bar_xfull + 8 * ((xi + 1) % p.rb), (uint32_t)(((xi + 1) / p.rb) & 1), 3u);
          }
          mbar_wait(bar_xfull + 8 * ((xi + 2) % p.rb), (uint32_t)(((xi + 2) / p.rb) & 1), 3u);
          const int s = gi % kWgGStages;
          mbar_wait(bar_gfull + 8 * s, (uint32_t)((gi / kWgGStages) & 1), 4u);
          tc_fence_after();
          const uint32_t a0 = xring + (uint32_t)(xi % p.rb) * p.slot_bytes;
          const uint32_t b0 = gst0 + (uint32_t)s * p.gstage_bytes;
          for (int ks = 0; ks < ksteps; ++ks) {
            const uint64_t bd = bdesc_t + (uint64_t)((b0 + ks * 256u) >> 4);
            for (int m = 0; m < p.mt; ++m) {
              const uint64_t ad = adesc_t + (uint64_t)((a0 + (uint32_t)m * 16u * group_bytes + ks * 256u) >> 4);
              umma_f16(tmem_base + (uint32_t)(m * N3), ad, bd, p.idesc, 1u);
            }
          }
          umma_commit(bar_gempty + 8 * s);
          umma_commit(bar_xempty + 8 * (xi % p.rb));          // row r-1 is not needed again
          if (r == yb - 1) {                                   // ... and neither are the last two rows of the segment
            umma_commit(bar_xempty + 8 * ((xi + 1) % p.rb));
            umma_commit(bar_xempty + 8 * ((xi + 2) % p.rb));
          }
          ++xi;
          ++gi;
        }
        xi += 2;
      }
      umma_commit(bar_done);
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ accumulator zero-fill, final read-out
    const int wq = warp & 3;
    const uint32_t tq = tmem_base + ((uint32_t)(wq * 32) << 16);
    const int cols = p.mt * N3;
    for (int c = 0; c < cols; c += 16) tmem_zero16(tq + (uint32_t)c);
    tc_wait_st();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_zero);
    mbar_wait(bar_done, 0u, 7u);
    tc_fence_after();
    float* out = p.part + (((size_t)blockIdx.x * p.mt) * 128 + wq * 32 + lane) * N3;
    for (int m = 0; m < p.mt; ++m) {
      for (int c = 0; c < N3; c += 16) {
        uint32_t r[16];
        tmem_ld16(tq + (uint32_t)(m * N3 + c), r);
        tc_wait_ld();
        float4* op = reinterpret_cast<float4*>(out + (size_t)m * 128 * N3 + c);
        op[0] = make_float4(__uint_as_float(r[0]), __uint_as_float(r[1]), __uint_as_float(r[2]), __uint_as_float(r[3]));
        op[1] = make_float4(__uint_as_float(r[4]), __uint_as_float(r[5]), __uint_as_float(r[6]), __uint_as_float(r[7]));
        op[2] = make_float4(__uint_as_float(r[8]), __uint_as_float(r[9]), __uint_as_float(r[10]), __uint_as_float(r[11]));
        op[3] = make_float4(__uint_as_float(r[12]), __uint_as_float(r[13]), __uint_as_float(r[14]), __uint_as_float(r[15]));
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512u);
  }
}

// dW[co][ci][ky][kx] (+)= scale * sum over CTAs of the partial accumulators.  One thread per weight.
__global__ void wgrad_reduce_kernel(const float* __restrict__ part, int ranges, int n_blocks, int mt, int nbn, int cp, int cout, int cin,
                                    int lead, float scale, int accumulate, float* __restrict__ dw) {
  const int total = cout * cin * 9;
  const int lead_pad = (lead + 7) / 8 * 8;
  const int N3 = 3 * nbn;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    int r = idx;
    const int kx = r % 3; r /= 3;
    const int ky = r % 3; r /= 3;
    const int ci = r % cin;
    const int co = r / cin;
    const int pc = ci < lead ? ci : ci - lead + lead_pad;       // channel position in plane space
    const int G = ky * cp + (pc >> 3);                            // (row, plane) group
    const int m = G >> 4, L = ((G & 15) << 3) + (pc & 7);
    const int nblk = co / nbn, col = kx * nbn + (co - nblk * nbn);
    float acc = 0.f;
    for (int rg = 0; rg < ranges; ++rg) {
      const size_t cta = (size_t)rg * n_blocks + nblk;
      acc += part[((cta * mt + m) * 128 + L) * N3 + col];
    }
    acc *= scale;
    dw[idx] = accumulate ? dw[idx] + acc : acc;
  }
}

// db[co] (+)= scale * sum_{n,y,x} gy[n][co][y][x]   (16-bit planes in, fp32 out).  grid = (chunks, planes)
__global__ void bias_grad_kernel(const uint16_t* __restrict__ gy, int dtype, int n, int gy_pt, int gy_po, int cout, size_t hw,
                                 float scale, float* __restrict__ db) {
  const int plane = blockIdx.y;
  float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const size_t total = (size_t)n * hw;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t img = i / hw, pix = i - img * hw;
    const uint4 q = __ldg(reinterpret_cast<const uint4*>(gy + ((img * gy_pt + gy_po + plane) * hw + pix) * 8));
    float a[8];
    unpack8(q, dtype, a);
#pragma unroll
    for (int k = 0; k < 8; ++k) s[k] += a[k];
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s[k] += __shfl_xor_sync(0xffffffffu, s[k], o);
  }
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int co = plane * 8 + k;
      if (co < cout) atomicAdd(db + co, s[k] * scale);
    }
  }
}

// out[c] += scale * sum_{n,y,x} src[n][c][y][x]  (NCHW fp32: bias gradient of the last conv straight from dL/dG, which
// has not been rounded to 16 bits).  grid = (chunks, c, n)
__global__ void sum_nchw_kernel(const float* __restrict__ src, int c, size_t hw, float scale, float* __restrict__ out) {
  const float* p = src + ((size_t)blockIdx.z * c + blockIdx.y) * hw;
  float s = 0.f;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < hw; i += (size_t)gridDim.x * blockDim.x) s += __ldg(p + i);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) atomicAdd(out + blockIdx.y, s * scale);
}

}  // namespace esr
