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
//     (v2, the default) The gradient row crosses L2 -> SM ONCE: eight loader warps read it with plain 16-byte loads (66 pixels:
//     the strip and one halo pixel each side), write the three shifted copies into shared memory, and add up the bias gradient
//     on the way (db = sum of gy over pixels); up to eight rows are in flight in their registers.  v1 (ESR_WGRAD_V1=1) loads the
//     three copies with three TMA boxes per 32 pixels: 3x the gradient traffic and eight TMA operations per row from one thread,
//     which bounded it (profiles/r02_ncu_full_conv3x3_wgrad_summary.csv: 0.8-1.0 us per row whatever Cin).
//   The sum over pixels is partitioned by OUTPUT row (y) and INPUT column (x): a CTA owns a 64-pixel strip of x
//   positions and a contiguous run of output rows, marches down it (every activation row is fetched once), and keeps
//   its partial dW in TMEM for the whole launch (ceil(3*Cin/128) chunks x 3*Cout_block fp32 columns).  At the end the
//   partials go to a workspace and a small kernel reduces them over CTAs into the OIHW fp32 gradient (the layout
//   torch.optim / the reference's optimizers read, models/SRRaGAN_model.py:163-188).
#pragma once
#include "conv3x3_rows.cuh"

namespace esr {

constexpr int kWgThreadsV1 = 256;     // v1: warp 0 producer, warps 1..3 MMA issuers, warps 4..7 TMEM zero-fill + final read-out
constexpr int kWgLoaders = 8;
constexpr int kWgThreads = 128 + 32 * kWgLoaders;   // warp 0 activation-row producer (TMA), warps 1..3 MMA issuers, warps 4..11 gradient-row
                                                    // loaders (+ bias gradient), TMEM zero-fill and final read-out
constexpr int kWgBiasStride = 64;     // floats per CTA in the bias partials (one n-block has at most 64 channels)
constexpr int kWgIssuers = 3;
constexpr int kWgBox = 32;            // pixels of one TMA box row (2 K steps of 16 pixels)
constexpr int kWgNB = 2;              // boxes per strip row: the per-row bookkeeping is amortised over 4 K steps
constexpr int kWgPW = kWgBox * kWgNB; // pixels of a strip
constexpr int kWgMaxRing = 8;
constexpr int kWgGStages = 3;

struct WgradParams {
  int n, h, w;
  int dtype;                                   // 0 fp16, 1 bf16 (operands)
  int strips, ranges, n_blocks;
  long long units;
  const uint8_t* x; int x_pt, x_po, cp;        // conv input (16-bit planes), cp planes used
  const uint8_t* gy; int gy_pt, gy_po, gyp;    // output gradient (16-bit planes), gyp planes exist
  int nbn, cpb;                                // channels / planes of one n-block
  int mt;                                      // M chunks of 16 (row, plane) groups
  int rb;                                      // ring rows (+2 duplicate slots)
  uint32_t slot_bytes, gstage_bytes;           // per 32-pixel box: cp * 512 and 3 * cpb * 512
  uint32_t ring_bytes;                         // one box ring: (rb + 2) slots + slack
  uint32_t idesc;
  int issuers;                                 // MMA issuer warps sharing the K steps (1..3; 1 = fixed accumulation order)
  int gstages, ngroups;                        // v2: gradient stages in shared memory (3..8); groups of four gradient planes per n-block (1, 2)
  float* part;                                 // [cta][mt][128][3*nbn] fp32 partial gradients
  float* bias_part;                            // [cta][64] fp32 partial bias gradients (v2; NULL = not wanted)
};

// position in a ring of `n` buffers with its mbarrier phase bit (no integer division in the per-row loops)
struct WgRing {
  int j;
  uint32_t ph;
  __device__ __forceinline__ void inc(int n) {
    if (++j == n) { j = 0; ph ^= 1u; }
  }
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

__global__ void __launch_bounds__(kWgThreadsV1, 1)
conv3x3_wgrad_kernel_v1(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmG, const __grid_constant__ WgradParams p) {
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
  const uint32_t group_bytes = kWgBox * 16u;                              // one plane of one box row
  // x: kWgNB rings [box][row slot][plane][32 px] (rows of one box are consecutive M groups); gy stages [box][kx][plane][32 px]
  const uint32_t gst0 = xring + kWgNB * p.ring_bytes;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int N3 = 3 * p.nbn;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.rb; ++s) {
      mbar_init(bar_xfull + 8 * s, 1);
      mbar_init(bar_xempty + 8 * s, (uint32_t)p.issuers);
    }
    for (int s = 0; s < kWgGStages; ++s) {
      mbar_init(bar_gfull + 8 * s, 1);
      mbar_init(bar_gempty + 8 * s, (uint32_t)p.issuers);
    }
    mbar_init(bar_done, (uint32_t)p.issuers);
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

  if (warp == 0) {
    // ------------------------------------------------------------------ producer
    // One TMA tensor load per activation row (box = 32 pixels x cp planes, landing as [plane][pixel] = consecutive M
    // groups) and three per gradient row (the kx-shifted copies); rows / pixels outside the image are zero-filled by the
    // TMA unit, which is exactly the convolution's zero padding.
    if (lane == 0) {
      tma_prefetch_desc(&tmX);
      tma_prefetch_desc(&tmG);
      WgRing xr = {0, 0u};   // next x row slot
      WgRing gr = {0, 0u};   // next gy stage
      long long u = u0;
      int img, x0, ya, yb;
      while (wg_next_segment(p, u, u1, img, x0, ya, yb)) {
        for (int row = ya - 1; row <= yb; ++row) {
          {
            const int j = xr.j;
            mbar_wait(bar_xempty + 8 * j, xr.ph ^ 1u, 1u);
            const bool dup = j < 2;                         // ring slots 0,1 are mirrored behind the last slot
            mbar_expect_tx(bar_xfull + 8 * j, kWgNB * p.slot_bytes * (dup ? 2u : 1u));
#pragma unroll
            for (int b = 0; b < kWgNB; ++b) {
              const uint32_t ring = xring + b * p.ring_bytes;
              tma_load_4d(ring + (uint32_t)j * p.slot_bytes, &tmX, bar_xfull + 8 * j, (x0 + b * kWgBox) * 8, row, p.x_po, img);
              if (dup) tma_load_4d(ring + (uint32_t)(p.rb + j) * p.slot_bytes, &tmX, bar_xfull + 8 * j, (x0 + b * kWgBox) * 8, row, p.x_po, img);
            }
            xr.inc(p.rb);
          }
          if (row >= ya && row < yb) {
            const int s = gr.j;
            mbar_wait(bar_gempty + 8 * s, gr.ph ^ 1u, 2u);
            const uint32_t dst = gst0 + (uint32_t)s * kWgNB * p.gstage_bytes;
            mbar_expect_tx(bar_gfull + 8 * s, kWgNB * p.gstage_bytes);
            // copy kx holds gy[row][x0 + k - kx + 1] at pixel k
#pragma unroll
            for (int b = 0; b < kWgNB; ++b)
              for (int kx = 0; kx < 3; ++kx)
                tma_load_4d(dst + b * p.gstage_bytes + (uint32_t)(kx * p.cpb) * group_bytes, &tmG, bar_gfull + 8 * s,
                            (x0 + b * kWgBox + 1 - kx) * 8, row, p.gy_po + nblk * p.cpb, img);
            gr.inc(kWgGStages);
          }
        }
      }
    }
  } else if (warp <= p.issuers) {
    // ------------------------------------------------------------------ MMA issuers (one thread per warp).  The K steps
    // of the launch are dealt round-robin to the issuers; every MMA accumulates into zero-initialised TMEM, so their
    // order is irrelevant, and every buffer is released when all issuers have committed it.
    const int iw = warp - 1;
    if (elect_one()) {
      int kmod = 0;
      if (u0 < u1) mbar_wait(bar_zero, 0u, 6u);
      tc_fence_after();
      const uint64_t adesc_t = make_smem_desc(0u, 128u, group_bytes);   // MN-major: LBO = K-block (8 px) stride, SBO = group stride
      const uint64_t bdesc_t = make_smem_desc(0u, 128u, group_bytes);
      WgRing cur = {0, 0u}, gr = {0, 0u};
      const uint64_t mstep = (uint64_t)((16u * group_bytes) >> 4);
      long long u = u0;
      int img, x0, ya, yb;
      while (wg_next_segment(p, u, u1, img, x0, ya, yb)) {
        const int valid = p.w - x0 < kWgPW ? p.w - x0 : kWgPW;
        const int ksteps = (valid + 15) >> 4;
        // window of output row r = x rows r-1, r, r+1 = three consecutive ring rows p0, p1, p2
        WgRing p0 = cur, p1 = cur, p2;
        p1.inc(p.rb);
        p2 = p1;
        p2.inc(p.rb);
        mbar_wait(bar_xfull + 8 * p0.j, p0.ph, 3u);
        mbar_wait(bar_xfull + 8 * p1.j, p1.ph, 3u);
        for (int r = ya; r < yb; ++r) {
          mbar_wait(bar_xfull + 8 * p2.j, p2.ph, 3u);
          mbar_wait(bar_gfull + 8 * gr.j, gr.ph, 4u);
          tc_fence_after();
          const uint64_t ad0 = adesc_t + (uint64_t)((xring + (uint32_t)p0.j * p.slot_bytes) >> 4);
          const uint64_t bd0 = bdesc_t + (uint64_t)((gst0 + (uint32_t)gr.j * kWgNB * p.gstage_bytes) >> 4);
          for (int ks = 0; ks < ksteps; ++ks) {
            if (kmod == iw) {
              // K step ks = 16 pixels: box ks / 2, half ks % 2 of its 32 pixels
              const uint64_t bd = bd0 + (uint64_t)((ks >> 1) * (p.gstage_bytes >> 4) + (ks & 1) * 16);
              uint64_t ad = ad0 + (uint64_t)((ks >> 1) * (p.ring_bytes >> 4) + (ks & 1) * 16);
              uint32_t d = tmem_base;
              for (int m = 0; m < p.mt; ++m, ad += mstep, d += (uint32_t)N3) umma_f16(d, ad, bd, p.idesc, 1u);
            }
            if (++kmod == p.issuers) kmod = 0;
          }
          umma_commit(bar_gempty + 8 * gr.j);
          umma_commit(bar_xempty + 8 * p0.j);                 // row r-1 is not needed again
          if (r == yb - 1) {                                   // ... and neither are the last two rows of the segment
            umma_commit(bar_xempty + 8 * p1.j);
            umma_commit(bar_xempty + 8 * p2.j);
          }
          p0 = p1;
          p1 = p2;
          p2.inc(p.rb);
          gr.inc(kWgGStages);
        }
        cur = p2;
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
    // the read-out warps have nothing to do until the whole range is accumulated: sleep between polls
    while (!mbar_try_wait(bar_done, 0u)) __nanosleep(2000);
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

// ---- v2 ---------------------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 ldg_nc_v4(const void* ptr) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(ptr));
  return v;
}
__device__ __forceinline__ void sts_v4(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

__global__ void __launch_bounds__(kWgThreads, 1)
conv3x3_wgrad_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 127u) & ~127u;
  // header: x row full[8] | x row empty[8] | gy full[8] | gy empty[8] | go[8] | done | zeroed | tmem pointer
  const uint32_t bar_xfull = smem_base;
  const uint32_t bar_xempty = smem_base + 64;
  const uint32_t bar_gfull = smem_base + 128;
  const uint32_t bar_gempty = smem_base + 192;
  const uint32_t bar_go = smem_base + 256;        // [8]: stage granted to loader warp i (by the producer thread, in consumption order)
  const uint32_t bar_done = smem_base + 320;
  const uint32_t bar_zero = smem_base + 328;
  const uint32_t tmem_slot = smem_base + 336;
  const uint32_t xring = smem_base + kSmemHeader;
  const uint32_t group_bytes = kWgBox * 16u;                              // one plane of one box row
  // x: kWgNB rings [box][row slot][plane][32 px] (rows of one box are consecutive M groups); gy stages [box][kx][plane][32 px]
  const uint32_t gst0 = xring + kWgNB * p.ring_bytes;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int N3 = 3 * p.nbn;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.rb; ++s) {
      mbar_init(bar_xfull + 8 * s, 1);
      mbar_init(bar_xempty + 8 * s, (uint32_t)p.issuers);
    }
    for (int s = 0; s < p.gstages; ++s) {
      mbar_init(bar_gfull + 8 * s, (uint32_t)p.ngroups);
      mbar_init(bar_gempty + 8 * s, (uint32_t)p.issuers);
    }
    mbar_init(bar_done, (uint32_t)p.issuers);
    mbar_init(bar_zero, kWgLoaders);
    for (int s = 0; s < kWgLoaders; ++s) mbar_init(bar_go + 8 * s, 1);
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

  if (warp == 0) {
    // ------------------------------------------------------------------ activation rows: one TMA tensor load per box row (box = 32
    // pixels x cp planes, landing as [plane][pixel] = consecutive M groups); rows / pixels outside the image are zero-filled by the
    // TMA unit, which is exactly the convolution's zero padding.
    if (elect_one()) {
      tma_prefetch_desc(&tmX);
      // Two independent duties polled by this one thread: (a) the next activation row as soon as its ring slot is free, (b) the next
      // gradient stage to the loader of the next row unit as soon as the MMAs have released it.  The stages are handed out here, in
      // consumption order, because a parity wait must not run more than one phase ahead of its barrier - which eight independent
      // loaders waiting on three stages could.
      WgRing xr = {0, 0u};
      WgRing gr = {0, 0u};
      long long q = 0;                                  // next work item of the loaders: (row unit, group of four gradient planes)
      const long long total = (u1 - u0) * p.ngroups;
      int gsub = 0;
      long long u = u0;
      int img, x0, ya, yb;
      bool x_pending = wg_next_segment(p, u, u1, img, x0, ya, yb);
      int row = x_pending ? ya - 1 : 0;
      unsigned long long t0 = 0ull;
      uint32_t idle = 0u;
      while (x_pending || q < total) {
        bool progressed = false;
        if (x_pending && mbar_test_wait(bar_xempty + 8 * xr.j, xr.ph ^ 1u)) {
          const int j = xr.j;
          const bool dup = j < 2;                         // ring slots 0,1 are mirrored behind the last slot
          mbar_expect_tx(bar_xfull + 8 * j, kWgNB * p.slot_bytes * (dup ? 2u : 1u));
#pragma unroll
          for (int b = 0; b < kWgNB; ++b) {
            const uint32_t ring = xring + b * p.ring_bytes;
            tma_load_4d(ring + (uint32_t)j * p.slot_bytes, &tmX, bar_xfull + 8 * j, (x0 + b * kWgBox) * 8, row, p.x_po, img);
            if (dup) tma_load_4d(ring + (uint32_t)(p.rb + j) * p.slot_bytes, &tmX, bar_xfull + 8 * j, (x0 + b * kWgBox) * 8, row, p.x_po, img);
          }
          xr.inc(p.rb);
          if (++row > yb) {
            x_pending = wg_next_segment(p, u, u1, img, x0, ya, yb);
            row = ya - 1;
          }
          progressed = true;
        }
        if (q < total && mbar_test_wait(bar_gempty + 8 * gr.j, gr.ph ^ 1u)) {
          mbar_arrive(bar_go + 8 * (int)(q & (kWgLoaders - 1)));
          ++q;
          if (++gsub == p.ngroups) {
            gsub = 0;
            gr.inc(p.gstages);
          }
          progressed = true;
        }
        if (progressed) {
          idle = 0;
        } else if (++idle >= 4096u) {       // (reading the global timer costs more than a poll: only while nothing moves)
          if (idle == 4096u) t0 = (unsigned long long)clock64();
          if (*(volatile unsigned int*)&g_watchdog[0]) break;
          if ((idle & 1023u) == 0u && (unsigned long long)clock64() - t0 > (unsigned long long)kWatchdogCycles) {
            if (atomicExch(&g_watchdog[0], 1u) == 0u) {
              g_watchdog[1] = blockIdx.x; g_watchdog[2] = threadIdx.x; g_watchdog[3] = bar_xempty + 8 * xr.j; g_watchdog[4] = xr.ph ^ 1u;
              g_watchdog[5] = 7u;
            }
            break;
          }
        }
      }
    }
  } else if (warp <= 3) {
    // ------------------------------------------------------------------ MMA issuers (one thread per warp).  The K steps
    // of the launch are dealt round-robin to the issuers; every MMA accumulates into zero-initialised TMEM, so their
    // order is irrelevant, and every buffer is released when all issuers have committed it.
    const int iw = warp - 1;
    if (iw < p.issuers && elect_one()) {
      int kmod = 0;
      if (u0 < u1) mbar_wait(bar_zero, 0u, 6u);
      tc_fence_after();
      const uint64_t adesc_t = make_smem_desc(0u, 128u, group_bytes);   // MN-major: LBO = K-block (8 px) stride, SBO = group stride
      const uint64_t bdesc_t = make_smem_desc(0u, 128u, group_bytes);
      WgRing cur = {0, 0u}, gr = {0, 0u};
      const uint64_t mstep = (uint64_t)((16u * group_bytes) >> 4);
      long long u = u0;
      int img, x0, ya, yb;
      while (wg_next_segment(p, u, u1, img, x0, ya, yb)) {
        const int valid = p.w - x0 < kWgPW ? p.w - x0 : kWgPW;
        const int ksteps = (valid + 15) >> 4;
        // window of output row r = x rows r-1, r, r+1 = three consecutive ring rows p0, p1, p2
        WgRing p0 = cur, p1 = cur, p2;
        p1.inc(p.rb);
        p2 = p1;
        p2.inc(p.rb);
        mbar_wait(bar_xfull + 8 * p0.j, p0.ph, 3u);
        mbar_wait(bar_xfull + 8 * p1.j, p1.ph, 3u);
        for (int r = ya; r < yb; ++r) {
          mbar_wait(bar_xfull + 8 * p2.j, p2.ph, 3u);
          mbar_wait(bar_gfull + 8 * gr.j, gr.ph, 4u);
          tc_fence_after();
          const uint64_t ad0 = adesc_t + (uint64_t)((xring + (uint32_t)p0.j * p.slot_bytes) >> 4);
          const uint64_t bd0 = bdesc_t + (uint64_t)((gst0 + (uint32_t)gr.j * kWgNB * p.gstage_bytes) >> 4);
          for (int ks = 0; ks < ksteps; ++ks) {
            if (kmod == iw) {
              // K step ks = 16 pixels: box ks / 2, half ks % 2 of its 32 pixels
              const uint64_t bd = bd0 + (uint64_t)((ks >> 1) * (p.gstage_bytes >> 4) + (ks & 1) * 16);
              uint64_t ad = ad0 + (uint64_t)((ks >> 1) * (p.ring_bytes >> 4) + (ks & 1) * 16);
              uint32_t d = tmem_base;
              for (int m = 0; m < p.mt; ++m, ad += mstep, d += (uint32_t)N3) umma_f16(d, ad, bd, p.idesc, 1u);
            }
            if (++kmod == p.issuers) kmod = 0;
          }
          umma_commit(bar_gempty + 8 * gr.j);
          umma_commit(bar_xempty + 8 * p0.j);                 // row r-1 is not needed again
          if (r == yb - 1) {                                   // ... and neither are the last two rows of the segment
            umma_commit(bar_xempty + 8 * p1.j);
            umma_commit(bar_xempty + 8 * p2.j);
          }
          p0 = p1;
          p1 = p2;
          p2.inc(p.rb);
          gr.inc(p.gstages);
        }
        cur = p2;
      }
      umma_commit(bar_done);
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ loaders: accumulator zero-fill, gradient rows, final read-out
    const int gw = warp - 4;                 // loader index
    const int wq = warp & 3;                 // TMEM lane quarter this warp may access
    const int half = gw >> 2;                // which 16 of every 32 accumulator columns
    const uint32_t tq = tmem_base + ((uint32_t)(wq * 32) << 16);
    const int cols = p.mt * N3;
    for (int c = half * 16; c < cols; c += 32) tmem_zero16(tq + (uint32_t)c);
    tc_wait_st();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_zero);

    // Work item t = (row unit q, group g of four gradient planes), in the order the issuers consume the units, is loaded by warp t % 8
    // into stage q % gstages: the loads of up to eight items are in flight (in registers) while the stages in shared memory only bridge
    // the loaders and the MMAs.  ngroups is 1 or 2, so a warp always gets the same group: its bias sums cover four fixed planes.
    //   copy kx holds gy[row][x0 + k - kx + 1] at pixel k; the loaded pixel j (x0 + j, j = -1..64) lands at k = j + kx - 1 of copy kx.
    float bsum[4][8];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 8; ++b) bsum[a][b] = 0.f;
    const size_t plane_elems = (size_t)p.h * p.w * 8;
    const bool want_bias = p.bias_part != nullptr;
    const int pg = (gw % p.ngroups) * 4;          // first plane of this warp's group
    {
      // this warp's items t = gw, gw + 8, ...: unit coordinates straight from the unit index (walking the other warps' units
      // cost more instructions than loading its own: ncu, 8 x 130 dependent integer instructions per item)
      const uint32_t nitems = (uint32_t)(u1 - u0) * (uint32_t)p.ngroups;
      uint32_t round = 0;                          // items this warp has taken so far
      for (uint32_t t = (uint32_t)gw; t < nitems; t += kWgLoaders, ++round) {
        const uint32_t q = p.ngroups == 2 ? (t >> 1) : t;
        const uint32_t U = (uint32_t)u0 + q;
        const uint32_t col = U / (uint32_t)p.h;
        const int row = (int)(U - col * (uint32_t)p.h);
        const int img = (int)(col / (uint32_t)p.strips);
        const int x0 = (int)(col - (uint32_t)img * (uint32_t)p.strips) * kWgPW;
        const int s = (int)(q % (uint32_t)p.gstages);
        const uint32_t stage = gst0 + (uint32_t)s * kWgNB * p.gstage_bytes;
        const uint16_t* grow = reinterpret_cast<const uint16_t*>(p.gy) +
                               (((size_t)img * p.gy_pt + p.gy_po + nblk * p.cpb + pg) * p.h + row) * (size_t)p.w * 8;
        uint4 v[4][3];
#pragma unroll
        for (int pp = 0; pp < 4; ++pp) {
          const bool okp = pg + pp < p.cpb && nblk * p.cpb + pg + pp < p.gyp;
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const int j = lane + 32 * c - 1, gx = x0 + j;
            v[pp][c] = make_uint4(0u, 0u, 0u, 0u);
            if (okp && j <= kWgPW && gx >= 0 && gx < p.w) v[pp][c] = ldg_nc_v4(grow + (size_t)pp * plane_elems + (size_t)gx * 8);
          }
        }
        mbar_wait(bar_go + 8 * gw, round & 1u, 5u);      // the stage is free (its previous MMAs have completed)
#pragma unroll
        for (int pp = 0; pp < 4; ++pp) {
          if (pg + pp < p.cpb) {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              const int j = lane + 32 * c - 1;
              if (j <= kWgPW) {
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                  const int k = j + kx - 1;
                  if (k >= 0 && k < kWgPW)
                    sts_v4(stage + (uint32_t)(k >> 5) * p.gstage_bytes + (uint32_t)((kx * p.cpb + pg + pp) * kWgBox + (k & 31)) * 16u, v[pp][c]);
                }
              }
            }
          }
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_gfull + 8 * s);
        if (want_bias) {      // the bias gradient from the registers, off the stage's critical path
#pragma unroll
          for (int pp = 0; pp < 4; ++pp) {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              const int j = lane + 32 * c - 1;
              if (j >= 0 && j < kWgPW) {
                float a[8];
                unpack8(v[pp][c], p.dtype, a);
#pragma unroll
                for (int e = 0; e < 8; ++e) bsum[pp][e] += a[e];
              }
            }
          }
        }
      }
    }
    // nothing more to do until the whole range is accumulated: sleep between polls
    while (!mbar_try_wait(bar_done, 0u)) __nanosleep(1000);
    tc_fence_after();
    if (want_bias) {
      // per-channel sums of this CTA's gradient rows: lanes -> warp (shuffles) -> CTA (through the now idle gradient stages)
      float* red = reinterpret_cast<float*>(smem_raw + (gst0 - smem_u32(smem_raw)));
#pragma unroll
      for (int pp = 0; pp < 4; ++pp) {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          float t = bsum[pp][e];
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
          if (lane == 0) red[gw * 32 + pp * 8 + e] = t;
        }
      }
      asm volatile("bar.sync 1, %0;" ::"n"(32 * kWgLoaders) : "memory");
      if (gw == 0) {
        for (int ch = lane; ch < p.cpb * 8; ch += 32) {
          float t = 0.f;
          const int grp = ch >> 5;                 // warps w with w % ngroups == grp hold this channel's group
          for (int w8 = grp; w8 < kWgLoaders; w8 += p.ngroups) t += red[w8 * 32 + (ch & 31)];
          p.bias_part[(size_t)blockIdx.x * kWgBiasStride + ch] = t;
        }
      }
    }
    float* out = p.part + (((size_t)blockIdx.x * p.mt) * 128 + wq * 32 + lane) * N3;
    for (int m = 0; m < p.mt; ++m) {
      for (int c = half * 16; c < N3; c += 32) {
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

// dW[co][ci][ky][kx] (+)= scale * sum over CTAs of the partial accumulators.  One thread per accumulator element in the
// partials' own memory order (coalesced reads of every CTA's block), scattered store into the OIHW gradient.
// grid = (ceil(mt*128*N3 / 256), n_blocks)
// bias_part (optional): db[co] (+)= scale * sum over CTAs of the v2 kernel's per-CTA bias partials, by the first block of every n-block.
__global__ void wgrad_reduce_kernel(const float* __restrict__ part, int ranges, int n_blocks, int mt, int nbn, int cp, int cout, int cin,
                                    int lead, float scale, int accumulate, float* __restrict__ dw, int cin_total, int cin_off,
                                    const float* __restrict__ bias_part, float* __restrict__ db) {
  const int lead_pad = (lead + 7) / 8 * 8;
  const int N3 = 3 * nbn;
  const int per_cta = mt * 128 * N3;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  const int nblk = blockIdx.y;
  if (bias_part && blockIdx.x == 0) {
    // 64 channels x 4 interleaved slices of the CTA list, four loads in flight per thread; combined in a fixed order
    __shared__ float bsh[4][kWgBiasStride];
    const int ch = threadIdx.x & (kWgBiasStride - 1), sl = threadIdx.x >> 6;
    float b4[4] = {0.f, 0.f, 0.f, 0.f};
    if (ch < nbn) {
      int rg = sl;
      for (; rg + 12 < ranges; rg += 16) {
#pragma unroll
        for (int k = 0; k < 4; ++k) b4[k] += __ldg(bias_part + (size_t)((rg + 4 * k) * n_blocks + nblk) * kWgBiasStride + ch);
      }
      for (; rg < ranges; rg += 4) b4[0] += __ldg(bias_part + (size_t)(rg * n_blocks + nblk) * kWgBiasStride + ch);
    }
    bsh[sl][ch] = (b4[0] + b4[1]) + (b4[2] + b4[3]);
    __syncthreads();
    if (sl == 0 && ch < nbn) {
      const int co = nblk * nbn + ch;
      if (co < cout) {
        const float b = ((bsh[0][ch] + bsh[1][ch]) + (bsh[2][ch] + bsh[3][ch])) * scale;
        db[co] = accumulate ? db[co] + b : b;
      }
    }
  }
  if (e >= per_cta) return;
  const int col = e % N3, L = (e / N3) % 128, m = e / (N3 * 128);
  const int G = (m << 4) + (L >> 3);
  const int ky = G / cp, plane = G - ky * cp;
  const int pc = plane * 8 + (L & 7);
  const int ci = pc < lead_pad ? (pc < lead ? pc : -1) : pc - lead_pad + lead;
  const int kx = col / nbn, co = nblk * nbn + (col - kx * nbn);
  if (ky >= 3 || ci < 0 || ci >= cin || co >= cout) return;
  // eight independent partial sums keep eight loads in flight per thread; the order is fixed (bit-reproducible)
  float a8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const float* src = part + (size_t)nblk * per_cta + e;
  const size_t rstride = (size_t)n_blocks * per_cta;
  int rg = 0;
  for (; rg + 8 <= ranges; rg += 8) {
#pragma unroll
    for (int k = 0; k < 8; ++k) a8[k] += __ldg(src + (size_t)(rg + k) * rstride);
  }
  for (; rg < ranges; ++rg) a8[0] += __ldg(src + (size_t)rg * rstride);
  float acc = ((a8[0] + a8[1]) + (a8[2] + a8[3])) + ((a8[4] + a8[5]) + (a8[6] + a8[7]));
  acc *= scale;
  float* dst = dw + (((size_t)co * cin_total + cin_off + ci) * 3 + ky) * 3 + kx;
  *dst = accumulate ? *dst + acc : acc;
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
  // block-level sum first: same-address atomics from every warp of a large grid serialise in L2
  __shared__ float red[8][8];
  const int wid = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int k = 0; k < 8; ++k) red[wid][k] = s[k];
  }
  __syncthreads();
  if (threadIdx.x < 8) {
    float t = 0.f;
    for (int w8 = 0; w8 < (int)(blockDim.x >> 5); ++w8) t += red[w8][threadIdx.x];
    const int co = plane * 8 + threadIdx.x;
    if (co < cout) atomicAdd(db + co, t * scale);
  }
}

// out[c] += scale * sum_{n,y,x} src[n][c][y][x]  (NCHW fp32: bias gradient of the last conv straight from dL/dG, which
// has not been rounded to 16 bits).  grid = (chunks, c, n)
// (n_loop > 0: one block per channel walks all n_loop images and adds its total in a fixed order - the deterministic variant)
__global__ void sum_nchw_kernel(const float* __restrict__ src, int c, size_t hw, float scale, float* __restrict__ out, int n_loop) {
  float s = 0.f;
  const int n0 = n_loop > 0 ? 0 : (int)blockIdx.z, n1 = n_loop > 0 ? n_loop : (int)blockIdx.z + 1;
  for (int img = n0; img < n1; ++img) {
    const float* p = src + ((size_t)img * c + blockIdx.y) * hw;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < hw; i += (size_t)gridDim.x * blockDim.x) s += __ldg(p + i);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (n_loop > 0) {
    __shared__ float red[8];
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.f;
      for (int w8 = 0; w8 < (int)(blockDim.x >> 5); ++w8) t += red[w8];
      out[blockIdx.y] += t * scale;
    }
  } else if ((threadIdx.x & 31) == 0) {
    atomicAdd(out + blockIdx.y, s * scale);
  }
}

}  // namespace esr
