// Row-streaming 3x3 convolution (sm_100a, tcgen05): the three vertical taps are merged into the MMA's N dimension.
//
//   For one INPUT row yi of a 128-pixel column strip and one horizontal tap dx
//       D'[x, (ky, co)] = sum_cin  A[yi, x + dx - 1, cin] * W[ky, dx, cin, co]            (M = 128, N = 3*Cout_block)
//   and block ky of D' belongs to OUTPUT row  y = yi + 1 - ky.  The accumulators of consecutive output rows sit in
//   consecutive TMEM column slots (descending row order), so ONE tcgen05.mma of N = 3*NBN adds the three blocks
//   straight into the accumulators of output rows yi+1, yi, yi-1: no partial sums are ever read back.
//
// Why: an SS-mode tcgen05.mma re-reads its 128x16 A tile (4 KB) from shared memory at 128 B/clk = 32 clk, while the
// tensor pipe needs only N/2 clk; at N = 32 (the dense block's growth convs, 54 % of the FLOPs) the tile kernel is
// operand-bandwidth bound at 40 %.  Merging the vertical taps triples N for the same A read (measured on B200:
// 56 clk per N=96 MMA vs 3 x 40 clk, 96 clk per N=192 MMA = tensor bound) and issues 3x fewer MMAs.  A CTA marches
// down its strip, so every activation row is fetched from L2 exactly once (no halo re-fetch in y), all 128 lanes of
// an M tile are real output pixels (130-pixel rows in shared memory carry the x halo), and the weights of the CTA's
// n-block stay resident in shared memory for the whole launch.
//
// Work = n * strips * h output rows, cut into `ranges` contiguous ranges; CTA b handles n-block b % n_blocks of range
// b / n_blocks (CTAs that share a range run side by side, so the second read of the activations hits L2).
// Roles as in the tile kernel: warp 0 bulk-copy producer, warp 1 MMA issuer, warps 4..11 epilogue (two per TMEM lane
// quarter, alternating output rows).  Pipelines: activation stages (one 8-plane K chunk of one input row each) and a
// ring of `slots` TMEM accumulator slots, one output row each, released row by row.  The epilogue zeroes a slot after
// draining it, so every MMA accumulates and no MMA has to be split to start an accumulator.
#pragma once
#include "conv3x3_tc.cuh"

namespace esr {

constexpr uint32_t kRowPx = 130;                              // 128 output pixels + one halo pixel per side
constexpr uint32_t kRowBytes = kRowPx * 16;                   // one plane of one input row in shared memory
constexpr int kRowsKch = 8;                                   // planes per K chunk (= pipeline stage): four K=16 steps
constexpr uint32_t kRowsStageBytes = kRowsKch * kRowBytes;    // 16640
constexpr int kRowsMaxStages = 12;
constexpr int kRowsMaxSlots = 16;

__device__ __forceinline__ void st_shared_zero16(uint32_t addr) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(addr), "r"(0u) : "memory");
}
// zero 32 lanes x 16 consecutive TMEM columns
__device__ __forceinline__ void tmem_zero16(uint32_t taddr) {
  const uint32_t z = 0u;
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(taddr),
      "r"(z)
      : "memory");
}
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// next run of rows of [u, u1) that lies inside one (image, strip) column: output rows [ya, yb) at x0
__device__ __forceinline__ bool rows_next_segment(const ConvParams& p, long long& u, long long u1, int& img, int& x0, int& ya,
                                                  int& yb) {
  if (u >= u1) return false;
  const long long col = u / p.h;
  ya = (int)(u - col * p.h);
  const long long rem = u1 - u;
  yb = (rem < (long long)(p.h - ya)) ? ya + (int)rem : p.h;
  img = (int)(col / p.strips);
  x0 = (int)(col - (long long)img * p.strips) * 128;
  u += yb - ya;
  return true;
}

// the (up to) twelve MMAs of one K chunk of one input row: K step j (plane pair) x horizontal tap dx, each over the row's
// accumulator blocks (one MMA, or two when the slot ring wraps inside the blocks).  Every descriptor is a base plus an
// immediate.
template <int N3, bool kTwo>
__device__ __forceinline__ void rows_issue_chunk(uint64_t ad0, uint64_t bdA, uint64_t bdB, uint32_t dA, uint32_t dB, uint32_t idA,
                                                 uint32_t idB, int nk) {
#pragma unroll
  for (int j = 0; j < kRowsKch / 2; ++j) {
    if (j < nk) {
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        const uint64_t ao = (uint64_t)(j * 2 * kRowPx + dx), bo = (uint64_t)((dx * kRowsKch + 2 * j) * N3);
        umma_f16(dA, ad0 + ao, bdA + bo, idA, 1u);
        if (kTwo) umma_f16(dB, ad0 + ao, bdB + bo, idB, 1u);
      }
    }
  }
}

template <int NBN, bool kBwd, int kEpi>
__global__ void __launch_bounds__(kConvThreads, 1) conv3x3_rows_kernel(const __grid_constant__ ConvParams p) {
  extern __shared__ uint8_t smem_raw[];
  constexpr int N3 = 3 * NBN;
  constexpr uint32_t kChunkW = 3u * kRowsKch * N3 * 16u;   // weights of one K chunk: [dx][plane][ky*NBN+co][8 cin]
  const uint32_t smem_base = (smem_u32(smem_raw) + 127u) & ~127u;
  // header: stage full[16] | stage empty[16] | row full[16] | row empty[16] | weights barrier | zeroed barrier | tmem pointer
  const uint32_t bar_full = smem_base;
  const uint32_t bar_empty = smem_base + 128;
  const uint32_t bar_rfull = smem_base + 256;
  const uint32_t bar_rempty = smem_base + 384;
  const uint32_t bar_w = smem_base + 512;
  const uint32_t bar_zero = smem_base + 520;
  const uint32_t tmem_slot = smem_base + 528;
  const uint32_t wres = smem_base + kSmemHeader;
  const uint32_t stage0 = wres + p.w_bytes;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int S = p.slots;                      // power of two
  const int smask = S - 1, sshift = 31 - __clz(S);

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int s = 0; s < S; ++s) {
      mbar_init(bar_rfull + 8 * s, (uint32_t)p.issuers);
      mbar_init(bar_rempty + 8 * s, 4);
    }
    mbar_init(bar_w, 1);
    mbar_init(bar_zero, 8);
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
  if (threadIdx.x == 0) pdl_launch_dependents();

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    if (warp == 0) {
      // ------------------------------------------------------------------ producer (bulk copies)
      // One lane per plane of a stage: a single thread issuing all copies of a chunk (address arithmetic included)
      // takes longer than the MMAs of that chunk.
      if (u0 < u1 && elect_one()) {
        mbar_expect_tx(bar_w, p.w_bytes);
        const uint8_t* wsrc = p.wts + (size_t)nblk * p.w_bytes;
        for (int c = 0; c < p.nchunks; ++c) bulk_load(wres + c * kChunkW, wsrc + (size_t)c * kChunkW, kChunkW, bar_w);
      }
      pdl_wait();  // weights do not depend on the previous launch, activations do
      // ONE thread issues every copy, with addresses advanced by additions only.  (One lane per plane looked parallel in the source
      // but compiled into an ELECT loop that moved every lane's addresses into uniform registers: ~150 cycles per copy, and with 8-24
      // copies per input row the producer warp - not the tensor pipe - set the pace of the whole kernel: ncu, round 2.)
      if (elect_one()) {
        int s = 0;
        uint32_t ph = 0;
        const size_t plane16 = hw * 16;
        // chunks per plane segment (split precision: three segments hi | lo | hi, else one); the last chunk of a segment may be
        // partial.  K steps take plane pairs: an odd tail re-loads its last plane (zero weights)
        const int cps = p.split ? p.cps : p.nchunks;
        const int npl_last = p.cin_planes - (cps - 1) * kRowsKch;
        const int nld_last = (npl_last + 1) & ~1;
        int u_img, x0, ya, yb;
        long long u = u0;
        while (rows_next_segment(p, u, u1, u_img, x0, ya, yb)) {
          const int yi0 = ya > 0 ? ya - 1 : 0, yi1 = yb < p.h ? yb : p.h - 1;
          const int xs = x0 > 0 ? x0 - 1 : 0, xe = x0 + 129 < p.w ? x0 + 129 : p.w;
          const uint32_t cnt_bytes = (uint32_t)(xe - xs) * 16u, dst_off = (uint32_t)(xs - (x0 - 1)) * 16u;
          const bool zl = x0 == 0, zr = x0 + 129 > p.w;   // image border inside this strip: the halo pixel is zero padding
          const uint32_t zr_off = (uint32_t)(p.w - (x0 - 1)) * 16u;
          const uint8_t* colp = p.in + (((size_t)u_img * p.in_pt) * hw + xs) * 16;
          int seg_uses = 0;
          for (int yi = yi0; yi <= yi1; ++yi) {
            const uint8_t* rowp = colp + (size_t)yi * p.w * 16;
            int cc = 0, sg = 0;
            for (int c = 0; c < p.nchunks; ++c) {
              const bool last = cc == cps - 1;
              const int nld = last ? nld_last : kRowsKch;
              const int npl = last ? npl_last : kRowsKch;
              const int plane0 = (p.split ? p.seg_base[sg] : p.in_plane_off) + cc * kRowsKch;
              mbar_wait(bar_empty + 8 * s, ph ^ 1u, 1u);
              const uint32_t sa = stage0 + s * kRowsStageBytes;
              if ((zl || zr) && seg_uses < p.stages) {
                // border pixels are never written by this segment's copies: zero each stage buffer once per segment
#pragma unroll
                for (int k = 0; k < kRowsKch; ++k) {
                  if (zl) st_shared_zero16(sa + (uint32_t)k * kRowBytes);
                  if (zr) st_shared_zero16(sa + zr_off + (uint32_t)k * kRowBytes);
                }
                fence_proxy_async();
              }
              const uint32_t full = bar_full + 8 * s;
              mbar_expect_tx(full, (uint32_t)nld * cnt_bytes);
              const uint8_t* src = rowp + (size_t)plane0 * plane16;
              uint32_t dst = sa + dst_off;
#pragma unroll
              for (int k = 0; k < kRowsKch; ++k) {
                if (k < nld) {
                  bulk_load(dst, src, cnt_bytes, full);
                  dst += kRowBytes;
                  if (k + 1 < npl) src += plane16;       // (the odd tail re-loads the last plane)
                }
              }
              ++seg_uses;
              if (++s == p.stages) { s = 0; ph ^= 1u; }
              if (++cc == cps) { cc = 0; ++sg; }
            }
          }
        }
      }
      __syncwarp();
    } else {
      // ------------------------------------------------------------------ MMA issuers: ONE thread per warp runs the role.
      // A single thread cannot issue (and book-keep) as fast as the tensor pipe retires N=96 MMAs, so up to three warps
      // share the stream of K chunks round-robin.  All of them walk the same (row, chunk) sequence; each waits for and
      // issues only the chunks it owns.  Every MMA accumulates (the epilogue hands slots back zeroed), so MMAs of
      // different issuers on the same accumulator commute; a row is complete when every issuer has committed it.
      const int iw = warp - 1, NI = p.issuers;
      if (iw < NI && elect_one()) {
        const uint64_t adesc_t = make_smem_desc(0u, kRowBytes, 128u);
        const uint64_t bdesc_t = make_smem_desc(0u, (uint32_t)N3 * 16u, 128u) + (uint64_t)(wres >> 4);
        const int nchunks = p.nchunks, stages = p.stages;
        const int cps = p.split ? p.cps : nchunks;
        const int nk_last = (p.cin_planes - (cps - 1) * kRowsKch + 1) >> 1;
        const uint32_t id1 = p.idesc_n[0], id2 = p.idesc_n[1], id3 = p.idesc_n[2];
        if (u0 < u1) {
          mbar_wait(bar_w, 0u, 5u);
          mbar_wait(bar_zero, 0u, 6u);
        }
        tc_fence_after();
        int s = 0, kmod = 0;
        uint32_t ph = 0;
        int g0 = 0;   // index of the segment's first output row in this CTA's sequence of rows: row g uses slot S-1-g%S
        long long u = u0;
        int img, x0, ya, yb;
        while (rows_next_segment(p, u, u1, img, x0, ya, yb)) {
          const int yi0 = ya > 0 ? ya - 1 : 0, yi1 = yb < p.h ? yb : p.h - 1;
          for (int yi = yi0; yi <= yi1; ++yi) {
            // vertical taps of this input row that land on output rows of the segment: y = yi + 1 - ky in [ya, yb)
            const int ky_lo = yi + 2 - yb > 0 ? yi + 2 - yb : 0;
            const int ky_hi = yi + 1 - ya < 2 ? yi + 1 - ya : 2;
            const int nbk = ky_hi - ky_lo + 1;
            const int g_top = g0 + (yi + 1 - ky_lo - ya);   // highest output row touched = first block
            const int slot0 = smask - (g_top & smask);
            // rows touched for the first time need their slot drained (and zeroed) by the epilogue
            if (yi == yi0) {
              for (int b = 0; b < nbk; ++b) {
                const int g = g_top - b;
                mbar_wait(bar_rempty + 8 * (smask - (g & smask)), (uint32_t)((g >> sshift) & 1) ^ 1u, 2u);
              }
            } else if (ky_lo == 0) {
              mbar_wait(bar_rempty + 8 * slot0, (uint32_t)((g_top >> sshift) & 1) ^ 1u, 2u);
            }
            tc_fence_after();
            // blocks [0, nA) go to slots slot0.., blocks [nA, nbk) wrap to slot 0
            const int wrapn = S - slot0;
            const int nA = nbk < wrapn ? nbk : wrapn, nB = nbk - nA;
            const uint32_t dA = tmem_base + (uint32_t)(slot0 * NBN), dB = tmem_base;
            const uint32_t idA = nA == 3 ? id3 : (nA == 2 ? id2 : id1), idB = nB == 2 ? id2 : id1;
            const uint64_t bA = bdesc_t + (uint64_t)(ky_lo * NBN), bB = bdesc_t + (uint64_t)((ky_lo + nA) * NBN);
            for (int c = 0, cc = 0; c < nchunks; ++c) {
              const bool last_of_seg = ++cc == cps;
              if (last_of_seg) cc = 0;
              if (kmod == iw) {
                mbar_wait(bar_full + 8 * s, ph, 3u);
                tc_fence_after();
                const uint64_t ad0 = adesc_t + (uint64_t)((stage0 + s * kRowsStageBytes) >> 4);
                const uint64_t wo = (uint64_t)(c * (kChunkW >> 4));
                const int nk = last_of_seg ? nk_last : kRowsKch / 2;
                if (nB == 0) rows_issue_chunk<N3, false>(ad0, bA + wo, bB + wo, dA, dB, idA, idB, nk);
                else rows_issue_chunk<N3, true>(ad0, bA + wo, bB + wo, dA, dB, idA, idB, nk);
                umma_commit(bar_empty + 8 * s);
              }
              if (++kmod == NI) kmod = 0;
              if (++s == stages) { s = 0; ph ^= 1u; }
            }
            // output row yi-1 has now seen all three input rows; the last row of the image completes with yi itself
            if (yi - 1 >= ya) umma_commit(bar_rfull + 8 * (smask - ((g0 + yi - 1 - ya) & smask)));
            if (yi == yi1 && yi < yb) umma_commit(bar_rfull + 8 * (smask - ((g0 + yi - ya) & smask)));
          }
          g0 += yb - ya;
        }
      }
      __syncwarp();
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
    // ------------------------------------------------------------------ epilogue (warps 4..11)
    const int wq = warp & 3;            // TMEM lane quarter this warp may touch = 32-pixel quarter of the strip
    const int eh = (warp - 4) >> 2;     // parity of the output rows this warp takes
    const uint32_t tq = tmem_base + ((uint32_t)(wq * 32) << 16);
    // all accumulators start from zero: this warp clears its lane quarter of one half of the 512 columns
#pragma unroll
    for (int k = 0; k < 16; ++k) tmem_zero16(tq + (uint32_t)(eh * 256 + k * 16));
    tc_wait_st();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_zero);
    pdl_wait();  // residual reads and all stores must not overtake the previous launch
    int g0 = 0;
    long long u = u0;
    int img, x0, ya, yb;
    while (rows_next_segment(p, u, u1, img, x0, ya, yb)) {
      const int x = x0 + wq * 32 + lane;
      const bool valid = x < p.w;
      for (int y = ya; y < yb; ++y) {
        const int g = g0 + (y - ya);
        if ((g & 1) != eh) continue;
        const int slot = smask - (g & smask);
        mbar_wait(bar_rfull + 8 * slot, (uint32_t)((g >> sshift) & 1), 4u);
        tc_fence_after();
        const uint32_t trow = tq + (uint32_t)(slot * NBN);
        if (kEpi == 0) conv_epilogue_px<NBN, kBwd>(p, trow, img, y, x, valid, nblk);
        else if (kEpi == 3) conv_epilogue_nchw<NBN>(p, trow, img, y, x, valid);
        else conv_epilogue_fast<NBN, kEpi>(p, trow, img, y, x, valid, nblk);
        // hand the slot back zeroed: the next output row that uses it accumulates from its first MMA on
#pragma unroll
        for (int k = 0; k < NBN / 16; ++k) tmem_zero16(trow + k * 16);
        tc_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_rempty + 8 * slot);
      }
      g0 += yb - ya;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512u);
  }
}

// ---------------------------------------------------------------------------------------------
// weight packing for the row kernel: OIHW fp32 -> [n_block][chunk][dx][plane-in-chunk (4)][ky*NBN + co][8 cin] 16-bit
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint16_t pack_weights_rows_elem(const float* __restrict__ w, int cout, int cin, int lead, int nb_n, int nchunks, int dtype,
                                                           int transpose_flip, size_t idx) {
  const int lc_out = transpose_flip ? cin : cout;
  const int lc_in = transpose_flip ? cout : cin;
  const int lead_pad = (lead + 7) / 8 * 8;
  const int n3 = 3 * nb_n;
  size_t r = idx;
  const int ci8 = r % 8; r /= 8;
  const int nn = r % n3; r /= n3;
  const int j = r % kRowsKch; r /= kRowsKch;
  const int dx = r % 3; r /= 3;
  int c = r % nchunks; r /= nchunks;
  const int nb = (int)r;
  int seg = 0;
  if (dtype == 2) { const int cps = nchunks / 3; seg = c / cps; c -= seg * cps; }
  const int ky = nn / nb_n;
  int o = nb * nb_n + (nn - ky * nb_n);
  int i = (c * kRowsKch + j) * 8 + ci8;
  if (!transpose_flip) {
    if (i < lead_pad) i = i < lead ? i : -1;
    else i = i - lead_pad + lead;
  } else {
    if (o < lead_pad) o = o < lead ? o : -1;
    else o = o - lead_pad + lead;
  }
  float val = 0.f;
  if (o >= 0 && o < lc_out && i >= 0 && i < lc_in) {
    if (!transpose_flip) val = w[(((size_t)o * cin + i) * 3 + ky) * 3 + dx];
    else val = w[(((size_t)i * cin + o) * 3 + (2 - ky)) * 3 + (2 - dx)];
  }
  if (dtype == 2) return split_weight_bits(val, seg);
  return (uint16_t)(pack2(val, 0.f, dtype) & 0xFFFFu);
}

__global__ void pack_weights_rows_kernel(const float* __restrict__ w, int cout, int cin, int lead, int nb_n, int nchunks, int dtype,
                                         int transpose_flip, uint16_t* __restrict__ dst, size_t total) {
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x)
    dst[idx] = pack_weights_rows_elem(w, cout, cin, lead, nb_n, nchunks, dtype, transpose_flip, idx);
}

// Every conv of a network re-packed by ONE launch (after an optimizer step): blockIdx.y = conv, the job table lives in device
// memory.  Per job: the tile-kernel image, the row-kernel image (optional) and the zero-padded fp32 bias.
struct PackJob {
  const float* w; int cout, cin, lead, kcp, nb_n, nchunks, dtype, transpose_flip;
  uint16_t* dst; unsigned long long total;
  int rows_nb_n, rows_nchunks; uint16_t* rows_dst; unsigned long long rows_total;
  float* bias_out; const float* bias_in; int cout_pad, bias_n;
};

// eight consecutive packed elements (one 16-byte core-matrix row: input channel positions i0 .. i0+7 of output position o at tap
// (ky, kx)) -> one uint4.  The element decomposition (divisions by run-time extents) is paid once per eight elements.
__device__ __forceinline__ uint4 pack_weights_vec8(const float* __restrict__ w, int cout, int cin, int lead, int dtype, int transpose_flip, int seg,
                                                   int o, int i0, int ky, int kx) {
  const int lc_out = transpose_flip ? cin : cout;
  const int lc_in = transpose_flip ? cout : cin;
  const int lead_pad = (lead + 7) / 8 * 8;
  if (transpose_flip) {
    if (o < lead_pad) o = o < lead ? o : -1;
    else o = o - lead_pad + lead;
  }
  uint32_t h[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    int i = i0 + k;
    if (!transpose_flip) {
      if (i < lead_pad) i = i < lead ? i : -1;
      else i = i - lead_pad + lead;
    }
    float val = 0.f;
    if (o >= 0 && o < lc_out && i >= 0 && i < lc_in) {
      if (!transpose_flip) val = __ldg(w + (((size_t)o * cin + i) * 3 + ky) * 3 + kx);
      else val = __ldg(w + (((size_t)i * cin + o) * 3 + (2 - ky)) * 3 + (2 - kx));
    }
    h[k] = dtype == 2 ? (uint32_t)split_weight_bits(val, seg) : (pack2(val, 0.f, dtype) & 0xFFFFu);
  }
  return make_uint4(h[0] | (h[1] << 16), h[2] | (h[3] << 16), h[4] | (h[5] << 16), h[6] | (h[7] << 16));
}

__global__ void pack_weights_batch_kernel(const PackJob* __restrict__ jobs) {
  const PackJob jb = jobs[blockIdx.y];
  const uint32_t v_tile = (uint32_t)(jb.total >> 3), v_rows = (uint32_t)(jb.rows_total >> 3);
  const uint32_t all = v_tile + v_rows + (uint32_t)((jb.cout_pad + 7) >> 3);
  for (uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x; idx < all; idx += gridDim.x * blockDim.x) {
    if (idx < v_tile) {
      // tile-kernel image: [n-block][chunk][tap][plane of the chunk][n][8 channels]
      uint32_t r = idx;
      const int n = r % jb.nb_n; r /= jb.nb_n;
      const int j = r % jb.kcp; r /= jb.kcp;
      const int tap = r % 9; r /= 9;
      int c = r % jb.nchunks; r /= jb.nchunks;
      int seg = 0;
      if (jb.dtype == 2) { const int cps = jb.nchunks / 3; seg = c / cps; c -= seg * cps; }
      reinterpret_cast<uint4*>(jb.dst)[idx] = pack_weights_vec8(jb.w, jb.cout, jb.cin, jb.lead, jb.dtype, jb.transpose_flip, seg,
                                                                (int)r * jb.nb_n + n, (c * jb.kcp + j) * 8, tap / 3, tap % 3);
    } else if (idx < v_tile + v_rows) {
      // row-kernel image: [n-block][chunk][dx][plane of the chunk][(ky, n)][8 channels]
      uint32_t r = idx - v_tile;
      const int n3 = 3 * jb.rows_nb_n;
      const int nn = r % n3; r /= n3;
      const int j = r % kRowsKch; r /= kRowsKch;
      const int dx = r % 3; r /= 3;
      int c = r % jb.rows_nchunks; r /= jb.rows_nchunks;
      int seg = 0;
      if (jb.dtype == 2) { const int cps = jb.rows_nchunks / 3; seg = c / cps; c -= seg * cps; }
      const int ky = nn / jb.rows_nb_n;
      reinterpret_cast<uint4*>(jb.rows_dst)[idx - v_tile] = pack_weights_vec8(jb.w, jb.cout, jb.cin, jb.lead, jb.dtype, jb.transpose_flip, seg,
                                                                             (int)r * jb.rows_nb_n + (nn - ky * jb.rows_nb_n),
                                                                             (c * kRowsKch + j) * 8, ky, dx);
    } else {
      const int k0 = (int)(idx - v_tile - v_rows) * 8;
      for (int k = k0; k < k0 + 8 && k < jb.cout_pad; ++k) jb.bias_out[k] = (jb.bias_in && k < jb.bias_n) ? jb.bias_in[k] : 0.f;
    }
  }
}

}  // namespace esr
