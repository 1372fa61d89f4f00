// C-ABI of the B200-native hot path (see include/esr_b200.h).  Host side: argument validation, TMA
// descriptor encoding, tile/pipeline sizing, launches.  No torch types, no exceptions across the boundary.
#include "../../include/esr_b200.h"

#include <cuda.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <mutex>
#include <set>
#include <vector>

#include "aux_kernels.cuh"
#include "cem_kernels.cuh"
#include "conv3x3_tc.cuh"
#include "conv3x3_rows.cuh"
#include "conv3x3_wgrad.cuh"
#include "disc_kernels.cuh"
#include "train_kernels.cuh"
#include "hist_kernels.cuh"

namespace {

thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

#define CUDA_TRY(expr)                                                                           \
  do {                                                                                           \
    cudaError_t _e = (expr);                                                                     \
    if (_e != cudaSuccess) return fail(ESR_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(_e)); \
  } while (0)

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

// ESR_DETERMINISTIC=1 / esr_set_deterministic(1): every reduction takes a fixed order (one MMA issuer per CTA in the row and wgrad kernels,
// no floating-point atomics across blocks): launches become bit-reproducible at ~15 % of the conv throughput.
std::atomic<int> g_deterministic{-1};
bool deterministic() {
  int v = g_deterministic.load();
  if (v < 0) {
    const char* e = getenv("ESR_DETERMINISTIC");
    v = (e && e[0] != '0') ? 1 : 0;
    g_deterministic.store(v);
  }
  return v != 0;
}

int grid_for(size_t total, int threads) {
  size_t b = (total + threads - 1) / threads;
  const size_t cap = 148 * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

int num_sms() {
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  return sms;
}

typedef void (*ConvKernelFn)(const CUtensorMap, const esr::ConvParams);

template <int P, int KCP, bool kBwd>
ConvKernelFn select_n(int nb_n) {
  switch (nb_n) {
    case 16: return esr::conv3x3_tc_kernel<P, KCP, 16, kBwd>;
    case 32: return esr::conv3x3_tc_kernel<P, KCP, 32, kBwd>;
    case 48: return esr::conv3x3_tc_kernel<P, KCP, 48, kBwd>;
    case 64: return esr::conv3x3_tc_kernel<P, KCP, 64, kBwd>;
  }
  return nullptr;
}
ConvKernelFn select_conv_kernel(int P, int kcp, int nb_n, bool bwd) {
  if (P != 32) return nullptr;
  if (kcp == 4) return bwd ? select_n<32, 4, true>(nb_n) : select_n<32, 4, false>(nb_n);
  if (kcp == 2) return bwd ? select_n<32, 2, true>(nb_n) : select_n<32, 2, false>(nb_n);
  return nullptr;
}

int nblock_for(int cout, int* cout_pad) {
  int nb_n;
  if (cout <= 64) {
    nb_n = (cout + 15) / 16 * 16;
    *cout_pad = nb_n;
  } else {
    nb_n = 64;
    *cout_pad = (cout + 63) / 64 * 64;
  }
  return nb_n;
}

const uint32_t kSmemMax = 232448u;

// n-block of the row-streaming kernel for (cin_planes, cout); 0 = this conv cannot use it
int rows_nbn_for(int cin_planes, int cout, int segs = 1) {
  const int nchunks = segs * ((cin_planes + esr::kRowsKch - 1) / esr::kRowsKch);
  auto fits = [&](int nbn, int min_stages) {
    const size_t wb = (size_t)nchunks * 3 * esr::kRowsKch * 3 * nbn * 16;
    return esr::kSmemHeader + 128u + wb + (size_t)min_stages * esr::kRowsStageBytes <= kSmemMax;
  };
  if (cout <= 16) return fits(16, 6) ? 16 : 0;
  if (cout > 32 && cout <= 64 && fits(64, 8)) return 64;
  return fits(32, 4) ? 32 : 0;
}

typedef void (*RowsKernelFn)(const esr::ConvParams);
// epi: 0 generic epilogue, 1 / 2 / 4 the specialised epilogues (conv_epilogue_fast), 3 the NCHW image store
RowsKernelFn select_rows_kernel(int nbn, bool bwd, int epi) {
  if (bwd) {
    switch (nbn) {
      case 16: return esr::conv3x3_rows_kernel<16, true, 0>;
      case 32: return esr::conv3x3_rows_kernel<32, true, 0>;
      case 64: return esr::conv3x3_rows_kernel<64, true, 0>;
    }
    return nullptr;
  }
  switch (nbn * 8 + epi) {
    case 16 * 8 + 0: return esr::conv3x3_rows_kernel<16, false, 0>;
    case 16 * 8 + 3: return esr::conv3x3_rows_kernel<16, false, 3>;
    case 32 * 8 + 0: return esr::conv3x3_rows_kernel<32, false, 0>;
    case 32 * 8 + 1: return esr::conv3x3_rows_kernel<32, false, 1>;
    case 32 * 8 + 2: return esr::conv3x3_rows_kernel<32, false, 2>;
    case 32 * 8 + 4: return esr::conv3x3_rows_kernel<32, false, 4>;
    case 32 * 8 + 5: return esr::conv3x3_rows_kernel<32, false, 5>;
    case 64 * 8 + 4: return esr::conv3x3_rows_kernel<64, false, 4>;
    case 64 * 8 + 5: return esr::conv3x3_rows_kernel<64, false, 5>;
    case 64 * 8 + 0: return esr::conv3x3_rows_kernel<64, false, 0>;
    case 64 * 8 + 1: return esr::conv3x3_rows_kernel<64, false, 1>;
    case 64 * 8 + 2: return esr::conv3x3_rows_kernel<64, false, 2>;
  }
  return nullptr;
}

int set_max_smem_once(const void* kern) {
  static std::mutex mu;
  static std::set<const void*> done;
  std::lock_guard<std::mutex> lk(mu);
  if (!done.count(kern)) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemMax);
    if (e != cudaSuccess) return fail(ESR_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    done.insert(kern);
  }
  return ESR_OK;
}

int launch_conv(const void* kern, int grid, uint32_t smem_bytes, void* stream, void** kargs) {
  static const bool use_pdl = [] { const char* e = getenv("ESR_PDL"); return !(e && e[0] == '0'); }();
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(esr::kConvThreads);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = use_pdl ? 1 : 0;
  CUDA_TRY(cudaLaunchKernelExC(&cfg, kern, kargs));
  g_launches++;
  CUDA_TRY(cudaGetLastError());
  return ESR_OK;
}

struct WgradPlan {
  int gyp, nbn, cpb, n_blocks, mt, rb, gstages;
  uint32_t slot_bytes, gstage_bytes, ring_bytes, smem_bytes;
};
// tiling of the weight-gradient kernel for (cin_planes, cout); returns false when it does not fit
bool wgrad_plan(int cp, int cout, WgradPlan* w) {
  if (cp < 1 || cp > 32 || cout < 1) return false;
  w->gyp = (cout + 7) / 8;
  w->mt = (3 * cp + 15) / 16;
  if (cout <= 8) w->nbn = 16;
  else if (cout >= 64 && w->mt * 192 <= 512) w->nbn = 64;
  else w->nbn = 32;
  if (w->mt * 3 * w->nbn > 512) return false;
  w->cpb = w->nbn / 8;
  w->n_blocks = (w->gyp + w->cpb - 1) / w->cpb;
  const uint32_t group = esr::kWgBox * 16u;
  w->slot_bytes = (uint32_t)cp * group;
  w->gstage_bytes = 3u * w->cpb * group;
  // activation ring as deep as fits next to three gradient stages, then as many more gradient stages as fit (v1 always uses three)
  int rb = esr::kWgMaxRing;
  for (; rb >= 4; --rb) {
    w->ring_bytes = (uint32_t)(rb + 2) * w->slot_bytes + 16u * group;   // + slack for the last M chunk
    w->smem_bytes = esr::kSmemHeader + 128u + esr::kWgNB * (w->ring_bytes + esr::kWgGStages * w->gstage_bytes);
    if (w->smem_bytes <= kSmemMax) break;
  }
  if (rb < 4) return false;
  w->gstages = esr::kWgGStages;
  const int max_stages = w->cpb > 4 ? 4 : 8;      // at most one outstanding stage grant per loader warp: stages x plane groups <= 8
  while (w->gstages < max_stages && w->smem_bytes + esr::kWgNB * w->gstage_bytes <= kSmemMax) {
    w->gstages++;
    w->smem_bytes += esr::kWgNB * w->gstage_bytes;
  }
  w->rb = rb;
  return true;
}

// The row-streaming kernel is used when its 128-pixel strips are mostly real pixels, and for forward-type launches on any image
// at least 32 pixels wide: they win 1.5-2x per launch even at 41 % lane fill (resident weights, no per-tile pipeline refill;
// tools/small_conv_time.py), while the generic backward epilogue with many n-blocks does not.  ESR_ROWS=0 forces the tile kernel, ESR_ROWS=2 the row kernel.
bool rows_shape_ok(int w, bool bwd) {
  static const int mode = [] { const char* e = getenv("ESR_ROWS"); return e ? atoi(e) : 1; }();   // 0 never, 1 by shape, 2 always
  const int strips = (w + 127) / 128;
  if (mode != 1) return mode == 2;
  return w * 100 >= strips * 128 * 80 || (!bwd && w >= 32);
}

}  // namespace

extern "C" {

const char* esr_last_error(void) { return g_err; }
int esr_version(void) { return 100; }
long long esr_launch_count(void) { return g_launches.load(); }

int esr_debug_watchdog(unsigned int* out8_host, int reset) {
  CUDA_TRY(cudaMemcpyFromSymbol(out8_host, esr::g_watchdog, 8 * sizeof(unsigned int)));
  if (reset) {
    unsigned int z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    CUDA_TRY(cudaMemcpyToSymbol(esr::g_watchdog, z, sizeof(z)));
  }
  return ESR_OK;
}

int esr_set_deterministic(int on) {
  g_deterministic.store(on ? 1 : 0);
  return ESR_OK;
}

int esr_device_check(void) {
  int dev = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  int major = 0, minor = 0;
  CUDA_TRY(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  CUDA_TRY(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  if (major != 10) return fail(ESR_ERR_UNSUPPORTED, "device is sm_%d%d; this library is built for sm_100a only", major, minor);
  if (!get_encode()) return fail(ESR_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  return ESR_OK;
}

size_t esr_conv3x3_packed_bytes_ex(int cin_planes, int cout, int kcp, int dtype, int* cout_pad_out) {
  int cout_pad = 0;
  nblock_for(cout, &cout_pad);
  if (cout_pad_out) *cout_pad_out = cout_pad;
  const int nchunks = (dtype == ESR_BF16X3 ? 3 : 1) * ((cin_planes + kcp - 1) / kcp);
  return (size_t)nchunks * 9 * kcp * cout_pad * 16;
}
size_t esr_conv3x3_packed_bytes(int cin_planes, int cout, int kcp, int* cout_pad_out) {
  return esr_conv3x3_packed_bytes_ex(cin_planes, cout, kcp, ESR_F16, cout_pad_out);
}

int esr_conv3x3_cin_planes(int cin, int lead) { return (lead + 7) / 8 + (cin - lead + 7) / 8; }

int esr_pack_conv3x3_weights(const float* w_oihw, int cout, int cin, int lead, int kcp, int dtype, int transpose_flip,
                             void* wpacked, float* bias_out, const float* bias_in, void* stream) {
  if (!w_oihw || !wpacked) return fail(ESR_ERR_INVALID, "pack_weights: null pointer");
  if (lead < 0 || lead > cin) return fail(ESR_ERR_INVALID, "pack_weights: bad lead %d", lead);
  if (kcp != 2 && kcp != 4) return fail(ESR_ERR_INVALID, "pack_weights: kcp must be 2 or 4");
  const int lc_out = transpose_flip ? cin : cout;
  const int lc_in = transpose_flip ? cout : cin;
  int cout_pad = 0;
  // dgrad: the conv's outputs are the forward inputs, laid out in plane space ([lead pad | rest])
  const int lc_out_planespace = transpose_flip ? esr_conv3x3_cin_planes(cin, lead) * 8 : lc_out;
  const int nb_n = nblock_for(lc_out_planespace, &cout_pad);
  const int n_blocks = cout_pad / nb_n;
  const int cin_planes = transpose_flip ? (lc_in + 7) / 8 : esr_conv3x3_cin_planes(cin, lead);
  const int nchunks = (dtype == ESR_BF16X3 ? 3 : 1) * ((cin_planes + kcp - 1) / kcp);
  const size_t total = (size_t)n_blocks * nchunks * 9 * kcp * nb_n * 8;
  cudaStream_t st = (cudaStream_t)stream;
  esr::pack_weights_kernel<<<grid_for(total, 256), 256, 0, st>>>(w_oihw, cout, cin, lead, kcp, nb_n, n_blocks, nchunks, dtype,
                                                                transpose_flip, (uint16_t*)wpacked, total);
  g_launches++;
  CUDA_TRY(cudaGetLastError());
  if (bias_out) {
    CUDA_TRY(cudaMemsetAsync(bias_out, 0, sizeof(float) * cout_pad, st));
    if (bias_in && !transpose_flip)
      CUDA_TRY(cudaMemcpyAsync(bias_out, bias_in, sizeof(float) * cout, cudaMemcpyDeviceToDevice, st));
  }
  return ESR_OK;
}

int esr_conv3x3_fwd(const esr_conv3x3_args* a, void* stream) {
  if (!a) return fail(ESR_ERR_INVALID, "conv3x3: null args");
  if (!a->in || !a->wpacked || !a->bias) return fail(ESR_ERR_INVALID, "conv3x3: null in/weights/bias");
  if (a->n <= 0 || a->h <= 0 || a->w <= 0) return fail(ESR_ERR_INVALID, "conv3x3: bad shape %dx%dx%d", a->n, a->h, a->w);
  if (a->kcp != 2 && a->kcp != 4) return fail(ESR_ERR_INVALID, "conv3x3: kcp must be 2 or 4");
  if (a->dtype != ESR_F16 && a->dtype != ESR_BF16 && a->dtype != ESR_BF16X3) return fail(ESR_ERR_INVALID, "conv3x3: bad dtype");
  const int split = a->dtype == ESR_BF16X3 ? 1 : 0;
  const int edt = split ? (int)ESR_BF16 : a->dtype;     // element format of the 16-bit operands
  const int segs = split ? 3 : 1;
  if (split && ((a->in_planes_total & 1) || (a->out16 && (a->out16_planes_total & 1)) || (a->res1 && a->res1_is16 && (a->res1_planes_total & 1))))
    return fail(ESR_ERR_INVALID, "conv3x3: split-precision tensors hold hi and lo halves: planes_total must be even");
  if (!a->out16 && !a->out32 && !a->out_nchw && !a->lead_acc) return fail(ESR_ERR_INVALID, "conv3x3: no output requested");
  if (a->out16_up2 && a->out16_pixel_shuffle) return fail(ESR_ERR_INVALID, "conv3x3: up2 and pixel_shuffle are exclusive");
  if (((uintptr_t)a->in & 15) || ((uintptr_t)a->wpacked & 15) || ((uintptr_t)a->bias & 15))
    return fail(ESR_ERR_INVALID, "conv3x3: pointers must be 16-byte aligned");
  EncodeTiledFn encode = get_encode();
  if (!encode) return fail(ESR_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");

  esr::ConvParams p;
  memset(&p, 0, sizeof(p));
  int cout_pad = 0;
  p.nb_n = nblock_for(a->cout, &cout_pad);
  if (a->cout_pad != cout_pad) return fail(ESR_ERR_INVALID, "conv3x3: cout_pad %d does not match packing (%d)", a->cout_pad, cout_pad);
  p.n_blocks = cout_pad / p.nb_n;
  p.n = a->n; p.h = a->h; p.w = a->w;
  p.P = a->tile_p ? a->tile_p : 32;
  if (p.P != 32) return fail(ESR_ERR_INVALID, "conv3x3: tile_p must be 32 (the 8*P-element TMA inner box is limited to 256 elements)");
  p.TW = p.P - 2;
  p.MT = a->tile_mt ? a->tile_mt : 4;
  if (p.MT != 1 && p.MT != 2 && p.MT != 4) return fail(ESR_ERR_INVALID, "conv3x3: tile_mt must be 1, 2 or 4");
  // do not use taller tiles than the image needs
  while (p.MT > 1 && (p.MT / 2) * 128 / p.P >= a->h) p.MT /= 2;
  p.R = p.MT * 128 / p.P;
  p.tiles_x = (a->w + p.TW - 1) / p.TW;
  p.tiles_y = (a->h + p.R - 1) / p.R;
  p.num_tiles = p.tiles_x * p.tiles_y * a->n;
  p.kcp = a->kcp;
  p.split = split;
  p.cps = (a->cin_planes + a->kcp - 1) / a->kcp;
  p.nchunks = segs * p.cps;
  p.in_plane_off = a->in_plane_off;
  p.seg_base[0] = p.seg_base[2] = a->in_plane_off;
  p.seg_base[1] = a->in_plane_off + a->in_planes_total / 2;
  {
    const size_t f = a->out16_up2 ? 2 : (a->out16_pixel_shuffle ? a->out16_pixel_shuffle : 1);
    p.out16_lo = (size_t)(a->out16_planes_total / 2) * f * f * a->h * a->w * 8;
    p.res1_lo = (size_t)(a->res1_planes_total / 2) * a->h * a->w * 8;
  }
  p.plane_stride = (uint32_t)(p.R + 2) * p.P * 16u;
  p.a_bytes = p.kcp * p.plane_stride;
  p.b_bytes = 9u * p.kcp * p.nb_n * 16u;
  p.a_alloc = (p.a_bytes + esr::kASlack + 127u) & ~127u;
  const uint32_t smem_max = 232448u;
  // weights stay resident in shared memory when all K chunks fit next to >= 3 activation stages
  p.w_bytes = (uint32_t)p.nchunks * p.b_bytes;
  p.w_resident = (p.n_blocks == 1 && esr::kSmemHeader + 128u + p.w_bytes + 3u * p.a_alloc <= smem_max) ? 1 : 0;
  p.stage_bytes = p.w_resident ? p.a_alloc : p.a_alloc + ((p.b_bytes + 127u) & ~127u);
  const uint32_t fixed = esr::kSmemHeader + 128u + (p.w_resident ? p.w_bytes : 0u);
  int stages = (int)((smem_max - fixed) / p.stage_bytes);
  if (stages > esr::kMaxStages) stages = esr::kMaxStages;
  if (stages > p.nchunks * 4) stages = p.nchunks * 4;  // no point in more buffers than a few items' worth
  if (stages < 2) return fail(ESR_ERR_INVALID, "conv3x3: stage of %u bytes does not fit twice in shared memory", p.stage_bytes);
  p.stages = stages;
  const uint32_t smem_bytes = fixed + (uint32_t)stages * p.stage_bytes;
  // instruction descriptor: D=f32, A/B = f16|bf16, K-major both, N, M=128
  p.idesc = (1u << 4) | ((uint32_t)edt << 7) | ((uint32_t)edt << 10) | ((uint32_t)(p.nb_n >> 3) << 17) |
            ((uint32_t)(128 >> 4) << 24);
  uint32_t cols = 2u * p.MT * p.nb_n, alloc = 32;
  while (alloc < cols) alloc <<= 1;
  if (alloc > 512) return fail(ESR_ERR_INVALID, "conv3x3: TMEM budget exceeded (%u columns)", cols);
  p.tmem_cols = alloc;
  p.wts = (const uint8_t*)a->wpacked;
  p.bias = a->bias;
  p.cout = a->cout;
  p.dtype = edt;
  p.lrelu = a->lrelu; p.slope = a->slope; p.alpha = a->alpha;
  p.res1 = a->res1; p.res1_is16 = a->res1_is16; p.res1_pt = a->res1_planes_total; p.res1_po = a->res1_plane_off; p.beta1 = a->beta1;
  p.res2 = a->res2; p.res2_pt = a->res2_planes_total; p.res2_po = a->res2_plane_off; p.beta2 = a->beta2;
  p.res3 = a->res3; p.res3_pt = a->res3_planes_total; p.res3_po = a->res3_plane_off; p.beta3 = a->beta3;
  p.lead_planes = a->lead_planes; p.lead_acc = a->lead_acc; p.lead_pt = a->lead_planes_total;
  if (p.lead_planes && !p.lead_acc) return fail(ESR_ERR_INVALID, "conv3x3: lead_planes without lead_acc");
  p.mask16 = (const uint16_t*)a->mask16; p.mask_pt = a->mask_planes_total; p.mask_po = a->mask_plane_off; p.mask_slope = a->mask_slope;
  p.tail_first = a->tail_first_plane;
  p.out16 = (uint16_t*)a->out16; p.out16_pt = a->out16_planes_total; p.out16_po = a->out16_plane_off;
  p.out16_up2 = a->out16_up2; p.out16_ps = a->out16_pixel_shuffle;
  p.out32 = a->out32; p.out32_pt = a->out32_planes_total; p.out32_po = a->out32_plane_off;
  p.out_nchw = a->out_nchw; p.out_nchw_c = a->out_nchw_c;

  const bool bwd = p.lead_planes > 0 || p.mask16 != nullptr || p.res3 != nullptr || p.tail_first > 0;
  // (the gradient-slice launches of the dense-block backward have the forward-type fast epilogue: same width rule as forward launches)
  const bool mask_only_shape = bwd && a->rows_nbn >= 32 && a->cout % 32 == 0 && a->mask16 && a->out16 && !a->lead_planes && !a->res1 && !a->res2 &&
                               !a->res3 && !a->out32 && !a->out_nchw && !a->out16_up2 && !a->out16_pixel_shuffle && a->tail_first_plane == 0 &&
                               a->alpha == 1.0f && !a->lrelu;
  const bool closing_shape = bwd && a->rows_nbn >= 32 && a->cout % 32 == 0 && !a->mask16 && !a->lead_planes && a->res3 && !a->res2 && a->out32 && a->out16 &&
                             !a->out_nchw && !a->out16_up2 && !a->out16_pixel_shuffle && a->tail_first_plane == 0 && a->alpha == 1.0f && !a->lrelu &&
                             (!a->res1 || !a->res1_is16);
  if (a->wpacked_rows && a->rows_nbn > 0 && a->rows_mode >= 0 &&
      (a->rows_mode > 0 || rows_shape_ok(a->w, bwd && !mask_only_shape && !closing_shape))) {
    // ---- row-streaming kernel (conv3x3_rows.cuh)
    const int nbn = a->rows_nbn;
    if (nbn != 16 && nbn != 32 && nbn != 64) return fail(ESR_ERR_INVALID, "conv3x3: rows_nbn must be 16, 32 or 64");
    if ((uintptr_t)a->wpacked_rows & 15) return fail(ESR_ERR_INVALID, "conv3x3: wpacked_rows must be 16-byte aligned");
    int epi = 0;
    if (!bwd && nbn >= 32 && a->cout % 32 == 0 && p.out16 && !p.out_nchw && !p.out16_ps && !p.res3) {
      if ((!p.res1 || !p.res1_is16) && !p.res2 && !p.out32 && p.alpha == 1.0f) epi = 1;
      else if (p.res1 && p.res1_is16 && !p.lrelu && !p.out16_up2) epi = 2;
    }
    if (!bwd && nbn == 16 && p.out_nchw && p.out_nchw_c <= 8 && !p.out16 && !p.out32 && !p.res1 && !p.res2 && !p.res3) epi = 3;
    // gradient slice of the dense-block backward: mask * acc -> 16-bit planes, nothing else
    const bool mask_only = bwd && nbn >= 32 && a->cout % 32 == 0 && p.mask16 && p.out16 && !p.lead_planes && !p.res1 && !p.res2 && !p.res3 &&
                           !p.out32 && !p.out_nchw && !p.out16_up2 && !p.out16_ps && p.tail_first == 0 && p.alpha == 1.0f && !p.lrelu;
    // closing launch of the dense-block backward: acc + beta3*res3 (+ beta1*res1 fp32) -> fp32 planes and 16-bit planes
    const bool closing = bwd && nbn >= 32 && a->cout % 32 == 0 && !p.mask16 && !p.lead_planes && p.res3 && !p.res2 && p.out32 && p.out16 && !p.out_nchw &&
                         !p.out16_up2 && !p.out16_ps && p.tail_first == 0 && p.alpha == 1.0f && !p.lrelu && (!p.res1 || !p.res1_is16);
    RowsKernelFn rk = mask_only ? select_rows_kernel(nbn, false, 4) : (closing ? select_rows_kernel(nbn, false, 5) : select_rows_kernel(nbn, bwd, epi));
    if (!rk) return fail(ESR_ERR_INVALID, "conv3x3: no row kernel for N block %d", nbn);
    p.nb_n = nbn;
    p.n_blocks = (a->cout + nbn - 1) / nbn;
    p.cps = (a->cin_planes + esr::kRowsKch - 1) / esr::kRowsKch;
    p.nchunks = segs * p.cps;
    p.cin_planes = a->cin_planes;
    p.w_bytes = (uint32_t)p.nchunks * 3u * esr::kRowsKch * 3u * nbn * 16u;
    const uint32_t fixed = esr::kSmemHeader + 128u + p.w_bytes;
    int stages = (int)((kSmemMax - fixed) / esr::kRowsStageBytes);
    if (stages > esr::kRowsMaxStages) stages = esr::kRowsMaxStages;
    if (stages < 4) return fail(ESR_ERR_INVALID, "conv3x3: row kernel weights (%u bytes) leave no room for the pipeline", p.w_bytes);
    p.stages = stages;
    p.slots = 512 / nbn < esr::kRowsMaxSlots ? 512 / nbn : esr::kRowsMaxSlots;
    {
      static const int issuers = [] { const char* e = getenv("ESR_ISSUERS"); int v = e ? atoi(e) : 3; return v < 1 ? 1 : (v > 3 ? 3 : v); }();
      p.issuers = deterministic() ? 1 : issuers;
    }
    p.strips = (a->w + 127) / 128;
    p.units = (long long)a->n * p.strips * a->h;
    int ranges = num_sms() / p.n_blocks;
    if (ranges < 1) return fail(ESR_ERR_INVALID, "conv3x3: too many n-blocks (%d) for the row kernel", p.n_blocks);
    if ((long long)ranges > p.units) ranges = (int)p.units;
    p.ranges = ranges;
    p.in = (const uint8_t*)a->in;
    p.in_pt = a->in_planes_total;
    p.wts = (const uint8_t*)a->wpacked_rows;
    for (int k = 0; k < 3; ++k)
      p.idesc_n[k] = (1u << 4) | ((uint32_t)edt << 7) | ((uint32_t)edt << 10) | ((uint32_t)(((k + 1) * nbn) >> 3) << 17) |
                     ((uint32_t)(128 >> 4) << 24);
    int rc = set_max_smem_once((const void*)rk);
    if (rc) return rc;
    void* kargs[1] = {(void*)&p};
    return launch_conv((const void*)rk, ranges * p.n_blocks, fixed + (uint32_t)stages * esr::kRowsStageBytes, stream, kargs);
  }

  // TMA descriptor over the input planes: dims (8ch*W, H, planes, N), box (8*P, R+2, kcp, 1).  A pixel row of a
  // plane is one contiguous run, so (channel-in-plane, x) is a single 512-byte inner box dimension.
  CUtensorMap tm;
  cuuint64_t gdim[4] = {(cuuint64_t)a->w * 8, (cuuint64_t)a->h, (cuuint64_t)a->in_planes_total, (cuuint64_t)a->n};
  cuuint64_t gstr[3] = {(cuuint64_t)a->w * 16, (cuuint64_t)a->w * a->h * 16, (cuuint64_t)a->w * a->h * 16 * a->in_planes_total};
  cuuint32_t box[4] = {(cuuint32_t)p.P * 8, (cuuint32_t)(p.R + 2), (cuuint32_t)p.kcp, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult cr = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT16, 4, const_cast<void*>(a->in), gdim, gstr, box, estr,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) return fail(ESR_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)cr);

  ConvKernelFn kern = select_conv_kernel(p.P, p.kcp, p.nb_n, bwd);
  if (!kern) return fail(ESR_ERR_INVALID, "conv3x3: no kernel instance for P=%d kcp=%d N=%d", p.P, p.kcp, p.nb_n);
  int rc = set_max_smem_once((const void*)kern);
  if (rc) return rc;
  int grid = p.num_tiles * p.n_blocks;
  if (grid > num_sms()) grid = num_sms();
  void* kargs[2] = {(void*)&tm, (void*)&p};
  return launch_conv((const void*)kern, grid, smem_bytes, stream, kargs);
}

int esr_conv3x3_fwd_batch(const esr_conv3x3_args* args, int count, void* stream, int* failed_index) {
  if (!args || count < 0) return fail(ESR_ERR_INVALID, "conv3x3 batch: bad arguments");
  for (int i = 0; i < count; ++i) {
    const int rc = esr_conv3x3_fwd(args + i, stream);
    if (rc != ESR_OK) {
      if (failed_index) *failed_index = i;
      return rc;
    }
  }
  return ESR_OK;
}

int esr_conv3x3_rows_config_ex(int cin_planes, int cout, int dtype, int* nbn_out, size_t* bytes_out) {
  const int segs = dtype == ESR_BF16X3 ? 3 : 1;
  const int nbn = rows_nbn_for(cin_planes, cout, segs);
  if (nbn_out) *nbn_out = nbn;
  if (!nbn) return fail(ESR_ERR_INVALID, "conv3x3 rows: weights of (cin_planes %d, cout %d) do not fit in shared memory", cin_planes, cout);
  const int nchunks = segs * ((cin_planes + esr::kRowsKch - 1) / esr::kRowsKch);
  const int n_blocks = (cout + nbn - 1) / nbn;
  if (bytes_out) *bytes_out = (size_t)n_blocks * nchunks * 3 * esr::kRowsKch * 3 * nbn * 16;
  return ESR_OK;
}
int esr_conv3x3_rows_config(int cin_planes, int cout, int* nbn_out, size_t* bytes_out) {
  return esr_conv3x3_rows_config_ex(cin_planes, cout, ESR_F16, nbn_out, bytes_out);
}

int esr_pack_conv3x3_weights_rows(const float* w_oihw, int cout, int cin, int lead, int dtype, int transpose_flip, int nbn,
                                  void* wpacked_rows, void* stream) {
  if (!w_oihw || !wpacked_rows) return fail(ESR_ERR_INVALID, "pack_weights_rows: null pointer");
  if (lead < 0 || lead > cin) return fail(ESR_ERR_INVALID, "pack_weights_rows: bad lead %d", lead);
  if (nbn != 16 && nbn != 32 && nbn != 64) return fail(ESR_ERR_INVALID, "pack_weights_rows: n-block must be 16, 32 or 64");
  const int lc_out = transpose_flip ? esr_conv3x3_cin_planes(cin, lead) * 8 : cout;
  const int lc_in = transpose_flip ? cout : cin;
  const int cin_planes = transpose_flip ? (lc_in + 7) / 8 : esr_conv3x3_cin_planes(cin, lead);
  const int nchunks = (dtype == ESR_BF16X3 ? 3 : 1) * ((cin_planes + esr::kRowsKch - 1) / esr::kRowsKch);
  const int n_blocks = (lc_out + nbn - 1) / nbn;
  const size_t total = (size_t)n_blocks * nchunks * 3 * esr::kRowsKch * 3 * nbn * 8;
  esr::pack_weights_rows_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(w_oihw, cout, cin, lead, nbn, nchunks, dtype,
                                                                                      transpose_flip, (uint16_t*)wpacked_rows, total);
  g_launches++;
  CUDA_TRY(cudaGetLastError());
  return ESR_OK;
}

size_t esr_pack_batch_scratch_bytes(int count) { return count > 0 ? (size_t)count * sizeof(esr::PackJob) : 0; }

int esr_pack_conv3x3_weights_batch(const esr_pack_item* items, int count, void* scratch, size_t scratch_bytes, void* stream,
                                   int* failed_index) {
  if (!items || count < 0) return fail(ESR_ERR_INVALID, "pack batch: bad arguments");
  if (count == 0) return ESR_OK;
  static const bool fused = [] { const char* e = getenv("ESR_PACK_FUSED"); return !(e && e[0] == '0'); }();
  if (!fused || !scratch || scratch_bytes < esr_pack_batch_scratch_bytes(count)) {
    // one launch pair per conv (the path single convs use)
    for (int i = 0; i < count; ++i) {
      const esr_pack_item* it = items + i;
      int rc = esr_pack_conv3x3_weights(it->w_oihw, it->cout, it->cin, it->lead, it->kcp, it->dtype, it->transpose_flip, it->wpacked,
                                        it->bias_out, it->bias_in, stream);
      if (rc == ESR_OK && it->wpacked_rows)
        rc = esr_pack_conv3x3_weights_rows(it->w_oihw, it->cout, it->cin, it->lead, it->dtype, it->transpose_flip, it->rows_nbn,
                                           it->wpacked_rows, stream);
      if (rc != ESR_OK) {
        if (failed_index) *failed_index = i;
        return rc;
      }
    }
    return ESR_OK;
  }
  // one launch for all of them: job table -> device scratch, blockIdx.y = conv
  std::vector<esr::PackJob> jobs((size_t)count);
  for (int i = 0; i < count; ++i) {
    const esr_pack_item* it = items + i;
    if (failed_index) *failed_index = i;
    if (!it->w_oihw || !it->wpacked) return fail(ESR_ERR_INVALID, "pack batch: null pointer in item %d", i);
    if (((uintptr_t)it->wpacked & 15) || ((uintptr_t)it->wpacked_rows & 15))
      return fail(ESR_ERR_INVALID, "pack batch: packed images must be 16-byte aligned (item %d)", i);
    if (it->lead < 0 || it->lead > it->cin) return fail(ESR_ERR_INVALID, "pack batch: bad lead %d in item %d", it->lead, i);
    if (it->kcp != 2 && it->kcp != 4) return fail(ESR_ERR_INVALID, "pack batch: kcp must be 2 or 4 (item %d)", i);
    esr::PackJob& jb = jobs[(size_t)i];
    memset(&jb, 0, sizeof(jb));
    // same derivation as esr_pack_conv3x3_weights / esr_pack_conv3x3_weights_rows
    const int lc_out = it->transpose_flip ? it->cin : it->cout;
    const int lc_in = it->transpose_flip ? it->cout : it->cin;
    const int lc_out_planespace = it->transpose_flip ? esr_conv3x3_cin_planes(it->cin, it->lead) * 8 : lc_out;
    int cout_pad = 0;
    const int nb_n = nblock_for(lc_out_planespace, &cout_pad);
    const int n_blocks = cout_pad / nb_n;
    const int cin_planes = it->transpose_flip ? (lc_in + 7) / 8 : esr_conv3x3_cin_planes(it->cin, it->lead);
    const int segs = it->dtype == ESR_BF16X3 ? 3 : 1;
    const int nchunks = segs * ((cin_planes + it->kcp - 1) / it->kcp);
    jb.w = it->w_oihw; jb.cout = it->cout; jb.cin = it->cin; jb.lead = it->lead; jb.kcp = it->kcp; jb.nb_n = nb_n; jb.nchunks = nchunks;
    jb.dtype = it->dtype; jb.transpose_flip = it->transpose_flip;
    jb.dst = (uint16_t*)it->wpacked;
    jb.total = (unsigned long long)n_blocks * nchunks * 9 * it->kcp * nb_n * 8;
    if (it->wpacked_rows) {
      const int nbn = it->rows_nbn;
      if (nbn != 16 && nbn != 32 && nbn != 64) return fail(ESR_ERR_INVALID, "pack batch: row n-block must be 16, 32 or 64 (item %d)", i);
      const int rows_out = it->transpose_flip ? esr_conv3x3_cin_planes(it->cin, it->lead) * 8 : it->cout;
      const int rows_chunks = segs * ((cin_planes + esr::kRowsKch - 1) / esr::kRowsKch);
      const int rows_blocks = (rows_out + nbn - 1) / nbn;
      jb.rows_nb_n = nbn; jb.rows_nchunks = rows_chunks; jb.rows_dst = (uint16_t*)it->wpacked_rows;
      jb.rows_total = (unsigned long long)rows_blocks * rows_chunks * 3 * esr::kRowsKch * 3 * nbn * 8;
    }
    if (it->bias_out) {
      jb.bias_out = it->bias_out; jb.cout_pad = cout_pad;
      jb.bias_in = it->transpose_flip ? nullptr : it->bias_in;
      jb.bias_n = it->cout;
    }
  }
  if (failed_index) *failed_index = -1;
  cudaStream_t st = (cudaStream_t)stream;
  CUDA_TRY(cudaMemcpyAsync(scratch, jobs.data(), sizeof(esr::PackJob) * (size_t)count, cudaMemcpyHostToDevice, st));
  for (int i0 = 0; i0 < count; i0 += 65535) {     // gridDim.y limit
    const int n = count - i0 < 65535 ? count - i0 : 65535;
    esr::pack_weights_batch_kernel<<<dim3(4, (unsigned)n), 256, 0, st>>>((const esr::PackJob*)scratch + i0);
    g_launches++;
    CUDA_TRY(cudaGetLastError());
  }
  return ESR_OK;
}

size_t esr_conv3x3_wgrad_workspace(int cin_planes, int cout) {
  WgradPlan w;
  if (!wgrad_plan(cin_planes, cout, &w)) return 0;
  return (size_t)num_sms() * (w.mt * 128 * 3 * w.nbn + esr::kWgBiasStride) * sizeof(float);
}

int esr_conv3x3_wgrad(const esr_conv3x3_wgrad_args* a, void* stream) {
  if (!a || !a->x || !a->gy || !a->workspace) return fail(ESR_ERR_INVALID, "wgrad: null pointer");
  if (a->dtype == ESR_BF16X3) {
    // split precision: dW = x_hi (*) gy_hi + x_lo (*) gy_hi + x_hi (*) gy_lo, three accumulating launches of the bf16 kernel
    if ((a->x_planes_total & 1) || (a->gy_planes_total & 1)) return fail(ESR_ERR_INVALID, "wgrad: split tensors need an even planes_total");
    esr_conv3x3_wgrad_args b = *a;
    b.dtype = ESR_BF16;
    int rc = esr_conv3x3_wgrad(&b, stream);
    if (rc) return rc;
    b.accumulate = 1;
    if (a->dw) {
      b.db = nullptr;
      b.x_plane_off = a->x_plane_off + a->x_planes_total / 2;
      rc = esr_conv3x3_wgrad(&b, stream);
      if (rc) return rc;
    }
    b.db = a->db;
    b.x_plane_off = a->x_plane_off;
    b.gy_plane_off = a->gy_plane_off + a->gy_planes_total / 2;
    return esr_conv3x3_wgrad(&b, stream);
  }
  if (!a->dw && !a->db) return fail(ESR_ERR_INVALID, "wgrad: no output requested");
  if (a->n <= 0 || a->h <= 0 || a->w <= 0) return fail(ESR_ERR_INVALID, "wgrad: bad shape %dx%dx%d", a->n, a->h, a->w);
  if (a->dtype != ESR_F16 && a->dtype != ESR_BF16) return fail(ESR_ERR_INVALID, "wgrad: bad dtype");
  if (((uintptr_t)a->x & 15) || ((uintptr_t)a->gy & 15)) return fail(ESR_ERR_INVALID, "wgrad: pointers must be 16-byte aligned");
  if (a->lead < 0 || a->lead > a->cin) return fail(ESR_ERR_INVALID, "wgrad: bad lead %d", a->lead);
  if (a->cin_total > 0 && (a->cin_off < 0 || a->cin_off + a->cin > a->cin_total)) return fail(ESR_ERR_INVALID, "wgrad: bad input-channel slice");
  const int cp = esr_conv3x3_cin_planes(a->cin, a->lead);
  if (a->x_plane_off + cp > a->x_planes_total) return fail(ESR_ERR_INVALID, "wgrad: input planes out of range");
  WgradPlan w;
  if (!wgrad_plan(cp, a->cout, &w)) return fail(ESR_ERR_INVALID, "wgrad: (cin %d, cout %d) does not fit the tensor-core tiling", a->cin, a->cout);
  if (a->gy_plane_off + w.gyp > a->gy_planes_total) return fail(ESR_ERR_INVALID, "wgrad: gradient planes out of range");
  cudaStream_t st = (cudaStream_t)stream;
  bool fused_bias = false;
  if (a->dw) {
    esr::WgradParams p;
    memset(&p, 0, sizeof(p));
    p.n = a->n; p.h = a->h; p.w = a->w;
    p.strips = (a->w + esr::kWgPW - 1) / esr::kWgPW;
    p.units = (long long)a->n * p.strips * a->h;
    if (p.units > 0x3fffffffLL) return fail(ESR_ERR_INVALID, "wgrad: too many row units (%lld)", p.units);
    p.n_blocks = w.n_blocks;
    int ranges = num_sms() / w.n_blocks;
    if (ranges < 1) return fail(ESR_ERR_INVALID, "wgrad: too many n-blocks (%d)", w.n_blocks);
    if ((long long)ranges > p.units) ranges = (int)p.units;
    // small images: every CTA dumps a full partial (mt*128*3*nbn floats) whatever its share of rows, and the reduce reads all of
    // them (profiles/r01c_launches_c3_step.csv: at ~1.4 rows per CTA the reduce cost as much as the wgrad itself).  Every CTA gets
    // at least ESR_WGRAD_MIN_ROWS row units (default 6: a row unit is 0.6-1.1 us of MMAs, a partial 60-250 KB).
    static const int min_rows = [] { const char* e = getenv("ESR_WGRAD_MIN_ROWS"); const int v = e ? atoi(e) : 6; return v < 1 ? 1 : v; }();
    if (min_rows > 1 && p.units / min_rows < ranges) ranges = (int)(p.units / min_rows > 0 ? p.units / min_rows : 1);
    p.ranges = ranges;
    p.x = (const uint8_t*)a->x; p.x_pt = a->x_planes_total; p.x_po = a->x_plane_off; p.cp = cp;
    p.gy = (const uint8_t*)a->gy; p.gy_pt = a->gy_planes_total; p.gy_po = a->gy_plane_off; p.gyp = w.gyp;
    p.nbn = w.nbn; p.cpb = w.cpb; p.mt = w.mt; p.rb = w.rb;
    p.slot_bytes = w.slot_bytes; p.gstage_bytes = w.gstage_bytes; p.ring_bytes = w.ring_bytes;
    // D=f32, A/B = f16|bf16, both MN-major (bits 15, 16), N = 3*nbn, M = 128
    p.issuers = deterministic() ? 1 : esr::kWgIssuers;
    p.gstages = w.gstages; p.ngroups = w.cpb > 4 ? 2 : 1;
    p.idesc = (1u << 4) | ((uint32_t)a->dtype << 7) | ((uint32_t)a->dtype << 10) | (1u << 15) | (1u << 16) |
              ((uint32_t)((3 * w.nbn) >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const int grid = ranges * w.n_blocks;
    const size_t part_floats = (size_t)grid * w.mt * 128 * 3 * w.nbn;
    const size_t need = (part_floats + (size_t)grid * esr::kWgBiasStride) * sizeof(float);
    if (a->workspace_bytes < need) return fail(ESR_ERR_INVALID, "wgrad: workspace of %zu bytes needed, %zu given", need, a->workspace_bytes);
    p.part = a->workspace;
    p.dtype = a->dtype;
    // v2 (default): the gradient row crosses L2 -> SM once and the bias gradient is summed on the way; ESR_WGRAD_V1=1: three TMA copies
    static const bool v1 = [] { const char* e = getenv("ESR_WGRAD_V1"); return e && atoi(e) != 0; }();
    fused_bias = !v1 && a->db != nullptr;
    p.bias_part = fused_bias ? a->workspace + part_floats : nullptr;
    // tensor maps over the planar-8 buffers: dims (8ch*W, H, planes, N); box = (8*32 elements, 1 row, planes of the operand, 1)
    EncodeTiledFn encode = get_encode();
    if (!encode) return fail(ESR_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    CUtensorMap tmx, tmg;
    {
      cuuint64_t gdim[4] = {(cuuint64_t)a->w * 8, (cuuint64_t)a->h, (cuuint64_t)a->x_planes_total, (cuuint64_t)a->n};
      cuuint64_t gstr[3] = {(cuuint64_t)a->w * 16, (cuuint64_t)a->w * a->h * 16, (cuuint64_t)a->w * a->h * 16 * a->x_planes_total};
      cuuint32_t box[4] = {(cuuint32_t)esr::kWgBox * 8, 1, (cuuint32_t)cp, 1};
      cuuint32_t estr[4] = {1, 1, 1, 1};
      CUresult cr = encode(&tmx, CU_TENSOR_MAP_DATA_TYPE_UINT16, 4, const_cast<void*>(a->x), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (cr != CUDA_SUCCESS) return fail(ESR_ERR_CUDA, "wgrad: cuTensorMapEncodeTiled(x) failed with CUresult %d", (int)cr);
      if (v1) {
        cuuint64_t gdim2[4] = {(cuuint64_t)a->w * 8, (cuuint64_t)a->h, (cuuint64_t)a->gy_planes_total, (cuuint64_t)a->n};
        cuuint64_t gstr2[3] = {(cuuint64_t)a->w * 16, (cuuint64_t)a->w * a->h * 16, (cuuint64_t)a->w * a->h * 16 * a->gy_planes_total};
        cuuint32_t box2[4] = {(cuuint32_t)esr::kWgBox * 8, 1, (cuuint32_t)w.cpb, 1};
        cr = encode(&tmg, CU_TENSOR_MAP_DATA_TYPE_UINT16, 4, const_cast<void*>(a->gy), gdim2, gstr2, box2, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (cr != CUDA_SUCCESS) return fail(ESR_ERR_CUDA, "wgrad: cuTensorMapEncodeTiled(gy) failed with CUresult %d", (int)cr);
      }
    }
    if (v1) {
      int rc = set_max_smem_once((const void*)esr::conv3x3_wgrad_kernel_v1);
      if (rc) return rc;
      esr::conv3x3_wgrad_kernel_v1<<<grid, esr::kWgThreadsV1, w.smem_bytes, st>>>(tmx, tmg, p);
    } else {
      int rc = set_max_smem_once((const void*)esr::conv3x3_wgrad_kernel);
      if (rc) return rc;
      esr::conv3x3_wgrad_kernel<<<grid, esr::kWgThreads, w.smem_bytes, st>>>(tmx, p);
    }
    g_launches++;
    CUDA_TRY(cudaGetLastError());
    const int per_cta = w.mt * 128 * 3 * w.nbn;
    esr::wgrad_reduce_kernel<<<dim3((unsigned)((per_cta + 255) / 256), (unsigned)w.n_blocks), 256, 0, st>>>(
        a->workspace, ranges, w.n_blocks, w.mt, w.nbn, cp, a->cout, a->cin, a->lead, a->scale, a->accumulate, a->dw,
        a->cin_total > 0 ? a->cin_total : a->cin, a->cin_off, p.bias_part, a->db);
    g_launches++;
    CUDA_TRY(cudaGetLastError());
  }
  if (a->db && !fused_bias) {
    if (!a->accumulate) CUDA_TRY(cudaMemsetAsync(a->db, 0, sizeof(float) * a->cout, st));
    const size_t hw = (size_t)a->h * a->w;
    size_t chunks = ((size_t)a->n * hw + 255) / 256;
    if (chunks > 148) chunks = 148;
    if (deterministic()) chunks = 1;       // one block per plane: a single atomicAdd per channel, fixed order
    dim3 grid((unsigned)chunks, (unsigned)w.gyp);
    esr::bias_grad_kernel<<<grid, 256, 0, st>>>((const uint16_t*)a->gy, a->dtype, a->n, a->gy_planes_total, a->gy_plane_off, a->cout, hw,
                                                a->scale, a->db);
    g_launches++;
    CUDA_TRY(cudaGetLastError());
  }
  return ESR_OK;
}

int esr_sum_nchw(const float* src, int n, int c, int h, int w, float scale, int accumulate, float* out, void* stream) {
  if (!src || !out || n <= 0 || c <= 0) return fail(ESR_ERR_INVALID, "sum_nchw: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  if (!accumulate) CUDA_TRY(cudaMemsetAsync(out, 0, sizeof(float) * c, st));
  const size_t hw = (size_t)h * w;
  size_t chunks = (hw + 1023) / 1024;
  if (chunks > 64) chunks = 64;
  if (deterministic())
    esr::sum_nchw_kernel<<<dim3(1, (unsigned)c, 1), 256, 0, st>>>(src, c, hw, scale, out, n);
  else
    esr::sum_nchw_kernel<<<dim3((unsigned)chunks, (unsigned)c, (unsigned)n), 256, 0, st>>>(src, c, hw, scale, out, 0);
  g_launches++;
  CUDA_TRY(cudaGetLastError());
  return ESR_OK;
}

int esr_pack_nchw(const float* src, int n, int c, int h, int w, int pad, int dtype, void* dst16, float* dst32,
                  int planes_total, int plane_off, void* stream) {
  if (!src || (!dst16 && !dst32)) return fail(ESR_ERR_INVALID, "pack_nchw: null pointer");
  const int planes = (c + 7) / 8;
  if (plane_off + planes > planes_total) return fail(ESR_ERR_INVALID, "pack_nchw: planes out of range");
  const size_t total = (size_t)n * planes * (h + 2 * pad) * (w + 2 * pad);
  const bool split = dtype == ESR_BF16X3 && dst16;
  if (split && ((planes_total & 1) || plane_off + planes > planes_total / 2)) return fail(ESR_ERR_INVALID, "pack_nchw: bad split layout");
  const size_t lo16 = split ? (size_t)(planes_total / 2) * (h + 2 * pad) * (w + 2 * pad) * 8 : 0;
  esr::pack_nchw_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(src, n, c, h, w, pad, split ? 1 : dtype, (uint16_t*)dst16,
                                                                              dst32, planes_total, plane_off, planes, lo16);
  g_launches++;
  CUDA_TRY(cudaGetLastError());
  return ESR_OK;
}

int esr_pack_nchw_affine(const float* src, int n, int c, int h, int w, const float* scale, const float* shift, int dtype, void* dst16,
                         int planes_total, int plane_off, void* stream) {
  if (!src || !scale || !shift || !dst16) return fail(ESR_ERR_INVALID, "pack_nchw_affine: null pointer");
  const int planes = (c + 7) / 8;
  if (plane_off + planes > planes_total) return fail(ESR_ERR_INVALID, "pack_nchw_affine: planes out of range");
  const size_t total = (size_t)n * planes * h * w;
  const bool split = dtype == ESR_BF16X3;
  if (split && ((planes_total & 1) || plane_off + planes > planes_total / 2)) return fail(ESR_ERR_INVALID, "pack_nchw_affine: bad split layout");
  esr::pack_nchw_affine_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(src, n, c, h, w, scale, shift, split ? 1 : dtype, (uint16_t*)dst16,
                                                                                     planes_total, plane_off, planes,
                                                                                     split ? (size_t)(planes_total / 2) * h * w * 8 : 0);
  g_launches++;
  CUDA_TRY(cudaGetLastError());
  return ESR_OK;
}

int esr_maxpool2x2_planes16(const void* src, int dtype, int n, int planes, int h, int w, void* dst, void* stream) {
  if (!src || !dst) return fail(ESR_ERR_INVALID, "maxpool2x2: null pointer");
  if ((h & 1) || (w & 1)) return fail(ESR_ERR_INVALID, "maxpool2x2: odd size %dx%d", h, w);
  if (dtype == ESR_BF16X3) {
    if (planes & 1) return fail(ESR_ERR_INVALID, "maxpool2x2: split tensors need an even plane count");
    const size_t tot = (size_t)n * (planes / 2) * (h / 2) * (w / 2);
    esr::maxpool2x2_split_kernel<<<grid_for(tot, 256), 256, 0, (cudaStream_t)stream>>>((const uint16_t*)src, (size_t)n * (planes / 2), planes / 2, h, w,
                                                                                    (uint16_t*)dst);
    g_launches++;
    CUDA_TRY(cudaGetLastError());
    return ESR_OK;
  }
  const size_t total = (size_t)n * planes * (h / 2) * (w / 2);
  esr::maxpool2x2_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>((const uint4*)src, (size_t)n * planes, h, w, dtype, (uint4*)dst);
  g_launches++;
  CUDA_TRY(cudaGetLastError());
  return ESR_OK;
}

int esr_maxpool2x2_bwd_planes16(const void* gout, const void* act, int dtype, int n, int planes, int h, int w, void* gin, void* stream) {
  if (!gout || !act || !gin) return fail(ESR_ERR_INVALID, "maxpool2x2_bwd: null pointer");
  if ((h & 1) || (w & 1)) return fail(ESR_ERR_INVALID, "maxpool2x2_bwd: odd size %dx%d", h, w);
  if (dtype == ESR_BF16X3) {
    if (planes & 1) return fail(ESR_ERR_INVALID, "maxpool2x2_bwd: split tensors need an even plane count");
    const size_t tot = (size_t)n * (planes / 2) * (h / 2) * (w / 2);
    esr::maxpool2x2_bwd_split_kernel<<<grid_for(tot, 256), 256, 0, (cudaStream_t)stream>>>((const uint16_t*)gout, (const uint16_t*)act,
                                                                                        (size_t)n * (planes / 2), planes / 2, h, w, (uint16_t*)gin);
    g_launches++;
    CUDA_TRY(cudaGetLastError());
    return ESR_OK;
  }
  const size_t total = (size_t)n * planes * (h / 2) * (w / 2);
  esr::maxpool2x2_bwd_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>((const uint4*)gout, (const uint4*)act, (size_t)n * planes, h, w,
                                                                                   dtype, (uint4*)gin);
  g_launches++;
  CUDA_TRY(cudaGetLastError());
  return ESR_OK;
}

int esr_unpack_planes16(const void* src16, int dtype, int n, int c, int h, int w, int planes_total, int plane_off,
                        float* dst, void* stream) {
  if (!src16 || !dst) return fail(ESR_ERR_INVALID, "unpack16: null pointer");
  const size_t total = (size_t)n * c * h * w;
  esr::unpack_planes_kernel<true><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
      src16, dtype == ESR_BF16X3 ? 1 : dtype, n, c, h, w, planes_total, plane_off, dst, dtype == ESR_BF16X3 ? (size_t)(planes_total / 2) * h * w * 8 : 0);
  g_launches++;
  CUDA_TRY(cudaGetLastError());
  return ESR_OK;
}

int esr_unpack_planes32(const float* src32, int n, int c, int h, int w, int planes_total, int plane_off, float* dst,
                        void* stream) {
  if (!src32 || !dst) return fail(ESR_ERR_INVALID, "unpack32: null pointer");
  const size_t total = (size_t)n * c * h * w;
  esr::unpack_planes_kernel<false><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(src32, 0, n, c, h, w,
                                                                                        planes_total, plane_off, dst, 0);
  g_launches++;
  CUDA_TRY(cudaGetLastError());
  return ESR_OK;
}

int esr_pixel_unshuffle2_planes16(const void* src, int n, int planes, int h, int w, int src_planes_total, int src_plane_off, void* dst,
                                  int dst_planes_total, int dst_plane_off, void* stream) {
  if (!src || !dst || n <= 0 || planes <= 0 || h <= 0 || w <= 0) return fail(ESR_ERR_INVALID, "pixel_unshuffle: bad arguments");
  if (src_plane_off < 0 || src_plane_off + planes > src_planes_total || dst_plane_off < 0 || dst_plane_off + 4 * planes > dst_planes_total)
    return fail(ESR_ERR_INVALID, "pixel_unshuffle: planes out of range");
  if (((uintptr_t)src & 15) || ((uintptr_t)dst & 15)) return fail(ESR_ERR_INVALID, "pixel_unshuffle: pointers must be 16-byte aligned");
  const size_t total = (size_t)n * planes * h * w;
  esr::pixel_unshuffle2_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>((const uint16_t*)src, n, planes, h, w, src_planes_total,
                                                                                     src_plane_off, (uint16_t*)dst, dst_planes_total, dst_plane_off);
  g_launches++;
  CUDA_TRY(cudaGetLastError());
  return ESR_OK;
}

int esr_upsample2x_planes16(const void* src, int n, int planes, int h, int w, void* dst, void* stream) {
  if (!src || !dst) return fail(ESR_ERR_INVALID, "upsample2x: null pointer");
  const size_t total = (size_t)n * planes * h * w;
  esr::upsample2x_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>((const uint4*)src, (size_t)n * planes, h, w,
                                                                               (uint4*)dst);
  g_launches++;
  CUDA_TRY(cudaGetLastError());
  return ESR_OK;
}

int esr_latent_downscale(const float* z_hr, int n, int c, int hh, int wh, int s, int pad_hr, float* out, void* stream) {
  if (!z_hr || !out) return fail(ESR_ERR_INVALID, "latent_downscale: null pointer");
  if (s < 1 || (hh + 2 * pad_hr) % s || (wh + 2 * pad_hr) % s) return fail(ESR_ERR_INVALID, "latent_downscale: size not divisible by %d", s);
  const size_t total = (size_t)n * c * ((hh + 2 * pad_hr) / s) * ((wh + 2 * pad_hr) / s);
  esr::latent_downscale_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(z_hr, (size_t)n * c, hh, wh, s, pad_hr, out);
  g_launches++;
  CUDA_TRY(cudaGetLastError());
  return ESR_OK;
}

int esr_downsum2x_planes(const float* src32, int n, int planes, int h, int w, const void* act16_hi, float slope, int dtype,
                         float* dst32, void* dst16, void* stream) {
  if (!src32 || (!dst32 && !dst16)) return fail(ESR_ERR_INVALID, "downsum2x: null pointer");
  const size_t total = (size_t)n * planes * h * w;
  esr::downsum2x_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>((const float4*)src32, (size_t)n * planes, h, w,
                                                                              (const uint4*)act16_hi, slope, dtype == ESR_BF16X3 ? 1 : dtype,
                                                                              (float4*)dst32, (uint4*)dst16, planes, dtype == ESR_BF16X3 ? 1 : 0);
  g_launches++;
  CUDA_TRY(cudaGetLastError());
  return ESR_OK;
}

int esr_planes_add(const float* a32, const float* b32, size_t n_groups8, int dtype, float* out32, void* out16, void* stream) {
  return esr_planes_add_ex(a32, b32, n_groups8, 0, dtype, out32, out16, stream);
}

int esr_planes_add_ex(const float* a32, const float* b32, size_t n_groups8, size_t per_image_groups8, int dtype, float* out32, void* out16,
                      void* stream) {
  if (!a32 || !b32 || (!out32 && !out16)) return fail(ESR_ERR_INVALID, "planes_add: null pointer");
  const bool split = dtype == ESR_BF16X3 && out16;
  if (split && (per_image_groups8 == 0 || n_groups8 % per_image_groups8)) return fail(ESR_ERR_INVALID, "planes_add: split output needs the per-image size");
  esr::planes_add_kernel<<<grid_for(n_groups8, 256), 256, 0, (cudaStream_t)stream>>>((const float4*)a32, (const float4*)b32, n_groups8,
                                                                                   split ? 1 : dtype, (float4*)out32, (uint4*)out16,
                                                                                   per_image_groups8, split ? 1 : 0);
  g_launches++;
  CUDA_TRY(cudaGetLastError());
  return ESR_OK;
}

int esr_sep_adjoint_1d(const float* gout, int imgs, int outer, int inner, int n_out, int o_lo, int o_cnt, int n_in, int n_store,
                       int a_stride, int c_off, int m_stride, int m_phase, const float* taps, int len, const float* sub_from,
                       int sub_crop, float* gin, void* stream) {
  if (!gout || !taps || !gin) return fail(ESR_ERR_INVALID, "sep_adjoint_1d: null pointer");
  if (a_stride < 1 || m_stride < 1 || len < 1) return fail(ESR_ERR_INVALID, "sep_adjoint_1d: bad strides");
  if (o_lo < 0 || o_cnt < 0 || o_lo + o_cnt > n_out) return fail(ESR_ERR_INVALID, "sep_adjoint_1d: bad stored range");
  if (sub_from && inner != 1) return fail(ESR_ERR_INVALID, "sep_adjoint_1d: sub_from needs the x-axis pass");
  const size_t total = (size_t)imgs * outer * n_store * inner;
  esr::sep_adjoint_1d_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(gout, imgs, outer, inner, n_out, o_lo, o_cnt, n_in,
                                                                                   n_store, a_stride, c_off, m_stride, m_phase, taps, len,
                                                                                   sub_from, sub_crop, gin);
  g_launches++;
  CUDA_TRY(cudaGetLastError());
  return ESR_OK;
}

int esr_latent_grad(const float* gz_hr_planes32, const float* gz_lr_planes32, int n, int c, int hh, int wh, int s, int pad_hr,
                    float* dst_nchw, void* stream) {
  if (!dst_nchw || c > 8) return fail(ESR_ERR_INVALID, "latent_grad: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  CUDA_TRY(cudaMemsetAsync(dst_nchw, 0, sizeof(float) * (size_t)n * c * hh * wh, st));
  if (gz_hr_planes32) {
    const size_t total = (size_t)n * (hh + 2 * pad_hr) * (wh + 2 * pad_hr);
    esr::latent_grad_hr_kernel<<<grid_for(total, 256), 256, 0, st>>>(gz_hr_planes32, n, c, hh, wh, pad_hr, dst_nchw);
    g_launches++;
  }
  if (gz_lr_planes32) {
    const size_t total = (size_t)n * ((hh + 2 * pad_hr) / s) * ((wh + 2 * pad_hr) / s);
    esr::latent_grad_lr_kernel<<<grid_for(total, 256), 256, 0, st>>>(gz_lr_planes32, n, c, hh, wh, s, pad_hr, dst_nchw);
    g_launches++;
  }
  CUDA_TRY(cudaGetLastError());
  return ESR_OK;
}

// rank-1 filters take the HBM-bound kernels of cem_kernels.cuh (ESR_CEM_FAST=0: the general kernels, which also cover rank > 1)
static bool cem_fast(int rank, int len) {
  static const bool on = [] { const char* e = getenv("ESR_CEM_FAST"); return !(e && e[0] == '0'); }();
  return on && rank == 1 && len <= esr::kCemMaxTaps;
}

static int set_smem_attr(const void* fn, size_t bytes) {
  if (bytes > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return fail(ESR_ERR_CUDA, "cudaFuncSetAttribute(%zu): %s", bytes, cudaGetErrorString(e));
  }
  return ESR_OK;
}

int esr_cem_down(const float* g, int n, int c, int hh, int wh, int s, int phase, const float* kd_v, const float* kd_h,
                 int kd_len, int rank, const float* sub_from, float* out_lr, void* stream) {
  if (!g || !kd_v || !kd_h || !out_lr) return fail(ESR_ERR_INVALID, "cem_down: null pointer");
  if (s < 1 || hh % s || wh % s) return fail(ESR_ERR_INVALID, "cem_down: HR size %dx%d not divisible by scale %d", hh, wh, s);
  if (phase < 0 || phase >= s || kd_len < 1 || rank < 1) return fail(ESR_ERR_INVALID, "cem_down: bad phase/filter");
  const int hl = hh / s, wl = wh / s;
  if (cem_fast(rank, kd_len)) {
    typedef void (*DownFn)(const float*, int, int, int, const float*, const float*, const float*, float*);
    DownFn fn = nullptr;
    if (s == 2 && kd_len == 9) fn = esr::cem_down_fast_kernel<2, 9>;
    else if (s == 3 && kd_len == 11) fn = esr::cem_down_fast_kernel<3, 11>;
    else if (s == 4 && kd_len == 17) fn = esr::cem_down_fast_kernel<4, 17>;
    else if (s == 8 && kd_len == 33) fn = esr::cem_down_fast_kernel<8, 33>;
    if (fn) {
      const int frows = (esr::kDnTI - 1) * s + kd_len, fcols = (esr::kDnTJ - 1) * s + kd_len;
      const int pitch = ((fcols + 3 + 3) & ~3) | 1;
      const size_t fsmem = sizeof(float) * ((size_t)frows * pitch + (size_t)frows * (esr::kDnTJ + 1));
      int rc = set_smem_attr((const void*)fn, fsmem);
      if (rc) return rc;
      dim3 grid((wl + esr::kDnTJ - 1) / esr::kDnTJ, (hl + esr::kDnTI - 1) / esr::kDnTI, n * c);
      fn<<<grid, 256, fsmem, (cudaStream_t)stream>>>(g, hh, wh, phase, kd_v, kd_h, sub_from, out_lr);
      g_launches++;
      CUDA_TRY(cudaGetLastError());
      return ESR_OK;
    }
  }
  const int rows = (esr::kCemTI - 1) * s + kd_len, cols = (esr::kCemTJ - 1) * s + kd_len;
  const size_t smem = sizeof(float) * ((size_t)rows * cols + (size_t)rows * esr::kCemTJ);
  if (smem > 227 * 1024) return fail(ESR_ERR_INVALID, "cem_down: filter too large for shared memory");
  int rc = set_smem_attr((const void*)esr::cem_down_kernel, smem);
  if (rc) return rc;
  dim3 grid((wl + esr::kCemTJ - 1) / esr::kCemTJ, (hl + esr::kCemTI - 1) / esr::kCemTI, n * c);
  esr::cem_down_kernel<<<grid, esr::kCemThreads, smem, (cudaStream_t)stream>>>(g, hh, wh, s, phase, kd_v, kd_h, kd_len, rank,
                                                                             sub_from, out_lr);
  g_launches++;
  CUDA_TRY(cudaGetLastError());
  return ESR_OK;
}

int esr_cem_inv(const float* e, int n, int c, int hl, int wl, const float* ki_v, const float* ki_h, int ki_len, int rank,
                float* out_lr, void* stream) {
  if (!e || !ki_v || !ki_h || !out_lr) return fail(ESR_ERR_INVALID, "cem_inv: null pointer");
  if (ki_len < 1 || rank < 1) return fail(ESR_ERR_INVALID, "cem_inv: bad filter");
  if (cem_fast(rank, ki_len) && ki_len == 27) {
    constexpr int LEN = 27;
    const int frows = esr::kInvTI + LEN - 1, fcols = esr::kInvTJ + LEN - 1;
    const size_t fsmem = sizeof(float) * ((size_t)frows * (fcols | 1) + (size_t)frows * (esr::kInvTJ + 1));
    int rc = set_smem_attr((const void*)esr::cem_inv_fast_kernel<LEN>, fsmem);
    if (rc) return rc;
    dim3 grid((wl + esr::kInvTJ - 1) / esr::kInvTJ, (hl + esr::kInvTI - 1) / esr::kInvTI, n * c);
    esr::cem_inv_fast_kernel<LEN><<<grid, 256, fsmem, (cudaStream_t)stream>>>(e, hl, wl, ki_v, ki_h, out_lr);
    g_launches++;
    CUDA_TRY(cudaGetLastError());
    return ESR_OK;
  }
  const int rows = esr::kCemTI + ki_len - 1, cols = esr::kCemTJ + ki_len - 1;
  const size_t smem = sizeof(float) * ((size_t)rows * cols + (size_t)rows * esr::kCemTJ);
  if (smem > 227 * 1024) return fail(ESR_ERR_INVALID, "cem_inv: filter too large for shared memory");
  int rc = set_smem_attr((const void*)esr::cem_inv_kernel, smem);
  if (rc) return rc;
  dim3 grid((wl + esr::kCemTJ - 1) / esr::kCemTJ, (hl + esr::kCemTI - 1) / esr::kCemTI, n * c);
  esr::cem_inv_kernel<<<grid, esr::kCemThreads, smem, (cudaStream_t)stream>>>(e, hl, wl, ki_v, ki_h, ki_len, rank, out_lr);
  g_launches++;
  CUDA_TRY(cudaGetLastError());
  return ESR_OK;
}

int esr_cem_up_add(const float* f, const float* g, int n, int c, int hl, int wl, int s, int phase, const float* ku_v,
                   const float* ku_h, int ku_len, int rank, int crop, float* out_hr, void* stream) {
  if (!f || !ku_v || !ku_h || !out_hr) return fail(ESR_ERR_INVALID, "cem_up_add: null pointer");
  if (s < 1 || phase < 0 || phase >= s || ku_len < 1 || rank < 1) return fail(ESR_ERR_INVALID, "cem_up_add: bad scale/phase/filter");
  const int hh = hl * s, wh = wl * s;
  if (crop < 0 || 2 * crop >= hh || 2 * crop >= wh) return fail(ESR_ERR_INVALID, "cem_up_add: crop %d too large", crop);
  const int ho = hh - 2 * crop, wo = wh - 2 * crop;
  if (cem_fast(rank, ku_len) && (s == 2 || s == 3 || s == 4 || s == 8)) {
    typedef void (*UpFn)(const float*, const float*, int, int, int, const float*, const float*, int, int, float*);
    UpFn fn = s == 2 ? esr::cem_up_add_fast_kernel<2> : (s == 3 ? esr::cem_up_add_fast_kernel<3> : (s == 4 ? esr::cem_up_add_fast_kernel<4> : esr::cem_up_add_fast_kernel<8>));
    const int maxni = (esr::kUpTY + ku_len) / s + 2, maxnj = (esr::kUpTX + ku_len) / s + 2;
    const size_t fsmem = sizeof(float) * ((size_t)maxni * maxnj + (size_t)maxni * esr::kUpTX);
    int rc = set_smem_attr((const void*)fn, fsmem);
    if (rc) return rc;
    dim3 grid((wo + esr::kUpTX - 1) / esr::kUpTX, (ho + esr::kUpTY - 1) / esr::kUpTY, n * c);
    fn<<<grid, 256, fsmem, (cudaStream_t)stream>>>(f, g, hl, wl, phase, ku_v, ku_h, ku_len, crop, out_hr);
    g_launches++;
    CUDA_TRY(cudaGetLastError());
    return ESR_OK;
  }
  const int maxn = (esr::kUpT + ku_len) / s + 2;
  const size_t smem = sizeof(float) * ((size_t)maxn * maxn + (size_t)maxn * esr::kUpT);
  int rc = set_smem_attr((const void*)esr::cem_up_add_kernel, smem);
  if (rc) return rc;
  dim3 grid((wo + esr::kUpT - 1) / esr::kUpT, (ho + esr::kUpT - 1) / esr::kUpT, n * c);
  esr::cem_up_add_kernel<<<grid, esr::kCemThreads, smem, (cudaStream_t)stream>>>(f, g, hl, wl, s, phase, ku_v, ku_h, ku_len,
                                                                               rank, crop, out_hr);
  g_launches++;
  CUDA_TRY(cudaGetLastError());
  return ESR_OK;
}

// ---- discriminator: BatchNorm2d (batch statistics) + LeakyReLU, space-to-depth, fully connected layers ----------------------------
static int bn_chunks(size_t m) {
  size_t c = (m + 1023) / 1024;
  if (c < 1) c = 1;
  if (c > (size_t)esr::kBnMaxChunks) c = esr::kBnMaxChunks;
  return (int)c;
}

size_t esr_bn_workspace_bytes(int planes) { return planes > 0 ? (size_t)planes * esr::kBnMaxChunks * 16 * sizeof(float) : 0; }

int esr_bn_stats(const float* y32, int n, int planes, int h, int w, int c, const float* gamma, const float* beta, float eps,
                 float momentum, int train, float* running_mean, float* running_var, float* save_mean, float* save_invstd,
                 float* scale, float* shift, float* workspace, size_t workspace_bytes, void* stream) {
  if (!y32 || !gamma || !beta || !save_mean || !save_invstd || !scale || !shift) return fail(ESR_ERR_INVALID, "bn_stats: null pointer");
  if (n <= 0 || planes <= 0 || h <= 0 || w <= 0 || c <= 0 || c > planes * 8) return fail(ESR_ERR_INVALID, "bn_stats: bad shape");
  if (!train && (!running_mean || !running_var)) return fail(ESR_ERR_INVALID, "bn_stats: eval mode needs running statistics");
  const size_t m = (size_t)n * h * w;
  const int chunks = bn_chunks(m);
  cudaStream_t st = (cudaStream_t)stream;
  if (train) {
    if (!workspace || workspace_bytes < esr_bn_workspace_bytes(planes)) return fail(ESR_ERR_INVALID, "bn_stats: workspace too small");
    esr::bn_partial_kernel<<<dim3((unsigned)chunks, (unsigned)planes), 256, 0, st>>>(y32, n, planes, h * w, workspace, chunks);
    g_launches++;
    CUDA_TRY(cudaGetLastError());
  }
  esr::bn_finalize_kernel<<<(c + 127) / 128, 128, 0, st>>>(workspace, chunks, c, (double)m, gamma, beta, eps, momentum, train, running_mean,
                                                          running_var, save_mean, save_invstd, scale, shift);
  g_launches++;
  CUDA_TRY(cudaGetLastError());
  return ESR_OK;
}

int esr_bn_lrelu_fwd(const float* y32, int n, int planes, int h, int w, int c, const float* scale, const float* shift, float slope,
                     int dtype, void* dst16, int space_to_depth, float* dst_nchw, void* stream) {
  if (!y32 || !scale || !shift || (!dst16 && !dst_nchw)) return fail(ESR_ERR_INVALID, "bn_lrelu_fwd: null pointer");
  if (n <= 0 || planes <= 0 || h <= 0 || w <= 0 || c <= 0 || c > planes * 8) return fail(ESR_ERR_INVALID, "bn_lrelu_fwd: bad shape");
  if (dtype != ESR_F16 && dtype != ESR_BF16 && dtype != ESR_BF16X3) return fail(ESR_ERR_INVALID, "bn_lrelu_fwd: bad dtype");
  if (space_to_depth && ((h & 1) || (w & 1))) return fail(ESR_ERR_INVALID, "bn_lrelu_fwd: space-to-depth needs an even size, got %dx%d", h, w);
  const size_t total = (size_t)n * planes * h * w;
  esr::bn_lrelu_apply_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(y32, scale, shift, slope, n, planes, h, w, c,
                                                                                  dtype == ESR_BF16X3 ? 1 : dtype, (uint16_t*)dst16, space_to_depth,
                                                                                  dst_nchw, dtype == ESR_BF16X3 ? 1 : 0);
  g_launches++;
  CUDA_TRY(cudaGetLastError());
  return ESR_OK;
}

int esr_space_to_depth_planes16(const void* src, int n, int planes, int h, int w, void* dst, void* stream) {
  if (!src || !dst) return fail(ESR_ERR_INVALID, "space_to_depth: null pointer");
  if (n <= 0 || planes <= 0 || h <= 0 || w <= 0 || (h & 1) || (w & 1)) return fail(ESR_ERR_INVALID, "space_to_depth: bad shape %dx%d", h, w);
  const size_t total = (size_t)n * planes * h * w;
  esr::s2d_planes16_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>((const uint4*)src, n, planes, h, w, (uint4*)dst);
  g_launches++;
  CUDA_TRY(cudaGetLastError());
  return ESR_OK;
}

int esr_bn_lrelu_bwd(const float* g, int g_layout, const float* y32, int n, int planes, int h, int w, int c, const float* scale,
                     const float* shift, const float* save_mean, const float* save_invstd, float slope, int has_bn, int train, float gscale,
                     int accumulate, float* dgamma, float* dbeta, float* c1, float* c2, int dtype, void* gy16, float* workspace,
                     size_t workspace_bytes, void* stream) {
  if (!g || !y32 || !scale || !shift || !save_mean || !save_invstd || !c1 || !c2 || !gy16) return fail(ESR_ERR_INVALID, "bn_lrelu_bwd: null pointer");
  if (n <= 0 || planes <= 0 || h <= 0 || w <= 0 || c <= 0 || c > planes * 8) return fail(ESR_ERR_INVALID, "bn_lrelu_bwd: bad shape");
  if (g_layout < 0 || g_layout > 2) return fail(ESR_ERR_INVALID, "bn_lrelu_bwd: bad gradient layout %d", g_layout);
  if (g_layout == 1 && ((h & 1) || (w & 1))) return fail(ESR_ERR_INVALID, "bn_lrelu_bwd: space-to-depth gradient needs an even size");
  if (dtype != ESR_F16 && dtype != ESR_BF16 && dtype != ESR_BF16X3) return fail(ESR_ERR_INVALID, "bn_lrelu_bwd: bad dtype");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t m = (size_t)n * h * w;
  if (has_bn) {
    if (!workspace || workspace_bytes < esr_bn_workspace_bytes(planes)) return fail(ESR_ERR_INVALID, "bn_lrelu_bwd: workspace too small");
    const int chunks = bn_chunks(m);
    esr::bn_bwd_partial_kernel<<<dim3((unsigned)chunks, (unsigned)planes), 256, 0, st>>>(g, g_layout, y32, scale, shift, save_mean, save_invstd,
                                                                                        slope, n, planes, h, w, c, workspace, chunks);
    g_launches++;
    CUDA_TRY(cudaGetLastError());
    esr::bn_bwd_finalize_kernel<<<(c + 127) / 128, 128, 0, st>>>(workspace, chunks, c, (double)m, train, gscale, accumulate, dgamma, dbeta, c1, c2);
    g_launches++;
    CUDA_TRY(cudaGetLastError());
  }
  const size_t total = (size_t)n * planes * h * w;
  esr::bn_bwd_apply_kernel<<<grid_for(total, 256), 256, 0, st>>>(g, g_layout, y32, scale, shift, save_mean, save_invstd, c1, c2, slope, n, planes,
                                                                h, w, c, dtype == ESR_BF16X3 ? 1 : dtype, (uint16_t*)gy16, dtype == ESR_BF16X3 ? 1 : 0);
  g_launches++;
  CUDA_TRY(cudaGetLastError());
  return ESR_OK;
}

// ---- gradient penalty (WGAN-GP): tangent forward and double backward of BatchNorm + LeakyReLU ----------------------------------------
size_t esr_bn_dbl_workspace_bytes(int planes) { return planes > 0 ? (size_t)planes * esr::kBnMaxChunks * 40 * sizeof(float) : 0; }

int esr_bn_tangent_fwd(const float* t32, const float* y32, int n, int planes, int h, int w, int c, const float* scale, const float* shift,
                       const float* save_mean, const float* save_invstd, float slope, int has_bn, float* c1, float* c2, int dtype, void* dst16,
                       int space_to_depth, float* dst_nchw, float* workspace, size_t workspace_bytes, void* stream) {
  if (!t32 || !y32 || !scale || !shift || !save_mean || !save_invstd || !c1 || !c2 || (!dst16 && !dst_nchw)) return fail(ESR_ERR_INVALID, "bn_tangent_fwd: null pointer");
  if (n <= 0 || planes <= 0 || h <= 0 || w <= 0 || c <= 0 || c > planes * 8) return fail(ESR_ERR_INVALID, "bn_tangent_fwd: bad shape");
  if (dtype != ESR_F16 && dtype != ESR_BF16 && dtype != ESR_BF16X3) return fail(ESR_ERR_INVALID, "bn_tangent_fwd: bad dtype");
  if (space_to_depth && ((h & 1) || (w & 1))) return fail(ESR_ERR_INVALID, "bn_tangent_fwd: space-to-depth needs an even size");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t m = (size_t)n * h * w;
  if (has_bn) {     // c1 = mean(t), c2 = mean(yh t): the reductions of the BatchNorm backward with the mask switched off (slope 1)
    if (!workspace || workspace_bytes < esr_bn_workspace_bytes(planes)) return fail(ESR_ERR_INVALID, "bn_tangent_fwd: workspace too small");
    const int chunks = bn_chunks(m);
    esr::bn_bwd_partial_kernel<<<dim3((unsigned)chunks, (unsigned)planes), 256, 0, st>>>(t32, 0, y32, scale, shift, save_mean, save_invstd, 1.0f, n,
                                                                                        planes, h, w, c, workspace, chunks);
    g_launches++;
    CUDA_TRY(cudaGetLastError());
    esr::bn_bwd_finalize_kernel<<<(c + 127) / 128, 128, 0, st>>>(workspace, chunks, c, (double)m, 1, 1.0f, 0, nullptr, nullptr, c1, c2);
    g_launches++;
    CUDA_TRY(cudaGetLastError());
  } else {
    CUDA_TRY(cudaMemsetAsync(c1, 0, sizeof(float) * c, st));
    CUDA_TRY(cudaMemsetAsync(c2, 0, sizeof(float) * c, st));
  }
  const size_t total = (size_t)n * planes * h * w;
  esr::bn_tangent_apply_kernel<<<grid_for(total, 256), 256, 0, st>>>(t32, y32, scale, shift, save_mean, save_invstd, c1, c2, slope, n, planes, h, w, c,
                                                                    dtype == ESR_BF16X3 ? 1 : dtype, (uint16_t*)dst16, space_to_depth, dst_nchw,
                                                                    dtype == ESR_BF16X3 ? 1 : 0);
  g_launches++;
  CUDA_TRY(cudaGetLastError());
  return ESR_OK;
}

int esr_bn_double_bwd(const float* zb, const float* wb, int g_layout, const float* y32, const float* t32, int n, int planes, int h, int w, int c,
                      const float* scale, const float* shift, const float* save_mean, const float* save_invstd, const float* c1, const float* c2,
                      float slope, int has_bn, float gscale, int accumulate, float* dgamma, float* dbeta, float* coef, int dtype, void* tb16,
                      void* yb16, float* workspace, size_t workspace_bytes, void* stream) {
  if ((!zb && !wb) || !y32 || !t32 || !scale || !shift || !save_mean || !save_invstd || !c1 || !c2 || !coef || (!tb16 && !yb16))
    return fail(ESR_ERR_INVALID, "bn_double_bwd: null pointer");
  if (n <= 0 || planes <= 0 || h <= 0 || w <= 0 || c <= 0 || c > planes * 8) return fail(ESR_ERR_INVALID, "bn_double_bwd: bad shape");
  if (g_layout < 0 || g_layout > 2) return fail(ESR_ERR_INVALID, "bn_double_bwd: bad gradient layout %d", g_layout);
  if (g_layout == 1 && ((h & 1) || (w & 1))) return fail(ESR_ERR_INVALID, "bn_double_bwd: space-to-depth gradient needs an even size");
  if (dtype != ESR_F16 && dtype != ESR_BF16 && dtype != ESR_BF16X3) return fail(ESR_ERR_INVALID, "bn_double_bwd: bad dtype");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t m = (size_t)n * h * w;
  if (has_bn) {
    if (!workspace || workspace_bytes < esr_bn_dbl_workspace_bytes(planes)) return fail(ESR_ERR_INVALID, "bn_double_bwd: workspace too small");
    const int chunks = bn_chunks(m);
    esr::bn_dbl_partial_kernel<<<dim3((unsigned)chunks, (unsigned)planes), 256, 0, st>>>(zb, wb, g_layout, y32, t32, scale, shift, save_mean, save_invstd,
                                                                                        c1, c2, slope, n, planes, h, w, c, workspace, chunks);
    g_launches++;
    CUDA_TRY(cudaGetLastError());
    esr::bn_dbl_finalize_kernel<<<(c + 127) / 128, 128, 0, st>>>(workspace, chunks, c, (double)m, save_invstd, 1, gscale, accumulate, dgamma, dbeta, coef);
    g_launches++;
    CUDA_TRY(cudaGetLastError());
  }
  const size_t total = (size_t)n * planes * h * w;
  esr::bn_dbl_apply_kernel<<<grid_for(total, 256), 256, 0, st>>>(zb, wb, g_layout, y32, t32, scale, shift, save_mean, save_invstd, c1, c2, coef, has_bn,
                                                                slope, n, planes, h, w, c, dtype == ESR_BF16X3 ? 1 : dtype, (uint16_t*)tb16,
                                                                (uint16_t*)yb16, dtype == ESR_BF16X3 ? 1 : 0);
  g_launches++;
  CUDA_TRY(cudaGetLastError());
  return ESR_OK;
}

int esr_linear_fwd(const float* x, const float* weight, const float* bias, int batch, int in_features, int out_features, int lrelu,
                   float slope, float* out, void* stream) {
  if (!x || !weight || !out) return fail(ESR_ERR_INVALID, "linear_fwd: null pointer");
  if (batch <= 0 || in_features <= 0 || out_features <= 0) return fail(ESR_ERR_INVALID, "linear_fwd: bad shape");
  esr::linear_fwd_kernel<<<out_features, 256, 0, (cudaStream_t)stream>>>(x, weight, bias, batch, in_features, out_features, lrelu, slope, out);
  g_launches++;
  CUDA_TRY(cudaGetLastError());
  return ESR_OK;
}

int esr_linear_bwd(const float* g, const float* act, float slope, const float* x, const float* weight, int batch, int in_features,
                   int out_features, float gscale, int accumulate, float* gx, float* dweight, float* dbias, void* stream) {
  if (!g || !x || !weight) return fail(ESR_ERR_INVALID, "linear_bwd: null pointer");
  if (batch <= 0 || in_features <= 0 || out_features <= 0) return fail(ESR_ERR_INVALID, "linear_bwd: bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  if (gx) {
    esr::linear_bwd_input_kernel<<<grid_for((size_t)batch * in_features, 256), 256, 0, st>>>(g, act, slope, weight, batch, in_features,
                                                                                            out_features, gx);
    g_launches++;
    CUDA_TRY(cudaGetLastError());
  }
  if (dweight || dbias) {
    esr::linear_bwd_weight_kernel<<<grid_for((size_t)out_features * in_features, 256), 256, 0, st>>>(g, act, slope, x, batch, in_features,
                                                                                                    out_features, gscale, accumulate, dweight, dbias);
    g_launches++;
    CUDA_TRY(cudaGetLastError());
  }
  return ESR_OK;
}

size_t esr_structure_tensor_workspace_bytes(int n) { return n > 0 ? (size_t)n * esr::kStChunks * 3 * sizeof(float) : 0; }

int esr_structure_tensor_fwd(const float* img, int n, int c, int h, int w, float* out, float* workspace, size_t workspace_bytes,
                             void* stream) {
  if (!img || !out || !workspace) return fail(ESR_ERR_INVALID, "structure_tensor_fwd: null pointer");
  if (n <= 0 || c <= 0 || h < 2 || w < 2) return fail(ESR_ERR_INVALID, "structure_tensor_fwd: bad shape %dx%dx%dx%d", n, c, h, w);
  if (workspace_bytes < esr_structure_tensor_workspace_bytes(n)) return fail(ESR_ERR_INVALID, "structure_tensor_fwd: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  esr::structure_tensor_partial_kernel<<<dim3(esr::kStChunks, (unsigned)n), 256, 0, st>>>(img, c, h, w, workspace);
  g_launches++;
  CUDA_TRY(cudaGetLastError());
  esr::structure_tensor_finalize_kernel<<<(n * 3 + 127) / 128, 128, 0, st>>>(workspace, n, 1.0 / ((double)c * (h - 1) * (w - 1)), out);
  g_launches++;
  CUDA_TRY(cudaGetLastError());
  return ESR_OK;
}

int esr_structure_tensor_bwd(const float* img, const float* g, int n, int c, int h, int w, float* grad_img, void* stream) {
  if (!img || !g || !grad_img) return fail(ESR_ERR_INVALID, "structure_tensor_bwd: null pointer");
  if (n <= 0 || c <= 0 || h < 2 || w < 2) return fail(ESR_ERR_INVALID, "structure_tensor_bwd: bad shape %dx%dx%dx%d", n, c, h, w);
  const size_t total = (size_t)n * c * h * w;
  esr::structure_tensor_bwd_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(img, g, n, c, h, w,
                                                                                       (float)(1.0 / ((double)c * (h - 1) * (w - 1))), grad_img);
  g_launches++;
  CUDA_TRY(cudaGetLastError());
  return ESR_OK;
}

// ---- training step outside the networks: Adam, L1, relativistic BCE --------------------------------------------------------------
size_t esr_adam_scratch_bytes(int count) { return count > 0 ? (size_t)count * sizeof(esr::AdamTensor) : 0; }

int esr_adam_multi(const esr_adam_tensor* tensors_host, int count, void* scratch, size_t scratch_bytes, float lr, float beta1, float beta2,
                   float eps, float weight_decay, int step, float grad_scale, void* stream) {
  static_assert(sizeof(esr_adam_tensor) == sizeof(esr::AdamTensor), "esr_adam_tensor layout");
  if (count <= 0) return ESR_OK;
  if (!scratch || scratch_bytes < esr_adam_scratch_bytes(count)) return fail(ESR_ERR_INVALID, "adam: scratch of %zu bytes needed", esr_adam_scratch_bytes(count));
  if (step < 1) return fail(ESR_ERR_INVALID, "adam: step counts from 1");
  cudaStream_t st = (cudaStream_t)stream;
  unsigned long long nmax = 0;
  if (tensors_host) {     // NULL: the table uploaded by an earlier call is still in `scratch` (same tensors, e.g. flat buffers)
    for (int i = 0; i < count; ++i) {
      if (!tensors_host[i].p || !tensors_host[i].g || !tensors_host[i].m || !tensors_host[i].v) return fail(ESR_ERR_INVALID, "adam: null pointer in tensor %d", i);
      if (tensors_host[i].n > nmax) nmax = tensors_host[i].n;
    }
    CUDA_TRY(cudaMemcpyAsync(scratch, tensors_host, sizeof(esr::AdamTensor) * (size_t)count, cudaMemcpyHostToDevice, st));
  } else {
    nmax = count == 1 ? (1ull << 40) : (1ull << 18);
  }
  const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
  const float step_size = (float)((double)lr / bc1), bc2_sqrt = (float)sqrt(bc2);
  // blocks per tensor: enough for the largest one at 4 elements per thread, capped so that the grid stays a few waves
  unsigned long long bx = (nmax / 4 + 255) / 256;
  const unsigned long long cap = count == 1 ? 148ull * 8 : 16ull;
  if (bx > cap) bx = cap;
  if (bx < 1) bx = 1;
  for (int i0 = 0; i0 < count; i0 += 65535) {
    const int n = count - i0 < 65535 ? count - i0 : 65535;
    esr::adam_multi_kernel<<<dim3((unsigned)bx, (unsigned)n), 256, 0, st>>>((const esr::AdamTensor*)scratch + i0, beta1, beta2, eps, weight_decay,
                                                                          step_size, bc2_sqrt, grad_scale);
    g_launches++;
    CUDA_TRY(cudaGetLastError());
  }
  return ESR_OK;
}

size_t esr_l1_workspace_bytes(void) { return (size_t)esr::kL1Blocks * sizeof(float); }

int esr_l1_reduce(const float* a, const float* b, size_t n, float* workspace, size_t workspace_bytes, float* out_mean, void* stream) {
  if (!a || !b || !workspace || !out_mean || n == 0) return fail(ESR_ERR_INVALID, "l1_reduce: bad arguments");
  if (workspace_bytes < esr_l1_workspace_bytes()) return fail(ESR_ERR_INVALID, "l1_reduce: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  int blocks = (int)((n / 4 + 255) / 256);
  if (blocks > esr::kL1Blocks) blocks = esr::kL1Blocks;
  if (blocks < 1) blocks = 1;
  esr::l1_partial_kernel<<<blocks, 256, 0, st>>>(a, b, n, workspace);
  g_launches++;
  CUDA_TRY(cudaGetLastError());
  esr::sum_partials_kernel<<<1, 256, 0, st>>>(workspace, blocks, 1.0 / (double)n, out_mean);
  g_launches++;
  CUDA_TRY(cudaGetLastError());
  return ESR_OK;
}

int esr_l1_grad(const float* a, const float* b, size_t n, const float* gout, float* ga, float* gb, void* stream) {
  if (!a || !b || !gout || (!ga && !gb) || n == 0) return fail(ESR_ERR_INVALID, "l1_grad: bad arguments");
  esr::l1_grad_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(a, b, n, gout, (float)(1.0 / (double)n), ga, gb);
  g_launches++;
  CUDA_TRY(cudaGetLastError());
  return ESR_OK;
}

int esr_bce_rel_loss(const float* a, const float* b, int n, const float* sums_global, float n_global, float target_a, float target_b, float* out4,
                     float* ea, float* eb, void* stream) {
  if (!a || !b || !out4 || !ea || !eb || n <= 0) return fail(ESR_ERR_INVALID, "bce_rel_loss: bad arguments");
  esr::bce_rel_fwd_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(a, b, n, sums_global, sums_global ? n_global : (float)n, target_a, target_b, out4, ea, eb);
  g_launches++;
  CUDA_TRY(cudaGetLastError());
  return ESR_OK;
}

int esr_bce_rel_loss_bwd(const float* ea, const float* eb, int n, const float* S_global, float n_global, const float* ga_up, const float* gb_up,
                         float* ga, float* gb, void* stream) {
  if (!ea || !eb || !S_global || (!ga && !gb) || n <= 0) return fail(ESR_ERR_INVALID, "bce_rel_loss_bwd: bad arguments");
  esr::bce_rel_bwd_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(ea, eb, n, S_global, n_global, ga_up, gb_up, ga, gb);
  g_launches++;
  CUDA_TRY(cudaGetLastError());
  return ESR_OK;
}

// ---- soft histogram / dictionary objective of the latent-exploration tools --------------------------------------------------------
int esr_soft_hist_fwd(const double* x, int dims, int n_samples, const double* bins, int n_bins, double vmax, double eps, double temperature,
                      int dictionary, double* out, double* sum_e, void* stream) {
  if (!x || !bins || !out) return fail(ESR_ERR_INVALID, "soft_hist_fwd: null pointer");
  if (dims < 1 || dims > esr::kHistMaxD || n_samples < 1 || n_bins < 1 || !(temperature > 0)) return fail(ESR_ERR_INVALID, "soft_hist_fwd: bad shape (dims %d, samples %d, bins %d)", dims, n_samples, n_bins);
  if (dictionary && !sum_e) return fail(ESR_ERR_INVALID, "soft_hist_fwd: dictionary mode needs sum_e");
  cudaStream_t st = (cudaStream_t)stream;
  const double inv_DT = 1.0 / ((double)dims * temperature);
  if (dictionary) {
    esr::soft_dict_fwd_kernel<<<(n_samples + 127) / 128, 128, 0, st>>>(x, dims, n_samples, bins, n_bins, vmax, eps, inv_DT, out, sum_e);
  } else {
    CUDA_TRY(cudaMemsetAsync(out, 0, sizeof(double) * n_bins, st));
    const int bx = (n_bins + esr::kHistTileB - 1) / esr::kHistTileB;
    int chunks = deterministic() ? 1 : (148 * 4 + bx - 1) / bx;  // a few waves of blocks whatever the bin count
    int per = (n_samples + chunks - 1) / chunks;
    per = (per + esr::kHistTileP - 1) / esr::kHistTileP * esr::kHistTileP;
    chunks = (n_samples + per - 1) / per;
    esr::soft_hist_fwd_kernel<<<dim3((unsigned)bx, (unsigned)chunks), esr::kHistTileB, 0, st>>>(x, dims, n_samples, bins, n_bins, vmax, eps, inv_DT, per, out);
  }
  g_launches++;
  CUDA_TRY(cudaGetLastError());
  return ESR_OK;
}

int esr_soft_hist_bwd(const double* x, int dims, int n_samples, const double* bins, int n_bins, double vmax, double eps, double temperature,
                      const double* g_hist, const double* g_out, const double* sum_e, double* dx, void* stream) {
  if (!x || !bins || !dx || (!g_hist && !g_out) || (g_out && !sum_e)) return fail(ESR_ERR_INVALID, "soft_hist_bwd: null pointer");
  if (dims < 1 || dims > esr::kHistMaxD || n_samples < 1 || n_bins < 1 || !(temperature > 0)) return fail(ESR_ERR_INVALID, "soft_hist_bwd: bad shape");
  const double inv_DT = 1.0 / ((double)dims * temperature);
  esr::soft_hist_bwd_kernel<<<(n_samples + 127) / 128, 128, 0, (cudaStream_t)stream>>>(x, dims, n_samples, bins, n_bins, vmax, eps, inv_DT, g_hist, g_out,
                                                                                     sum_e, dx);
  g_launches++;
  CUDA_TRY(cudaGetLastError());
  return ESR_OK;
}

}  // extern "C"
