/*
 * esr_b200.h — C-ABI of the B200-native hot path of Explorable-Super-Resolution.
 *
 * The reference (YuvalBahat/Explorable-Super-Resolution) is pure Python/PyTorch and has no FFI of its
 * own; the "operator API" of the hot path is a set of torch.nn modules.  Each entry point below names
 * the reference call site it replaces (paths relative to the reference's codes/ directory).
 *
 * Conventions
 *   - every function returns 0 on success, a negative esr_status otherwise; esr_last_error() returns a
 *     human readable string for the calling thread's last failure.  No exception crosses this boundary.
 *   - all pointers are DEVICE pointers unless the name ends in _host.  The caller owns every buffer.
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream).
 *   - activation tensors use the planar-8 layout  [N][C/8][H][W][8]  ("planes"): channel c of pixel
 *     (n,y,x) lives at  (((n*planes_total + c/8)*H + y)*W + x)*8 + c%8.  16-bit planes carry the
 *     tensor-core operands (fp16 or bf16, selected by `dtype`), fp32 planes carry the residual trunk.
 *   - images at the boundary are NCHW fp32, exactly the reference's tensors.
 */
#ifndef ESR_B200_H
#define ESR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  ESR_OK = 0,
  ESR_ERR_INVALID = -1,     /* bad argument (shape / alignment / unsupported combination) */
  ESR_ERR_CUDA = -2,        /* a CUDA runtime / driver call failed                         */
  ESR_ERR_UNSUPPORTED = -3  /* device is not sm_100                                        */
} esr_status;

/* element format of the 16-bit tensor-core operands.
 * ESR_BF16X3 ("split precision", the parity mode): every 16-bit tensor holds planes_total/2 bf16 "hi" planes per image followed by
 * as many bf16 "lo" planes (lo = bf16(v - hi)); plane offsets always address the hi half.  Convolutions run
 * x*w ~= x_hi*w_hi + x_lo*w_hi + x_hi*w_lo on the kind::f16 tensor pipe with fp32 accumulation (16 mantissa bits per operand,
 * fp32 exponent range): the arithmetic class of the reference's fp32 convs (models/modules/block.py:141-146), 3x the MMAs. */
typedef enum { ESR_F16 = 0, ESR_BF16 = 1, ESR_BF16X3 = 2 } esr_dtype;

const char* esr_last_error(void);
int esr_version(void);
/* number of kernels launched by this library in this process (bench.py's gpu_launches claim) */
long long esr_launch_count(void);
/* debugging aid: copies the 8-word pipeline watchdog record (word 0 != 0: some mbarrier wait timed out;
 * then block, thread, barrier address, parity, role tag) to host memory; optionally clears it.  Synchronises. */
int esr_debug_watchdog(unsigned int* out8_host, int reset);
/* Bit-reproducible launches (also ESR_DETERMINISTIC=1 in the environment): one MMA issuer per CTA in the row-streaming and wgrad
 * kernels (three issuers add into one TMEM accumulator in arrival order otherwise), no floating-point atomics across blocks in the
 * bias / channel-sum / histogram reductions.  Costs about 15 % of the convolution throughput.  (Not covered: the scatter of the
 * latent map's gradient in esr_latent_grad, which folds replicate-padded pixels onto the border with atomics.) */
int esr_set_deterministic(int on);
/* 0 if the current device can run the sm_100a kernels */
int esr_device_check(void);

/* ------------------------------------------------------------------------------------------------
 * 3x3 stride-1 zero-pad-1 convolution as a tcgen05 implicit GEMM with fused epilogue.
 * Replaces  models/modules/block.py:129-146 (conv_block: nn.Conv2d + LeakyReLU(0.2)),
 *           block.py:230-235 (torch.cat of the dense block + `x5*0.2 + x`),
 *           block.py:262-270 (RRDB `out*0.2 + x`), block.py:96 (ShortcutBlock add),
 *           block.py:299-300 (nearest x2 Upsampler, folded into the producer's store),
 *           block.py:287 (nn.PixelShuffle, folded into the store addressing).
 *
 *   acc = conv3x3(in[:, in_plane_off*8 : +cin_planes*8]) + bias
 *   if lrelu: acc = acc > 0 ? acc : slope*acc
 *   v = alpha*acc + beta1*res1 + beta2*res2
 *   stores (any subset): out16 planes, out32 planes, out_nchw (fp32 NCHW, first out_nchw_c channels)
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  int n, h, w;               /* batch, height, width (input == output spatial size)          */
  int dtype;                 /* esr_dtype of `in`, `wpacked`, `out16`                        */
  /* input (16-bit planes) */
  const void* in;
  int in_planes_total;       /* planes in the input buffer                                    */
  int in_plane_off;          /* first plane consumed                                          */
  int cin_planes;            /* planes consumed (Cin/8 rounded up)                            */
  /* weights packed by esr_pack_conv3x3_weights, bias fp32 [cout_pad] */
  const void* wpacked;
  const float* bias;
  int cout;                  /* real output channels                                          */
  int cout_pad;              /* padded to a multiple of the n-block (16/32/64)                */
  int kcp;                   /* planes per K chunk used when packing (2 or 4)                 */
  /* epilogue */
  int lrelu; float slope;
  float alpha;
  const void* res1; int res1_is16;   /* residual 1: fp32 planes, or 16-bit planes (`dtype`) if res1_is16 */
  int res1_planes_total; int res1_plane_off; float beta1;
  const float* res2; int res2_planes_total; int res2_plane_off; float beta2;  /* may alias out32 (in-place +=) */
  const float* res3; int res3_planes_total; int res3_plane_off; float beta3;
  /* dgrad helpers (backward of the same convs, weights packed with transpose_flip=1):
   *  - the first `lead_planes` output planes (gradient of the latent channels every conv sees) are accumulated,
   *    lead_acc[plane] += alpha*acc, and skip the rest of the epilogue; all other plane offsets in this struct
   *    are relative to the first non-lead plane;
   *  - planes >= tail_first_plane are multiplied by the LeakyReLU derivative taken from the saved activation
   *    `mask16` (act > 0 ? 1 : mask_slope) before the 16-bit store, and the 16-bit store is limited to them. */
  int lead_planes; float* lead_acc; int lead_planes_total;
  const void* mask16; int mask_planes_total; int mask_plane_off; float mask_slope; int tail_first_plane;
  /* outputs */
  void* out16; int out16_planes_total; int out16_plane_off;
  int out16_up2;             /* 1: out16 is [N][planes][2H][2W][8], each pixel replicated 2x2 */
  int out16_pixel_shuffle;   /* r (0 = off): out16 is [N][planes][rH][rW][8], conv channel
                                c*r*r+i*r+j goes to channel c at (r*y+i, r*x+j)               */
  float* out32; int out32_planes_total; int out32_plane_off;
  float* out_nchw; int out_nchw_c;
  /* tiling hints (0 = library default) */
  int tile_p;                /* tile pitch 32 or 64 (valid width = pitch-2)                   */
  int tile_mt;               /* 128-row M tiles per CTA tile: 1, 2 or 4                        */
  /* optional second weight image for the row-streaming kernel (esr_pack_conv3x3_weights_rows); when given and the
   * image is wide enough for 128-pixel strips the launch uses that kernel, otherwise the tile kernel above */
  const void* wpacked_rows;
  int rows_nbn;              /* n-block the row image was packed with (esr_conv3x3_rows_config) */
  int rows_mode;             /* 0: library decides by image width; 1: always the row kernel; -1: never */
} esr_conv3x3_args;

int esr_conv3x3_fwd(const esr_conv3x3_args* a, void* stream);
/* `count` launches in order on one stream from one host call: a generator forward is ~350 fused convs with fixed arguments
 * (cached buffers, packed weights), so the host side replays a recorded argument array instead of rebuilding every struct
 * (the per-call host work, not the GPU, bounds single-image inference).  Stops at the first failure and returns its status;
 * *failed_index (may be NULL) receives its position. */
int esr_conv3x3_fwd_batch(const esr_conv3x3_args* args, int count, void* stream, int* failed_index);

/* bytes of the packed weight image for (cin_planes, cout) with chunk size kcp (_ex: for a given esr_dtype; ESR_BF16X3 images hold
 * the three segments w_hi | w_hi | w_lo) */
size_t esr_conv3x3_packed_bytes(int cin_planes, int cout, int kcp, int* cout_pad_out);
size_t esr_conv3x3_packed_bytes_ex(int cin_planes, int cout, int kcp, int dtype, int* cout_pad_out);
/* OIHW fp32 [cout][cin][3][3] (device) -> packed tensor-core image (device).
 * Layout [n_block][chunk][tap][plane-in-chunk][cout-in-block][8 cin], zero padded.
 * `transpose_flip` = 1 packs the dgrad operand (I/O swapped, taps rotated 180 degrees).
 * `lead`: the first `lead` input channels (the latent z the reference concatenates IN FRONT of every
 * conv input, block.py:92,265,268 / architecture.py:287,300) occupy their own zero-padded plane group;
 * the remaining cin-lead channels start at the next plane boundary.  0 for plain convs. */
int esr_pack_conv3x3_weights(const float* w_oihw, int cout, int cin, int lead, int kcp, int dtype,
                             int transpose_flip, void* wpacked, float* bias_out,
                             const float* bias_in, void* stream);
/* Row-streaming variant of the same convolution (csrc/conv3x3_rows.cuh): the three vertical taps are merged into the
 * MMA's N dimension (N = 3 x n-block), the weights of one n-block stay resident in shared memory and a CTA marches
 * down a 128-pixel column strip.  rows_config reports the n-block (16/32/64; 0 = the weights do not fit, use the tile
 * kernel) and the size of the packed image [n_block][chunk][dx][plane-in-chunk (4)][ky*nbn + cout-in-block][8 cin];
 * for a dgrad operand pass cout = 8 * esr_conv3x3_cin_planes(cin, lead) and cin_planes = ceil(cout_fwd / 8). */
int esr_conv3x3_rows_config(int cin_planes, int cout, int* nbn_out, size_t* bytes_out);
int esr_conv3x3_rows_config_ex(int cin_planes, int cout, int dtype, int* nbn_out, size_t* bytes_out);
int esr_pack_conv3x3_weights_rows(const float* w_oihw, int cout, int cin, int lead, int dtype, int transpose_flip,
                                  int nbn, void* wpacked_rows, void* stream);
/* ------------------------------------------------------------------------------------------------
 * Weight / bias gradient of the same convolution (what autograd computes for nn.Conv2d inside
 * models/modules/block.py:129-146 when models/SRRaGAN_model.py:436-500 calls l_g_total.backward()):
 *   dw[co][ci][ky][kx] (+)= scale * sum_{n,y,x} gy[n][co][y][x] * x[n][ci][y+ky-1][x+kx-1]     (OIHW fp32)
 *   db[co]             (+)= scale * sum_{n,y,x} gy[n][co][y][x]
 * x: the conv's input activations (16-bit planes, the buffers the forward pass wrote; latent channels in their own
 * leading plane group as for esr_pack_conv3x3_weights), gy: gradient w.r.t. the conv output BEFORE the activation
 * (16-bit planes).  K = pixels on the tensor cores, vertical taps merged into M, horizontal taps into N
 * (csrc/conv3x3_wgrad.cuh).  `workspace` holds per-CTA partial sums (esr_conv3x3_wgrad_workspace bytes).
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  int n, h, w;
  int dtype;                 /* esr_dtype of x and gy */
  const void* x; int x_planes_total; int x_plane_off;
  const void* gy; int gy_planes_total; int gy_plane_off;
  int cout, cin, lead;       /* real channel counts of the weight tensor; `lead` latent input channels */
  float* dw;                 /* [cout][cin][3][3] fp32 or NULL */
  float* db;                 /* [cout] fp32 or NULL */
  float scale;
  int accumulate;            /* 1: add into dw/db (gradient accumulation), 0: overwrite */
  float* workspace; size_t workspace_bytes;
  /* input-channel slicing for wide convs (3*cin/8 groups must fit 5 M chunks, i.e. cin <= 208 per launch): this launch covers
   * the `cin` channels starting at `cin_off` of a weight tensor with `cin_total` input channels (0 = cin); x_plane_off
   * points at the first plane of the slice */
  int cin_total, cin_off;
} esr_conv3x3_wgrad_args;
size_t esr_conv3x3_wgrad_workspace(int cin_planes, int cout);
int esr_conv3x3_wgrad(const esr_conv3x3_wgrad_args* a, void* stream);

/* out[c] (+)= scale * sum over n, y, x of src[n][c][y][x] (NCHW fp32): the bias gradient of the generator's last conv
 * taken from dL/dG itself, before it is rounded to 16-bit planes */
int esr_sum_nchw(const float* src, int n, int c, int h, int w, float scale, int accumulate, float* out, void* stream);

/* planes consumed by a conv with `cin` input channels of which the first `lead` are latent */
int esr_conv3x3_cin_planes(int cin, int lead);

/* ------------------------------------------------------------------------------------------------
 * layout conversion at the boundary (reference tensors are NCHW fp32)
 * pack:   NCHW fp32 -> planes (16-bit and/or fp32); optional replicate padding of `pad` pixels per
 *         side (CEM.CEMnet.py:70-71 LR_padder/HR_padder, applied at :286-295 in eval mode).
 * unpack: planes -> NCHW fp32
 * ---------------------------------------------------------------------------------------------- */
int esr_pack_nchw(const float* src, int n, int c, int h, int w, int pad, int dtype,
                  void* dst16, float* dst32, int planes_total, int plane_off, void* stream);
int esr_unpack_planes16(const void* src16, int dtype, int n, int c, int h, int w,
                        int planes_total, int plane_off, float* dst, void* stream);
int esr_unpack_planes32(const float* src32, int n, int c, int h, int w,
                        int planes_total, int plane_off, float* dst, void* stream);

/* ------------------------------------------------------------------------------------------------
 * VGG feature extractor (models/modules/architecture.py:658-724, the perceptual loss of SRRaGAN_model.py:448-451).
 * Its 3x3 convolutions + ReLU are esr_conv3x3_fwd launches (LeakyReLU with slope 0); these are the remaining pieces:
 *   pack_nchw_affine : (x - mean) / std input normalisation (:719-720) fused into the layout conversion, dst = v*scale[c]+shift[c]
 *   maxpool2x2       : nn.MaxPool2d(2, 2) of torchvision's vgg19.features on 16-bit planes
 *   maxpool2x2_bwd   : its backward fused with the preceding ReLU's (first maximum in row-major order, as torch)
 * ---------------------------------------------------------------------------------------------- */
int esr_pack_nchw_affine(const float* src, int n, int c, int h, int w, const float* scale, const float* shift, int dtype,
                         void* dst16, int planes_total, int plane_off, void* stream);
int esr_maxpool2x2_planes16(const void* src, int dtype, int n, int planes, int h, int w, void* dst, void* stream);
int esr_maxpool2x2_bwd_planes16(const void* gout, const void* act, int dtype, int n, int planes, int h, int w, void* gin,
                                void* stream);

/* ------------------------------------------------------------------------------------------------
 * Consistency-Enforcing Module (CEM/CEMnet.py:254-311).  Filters are the separable factors of the
 * reference's numpy-designed kernels (CEMnet.py:22-33,186-206): ds_kernel = outer(kd_v, kd_h) and
 * inv_hTh = outer(ki_v, ki_h), computed on the host by the Python layer.  Every filter is passed as
 * `rank` separable terms  K[a][b] = sum_r kv[r*len + a] * kh[r*len + b]  (rank 1 for the reference's
 * bicubic kernels; more terms represent estimated, non-separable kernels exactly).
 *
 * esr_cem_down : DownscaleOP, CEMnet.py:273-275 (replicate-pad k/2, correlate with rot90(ds_kernel,2),
 *                keep phase `phase` of every s x s cell).  Optionally returns x_lr - Down(g) when
 *                `sub_from` is given (fused residual for the projection).
 * esr_cem_inv  : Conv_LR_with_Inv_hTh_OP, CEMnet.py:262-264 (replicate-pad, correlate).
 * esr_cem_up_add: out = g + Upscale_OP(f) (CEMnet.py:266-272,305-310), polyphase, with optional crop
 *                of `crop` pixels per side (HR_unpadder, CEMnet.py:72,311).  g may be NULL (pure Up).
 * All images NCHW fp32.
 * ---------------------------------------------------------------------------------------------- */
int esr_cem_down(const float* g, int n, int c, int hh, int wh, int s, int phase,
                 const float* kd_v, const float* kd_h, int kd_len, int rank,
                 const float* sub_from, float* out_lr, void* stream);
int esr_cem_inv(const float* e, int n, int c, int hl, int wl,
                const float* ki_v, const float* ki_h, int ki_len, int rank, float* out_lr,
                void* stream);
int esr_cem_up_add(const float* f, const float* g, int n, int c, int hl, int wl, int s, int phase,
                   const float* ku_v, const float* ku_h, int ku_len, int rank, int crop,
                   float* out_hr, void* stream);

/* latent map: replicate-pad `pad_hr` (eval mode, CEMnet.py:290-292) then bilinear align_corners=False
 * resize by 1/s (models/modules/architecture.py:284).  NCHW fp32 in/out. */
int esr_latent_downscale(const float* z_hr, int n, int c, int hh, int wh, int s, int pad_hr, float* out,
                         void* stream);

/* ------------------------------------------------------------------------------------------------
 * backward helpers (gradients of the operators above; Z_optimization.py:673-749 drives them through autograd)
 * ---------------------------------------------------------------------------------------------- */
/* adjoint of nearest x2 (+ optional LeakyReLU derivative from the saved hi-res activation) */
int esr_downsum2x_planes(const float* src32, int n, int planes, int h, int w, const void* act16_hi, float slope,
                         int dtype, float* dst32, void* dst16, void* stream);
/* out = a + b on fp32 planes (n_groups8 groups of 8 floats), optional 16-bit copy */
int esr_planes_add(const float* a32, const float* b32, size_t n_groups8, int dtype, float* out32, void* out16,
                   void* stream);
/* same with the groups of one image stated (a split-precision out16 holds hi and lo halves per image) */
int esr_planes_add_ex(const float* a32, const float* b32, size_t n_groups8, size_t per_image_groups8, int dtype, float* out32,
                      void* out16, void* stream);
/* adjoint of one 1-D pass of a clamp-addressed strided filter (the building block of the CEM backward):
 *   forward out[o] = sum_t k[t]*in[clamp(a_stride*o + t + c_off, 0, n_in-1)];  produces gin at logical positions
 *   m = m_stride*mi + m_phase, mi < n_store.  Tensor viewed as [imgs][outer][axis][inner].
 *   The stored gout covers logical o in [o_lo, o_lo+o_cnt) only (adjoint of the HR_unpadder crop: zeros elsewhere).
 *   sub_from (x-axis pass only): gin = crop_adjoint(sub_from, sub_crop) - result, i.e. g_G = g_out - Down^T(...). */
int esr_sep_adjoint_1d(const float* gout, int imgs, int outer, int inner, int n_out, int o_lo, int o_cnt, int n_in,
                       int n_store, int a_stride, int c_off, int m_stride, int m_phase, const float* taps, int len,
                       const float* sub_from, int sub_crop, float* gin, void* stream);
/* gradient of the latent map Z[n][c][hh][wh] from the HR-resolution (padded planes32) and LR-resolution latent
 * gradients accumulated by the dgrad convs (adjoint of replicate pad + bilinear 1/s resize) */
int esr_latent_grad(const float* gz_hr_planes32, const float* gz_lr_planes32, int n, int c, int hh, int wh, int s,
                    int pad_hr, float* dst_nchw, void* stream);

/* nearest x2 of 16-bit planes (models/modules/block.py:299-300), for callers that cannot fold it */
int esr_upsample2x_planes16(const void* src, int n, int planes, int h, int w, void* dst, void* stream);
/* adjoint of nn.PixelShuffle(2) (models/modules/block.py:287, the 'pixelshuffle' upsample_mode of RRDBNet) on 16-bit planes: dst channel
 * 4c + 2dy + dx at (y, x) = src channel c at (2y + dy, 2x + dx); src is [n][src_planes_total][2h][2w][8], dst [n][dst_planes_total][h][w][8],
 * `planes` source planes from src_plane_off land in 4 * planes destination planes from dst_plane_off.  Bit-exact permutation. */
int esr_pixel_unshuffle2_planes16(const void* src, int n, int planes, int h, int w, int src_planes_total, int src_plane_off, void* dst,
                                  int dst_planes_total, int dst_plane_off, void* stream);

/* Re-packing after an optimizer step: every conv of a network in ONE host call (the same two kernels per item as
 * esr_pack_conv3x3_weights / esr_pack_conv3x3_weights_rows; wpacked_rows == NULL skips the row-kernel image).  The packed
 * buffers are caller-owned and can be re-used in place across steps, so recorded launch arguments stay valid. */
typedef struct {
  const float* w_oihw; int cout, cin, lead, kcp, dtype, transpose_flip;
  void* wpacked; float* bias_out; const float* bias_in;
  void* wpacked_rows; int rows_nbn;
} esr_pack_item;
size_t esr_pack_batch_scratch_bytes(int count);
/* scratch: caller-owned device memory of esr_pack_batch_scratch_bytes(count) bytes for the job table of the single fused launch
 * (NULL, or ESR_PACK_FUSED=0 in the environment: one launch pair per conv instead) */
int esr_pack_conv3x3_weights_batch(const esr_pack_item* items, int count, void* scratch, size_t scratch_bytes, void* stream,
                                   int* failed_index);

/* ------------------------------------------------------------------------------------------------
 * Discriminator_VGG_128 (models/modules/architecture.py:446-508; D step SRRaGAN_model.py:342-395, G-side GAN term :466-477).
 * Its 3x3 convolutions are esr_conv3x3_fwd launches; a 4x4 stride-2 pad-1 convolution (block.py:129-146 with kernel_size=4,
 * stride=2) is the 3x3 convolution of the 2x2 space-to-depth image with re-indexed weights
 *     W'[o][(py*2+px)*C + c][ty][tx] = W[o][c][2*ty+py-1][2*tx+px-1]   (zero where the 4x4 index falls outside 0..3),
 * so forward, input-gradient and weight-gradient launches are the same tensor-core kernels.  The remaining pieces:
 *   esr_bn_stats      nn.BatchNorm2d batch statistics of the conv output (fp32 planes) -> save_mean / save_invstd, the affine
 *                     scale = gamma*invstd, shift = beta - mean*scale, and the running-statistics update (unbiased variance);
 *                     train = 0 takes the statistics from running_mean / running_var instead (module.eval()).
 *   esr_bn_lrelu_fwd  LeakyReLU(y*scale[c] + shift[c]) -> 16-bit planes, plain or space-to-depth ([n][4*planes][h/2][w/2][8],
 *                     plane (py*2+px)*planes + g) for a following stride-2 conv, and / or NCHW fp32 (classifier input, :505).
 *   esr_bn_lrelu_bwd  gradient of the conv output from the gradient `g` of the activation: dgamma / dbeta (scaled by gscale,
 *                     optional accumulate), c1 = dbeta/M, c2 = dgamma/M scratch, gy16 = scale*(g*lrelu' - c1 - xhat*c2).
 *                     g_layout: 0 fp32 planes, 1 fp32 planes of the space-to-depth image (output of a stride-2 conv's
 *                     transposed launch), 2 NCHW fp32.  has_bn = 0: activation only (scale=1, shift=0, mean=0, invstd=1,
 *                     c1=c2=0 arrays supplied by the caller; no reductions run).
 *   esr_linear_fwd / esr_linear_bwd  nn.Linear (+ LeakyReLU) in fp32; bwd masks g with the layer's own activation `act`.
 * ---------------------------------------------------------------------------------------------- */
size_t esr_bn_workspace_bytes(int planes);
int esr_bn_stats(const float* y32, int n, int planes, int h, int w, int c, const float* gamma, const float* beta, float eps,
                 float momentum, int train, float* running_mean, float* running_var, float* save_mean, float* save_invstd,
                 float* scale, float* shift, float* workspace, size_t workspace_bytes, void* stream);
int esr_bn_lrelu_fwd(const float* y32, int n, int planes, int h, int w, int c, const float* scale, const float* shift, float slope,
                     int dtype, void* dst16, int space_to_depth, float* dst_nchw, void* stream);
int esr_space_to_depth_planes16(const void* src, int n, int planes, int h, int w, void* dst, void* stream);
int esr_bn_lrelu_bwd(const float* g, int g_layout, const float* y32, int n, int planes, int h, int w, int c, const float* scale,
                     const float* shift, const float* save_mean, const float* save_invstd, float slope, int has_bn, int train,
                     float gscale, int accumulate, float* dgamma, float* dbeta, float* c1, float* c2, int dtype, void* gy16,
                     float* workspace, size_t workspace_bytes, void* stream);
int esr_linear_fwd(const float* x, const float* weight, const float* bias, int batch, int in_features, int out_features, int lrelu,
                   float slope, float* out, void* stream);
int esr_linear_bwd(const float* g, const float* act, float slope, const float* x, const float* weight, int batch, int in_features,
                   int out_features, float gscale, int accumulate, float* gx, float* dweight, float* dbias, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Latent-control loss L_struct (models/modules/loss.py:27-209, FilterLoss with structure-tensor latent channels; used at
 * SRRaGAN_model.py:455-460).  The reference applies two 2x2 depth-wise finite-difference filters (:51-62), squares /
 * multiplies their outputs and averages per image (:140-147).  One pass here:
 *   out[n] = (mean dx^2, mean dy^2, mean dx*dy),  dx = x[i][j+1]-x[i][j], dy = x[i+1][j]-x[i][j], over c x (h-1) x (w-1)
 *   bwd: gradient of sum_k g[n][k]*out[n][k] with respect to the image (NCHW fp32, same shape).
 * ---------------------------------------------------------------------------------------------- */
size_t esr_structure_tensor_workspace_bytes(int n);
int esr_structure_tensor_fwd(const float* img, int n, int c, int h, int w, float* out, float* workspace, size_t workspace_bytes,
                             void* stream);
int esr_structure_tensor_bwd(const float* img, const float* g, int n, int c, int h, int w, float* grad_img, void* stream);

/* Second-order pass of the gradient penalty (WGAN-GP: GradientPenaltyLoss, models/modules/loss.py:260-279, differentiates THROUGH the
 * critic's input gradient; models/SRRaGAN_model.py:362-371).  d L_gp / d theta = grad_theta of the critic's directional derivative along
 * v = dL_gp/dg: a forward-mode tangent through the layers, then an ordinary backward over the (primal, tangent) pair.  Convolutions and
 * linear layers reuse the launches above; these two entry points are the BatchNorm2d (batch statistics) + LeakyReLU part:
 *   esr_bn_tangent_fwd : w = LeakyReLU'(z) * gamma r (t - mean(t) - yh mean(yh t)) from the conv's tangent t (fp32 planes) and the saved
 *                        primal output y32 / statistics; also returns c1 = mean(t), c2 = mean(yh t) per channel.  has_bn = 0: w = mask * t.
 *   esr_bn_double_bwd  : adjoints of the conv's output (yb16) and of the conv's tangent (tb16) from the adjoints zb / wb of the primal and
 *                        tangent activations (gradient layouts as esr_bn_lrelu_bwd; either may be NULL = zero), plus dgamma / dbeta.
 *                        coef: [5*c] fp32 scratch for the per-channel sums. */
size_t esr_bn_dbl_workspace_bytes(int planes);
int esr_bn_tangent_fwd(const float* t32, const float* y32, int n, int planes, int h, int w, int c, const float* scale, const float* shift,
                       const float* save_mean, const float* save_invstd, float slope, int has_bn, float* c1, float* c2, int dtype, void* dst16,
                       int space_to_depth, float* dst_nchw, float* workspace, size_t workspace_bytes, void* stream);
int esr_bn_double_bwd(const float* zb, const float* wb, int g_layout, const float* y32, const float* t32, int n, int planes, int h, int w, int c,
                      const float* scale, const float* shift, const float* save_mean, const float* save_invstd, const float* c1, const float* c2,
                      float slope, int has_bn, float gscale, int accumulate, float* dgamma, float* dbeta, float* coef, int dtype, void* tb16,
                      void* yb16, float* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * The training step outside the networks.
 * ---------------------------------------------------------------------------------------------- */
/* Fused multi-tensor Adam: ONE launch per optimizer step (models/SRRaGAN_model.py:182,188 build torch.optim.Adam for G and D; :403,499
 * step them).  Exactly torch.optim.Adam's arithmetic (amsgrad off): g' = grad_scale*g + wd*p; m += (g'-m)(1-b1); v = b2 v + (1-b2) g'^2;
 * p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps).  grad_scale folds the 1/world of a summed all-reduce.  `tensors_host` is a HOST
 * array copied into `scratch` (device, esr_adam_scratch_bytes); pass NULL to reuse the table a previous call uploaded. */
typedef struct { float* p; const float* g; float* m; float* v; unsigned long long n; } esr_adam_tensor;
size_t esr_adam_scratch_bytes(int count);
int esr_adam_multi(const esr_adam_tensor* tensors_host, int count, void* scratch, size_t scratch_bytes, float lr, float beta1, float beta2,
                   float eps, float weight_decay, int step, float grad_scale, void* stream);
/* nn.L1Loss (mean) of the pixel and VGG-feature criteria (models/SRRaGAN_model.py:98,129,434,448-451): out_mean[0] = mean |a - b|
 * (two-stage, deterministic), and its gradient ga = sign(a-b) * gout[0] / n, gb = -ga (either may be NULL). */
size_t esr_l1_workspace_bytes(void);
int esr_l1_reduce(const float* a, const float* b, size_t n, float* workspace, size_t workspace_bytes, float* out_mean, void* stream);
int esr_l1_grad(const float* a, const float* b, size_t n, const float* gout, float* ga, float* gb, void* stream);
/* Relativistic average GAN terms on logits (models/SRRaGAN_model.py:353-354 D step, :475-476 G step; GANLoss 'vanilla' =
 * BCEWithLogitsLoss, models/modules/loss.py:212-246):  la = mean_i bce(a_i - mean(b), target_a), lb = mean_i bce(b_i - mean(a), target_b).
 * out4 = [sum_i bce_a, sum_i bce_b, S_a, S_b] over the n LOCAL samples (S = sum of the bce derivatives, kept per sample in ea / eb);
 * sums_global = [sum a, sum b] over the global batch of n_global samples (NULL: local).  The backward gives d(ga_up*la + gb_up*lb)
 * with respect to a and b (either output may be NULL = that input is detached), S_global = [S_a, S_b] over the global batch. */
int esr_bce_rel_loss(const float* a, const float* b, int n, const float* sums_global, float n_global, float target_a, float target_b,
                     float* out4, float* ea, float* eb, void* stream);
int esr_bce_rel_loss_bwd(const float* ea, const float* eb, int n, const float* S_global, float n_global, const float* ga_up,
                         const float* gb_up, float* ga, float* gb, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Kernel-density soft histogram / dictionary distance (SoftHistogramLoss.ComputeSoftHistogram, Z_optimization.py:180-216: the `hist` /
 * `dict` objectives of the GUI's imprinting tools).  x: [dims][n_samples] fp64 (grey levels, or dims = patch_size^2 patch entries),
 * bins: [dims][n_bins] fp64.  E[p][b] = exp(-(1/dims) sum_d (cyclic|x - bins| + eps)^2 / temperature).
 *   dictionary == 0: out[n_bins]    = mean_p E[p][b]
 *   dictionary == 1: out[n_samples] = -log(mean_b E[p][b]),  sum_e[n_samples] = sum_b E[p][b] (kept for the backward)
 * The backward returns d/dx of  sum_b g_hist[b] out[b]  (histogram)  or  sum_p g_out[p] out[p]  (dictionary).  The [dims, P, B]
 * distance tensor the reference materialises never exists.
 * ---------------------------------------------------------------------------------------------- */
int esr_soft_hist_fwd(const double* x, int dims, int n_samples, const double* bins, int n_bins, double vmax, double eps, double temperature,
                      int dictionary, double* out, double* sum_e, void* stream);
int esr_soft_hist_bwd(const double* x, int dims, int n_samples, const double* bins, int n_bins, double vmax, double eps, double temperature,
                      const double* g_hist, const double* g_out, const double* sum_e, double* dx, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ESR_B200_H */
