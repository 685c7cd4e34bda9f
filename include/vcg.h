/*
 * vcg.h -- C ABI of libvcg_b200.so: the sm_100a kernels behind the VAE-CycleGAN training step.
 *
 * The reference (Baverne/VAE-CYCLEGAN-Implementation) is pure Python/PyTorch and has no FFI of
 * its own; every entry point below replaces an ATen call that the reference issues implicitly
 * from Networks.py / Losses.py / torch.optim.Adam.  The reference call site each one stands
 * in for is cited (file:line relative to the reference checkout; "torch/" = site-packages/torch).
 *
 * Conventions
 *   - plain C types only; every function returns 0 on success, <0 (VCG_E_*) on failure, never
 *     throws/aborts; the message is available through vcg_last_error() (thread local).
 *   - the caller owns every buffer (device pointers from tensor.data_ptr()); the library never
 *     allocates, frees or retains device memory.
 *   - all work is enqueued on the cudaStream_t passed as `void* stream`; no hidden sync, no
 *     default-stream use; every call is CUDA-graph capturable.
 *   - unsupported shape / dtype => VCG_E_UNSUPPORTED.  There is no CPU fallback.
 *
 * Activation layout in HBM: NHWC, element type VCG_F32 (parity mode) or VCG_BF16 (tensor-core
 * mode).  A convolution reads a *materialised* halo: its input buffer is [N, Hp, Wp, C] with the
 * reflect (forward) or zero (data-gradient) border already written by vcg_xform_fwd / bwd, so
 * every convolution is a "valid" stride-1 correlation (stride-2 4x4 convolutions are expressed
 * as 2x2 stride-1 ones over a space-to-depth input; PixelUnshuffle/PixelShuffle are folded into
 * the addressing of vcg_xform_*).
 */
#ifndef VCG_H_
#define VCG_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VCG_ABI_VERSION 2
#define VCG_API __attribute__((visibility("default")))

enum { VCG_OK = 0, VCG_E_INVALID = -1, VCG_E_UNSUPPORTED = -2, VCG_E_CUDA = -3, VCG_E_DRIVER = -4 };
enum { VCG_F32 = 0, VCG_BF16 = 1 };
enum { VCG_ACT_NONE = 0, VCG_ACT_RELU = 1, VCG_ACT_LEAKY = 2,           /* LeakyReLU slope 0.2 */
       VCG_ACT_TANH = 3, VCG_ACT_SIGMOID = 4 };                          /* CaSb's other choices, Networks.py:63-71 */
/* vcg_xform_* addressing modes (destination domain of the forward transform) */
enum { VCG_MODE_PLAIN = 0,      /* dst[h,w,c]            = src[h,w,c]                              */
       VCG_MODE_SHUFFLE = 1,    /* dst[2h+i,2w+j,c]      = src[h,w,c*4+i*2+j]   (nn.PixelShuffle)   */
       VCG_MODE_UNSHUFFLE = 2,  /* dst[h,w,(i*2+j)*C+c]  = src[2h+i,2w+j,c]     (nn.PixelUnshuffle) */
       VCG_MODE_PAD_S2D = 3 };  /* reflect-pad first, then space-to-depth (stride-2 conv input)    */
/* weight (un)packing channel maps */
enum { VCG_WMAP_PLAIN = 0, VCG_WMAP_UNSHUFFLE = 1, VCG_WMAP_S2D = 2 };

VCG_API int vcg_version(void);
VCG_API const char* vcg_last_error(void);
/* number of kernels launched by this library since load (bench.py's gpu_launches claim) */
VCG_API long long vcg_launch_count(void);
/* Grid budget of the persistent tensor-core kernels (process-wide, read at launch time): at most `sms` CTAs per
 * launch; 0 = every SM.  The host side runs two independent network passes (e.g. G(x) and F(y) of
 * Networks.py:1909-1914) on two streams and gives each half of the machine, so that small-batch layers whose
 * tile count cannot fill 148 SMs run side by side instead of one after the other.                              */
VCG_API int vcg_set_sm_budget(int32_t sms);
/* L2 read-ahead of the streaming backward transforms (process-wide, read at launch time): every block issues
 * cp.async.bulk.prefetch.L2 for the inputs of the pixels it will touch `chunks` iterations ahead (0 = off).  Their
 * demand loads are register-limited (8 x 16 B in flight per thread); the prefetch keeps HBM -> L2 streaming ahead.  */
VCG_API int vcg_set_l2_prefetch(int32_t chunks);

/* Multi-tensor variants: ONE launch packs (both layouts) or unpacks every filter of a network.
 * The caller fills d / pointers / accumulate for each job, runs vcg_wjob_plan on the HOST array (it assigns
 * the tile ranges), copies the array to the device and passes the device pointer.  Packed buffers must be
 * zero-initialised once (padding rows / columns are never written).  vcg_wunpack_multi re-zeroes the
 * fp32 accumulators it reads, so the weight-gradient GEMMs can keep accumulating into them.            */
typedef struct vcg_wjob {
  int32_t co, ci, kh, kw, wmap, c_phys, co_phys, rows_pad, pkh, pkw, kwc_pad, transpose_flip; /* as vcg_wpack_desc minus dtype */
  float* oihw;               /* pack: fp32 OIHW master (read); unpack: OIHW gradient (written) */
  void* packed;              /* forward layout [rows_pad][pkh][kwc_pad] */
  void* packed_t;            /* pack only: data-gradient layout [rup16(c_phys)][pkh][t_kwc_pad] or NULL */
  int32_t t_kwc_pad;
  int32_t accumulate;        /* unpack: grad += */
  int32_t co_t, tiles_ci, tile0, ntiles;   /* filled by vcg_wjob_plan */
  int32_t vec, reserved;                   /* filled by vcg_wjob_plan: 16-byte vector path (3x3, full 32x32 tiles) */
} vcg_wjob;
VCG_API int vcg_wjob_plan(vcg_wjob* jobs_host, int32_t njobs, int32_t* total_tiles);
VCG_API int vcg_wpack_multi(int32_t dtype, const vcg_wjob* jobs_dev, int32_t njobs, int32_t total_tiles, void* stream);
VCG_API int vcg_wunpack_multi(const vcg_wjob* jobs_dev, int32_t njobs, int32_t total_tiles, void* stream);
/* dst[i] += src[i]; src[i] = 0 for every job (bias-gradient accumulators -> .grad); jobs: device array */
typedef struct vcg_vecjob { float* src; float* dst; int32_t n, pad; } vcg_vecjob;
VCG_API int vcg_vecflush_multi(const vcg_vecjob* jobs_dev, int32_t njobs, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* Convolution as implicit GEMM.  Replaces F.pad(reflect)+F.conv2d issued by every
 * nn.Conv2d(padding_mode='reflect') (Networks.py:60,87,101,104,122,136,145;
 * torch/nn/modules/conv.py:534-550) and the autograd convolution_backward behind it.        */
typedef struct vcg_conv_desc {
  int32_t dtype;      /* VCG_F32: SIMT FFMA kernels; VCG_BF16: tcgen05/TMEM kernels, fp32 accumulate */
  int32_t n;          /* images */
  int32_t hp, wp;     /* input buffer spatial dims (halo included) */
  int32_t c;          /* physical channels per input pixel (multiple of 8) */
  int32_t kh, kw;     /* filter taps (stride 1) */
  int32_t kwc_pad;    /* packed filter row length per kh: roundup(kw*c, 64) */
  int32_t cout;       /* logical output channels (bias/activation/stats bound); stores are rounded up to 8 */
  int32_t cout_pad;   /* packed filter rows (multiple of 16, >= cout) */
  int32_t out_c;      /* physical channels per output pixel */
  int32_t act;        /* VCG_ACT_* applied after bias, before stats */
  int32_t stats;      /* !=0: accumulate per-(n,cout) sum / sum-of-squares of the stored values */
  int32_t flat;       /* vcg_conv_wgrad: VCG_WGRAD_ALLOW_SIMT or 0.  vcg_conv_fwd:
                         !=0: tile the output in flattened input-pitch order (data-gradient); 2: the caller also
                         guarantees that the outer (kh-1, kw-1) border of x is zero, so taps that only see the
                         border may be skipped (interior + ring schedule of conv_tc2.cu) */
  int32_t out_f32;    /* !=0: store y as fp32 even when dtype is VCG_BF16 (final image layer) */
} vcg_conv_desc;

/* y[n,h,w,co] = act(bias[co] + sum_{kh,j} x[n,h+kh,w*c + j] * w[co,kh,j]),  h<hp-kh+1, w<wp-kw+1.
 * x: [n,hp,wp,c]; w: [cout_pad,kh,kwc_pad] (dtype); bias: fp32[cout] or NULL; y: [n,ho,wo,out_c];
 * stats: fp32 [n,cout,2] accumulators (+=) or NULL.  The data-gradient pass is the same call
 * with the flipped/transposed packed filter and a zero-haloed dy as x.                       */
VCG_API int vcg_conv_fwd(const vcg_conv_desc* d, const void* x, const void* w, const float* bias,
                 void* y, float* stats, void* stream);

/* desc->flat bit for vcg_conv_wgrad: maps that no tensor-core kernel can tile (below 8x8, i.e. inputs smaller than
 * 256x256) may run on the ~50x slower SIMT kernel; without the bit such a shape is VCG_E_UNSUPPORTED.          */
enum { VCG_WGRAD_ALLOW_SIMT = 4 };
/* dw[co,kh,j] += sum_{n,h,w} dy[n,h+kh-1.., ...] ...: weight gradient, fp32 packed layout
 * [cout_pad,kh,kwc_pad], ACCUMULATED (split-K reductions use red.global.add).
 * x: the forward input [n,hp,wp,c]; dy: [n, ho+2*dy_halo, wo+2*dy_halo, dy_c] zero-haloed.      */
VCG_API int vcg_conv_wgrad(const vcg_conv_desc* d, const void* x, const void* dy, int32_t dy_halo,
                   int32_t dy_c, float* dw, void* stream);

/* Filter packing: reference OIHW fp32 master weight -> packed kernel layout, and back for grads.
 * transpose_flip=0: rows=co, K=(kh,kw,phys cin) (forward); 1: rows=phys cin, K=(flipped kh,kw,co)
 * (data-gradient).  wmap says how physical input channels map to (ci, sub-tap).              */
typedef struct vcg_wpack_desc {
  int32_t dtype;            /* element type of the packed filter */
  int32_t co, ci, kh, kw;   /* reference OIHW dims */
  int32_t wmap;             /* VCG_WMAP_* */
  int32_t c_phys;           /* physical input channels of the conv as executed (>= mapped channels) */
  int32_t co_phys;          /* physical output channels (data-gradient K extent) */
  int32_t rows_pad;         /* packed rows */
  int32_t pkh, pkw;         /* taps as executed (kh/2, kw/2 for S2D) */
  int32_t kwc_pad;          /* packed row length per tap row */
  int32_t transpose_flip;
} vcg_wpack_desc;
VCG_API int vcg_wpack(const vcg_wpack_desc* d, const float* w_oihw, void* packed, void* stream);
/* grad_oihw (=|+=) unpack(dw_packed fp32 forward layout) */
VCG_API int vcg_wunpack_grad(const vcg_wpack_desc* d, const float* dw_packed, float* grad_oihw,
                     int32_t accumulate, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* InstanceNorm statistics (nn.InstanceNorm2d, Networks.py:61,88,102,105,123 -> native_batch_norm) */
/* mean/rstd[n,c] from a dense NHWC tensor (fp64 accumulation; parity mode and tests).
 * mean_rstd must have room for n*c*2 floats FOLLOWED BY n*c*2 doubles of scratch (3x the floats). */
VCG_API int vcg_in_stats(int32_t dtype, const void* y, int32_t n, int32_t hw, int32_t c, int32_t c_pitch,
                 float* mean_rstd, void* stream);
/* mean/rstd[n,c] from the (sum,sumsq) accumulators a conv epilogue produced; eps=1e-5, biased var */
VCG_API int vcg_in_finalize(const float* sums, int32_t nc, int32_t hw, float* mean_rstd, void* stream);

/* The memory-bound pass between two convolutions: normalise -> activation -> (+residual) ->
 * pixel (un)shuffle / space-to-depth addressing -> reflect halo -> cast.  Replaces
 * instance_norm, ReLU/LeakyReLU, pixel_(un)shuffle, reflection_pad2d and the residual add
 * (Networks.py:76-81, 91-96, 108-116, 126-131).                                               */
typedef struct vcg_xform_desc {
  int32_t dtype;
  int32_t n, h, w, c;        /* source logical dims */
  int32_t src_c;             /* source physical channel pitch */
  int32_t norm;              /* use mean_rstd[n,c] */
  int32_t act;               /* VCG_ACT_* after the norm */
  int32_t mode;              /* VCG_MODE_* */
  int32_t pad;               /* reflect halo width (destination domain; PAD_S2D: source domain) */
  int32_t dst_c;             /* destination physical channel pitch (extra channels zero-filled) */
  int32_t res_hp, res_wp, res_c, res_off;  /* residual: [n,res_hp,res_wp,res_c] read at (+off,+off) */
  int32_t stats_hw;          /* >0: mean_rstd holds the RAW {sum, sum of squares} pairs of vcg_conv_fwd's statistics
                                output over stats_hw pixels; mean / rstd (eps 1e-5, biased variance) are derived on load */
} vcg_xform_desc;
VCG_API int vcg_xform_fwd(const vcg_xform_desc* d, const void* src, const float* mean_rstd,
                  const void* residual, void* dst, void* stream);

/* Backward of the same pass, phase 1: g = sum_k gather(dxp_k) (halo folded, shuffle inverted),
 * then the activation derivative; writes g into the interior of the zero-haloed dy buffer
 * [n,h+2*dy_halo,w+2*dy_halo,dy_c] and accumulates sum(g), sum(g*zhat) per (n,c) for phase 2.
 * Without norm it also applies the pre-norm activation mask (y>0) and accumulates dbias.       */
typedef struct vcg_gsrc {      /* one gradient source = padded-input gradient of a consumer conv */
  const void* dxp;             /* [n, *, *, c_pitch] in the consumer's destination domain */
  int32_t mode, pad, c_pitch;
  int32_t folded;              /* 1: vcg_fold_halo already added the reflect halo into the interior */
} vcg_gsrc;
typedef struct vcg_xbwd_desc {
  int32_t dtype;
  int32_t n, h, w, c;
  int32_t y_c;               /* physical pitch of the saved conv output y */
  int32_t norm, act;         /* as in the forward desc (act = post-norm activation) */
  int32_t pre_act;           /* activation fused in the conv epilogue (before the norm) */
  int32_t dy_halo, dy_c;     /* destination geometry */
  int32_t nsrc;              /* 1..3 */
  int32_t stats_hw;          /* as in vcg_xform_desc */
  int32_t clear_halo;        /* gather only: also zero the halo ring of dy (saves a separate vcg_zero_halo launch) */
} vcg_xbwd_desc;
VCG_API int vcg_xform_bwd_gather(const vcg_xbwd_desc* d, const vcg_gsrc* srcs, const void* y,
                         const float* mean_rstd, void* dy, float* gsums /*[n,c,2]*/,
                         float* dbias /*[c] or NULL*/, void* stream);
/* Fold the reflect halo of a consumer's padded-input gradient into its interior, IN PLACE (adjoint of the
 * reflect padding, F.pad(mode='reflect') backward): afterwards every activation pixel reads exactly one
 * position of dxp.  (h, w, c) are the ACTIVATION dims as in vcg_xbwd_desc; mode/pad/c_pitch as in vcg_gsrc.
 * Only the ring of interior pixels within `pad` of the border is touched.                                   */
VCG_API int vcg_fold_halo(int32_t dtype, void* dxp, int32_t n, int32_t h, int32_t w, int32_t c, int32_t mode,
                  int32_t pad, int32_t c_pitch, void* stream);
/* phase 2 (norm only), in place on dy: v = rstd*(g - mean(g) - zhat*mean(g*zhat)) * pre_act'(y) */
VCG_API int vcg_xform_bwd_norm(const vcg_xbwd_desc* d, const void* y, const float* mean_rstd,
                       const float* gsums, void* dy, float* dbias, void* stream);

/* NCHW fp32 <-> NHWC(dtype) with channel pitch/offset: the API boundary of a network */
VCG_API int vcg_pack_nchw(int32_t dtype, const float* src, int32_t n, int32_t c, int32_t h, int32_t w,
                  void* dst, int32_t dst_c, int32_t dst_halo, void* stream);  /* zero halo + pad channels */
VCG_API int vcg_unpack_nchw(int32_t dtype, const void* src, int32_t src_c, int32_t n, int32_t c,
                    int32_t h, int32_t w, float* dst, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* VAE reparameterisation + KL (Networks.py:219-227, Losses.py:115-121).                       */
/* z = mu + eps*exp(0.5*clamp(lv,-10,10)); logvar_out = clamp(lv); kl_sum += sum(1+lv-mu^2-e^lv).
 * mu / lv are the fp32 NHWC outputs of the mu and logvar convolutions (kept in fp32 in both modes: logvar
 * reaches +-10 and feeds exp()); dtype is the element type of z / dz / dmu / dlv.                          */
VCG_API int vcg_reparam_fwd(int32_t dtype, const float* mu, int32_t mu_pitch, const float* lv, int32_t lv_pitch,
                    const float* eps /*NCHW fp32*/, int32_t n, int32_t hw, int32_t c,
                    void* z /*dense NHWC c*/, float* mu_out /*NCHW*/, float* lv_out /*NCHW*/,
                    float* kl_sum, void* stream);
/* dmu = dz + gmu_ext + kl_scale*mu ; dlv = mask*(dz*eps*0.5*std + glv_ext - 0.5*kl_scale*(1-e^lv)) */
VCG_API int vcg_reparam_bwd(int32_t dtype, const float* mu, int32_t mu_pitch, const float* lv, int32_t lv_pitch,
                    const float* eps, const void* dz, int32_t dz_pitch,
                    const float* gmu_ext /*NCHW or NULL*/, const float* glv_ext /*NCHW or NULL*/,
                    float kl_scale, int32_t n, int32_t hw, int32_t c,
                    void* dmu, int32_t dmu_pitch, void* dlv, int32_t dlv_pitch, void* stream);

/* Fused losses on fp32 tensors (Losses.py:14-121).  Each writes the *sum* into out[0] (+=) and,
 * when grad != NULL, grad = scale * d(sum)/d(a).                                              */
VCG_API int vcg_l1_fwd_bwd(const float* a, const float* b, int64_t numel, float scale, float* out_sum,
                   float* grad_a, void* stream);                      /* sum |a-b|, scale*sign(a-b) */
VCG_API int vcg_mse_const_fwd_bwd(const float* d, int64_t numel, float target, float scale,
                          float* out_sum, float* grad_d, void* stream); /* sum (d-t)^2, scale*2(d-t) */
VCG_API int vcg_kl_fwd_bwd(const float* mu, const float* lv, int64_t numel, float scale, float* out_sum,
                   float* gmu, float* glv, void* stream);

/* Spectral-normalised 512->1 16x16 head (Networks.py:248,267-269;
 * torch/nn/utils/spectral_norm.py:92-114): score[n] = bias + <x[n], w>/||w||.                   */
VCG_API int vcg_dhead_fwd(int32_t dtype, const void* x /*[n,k] NHWC-flattened*/, const float* w_khwc,
                  const float* bias, int32_t n, int32_t k, float* score, float* wnorm2 /*[1]*/,
                  void* stream);
/* dx[n,:] = gs[n]*w/||w||; dw += (G - (G.w_hat) w_hat)/||w|| with G = sum_n gs[n] x[n]; dbias += sum gs */
VCG_API int vcg_dhead_bwd(int32_t dtype, const void* x, const float* w_khwc, const float* wnorm2,
                  const float* gscore, int32_t n, int32_t k, void* dx, float* dw, float* dbias,
                  float* scratch /*[k+1]*/, int32_t dw_c /*0: dw in (h,w,c) order; >0: OIHW gradient of [1,dw_c,kh,kw]*/,
                  void* stream);
/* One launch per discriminator and step: the power iteration every training forward of the reference runs
 * (torch/nn/utils/spectral_norm.py:92-114: v = normalize(W^T u), u = normalize(W v), in place, when do_iter),
 * the (h, w, c)-ordered filter copy the two calls above read, and aux = {sigma = u . (W v), |W|} (eval mode divides by
 * the sigma of the stale u, v: spectral_norm.py:125-130).  w_oihw: [1, c, kh, kw] fp32, hw = kh*kw.           */
VCG_API int vcg_dhead_prepare(const float* w_oihw, int32_t c, int32_t hw, float* u, float* v, float* w_hwc /*or NULL*/,
                      float* aux /*[2] or NULL*/, double* scratch /*[3], caller-owned*/, int32_t do_iter, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* Multi-tensor Adam (torch/optim/adam.py:457-547 as called from Networks.py:312,894,1032-1033,
 * 1669-1676,1928-1935): one launch over a table of tensors.                                    */
typedef struct vcg_adam_chunk {  /* one <=65536-element slice of one tensor */
  float* p; const float* g; float* m; float* v; int32_t numel;
} vcg_adam_chunk;
/* state_dev: 4 device floats {step, lr/bias_corr1, sqrt(bias_corr2), -}: the step counter is advanced
 * and the bias corrections are recomputed ON THE DEVICE by each call (CUDA-graph replayable).          */
enum { VCG_ADAM_TICK = 1,        /* advance the step counter / bias corrections (once per optimiser step) */
       VCG_ADAM_GRAD_BF16 = 2,   /* chunk.g points to bfloat16 gradients (all-reduced wire buffer) */
       VCG_ADAM_ZERO_GRAD = 4 }; /* zero the fp32 gradient after reading it (next zero_grad() is free) */
VCG_API int vcg_adam_multi(const vcg_adam_chunk* chunks_dev, int32_t nchunks, float* state_dev, float lr,
                           float beta1, float beta2, float eps, float grad_scale, int32_t flags, void* stream);
/* dst(bf16)[i] = src(fp32)[i] (round to nearest even), optionally src[i] = 0: packs a slice of the flat gradient
 * buffer into the bf16 wire buffer of the data-parallel all-reduce (the reference has no collective at all;
 * this is the exchange step SURVEY.md 8e adds).                                                             */
VCG_API int vcg_cast_bf16(float* src, void* dst, int64_t n, int32_t zero_src, void* stream);

/* test hook: encode a bf16 SWIZZLE_128B tiled TMA descriptor only (no launch) */
VCG_API int vcg_probe_tmap(const void* base, int32_t rank, const uint64_t* dims,
                           const uint64_t* strides_bytes, const uint32_t* box);

/* zero only the halo ring of an NHWC buffer [n, h+2*halo, w+2*halo, c] (dY buffers: the interior is
 * fully overwritten by vcg_xform_bwd_gather)                                                        */
VCG_API int vcg_zero_halo(int32_t dtype, void* buf, int32_t n, int32_t h, int32_t w, int32_t c,
                          int32_t halo, void* stream);

/* utility: fill fp32 zeros (graph-capturable memset wrapper) */
VCG_API int vcg_zero(void* p, size_t bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VCG_H_ */
