/*
 * hsidm.h - C ABI of the B200-native HSI-DMGASR inference hot path.
 *
 * The reference (handsomewzy/HSI-DMGASR) is pure Python/PyTorch and has no FFI of its own;
 * the drop-in boundary is its Python object API, and every entry point below is what the Python mirror of
 * that API (hsi_dmgasr_b200/*.py, bound through ctypes) calls.  Each declaration cites the reference
 * interface it replaces as file:line relative to the reference tree.
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / C++ types cross this boundary
 *   - every tensor argument is a DEVICE pointer to dense fp32 in the reference's own layout (NCHW) unless
 *     stated otherwise; the library never takes ownership of caller tensors and never mutates inputs
 *   - all work is enqueued on the caller-supplied stream; no entry point synchronises the device except
 *     *_create / *_commit / *_destroy and workspace growth on the first call at a new shape
 *   - return value: 0 on success, negative hsidm_status on failure; hsidm_last_error() returns the
 *     thread-local message of the most recent failure.  No exceptions, no abort().
 *   - one context per (process, device); a context is not re-entrant.
 *   - there is NO CPU fallback: without a CUDA device every compute entry point returns HSIDM_CUDA_ERROR.
 */
#ifndef HSIDM_H_
#define HSIDM_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define HSIDM_API __attribute__((visibility("default")))
#else
#define HSIDM_API
#endif

#define HSIDM_VERSION 100 /* major*10000 + minor*100 + patch */

typedef enum hsidm_status {
  HSIDM_OK = 0,
  HSIDM_BAD_SHAPE = -1,
  HSIDM_BAD_DTYPE = -2,
  HSIDM_UNSUPPORTED_CFG = -3,
  HSIDM_CUDA_ERROR = -4,
  HSIDM_OOM_WORKSPACE = -5,
  HSIDM_BAD_ARG = -6,
  HSIDM_BAD_STATE = -7
} hsidm_status;

/* Arithmetic mode of a context.  F32: every contraction on CUDA cores in fp32 (parity gate 1e-4 on eps).
 * BF16: bf16 operands, fp32 accumulation, tcgen05 tensor cores for the convolutions (parity gate 2e-2). */
typedef enum hsidm_precision { HSIDM_F32 = 0, HSIDM_BF16 = 1 } hsidm_precision;

typedef void* hsidm_stream; /* a cudaStream_t */

typedef struct hsidm_ctx hsidm_ctx; /* SR3 UNet + diffusion schedule */
typedef struct hsidm_gae hsidm_gae; /* group autoencoder */

#define HSIDM_MAX_LEVELS 8

/* Constructor arguments of UNet (model/sr3_modules/unet.py:163-176) as passed by define_G
 * (model/networks.py:91-101). dropout is accepted for schema compatibility; inference is eval() (model/model.py:62). */
typedef struct hsidm_unet_cfg {
  int32_t in_channel;
  int32_t out_channel;
  int32_t inner_channel;
  int32_t norm_groups;
  int32_t n_mults;
  int32_t channel_mults[HSIDM_MAX_LEVELS];
  int32_t n_attn_res;
  int32_t attn_res[HSIDM_MAX_LEVELS];
  int32_t res_blocks;
  float dropout;
  int32_t image_size; /* model.diffusion.image_size: drives attention placement only (unet.py:195-233) */
  int32_t precision;  /* hsidm_precision */
} hsidm_unet_cfg;

/* Constructor arguments of GAE (AE.py:256-280) plus the widths found in the shipped checkpoints. */
typedef struct hsidm_gae_cfg {
  int32_t n_colors;
  int32_t n_subs;
  int32_t n_ovls;
  int32_t n_feats;      /* Encoder/Decoder width (64 in GAE_pretrained/GAE_4_*.pth) */
  int32_t trunk_feats;  /* 32, AE.py:268 */
  int32_t n_blocks;     /* SSB blocks per Encoder/Decoder branch: 3, AE.py:192,225 */
  int32_t trunk_blocks; /* 2, AE.py:268 */
  int32_t latent;       /* 3 latent channels, AE.py:259-260 */
  int32_t precision;    /* hsidm_precision */
} hsidm_gae_cfg;

HSIDM_API int hsidm_version(void);
HSIDM_API const char* hsidm_last_error(void);
/* Number of CUDA kernels this library has launched in this process (all contexts); used by bench.py. */
HSIDM_API int64_t hsidm_launch_count(void);
/* The tensor-core kernels bound every pipeline-barrier wait: a protocol or descriptor fault raises a device flag instead of
 * hanging the GPU, and the launch then leaves incomplete output.  hsidm_check_health() synchronises the current device,
 * reads and clears that flag and returns HSIDM_CUDA_ERROR if it was raised by any launch since the previous check
 * (HSIDM_OK otherwise).  The Python mirror calls it before results leave the library (SRPipeline, DDPM.test). */
HSIDM_API int hsidm_check_health(void);
/* Opt-in per-kernel-class timing used by bench.py's roofline leg: while enabled, every launch outside graph capture
 * is bracketed by CUDA events on its own stream. kind: 0 tensor-core conv, 1 CUDA-core conv, 2 GroupNorm stats,
 * 3 GroupNorm apply, 4 attention GEMM, 5 posterior step. read() synchronises and returns the summed event time (ms),
 * algorithmic work (FLOPs for 0/1/4, bytes for 2/3/5) and launch count since enable(1). */
HSIDM_API int hsidm_prof_enable(int on);
HSIDM_API int hsidm_prof_read(int kind, double* ms, double* work, int64_t* launches);
/* Writes one CSV line per recorded launch (kind,tag,work,ms); synchronises. */
HSIDM_API int hsidm_prof_dump(const char* path);

/* ---- UNet + diffusion -------------------------------------------------------------------------- */

/* UNet.__init__ (unet.py:163-236) + GaussianDiffusion.__init__ (diffusion.py:64-84). */
HSIDM_API int hsidm_ctx_create(const hsidm_unet_cfg* cfg, int device, hsidm_ctx** out);
HSIDM_API int hsidm_ctx_destroy(hsidm_ctx* ctx);

/* Enumerate the state_dict keys (relative to "denoise_fn.") the context expects, in the reference's
 * state_dict order. Returns the count; name(i) is NULL out of range. */
HSIDM_API int hsidm_unet_param_count(const hsidm_ctx* ctx);
HSIDM_API const char* hsidm_unet_param_name(const hsidm_ctx* ctx, int index);

/* nn.Module.load_state_dict for one tensor (model/model.py:177-202 loads "<prefix>_gen.pth").
 * `data` is fp32, host or device (copied); shape must match the reference parameter's shape. */
HSIDM_API int hsidm_unet_set_param(hsidm_ctx* ctx, const char* key, const float* data, const int64_t* shape, int ndim);
/* Whether the caller's current parameter tensors still equal the copies uploaded by set_param (bitwise).  In-place edits
 * through `.data` (networks.py:13-74 init_weights, model.py finetune_norm, EMA swaps) do not bump a tensor's autograd
 * version, so the Python mirror asks the device: `table_dev` is a DEVICE array of hsidm_unet_param_count() device
 * pointers in parameter order.  *changed (host) = 1 if any element differs or nothing was uploaded yet.  Synchronises
 * `stream`. */
HSIDM_API int hsidm_unet_params_changed(hsidm_ctx* ctx, const void* const* table_dev, int n, int* changed, hsidm_stream stream);
/* Pack weights into kernel layouts (bf16 K-major for tcgen05, fp32 [K][Cout] for the fp32 path). Must be
 * called after the last set_param and before forward/sample; fails if a parameter was never set. */
HSIDM_API int hsidm_unet_commit(hsidm_ctx* ctx);

/* GaussianDiffusion.set_new_noise_schedule (diffusion.py:93-140) given float64 betas from
 * make_beta_schedule (diffusion.py:19-49). Builds the per-timestep coefficient and noise-embedding tables. */
HSIDM_API int hsidm_set_schedule(hsidm_ctx* ctx, const double* betas, int T);

/* UNet.forward(x, time) (unet.py:239-263). The 6-channel input may be given as one tensor (x1 == NULL,
 * c0 == in_channel) or as the two halves of torch.cat([condition_x, x], dim=1) (diffusion.py:158) with
 * c0 + c1 == in_channel. noise_level: N fp32 values, element n at noise_level[n * level_stride]
 * (stride 0 = one shared level). eps_out: [N, out_channel, H, W]. H and W must be multiples of 2^(levels-1). */
HSIDM_API int hsidm_unet_forward(hsidm_ctx* ctx, const float* x0, int c0, const float* x1, int c1, const float* noise_level,
                       int level_stride, float* eps_out, int N, int H, int W, hsidm_stream stream);

/* p_mean_variance + p_sample minus the UNet call (diffusion.py:142-175): x0 = a_t*x_t - b_t*eps; clamp[-1,1];
 * mean = c1_t*x0 + c2_t*x_t; x_prev = mean + noise*exp(0.5*logvar_t). noise may be NULL (t == 0). n = elements. */
HSIDM_API int hsidm_posterior_step(hsidm_ctx* ctx, int t, const float* x_t, const float* eps, const float* noise,
                         float* x_prev, int64_t n, hsidm_stream stream);

/* GaussianDiffusion.p_sample_loop, conditional branch (diffusion.py:188-201), batched over N latent images.
 *   cond  [N,c,H,W]   condition_x           x_T [N,c,H,W] the initial torch.randn draw (diffusion.py:192)
 *   noise_tape: NULL -> per-step noise from the built-in counter-based Philox generator keyed by `seed`;
 *               else element (n, j) at noise_tape + n*tape_image_stride + j*tape_step_stride is the
 *               randn_like draw used at loop index i = T-1-j (j = 0..T-2) (diffusion.py:174).
 *   out   [N,c,H,W]   final x_0
 *   snapshots: NULL or [n_snap, N, c, H, W]; snapshot k is img after loop index i with i % (1|(T/10)) == 0,
 *              in loop order (diffusion.py:193-197); n_snap = hsidm_snapshot_count(ctx).
 * Steps are replayed from a CUDA graph after the first call at a given (N,H,W). */
HSIDM_API int hsidm_sample(hsidm_ctx* ctx, const float* cond, const float* x_T, const float* noise_tape,
                 int64_t tape_image_stride, int64_t tape_step_stride, uint64_t seed, float* out, float* snapshots,
                 int N, int H, int W, hsidm_stream stream);
/* Same, for a batch that is the slice [first_image, first_image + N) of a longer list of latent images sampled in several
 * calls or on several GPUs: the built-in generator's counters are offset so that image k of the list sees the same
 * per-step noise whatever batch or rank it lands in (tile-sharded scenes, SURVEY 8e/8f N1, give bit-identical results on
 * 1 and N GPUs).  first_image is ignored when a noise tape is injected. */
HSIDM_API int hsidm_sample_at(hsidm_ctx* ctx, const float* cond, const float* x_T, const float* noise_tape,
                    int64_t tape_image_stride, int64_t tape_step_stride, uint64_t seed, int64_t first_image, float* out,
                    float* snapshots, int N, int H, int W, hsidm_stream stream);
/* n standard-normal draws (n and first_element multiples of 4) from the same counter-based generator, stream `seed`,
 * starting at element first_element of that stream: the x_T draw of torch.randn(shape) (diffusion.py:192) made
 * independent of batch composition.  out is a device pointer. */
HSIDM_API int hsidm_randn(float* out, int64_t n, uint64_t seed, int64_t first_element, hsidm_stream stream);
HSIDM_API int hsidm_snapshot_count(const hsidm_ctx* ctx);
HSIDM_API int hsidm_num_timesteps(const hsidm_ctx* ctx);
/* Workspace bytes currently held by the context (arena + packed weights + tables). */
HSIDM_API int64_t hsidm_ctx_bytes(const hsidm_ctx* ctx);

/* ---- training step ---------------------------------------------------------------------------------- */

/* GaussianDiffusion.p_losses (diffusion.py:222-250) for one batch: x_noisy = level*HR + sqrt(1-level^2)*noise (q_sample,
 * diffusion.py:213-220), eps = UNet(cat([SR, x_noisy]), level), loss_sum = sum|noise - eps| (loss_type 0, "l1") or
 * sum (noise - eps)^2 (loss_type 1, "l2").  hr / sr / noise: device [B, out_channel, H, W] fp32 (the GAE latents of the HR cube,
 * of the bicubic cube, and the N(0,1) draw of diffusion.py:238); levels: B continuous sqrt(alpha_bar) values (host or device;
 * the np.random.uniform draw of diffusion.py:228-234); dropout_seed keys the counter-based dropout mask of block2
 * (unet.py:102; the mask is regenerated, not stored, by the backward).  Saves every activation the backward needs in a
 * workspace owned by the context; grad_slab (device, hsidm_train_grad_numel floats) is where hsidm_train_backward will leave
 * the gradients.  loss_sum: one float (host or device).  fp32 on CUDA cores, deterministic (no atomics). */
HSIDM_API int hsidm_train_forward(hsidm_ctx* ctx, const float* hr, const float* sr, const float* noise, const float* levels, int B,
                        int H, int W, int loss_type, uint64_t dropout_seed, float* grad_slab, float* loss_sum, hsidm_stream stream);
/* Backward of upstream * loss_sum / (B*out_channel*H*W) (DDPM.optimize_parameters, model.py:49-55): gradients of every
 * parameter into the slab given to the forward; parameter i (hsidm_unet_param_name order) starts at element
 * hsidm_train_grad_offset(ctx, i) and has the parameter's own shape and layout.  The slab is contiguous so that a
 * data-parallel job all-reduces it with ONE NCCL call (BASELINE configs[4]).  Consumes the saved forward. */
HSIDM_API int hsidm_train_backward(hsidm_ctx* ctx, float upstream, hsidm_stream stream);
HSIDM_API int64_t hsidm_train_grad_numel(const hsidm_ctx* ctx);
HSIDM_API int64_t hsidm_train_grad_offset(const hsidm_ctx* ctx, int index);

/* ---- group autoencoder ------------------------------------------------------------------------------ */

/* GAE.__init__ (AE.py:256-280). */
HSIDM_API int hsidm_gae_create(const hsidm_gae_cfg* cfg, int device, hsidm_gae** out);
HSIDM_API int hsidm_gae_destroy(hsidm_gae* gae);
HSIDM_API int hsidm_gae_param_count(const hsidm_gae* gae);
HSIDM_API const char* hsidm_gae_param_name(const hsidm_gae* gae, int index);
HSIDM_API int hsidm_gae_set_param(hsidm_gae* gae, const char* key, const float* data, const int64_t* shape, int ndim);
HSIDM_API int hsidm_gae_commit(hsidm_gae* gae);
/* Same contract as hsidm_unet_params_changed, over hsidm_gae_param_count() pointers. */
HSIDM_API int hsidm_gae_params_changed(hsidm_gae* gae, const void* const* table_dev, int n, int* changed, hsidm_stream stream);
/* Group count and band ranges computed by AE.py:264-280. start/end receive G entries each (may be NULL). */
HSIDM_API int hsidm_gae_groups(const hsidm_gae* gae, int32_t* start, int32_t* end);

/* GAE.encode (AE.py:310-324): x [B,n_colors,H,W] -> z [B*G, latent, H, W]; latent image of (cube b, group g)
 * is z[b*G + g]. All B*G band groups go through the Encoder as one batch. */
HSIDM_API int hsidm_gae_encode(hsidm_gae* gae, const float* x, float* z, int B, int H, int W, hsidm_stream stream);
/* GAE.decode (AE.py:283-308): z [B*G, latent, H, W] -> y [B,n_colors,H,W] = avg-overlap(Decoder(z)) passed
 * through the residual trunk. clamp01 != 0 additionally applies the driver's clamp to [0,1] (sr_gae.py:474-475). */
HSIDM_API int hsidm_gae_decode(hsidm_gae* gae, const float* z, float* y, int B, int H, int W, int clamp01, hsidm_stream stream);

/* ---- the steps either side of the path ------------------------------------------------------------ */

/* Bicubic x`scale` pre-upsampling of the low-resolution cube, NCHW fp32 on the device: replaces
 * torch.nn.functional.interpolate(img_LR, scale_factor=4, mode='bicubic') (sr_gae.py:72, :118).  PyTorch's convention:
 * align_corners=False, A = -0.75, border indices clamped.  lr [N,C,h,w] -> sr [N,C,h*scale,w*scale]; clamp01 != 0 also
 * applies the dataset's clamp to [0,1] (HStest.py:59-60). */
HSIDM_API int hsidm_bicubic_upsample(const float* lr, float* sr, int N, int C, int h, int w, int scale, int clamp01, hsidm_stream stream);

/* Per-cube validation metrics on the device, after the driver's clamp of both cubes to [0,1] (sr_gae.py:474-475):
 * out[n] = (MPSNR in dB with data_range 1 - eval_hsi.py:110-121, SAM in degrees - eval_hsi.py:47-65).
 * truth / pred: [N,C,H,W] fp32 device pointers; out: [N][2] fp32 device pointer.  Deterministic (fixed-order folds in
 * float64); allocates its scratch with cudaMallocAsync on `stream`. */
HSIDM_API int hsidm_quality_metrics(const float* truth, const float* pred, int N, int C, int H, int W, float* out, hsidm_stream stream);

/* MATLAB-style imresize of the dataset code (GAE/imsize.py:116-158 `imresize(I, output_shape=...)`, called by
 * HStest.py:44-45 and HStrain.py:61-63 for the x4 degradation and the pre-upsampling): separable resampling, antialiased
 * when shrinking (kernel stretched by 1/scale), taps mirrored at the borders, weights and sums in float64.
 * in [planes,h,w] -> out [planes,out_h,out_w] fp32 device pointers (a plane = one band of one cube; the reference's HWC
 * arrays are resized band by band, so the layouts are interchangeable).  method: 0 = 'bicubic' (a = -0.5), 1 = 'bilinear'
 * (both with the reference's kernel width 4).  scale_h / scale_w: the resampling scale per axis - pass 0 for the
 * `output_shape` form (scale = out / in, imsize.py:10-14); the `scalar_scale` form passes the scalar itself with
 * out = ceil(scale * in) (imsize.py:3-7), which is not out / in in general.  Allocates its tap tables with cudaMallocAsync
 * on `stream`. */
HSIDM_API int hsidm_imresize(const float* in, float* out, int planes, int h, int w, int out_h, int out_w, double scale_h, double scale_w,
                   int method, hsidm_stream stream);

/* quality_assessment (eval_hsi.py:217-238) per cube on the device, after the driver's clamp of both cubes to [0,1]
 * (sr_gae.py:474-475).  out [N][6] fp32 in the reference dict's key order: MPSNR (eval_hsi.py:110-121), MSSIM (:124-135,
 * skimage.metrics.structural_similarity with its defaults for float images: 7x7 uniform window, K1 = .01, K2 = .03, sample
 * covariance; NaN when H or W < 7), ERGAS (:18-35, `ratio` = the SR factor), SAM in degrees (:47-65), CrossCorrelation
 * (:58-70), RMSE (:88-96); data_range 1.  truth / pred: [N,C,H,W] fp32 device pointers, N*C <= 65535.  Deterministic. */
HSIDM_API int hsidm_quality_assessment(const float* truth, const float* pred, int N, int C, int H, int W, float ratio, float* out,
                             hsidm_stream stream);

/* Overlapping-tile scene driver (the reference only crops non-overlapping 128x128 blocks offline,
 * GAE/crop.py:12-36, HStest.py:33-45): feathered overlap-add of super-resolved tiles back into the scene on the device.
 * tiles [ny*nx, C, tile, tile] fp32 in row-major tile order, tile (iy, ix) at origin (ys[iy], xs[ix]); ys / xs are DEVICE
 * int32 arrays; weights ramp linearly over `overlap` pixels at every tile border; out [C, H, W].  Deterministic. */
HSIDM_API int hsidm_blend_tiles(const float* tiles, const int32_t* ys, int ny, const int32_t* xs, int nx, int C, int tile,
                      int overlap, int H, int W, float* out, hsidm_stream stream);

#ifdef __cplusplus
}
#endif
#endif /* HSIDM_H_ */
