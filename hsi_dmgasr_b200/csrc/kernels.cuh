// Host-side launch interface of every hsidm kernel.  All launchers enqueue on `stream` and return an
// hsidm_status; none synchronises.  `prec` selects the activation element type (HSIDM_F32 / HSIDM_BF16).
#pragma once
#include "common.cuh"

namespace hsidm {

enum ActFn { ACT_NONE = 0, ACT_LRELU = 1 };
enum TensorLayout { L_NHWC = 0, L_NCHW_F32 = 1 };

// One input of a convolution.  NHWC sources hold the context's activation type; NCHW sources are the
// caller's fp32 tensors in the reference layout (first UNet conv, GAE head convs).
struct ConvSrc {
  const void* p = nullptr;
  int C = 0;
  int layout = L_NHWC;
  const int64_t* img_off = nullptr;  // NCHW only: per-image element offset (device); null -> n*C*H*W
};

// GroupNorm over cat(src0, src1) as its CONSUMER evaluates it (no kernel between the producing convolution and the consumer):
// either from the per-slot partial sums the producing convolutions left ([N][slots][C][2] = (sum, sumsq), folded in slot
// order in float64 by the consumer's prologue), or from precomputed statistics [N][groups][2] = (mean, rstd).
struct GnIn {
  const float* part[2] = {nullptr, nullptr};   // per-slot partial sums of src0 / src1 (part[1] unused when C[1] == 0)
  int slots[2] = {0, 0};
  int C[2] = {0, 0};
  const float* stats = nullptr;                // alternative to part[]: (mean, rstd) per image and group
  const float* gamma = nullptr;
  const float* beta = nullptr;
  int groups = 0;
  float eps = 0.f;
  int swish = 0;
  bool on() const { return gamma && (stats || part[0]); }
};

// out = [clamp01]( act(conv(cat(src0,src1)) + bias + nbias[n]) * scale + resid )
struct ConvOp {
  ConvSrc src[2];
  // optional shortcut sources (halo tensor-core kernel only): un-normalised NHWC tensors whose 1x1 projection (or
  // identity) is accumulated into the same output: their channels are extra K columns after the k*k*(C0+C1) ones
  ConvSrc rsrc[2];
  int N = 0, Hin = 0, Win = 0;  // source spatial size
  int up = 0;                   // 1: source is nearest-2x upsampled on the fly (unet.py:58-65)
  // >= 0: sub-pixel form of up + 3x3 conv (halo tensor-core kernel): parity p = 2*py + px computes the output pixels
  // (2i+py, 2j+px) as a 2x2-tap conv over the LOW-RES source with pre-summed weights (w_bf16 = [Cout][4 parities][4 taps][C]);
  // Hin/Win are the source size, Hout/Wout twice that.  2.25x fewer FLOPs and no upsampled tensor.
  int up_parity = -1;
  // 1: stride-2 3x3 conv in its phase-lattice form (halo tensor-core kernel): src[0] is the FULL-RESOLUTION tensor
  // [N, 2*Hin, 2*Win, C], Hin/Win = Hout/Wout are the output size, w_bf16 = ConvW::w_s2.  No im2col buffer.
  int s2 = 0;
  int ksize = 3, stride = 1;    // padding = ksize/2
  int Hout = 0, Wout = 0, Cout = 0;
  const float* w_f32 = nullptr;  // [K][Cout], k = tap*(C0+C1) + c
  const bf16* w_bf16 = nullptr;  // [Cout][K]   (tensor-core path)
  const float* bias = nullptr;
  const float* nbias = nullptr;  // noise-embedding bias, element (n, co) at nbias[t*nbias_t_stride + n*nbias_stride + co]
  int64_t nbias_stride = 0;
  const int* nbias_t = nullptr;  // optional device-side timestep index t (CUDA-graph replay); null -> t = 0
  int64_t nbias_t_stride = 0;
  int act = ACT_NONE;
  float scale = 1.0f;
  const void* resid = nullptr;  // NHWC activation [N,Hout,Wout,Cout]
  void* out = nullptr;
  int out_layout = L_NHWC;
  int clamp01 = 0;
  // optional fused GroupNorm statistics of the output: [N][stats_slots][Cout][2]; stats_slots must equal
  // conv_tc_stats_slots(op) (tensor-core kernels only)
  float* stats_out = nullptr;
  int stats_slots = 0;
  // optional GroupNorm(+Swish) of the INPUT fused into the load path (halo tensor-core kernel only): src[] hold the
  // un-normalised tensors; the kernel derives the group statistics of the images it works on from gn (see GnIn) and
  // applies act(x*A + B) to each halo tile in shared memory, zero padding AFTER the activation, as in
  // GroupNorm -> Swish -> Conv2d.  The shortcut sources rsrc[] are never normalised.
  GnIn gn;
  int K() const {
    if (up_parity >= 0) return 16 * (src[0].C + src[1].C);
    return ksize * ksize * (src[0].C + src[1].C) + rsrc[0].C + rsrc[1].C;
  }
};

// conv_simt.cu : CUDA-core implicit GEMM, fp32 accumulate; handles every ConvOp.
int conv_simt(const ConvOp& op, int prec, cudaStream_t stream);
// conv_tc.cu : tcgen05/TMEM/TMA implicit GEMM (bf16 operands).  conv_tc_supported() says whether an op
// fits the tensor-core kernel's constraints; conv_tc() fails loudly otherwise.
bool conv_tc_supported(const ConvOp& op, int prec);
int conv_tc(const ConvOp& op, cudaStream_t stream);
bool conv_halo_ok(const ConvOp& op);         // the halo kernel (and with it fused shortcut sources) takes this op
int conv_tc_stats_slots(const ConvOp& op);  // slots per image the kernel chosen for `op` fills (0 = no fused statistics)
int conv_tc_init();               // resolves cuTensorMapEncodeTiled, sets kernel attributes
int conv_tc_bn_rows(int Cout);    // N-tile height; packed bf16 weights are padded to a multiple of it (0 = unsupported)
void conv_tc_set_mode(int no_halo, int variant);  // test knobs
int conv_tc_variant();                             // current variant bits (8 = no fused input GroupNorm)
int conv_tc_route_gen();                           // bumped by every conv_tc_set_mode call (part of the sampler's graph key)
void conv_halo_set_timing(long long* device_counters);   // developer probe, see hsidm_debug_halo_timing
int conv_tc_error_flag(int* v);   // barrier-timeout flag of the tensor-core kernel (synchronises; tests only)

// Batched GEMM on CUDA cores: C[b] = alpha * A[b] (MxK, row-major lda) * op(B[b]); B is [N][K] (transB=1)
// or [K][N] (transB=0). A/B hold the activation type, C is fp32 (c_f32=1) or the activation type.
struct GemmOp {
  const void* A = nullptr;
  const void* B = nullptr;
  void* C = nullptr;
  int M = 0, N = 0, K = 0;
  int64_t lda = 0, ldb = 0, ldc = 0, sA = 0, sB = 0, sC = 0;
  int batch = 1, transB = 1, a_f32 = 0, c_f32 = 0;
  float alpha = 1.0f;
};
int gemm_simt(const GemmOp& op, int prec, cudaStream_t stream);
int softmax_rows(float* x, int64_t rows, int cols, cudaStream_t stream);  // in place, fp32
int softmax_rows_bf16(const float* x, void* p_bf16, int64_t rows, int cols, cudaStream_t stream);  // fp32 in, bf16 out

// gemm_tc.cu : the same contraction on tcgen05 (bf16 K-major operands, A/B batch stride 0 = shared operand).
struct GemmTcOp {
  const void* A = nullptr;
  const void* B = nullptr;
  void* C = nullptr;
  int M = 0, N = 0, K = 0, batch = 1;
  int64_t lda = 0, ldb = 0, ldc = 0, sA = 0, sB = 0, sC = 0;
  int c_f32 = 0;
  float alpha = 1.0f;
  // 1: C = softmax over each row of alpha * A * B^T, written as bf16 (attention probabilities).  Needs the whole row in
  // one tile: N <= 256, N % 64 == 0, c_f32 == 0.  The scores never leave tensor memory.
  int row_softmax = 0;
};
// attn_flash.cu : softmax(alpha * Q K^T) V for up to 256 tokens in one kernel (scores in tensor memory, probabilities in
// shared memory).  Q, K: [batch][S][Ck] (row pitch ldq / ldk, batch stride sQ / sK), Vt: [batch][Cv][S] (V transposed, row
// pitch S), Y: [batch][S][Cv] (row pitch ldy); all bf16, element strides.  Cv a multiple of 256, Ck of 64, S in {64, 128, 256}.
struct AttnFlashOp {
  const void* Q = nullptr;
  const void* K = nullptr;
  const void* Vt = nullptr;
  void* Y = nullptr;
  int S = 0, Ck = 0, Cv = 0, batch = 1;
  int64_t ldq = 0, ldk = 0, ldy = 0, sQ = 0, sK = 0, sVt = 0, sY = 0;
  float alpha = 1.0f;
};
bool attn_flash_supported(const AttnFlashOp& op);
int attn_flash(const AttnFlashOp& op, cudaStream_t stream);
int attn_flash_init();

bool gemm_tc_supported(const GemmTcOp& op);
int gemm_tc(const GemmTcOp& op, cudaStream_t stream);
int gemm_tc_init();

// GroupNorm over the (virtual) concatenation of two NHWC tensors with the same N,H,W (deterministic reduction).
// scratch: gn_scratch_bytes() bytes = per-block partials followed by the published (mean, rstd) pairs.
// tickets: N zero-initialised counters, left zero again by gn_stats (one array per stream is enough).
struct GnGeo {
  int CV, lanes, threads, pix_per_block, slabs;
};
int gn_geometry(int C0, int C1, int N, int HW, GnGeo* g);
int64_t gn_scratch_bytes(int C0, int C1, int N, int HW, int groups);
inline float* gn_stats_ptr(void* scratch, int N, int slabs, int groups) {
  return reinterpret_cast<float*>(static_cast<double*>(scratch) + (int64_t)2 * N * slabs * groups);
}
// Standalone statistics pass (reads the tensors): stats[N][groups][2] = (mean, rstd).
int gn_stats(const void* x0, int C0, const void* x1, int C1, int N, int HW, int groups, float eps, void* scratch,
             unsigned* tickets, float* stats, int prec, cudaStream_t stream);
// Statistics from the per-slot partial sums the producing convolutions left behind (no pass over the tensors).
// part0/part1: [N][slots][C][2] (sum, sumsq) for the two concatenated sources.
// gamma/beta/ab optional: when given, also writes the per-channel affine ab[n][c] = (A, B) with GN(x) = x*A + B.
int gn_finalize(const float* part0, int slots0, int C0, const float* part1, int slots1, int C1, int N, int HW, int groups,
                float eps, float* stats, cudaStream_t stream, const float* gamma = nullptr, const float* beta = nullptr,
                float* ab = nullptr);
int gn_apply(const void* x0, int C0, const void* x1, int C1, int N, int HW, int groups, const float* stats,
             const float* gamma, const float* beta, int swish, void* out, int prec, cudaStream_t stream);
// The same with the statistics taken from `gn` (per-slot partial sums folded in each block's prologue, or gn.stats):
// no statistics / finalize kernel in front.  gn.C[] are the channel counts of x0 / x1.
int gn_apply_fused(const void* x0, const void* x1, int N, int HW, const GnIn& gn, void* out, int prec, cudaStream_t stream);
// bf16 only: y [N][S][C] = GroupNorm(x) (no activation) and the same values transposed, yt [N][C][S] (attention block).
int gn_apply_transposed(const void* x, int N, int S, const GnIn& gn, void* y, void* yt, cudaStream_t stream);

int upsample2x(const void* x, void* out, int N, int H, int W, int C, int prec, cudaStream_t stream);

// prepost.cu : bicubic pre-upsampling of NCHW fp32 planes and the per-cube MPSNR / SAM metrics
int bicubic_upsample(const float* src, float* dst, int planes, int h, int w, int scale, int clamp01, cudaStream_t stream);
// MATLAB-style imresize (imsize.py:116-158) of `planes` fp32 planes h x w -> H x W with the given per-axis scales;
// method 0 = bicubic, 1 = bilinear.
int64_t imresize_scratch_bytes(int H, int W, double scale_h, double scale_w);
int imresize(const float* src, float* dst, int planes, int h, int w, int H, int W, double scale_h, double scale_w, int method,
             void* scratch, cudaStream_t stream);
// quality_assessment (eval_hsi.py:217-238): out [N][6] = (MPSNR, MSSIM, ERGAS, SAM, CrossCorrelation, RMSE)
int64_t quality_assessment_scratch_bytes(int N, int C, int H, int W);
int quality_assessment(const float* truth, const float* pred, int N, int C, int H, int W, float ratio, void* scratch, float* out,
                       cudaStream_t stream);
int64_t quality_metrics_scratch_bytes(int N, int C, int HW);
int quality_metrics(const float* truth, const float* pred, int N, int C, int HW, float data_range, void* scratch, float* out,
                    cudaStream_t stream);
// 3x3/pad-1 im2col of up to two NCHW fp32 sources with 9*(C0+C1) <= 64 into [N,H,W,64] bf16 rows (k = tap*(C0+C1) + c,
// zero padded): lets the first UNet conv (6 -> 64) run as a K = 64 tensor-core GEMM.
int im2col_small(const float* x0, int C0, const float* x1, int C1, void* out, int N, int H, int W, cudaStream_t stream);
// Stride-2 3x3 im2col for the tensor-core path: out [N,H/2,W/2,9*C] (bf16).
int im2col_s2(const void* x, void* out, int N, int H, int W, int C, cudaStream_t stream);

// Noise-level embedding (unet.py:18-31, 182-187) followed by every FeatureWiseAffine linear (unet.py:34-50).
struct NoiseLayer {
  const float* w;      // [C][dim]
  const float* b;      // [C]
  const float* cbias;  // bias of the conv the embedding is added to (folded in so the epilogue adds one vector), or null
  int C, off;
};
int noise_embed(const float* level, int level_stride, int n, int dim, const float* w1, const float* b1,
                const float* w3, const float* b3, const NoiseLayer* layers_dev, int n_layers, int total,
                float* nbias /*[n][total]*/, cudaStream_t stream);

// Fused DDPM posterior step (diffusion.py:142-175). coef: device [T][5] = {recip, recipm1, c1, c2, sigma}.
// t_dev (optional) overrides t with *t_dev (graph replay). noise == null -> zeros; philox != 0 -> in-kernel RNG.
struct PosteriorArgs {
  const float* x_t;
  const float* eps;
  const float* noise;
  float* x_prev;
  int64_t n;           // elements
  int64_t per_image;   // elements per image (C*H*W), used to index the tape
  int64_t tape_image_stride, tape_step_stride;
  const float* coef;
  int T;
  int t;               // loop index (host known) or -1 -> read *t_dev
  const int* t_dev;
  uint64_t seed;
  const unsigned long long* seed_dev;  // optional: overrides seed (graph replay); seed_dev[1] = first element / 4 of this
                                       // batch in the caller's global image order (Philox counter offset)
  int use_philox;
  float* snapshot_base;   // optional: [n_snap][n] written when i % inter == 0
  int inter;
};
int posterior_step(const PosteriorArgs& a, cudaStream_t stream);
int step_counter_dec(int* t_dev, cudaStream_t stream);  // *t_dev -= 1
// state[0] = t (int), state[2..3] = seed (uint64), state[4..5] = Philox counter offset in float4 units (uint64): set from
// kernel arguments so no host buffer has to outlive the call
int sampler_state_set(int* state, int t, uint64_t seed, uint64_t offset4, cudaStream_t stream);
// out[i] ~ N(0,1), i in [0, n): the Philox stream (seed, step) at counter first4 + i/4 (n and out 16-byte multiples)
int randn_fill(float* out, int64_t n, uint64_t seed, uint32_t step, uint64_t first4, cudaStream_t stream);
// Feathered overlap-add of square tiles back into a scene (pipeline.blend_tiles): tiles [ny*nx][C][t][t] fp32 in row-major
// tile order at origins (ys[iy], xs[ix]) (device arrays), weights = separable linear ramps over `overlap` pixels, per-pixel
// normalisation; out [C][H][W].  Gather form, fixed summation order (iy, ix ascending): deterministic.
int blend_tiles(const float* tiles, const int* ys, int ny, const int* xs, int nx, int C, int tile, int overlap, int H, int W,
                float* out, cudaStream_t stream);

// GAE pieces (common.py:231-271, AE.py:288-295).
int channel_mean(const void* x, int N, int HW, int C, float* mean /*[N][C]*/, int prec, cudaStream_t stream);
int ca_gate(const float* mean, int N, int C, int Cr, const float* w0, const float* b0, const float* w1,
            const float* b1, float* gate /*[N][C]*/, cudaStream_t stream);
// out = x * gate[n][c] * scale + resid
int scale_residual(const void* x, const float* gate, float scale, const void* resid, void* out, int N, int HW, int C,
                   int prec, cudaStream_t stream);
// y[b, band] = sum_g dec[b*G+g, band - start[g]] / count[band]; dec NHWC [B*G,H,W,n_subs]; writes NHWC act
// [B,H,W,n_colors] (trunk input) and keeps fp32 precision by also writing y_f32 NHWC when non-null.
int overlap_average(const void* dec, int B, int G, int HW, int n_subs, int n_colors, const int* start_dev,
                    const float* inv_count_dev, void* y, int prec, cudaStream_t stream);

}  // namespace hsidm
