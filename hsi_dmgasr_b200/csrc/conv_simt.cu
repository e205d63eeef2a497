// CUDA-core implicit-GEMM convolution (fp32 accumulate).
//
// This is the arithmetic of the F32 precision mode (every nn.Conv2d of unet.py / AE.py / common.py evaluated
// in true fp32, parity gate 1e-4) and, in BF16 mode, the kernel for the few layers whose channel counts do not
// fit a tensor-core tile (first 6->64 conv, GAE head/trunk convs).  It gathers its A operand by index
// arithmetic, so stride-2, nearest-2x upsampling, two-source channel concatenation and NCHW fp32 sources all
// come for free.
//
// GEMM view: M = N*Hout*Wout output pixels, N = Cout, K = taps*(C0+C1); tile BM x BN x 16, 256 threads,
// 4x4 outputs per thread.
#include "kernels.cuh"

namespace hsidm {
namespace {

struct SimtP {
  const void* s[2];
  const int64_t* off[2];
  int C[2], lay[2];
  int N, Hin, Win, up, ks, stride, Hout, Wout, Cout;
  const float* w;
  const float* bias;
  const float* nbias;
  int64_t nbs;
  const int* nb_t;
  int64_t nb_ts;
  int act;
  float scale;
  const void* resid;
  void* out;
  int out_layout, clamp01;
  int Ctot, Ktot, M;
};

constexpr int BK = 16;

template <typename AT>
__device__ __forceinline__ float load_src(const SimtP& p, int which, int n, int y, int x, int c) {
  if (p.lay[which] == L_NHWC) {
    const AT* b = static_cast<const AT*>(p.s[which]);
    return to_f32(b[(((int64_t)n * p.Hin + y) * p.Win + x) * p.C[which] + c]);
  }
  const float* b = static_cast<const float*>(p.s[which]);
  int64_t base = p.off[which] ? p.off[which][n] : (int64_t)n * p.C[which] * p.Hin * p.Win;
  return b[base + ((int64_t)c * p.Hin + y) * p.Win + x];
}

template <typename AT, int BM, int BN, bool VEC>
__global__ void __launch_bounds__(256) conv_simt_kernel(const SimtP p) {
  constexpr int TX = BN / 4;   // threads along cout
  constexpr int TY = BM / 4;   // threads along pixels
  static_assert(TX * TY == 256, "tile/thread mismatch");
  constexpr int A_PER_T = BM * BK / 256;  // consecutive k's per thread for one pixel
  constexpr int A_TPP = BK / A_PER_T;     // threads per pixel
  constexpr int B_PER_T = BN * BK / 256;  // consecutive couts per thread for one k
  constexpr int B_TPK = BN / B_PER_T;
  static_assert(A_PER_T == 4 || A_PER_T == 16, "A mapping");

  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];

  const int tid = threadIdx.x;
  const int tx = tid % TX, ty = tid / TX;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;

  // pixel this thread gathers for
  const int a_pix = tid / A_TPP, a_k = (tid % A_TPP) * A_PER_T;
  const int am = m0 + a_pix;
  const bool am_ok = am < p.M;
  int an = 0, aoy = 0, aox = 0;
  if (am_ok) {
    an = am / (p.Hout * p.Wout);
    int r = am - an * p.Hout * p.Wout;
    aoy = r / p.Wout;
    aox = r - aoy * p.Wout;
  }
  const int pad = p.ks >> 1;
  const int He = p.up ? p.Hin * 2 : p.Hin, We = p.up ? p.Win * 2 : p.Win;

  const int b_k = tid / B_TPK, b_c = (tid % B_TPK) * B_PER_T;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < p.Ktot; k0 += BK) {
    // ---- gather A ----
    {
      float v[A_PER_T];
#pragma unroll
      for (int j = 0; j < A_PER_T; ++j) v[j] = 0.f;
      if (VEC) {
#pragma unroll
        for (int j4 = 0; j4 < A_PER_T; j4 += 4) {
          const int k = k0 + a_k + j4;
          if (am_ok && k < p.Ktot) {
            const int tap = k / p.Ctot, c = k - tap * p.Ctot;
            const int ky = tap / p.ks, kx = tap - ky * p.ks;
            int iy = aoy * p.stride + ky - pad, ix = aox * p.stride + kx - pad;
            if (iy >= 0 && iy < He && ix >= 0 && ix < We) {
              if (p.up) iy >>= 1, ix >>= 1;
              const int which = c < p.C[0] ? 0 : 1;
              const int cc = which ? c - p.C[0] : c;
              const AT* src = static_cast<const AT*>(p.s[which]) +
                              (((int64_t)an * p.Hin + iy) * p.Win + ix) * p.C[which] + cc;
              if (sizeof(AT) == 4) {
                float4 q = *reinterpret_cast<const float4*>(src);
                v[j4] = q.x, v[j4 + 1] = q.y, v[j4 + 2] = q.z, v[j4 + 3] = q.w;
              } else {
                uint2 q = *reinterpret_cast<const uint2*>(src);
                const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
                float2 f0 = __bfloat1622float2(h[0]), f1 = __bfloat1622float2(h[1]);
                v[j4] = f0.x, v[j4 + 1] = f0.y, v[j4 + 2] = f1.x, v[j4 + 3] = f1.y;
              }
            }
          }
        }
      } else {
#pragma unroll
        for (int j = 0; j < A_PER_T; ++j) {
          const int k = k0 + a_k + j;
          if (am_ok && k < p.Ktot) {
            const int tap = k / p.Ctot, c = k - tap * p.Ctot;
            const int ky = tap / p.ks, kx = tap - ky * p.ks;
            int iy = aoy * p.stride + ky - pad, ix = aox * p.stride + kx - pad;
            if (iy >= 0 && iy < He && ix >= 0 && ix < We) {
              if (p.up) iy >>= 1, ix >>= 1;
              const int which = c < p.C[0] ? 0 : 1;
              v[j] = load_src<AT>(p, which, an, iy, ix, which ? c - p.C[0] : c);
            }
          }
        }
      }
#pragma unroll
      for (int j = 0; j < A_PER_T; ++j) As[a_k + j][a_pix] = v[j];
    }
    // ---- load B ----
    {
      const int k = k0 + b_k;
#pragma unroll
      for (int j = 0; j < B_PER_T; ++j) {
        const int co = n0 + b_c + j;
        Bs[b_k][b_c + j] = (k < p.Ktot && co < p.Cout) ? __ldg(p.w + (int64_t)k * p.Cout + co) : 0.f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w};
      const float b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

  // ---- epilogue ----
  const int HWo = p.Hout * p.Wout;
  const float* nbias = p.nbias ? p.nbias + (p.nb_t ? (int64_t)(*p.nb_t) * p.nb_ts : 0) : nullptr;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= p.M) continue;
    const int n = m / HWo;
    const int r = m - n * HWo;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int co = n0 + tx * 4 + j;
      if (co >= p.Cout) continue;
      float v = acc[i][j];
      if (p.bias) v += __ldg(p.bias + co);
      if (nbias) v += __ldg(nbias + (int64_t)n * p.nbs + co);
      if (p.act == ACT_LRELU) v = v > 0.f ? v : 0.01f * v;
      v *= p.scale;
      if (p.resid) v += to_f32(static_cast<const AT*>(p.resid)[(int64_t)m * p.Cout + co]);
      if (p.clamp01) v = fminf(fmaxf(v, 0.f), 1.f);
      if (p.out_layout == L_NHWC)
        static_cast<AT*>(p.out)[(int64_t)m * p.Cout + co] = from_f32<AT>(v);
      else
        static_cast<float*>(p.out)[((int64_t)n * p.Cout + co) * HWo + r] = v;
    }
  }
}

template <typename AT>
int launch(const SimtP& p, cudaStream_t stream) {
  const bool vec = p.lay[0] == L_NHWC && (p.C[1] == 0 || p.lay[1] == L_NHWC) && p.C[0] % 4 == 0 && p.C[1] % 4 == 0;
  if (p.Cout <= 16) {
    dim3 grid((unsigned)ceil_div(p.M, 256), (unsigned)ceil_div(p.Cout, 16));
    if (vec)
      conv_simt_kernel<AT, 256, 16, true><<<grid, 256, 0, stream>>>(p);
    else
      conv_simt_kernel<AT, 256, 16, false><<<grid, 256, 0, stream>>>(p);
  } else {
    dim3 grid((unsigned)ceil_div(p.M, 64), (unsigned)ceil_div(p.Cout, 64));
    if (vec)
      conv_simt_kernel<AT, 64, 64, true><<<grid, 256, 0, stream>>>(p);
    else
      conv_simt_kernel<AT, 64, 64, false><<<grid, 256, 0, stream>>>(p);
  }
  return after_launch("conv_simt_kernel");
}

}  // namespace

int conv_simt(const ConvOp& op, int prec, cudaStream_t stream) {
  if (!op.w_f32) HSIDM_FAIL(HSIDM_BAD_STATE, "conv_simt: fp32 weights were not packed");
  if (op.ksize != 1 && op.ksize != 3) HSIDM_FAIL(HSIDM_UNSUPPORTED_CFG, "conv_simt: kernel size %d", op.ksize);
  if (op.rsrc[0].C || op.rsrc[1].C) HSIDM_FAIL(HSIDM_UNSUPPORTED_CFG, "conv_simt: fused shortcut sources are a tensor-core feature");
  SimtP p;
  for (int i = 0; i < 2; ++i) {
    p.s[i] = op.src[i].p;
    p.off[i] = op.src[i].img_off;
    p.C[i] = op.src[i].C;
    p.lay[i] = op.src[i].layout;
  }
  p.N = op.N, p.Hin = op.Hin, p.Win = op.Win, p.up = op.up, p.ks = op.ksize, p.stride = op.stride;
  p.Hout = op.Hout, p.Wout = op.Wout, p.Cout = op.Cout;
  p.w = op.w_f32, p.bias = op.bias, p.nbias = op.nbias, p.nbs = op.nbias_stride;
  p.nb_t = op.nbias_t, p.nb_ts = op.nbias_t_stride;
  p.act = op.act, p.scale = op.scale, p.resid = op.resid, p.out = op.out, p.out_layout = op.out_layout;
  p.clamp01 = op.clamp01;
  p.Ctot = op.src[0].C + op.src[1].C;
  p.Ktot = op.K();
  int64_t M = (int64_t)op.N * op.Hout * op.Wout;
  if (M <= 0 || M > INT32_MAX) HSIDM_FAIL(HSIDM_BAD_SHAPE, "conv_simt: M=%lld out of range", (long long)M);
  p.M = (int)M;
  ProfScope prof(PROF_CONV_SIMT, 2.0 * (double)M * op.Cout * p.Ktot, stream);
  return prec == HSIDM_BF16 ? launch<bf16>(p, stream) : launch<float>(p, stream);
}

}  // namespace hsidm
