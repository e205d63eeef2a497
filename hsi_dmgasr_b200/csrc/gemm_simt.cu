// Batched CUDA-core GEMM used by the F32-mode self-attention (unet.py:133-140: the two einsums).
// C[b] = alpha * A[b] * op(B[b]);  64x64x16 tiles, 256 threads, 4x4 outputs per thread, fp32 accumulate.
#include "kernels.cuh"

namespace hsidm {
namespace {

struct GemmP {
  const void* A;
  const void* B;
  void* C;
  int M, N, K;
  int64_t lda, ldb, ldc, sA, sB, sC;
  int transB, a_f32, c_f32;
  float alpha;
};

template <typename AT>
__global__ void __launch_bounds__(256) gemm_simt_kernel(const GemmP p) {
  __shared__ __align__(16) float As[16][64 + 4];
  __shared__ __align__(16) float Bs[16][64 + 4];
  const int b = blockIdx.z;
  const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
  const int m0 = blockIdx.x * 64, n0 = blockIdx.y * 64;
  const AT* Ab = static_cast<const AT*>(p.A) + (p.a_f32 ? 0 : b * p.sA);
  const float* Af = static_cast<const float*>(p.A) + (p.a_f32 ? b * p.sA : 0);
  const AT* Bb = static_cast<const AT*>(p.B) + b * p.sB;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < p.K; k0 += 16) {
    {  // A: 64 rows x 16 k ; thread -> row tid/4, k (tid%4)*4..+3
      const int r = tid / 4, kq = (tid % 4) * 4;
      const int m = m0 + r;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int k = k0 + kq + j;
        float v = 0.f;
        if (m < p.M && k < p.K) v = p.a_f32 ? Af[(int64_t)m * p.lda + k] : to_f32(Ab[(int64_t)m * p.lda + k]);
        As[kq + j][r] = v;
      }
    }
    if (p.transB) {  // B is [N][K]: thread -> col tid/4, k (tid%4)*4..+3
      const int c = tid / 4, kq = (tid % 4) * 4;
      const int n = n0 + c;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int k = k0 + kq + j;
        Bs[kq + j][c] = (n < p.N && k < p.K) ? to_f32(Bb[(int64_t)n * p.ldb + k]) : 0.f;
      }
    } else {  // B is [K][N]: thread -> k tid/16, cols (tid%16)*4..+3
      const int kk = tid / 16, cq = (tid % 16) * 4;
      const int k = k0 + kk;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int n = n0 + cq + j;
        Bs[kk][cq + j] = (n < p.N && k < p.K) ? to_f32(Bb[(int64_t)k * p.ldb + n]) : 0.f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w};
      const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= p.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= p.N) continue;
      const float v = acc[i][j] * p.alpha;
      const int64_t o = b * p.sC + (int64_t)m * p.ldc + n;
      if (p.c_f32)
        static_cast<float*>(p.C)[o] = v;
      else
        static_cast<AT*>(p.C)[o] = from_f32<AT>(v);
    }
  }
}

}  // namespace

int gemm_simt(const GemmOp& op, int prec, cudaStream_t stream) {
  GemmP p{op.A, op.B, op.C, op.M, op.N, op.K, op.lda, op.ldb, op.ldc, op.sA, op.sB, op.sC, op.transB, op.a_f32, op.c_f32,
          op.alpha};
  dim3 grid((unsigned)ceil_div(op.M, 64), (unsigned)ceil_div(op.N, 64), op.batch);
  ProfScope prof(PROF_GEMM, 2.0 * op.M * (double)op.N * op.K * op.batch, stream);
  if (prec == HSIDM_BF16)
    gemm_simt_kernel<bf16><<<grid, 256, 0, stream>>>(p);
  else
    gemm_simt_kernel<float><<<grid, 256, 0, stream>>>(p);
  return after_launch("gemm_simt_kernel");
}

}  // namespace hsidm
