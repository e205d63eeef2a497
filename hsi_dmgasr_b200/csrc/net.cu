#include "net.cuh"

#include <cstdlib>
#include <cstring>

namespace hsidm {

// ---- ParamStore ---------------------------------------------------------------------------------------------
ParamStore::~ParamStore() {
  if (slab_) cudaFree(slab_);
  if (cmp_meta_) cudaFree(cmp_meta_);
  if (cmp_flag_) cudaFree(cmp_flag_);
}

namespace {
// block (i, j): elements j*blockDim.x + t, stride gridDim.y*blockDim.x of parameter i, compared as raw 32-bit words
__global__ void params_differ_kernel(const uint32_t* __restrict__ slab, const int64_t* __restrict__ meta,
                                     const void* const* __restrict__ table, int* __restrict__ flag) {
  const int i = blockIdx.x;
  const uint32_t* a = slab + meta[2 * i];
  const uint32_t* b = static_cast<const uint32_t*>(table[i]);
  const int64_t n = meta[2 * i + 1];
  bool diff = false;
  for (int64_t k = blockIdx.y * (int64_t)blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.y * blockDim.x)
    diff |= a[k] != b[k];
  if (__syncthreads_or(diff) && threadIdx.x == 0) atomicOr(flag, 1);
}
}  // namespace

int ParamStore::differs(const void* const* table_dev, int n, int* changed, cudaStream_t stream) {
  if (!table_dev || !changed) HSIDM_FAIL(HSIDM_BAD_ARG, "params_changed: null argument");
  if (n != size()) HSIDM_FAIL(HSIDM_BAD_ARG, "params_changed: table has %d entries, the context has %d parameters", n, size());
  for (auto& p : params_)
    if (!p.set) {   // nothing uploaded yet: everything counts as changed
      *changed = 1;
      return HSIDM_OK;
    }
  if (!cmp_meta_) {
    std::vector<int64_t> meta(2 * params_.size());
    for (size_t i = 0; i < params_.size(); ++i) meta[2 * i] = params_[i].dev - slab_, meta[2 * i + 1] = params_[i].numel();
    HSIDM_CUDA(cudaMalloc(&cmp_meta_, sizeof(int64_t) * meta.size()));
    HSIDM_CUDA(cudaMemcpy(cmp_meta_, meta.data(), sizeof(int64_t) * meta.size(), cudaMemcpyHostToDevice));
    HSIDM_CUDA(cudaMalloc(&cmp_flag_, sizeof(int)));
  }
  HSIDM_CUDA(cudaMemsetAsync(cmp_flag_, 0, sizeof(int), stream));
  params_differ_kernel<<<dim3(n, 16), 256, 0, stream>>>(reinterpret_cast<const uint32_t*>(slab_), cmp_meta_, table_dev, cmp_flag_);
  HSIDM_TRY(after_launch("params_differ_kernel"));
  HSIDM_CUDA(cudaMemcpyAsync(changed, cmp_flag_, sizeof(int), cudaMemcpyDeviceToHost, stream));
  HSIDM_CUDA(cudaStreamSynchronize(stream));
  return HSIDM_OK;
}

int ParamStore::add(const std::string& key, std::vector<int64_t> shape) {
  Param p;
  p.key = key;
  p.shape = std::move(shape);
  params_.push_back(std::move(p));
  return (int)params_.size() - 1;
}

int ParamStore::find(const std::string& key) const {
  for (int i = 0; i < (int)params_.size(); ++i)
    if (params_[i].key == key) return i;
  return -1;
}

int ParamStore::alloc_all() {
  int64_t total = 0;
  for (auto& p : params_) total += round_up(p.numel(), 64);
  slab_bytes_ = total * (int64_t)sizeof(float);
  HSIDM_CUDA(cudaMalloc(&slab_, slab_bytes_));
  HSIDM_CUDA(cudaMemset(slab_, 0, slab_bytes_));
  int64_t off = 0;
  for (auto& p : params_) {
    p.dev = slab_ + off;
    off += round_up(p.numel(), 64);
  }
  return HSIDM_OK;
}

int ParamStore::set(const char* key, const float* data, const int64_t* shape, int ndim) {
  if (!key || !data) HSIDM_FAIL(HSIDM_BAD_ARG, "set_param: null key or data");
  int idx = find(key);
  if (idx < 0) HSIDM_FAIL(HSIDM_BAD_ARG, "set_param: unknown parameter '%s'", key);
  Param& p = params_[idx];
  bool same = ndim == (int)p.shape.size();
  for (int i = 0; same && i < ndim; ++i) same = shape[i] == p.shape[i];
  if (!same) {
    std::string want, got;
    for (auto d : p.shape) want += std::to_string(d) + ",";
    for (int i = 0; i < ndim; ++i) got += std::to_string(shape[i]) + ",";
    HSIDM_FAIL(HSIDM_BAD_SHAPE, "set_param: '%s' expects shape [%s] but got [%s]", key, want.c_str(), got.c_str());
  }
  HSIDM_CUDA(cudaMemcpy(p.dev, data, sizeof(float) * p.numel(), cudaMemcpyDefault));
  p.set = true;
  return HSIDM_OK;
}

int ParamStore::check_all_set() const {
  for (auto& p : params_)
    if (!p.set) HSIDM_FAIL(HSIDM_BAD_STATE, "commit: parameter '%s' was never set", p.key.c_str());
  return HSIDM_OK;
}

Exec::~Exec() {
  if (tickets) cudaFree(tickets);
}

int Exec::ensure_tickets(int n) {
  if (n <= tickets_cap) return HSIDM_OK;
  if (tickets) {
    HSIDM_CUDA(cudaDeviceSynchronize());
    cudaFree(tickets);
    tickets = nullptr;
  }
  const int cap = (int)round_up(n, 256);
  HSIDM_CUDA(cudaMalloc(&tickets, sizeof(unsigned) * cap));
  HSIDM_CUDA(cudaMemset(tickets, 0, sizeof(unsigned) * cap));
  tickets_cap = cap;
  return HSIDM_OK;
}

// ---- conv weights -------------------------------------------------------------------------------------------
ConvW make_conv(ParamStore& ps, const std::string& prefix, int Cin, int Cout, int ks, bool bias) {
  ConvW c;
  c.Cin = Cin, c.Cout = Cout, c.ks = ks;
  c.pw = ps.add(prefix + ".weight", {Cout, Cin, ks, ks});
  if (bias) c.pb = ps.add(prefix + ".bias", {Cout});
  return c;
}

namespace {
// w[co][ci][tap] -> f32[(tap*Cin+ci)*Cout + co] ; bf16[co*K + tap*Cin + ci]
__global__ void pack_kernel(const float* __restrict__ w, int Cout, int Cin, int taps, float* __restrict__ o32,
                            bf16* __restrict__ o16) {
  const int64_t total = (int64_t)Cout * Cin * taps;
  const int K = Cin * taps;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int tap = (int)(i % taps);
    const int ci = (int)((i / taps) % Cin);
    const int co = (int)(i / ((int64_t)taps * Cin));
    const float v = w[i];
    o32[((int64_t)tap * Cin + ci) * Cout + co] = v;
    if (o16) o16[(int64_t)co * K + tap * Cin + ci] = __float2bfloat16_rn(v);
  }
}
// Sub-pixel form of (nearest-2x upsample -> 3x3 conv, unet.py:58-65): output pixel (2i+py, 2j+px) only ever sees a 2x2
// block of source pixels, so the 3x3 taps that land on the same source pixel are summed (in fp32) once, here.
// o[co][((py*2+px)*4 + a*2+b)*Cin + ci] = sum_{dy in D(py,a)} sum_{dx in D(px,b)} w[co][ci][dy][dx],
// D(0,0)={0}, D(0,1)={1,2}, D(1,0)={0,1}, D(1,1)={2}.
__global__ void pack_up_kernel(const float* __restrict__ w, int Cout, int Cin, bf16* __restrict__ o) {
  const int64_t total = (int64_t)Cout * 16 * Cin;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int ci = (int)(i % Cin);
    const int t = (int)((i / Cin) % 16);
    const int co = (int)(i / ((int64_t)16 * Cin));
    const int py = t >> 3, px = (t >> 2) & 1, a = (t >> 1) & 1, b = t & 1;
    const int y0 = py == 0 ? (a == 0 ? 0 : 1) : (a == 0 ? 0 : 2), y1 = py == 0 ? (a == 0 ? 0 : 2) : (a == 0 ? 1 : 2);
    const int x0 = px == 0 ? (b == 0 ? 0 : 1) : (b == 0 ? 0 : 2), x1 = px == 0 ? (b == 0 ? 0 : 2) : (b == 0 ? 1 : 2);
    float acc = 0.f;
    for (int dy = y0; dy <= y1; ++dy)
      for (int dx = x0; dx <= x1; ++dx) acc += w[((int64_t)co * Cin + ci) * 9 + dy * 3 + dx];
    o[i] = __float2bfloat16_rn(acc);
  }
}

// Stride-2 halo form: k-blocks in the order the kernel consumes them - for phase (py,px) in (1,1) (1,0) (0,1) (0,0), for
// each 64-channel block, for each tap of that phase (lattice offsets in row-major order):
//   phase parity 1 -> kernel rows/cols {0, 2} (offsets -1, 0), parity 0 -> {1} (offset 0).
__global__ void pack_s2_kernel(const float* __restrict__ w, int Cout, int Cin, bf16* __restrict__ o) {
  const int cpp = Cin / 64;
  const int64_t K = 9LL * Cin, total = (int64_t)Cout * K;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int co = (int)(i / K);
    int kb = (int)((i % K) / 64);
    const int c64 = (int)(i % 64);
    // locate (phase, channel block, tap) of k-block kb: phases hold 4, 2, 2, 1 taps per channel block
    const int per_phase[4] = {4, 2, 2, 1};
    int ph = 0;
    while (kb >= per_phase[ph] * cpp) kb -= per_phase[ph] * cpp, ++ph;
    const int cb = kb / per_phase[ph], t = kb % per_phase[ph];
    const int py = ph < 2 ? 1 : 0, px = (ph == 0 || ph == 2) ? 1 : 0;
    const int ny = py ? 2 : 1, nx = px ? 2 : 1;        // taps along y / x for this phase
    const int ty = t / nx, tx = t % nx;
    const int ky = py ? 2 * ty : 1, kx = px ? 2 * tx : 1;
    (void)ny;
    o[i] = __float2bfloat16_rn(w[((int64_t)co * Cin + cb * 64 + c64) * 9 + ky * 3 + kx]);
  }
}

// w[co][ci][tap] -> col[co*64 + tap*Cin + ci], zero padded to 64 columns
__global__ void pack_col_kernel(const float* __restrict__ w, int Cout, int Cin, bf16* __restrict__ o) {
  const int total = Cout * 64;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int co = i / 64, k = i % 64;
    float v = 0.f;
    if (k < 9 * Cin) {
      const int tap = k / Cin, ci = k - tap * Cin;
      v = w[((int64_t)co * Cin + ci) * 9 + tap];
    }
    o[i] = __float2bfloat16_rn(v);
  }
}
}  // namespace

void free_conv(ConvW& c) {
  if (c.w_f32) cudaFree(c.w_f32);
  if (c.w_bf16) cudaFree(c.w_bf16);
  if (c.w_col) cudaFree(c.w_col);
  if (c.w_up) cudaFree(c.w_up);
  if (c.w_s2) cudaFree(c.w_s2);
  c.w_f32 = nullptr, c.w_bf16 = nullptr, c.w_col = nullptr, c.w_up = nullptr, c.w_s2 = nullptr, c.packed_bytes = 0;
}

int pack_conv_s2(const ParamStore& ps, ConvW& c) {
  const int bn = conv_tc_bn_rows(c.Cout);
  if (c.ks != 3 || bn <= 0 || c.Cin % 64 || c.Cout % 64) return HSIDM_OK;
  const int64_t rows = round_up(c.Cout, bn), K = 9LL * c.Cin;
  if (c.w_s2) cudaFree(c.w_s2);
  HSIDM_CUDA(cudaMalloc(&c.w_s2, sizeof(bf16) * rows * K));
  HSIDM_CUDA(cudaMemset(c.w_s2, 0, sizeof(bf16) * rows * K));
  c.packed_bytes += sizeof(bf16) * rows * K;
  pack_s2_kernel<<<(unsigned)std::min<int64_t>(ceil_div(c.Cout * K, 256), 4096), 256>>>(ps.dev(c.pw), c.Cout, c.Cin, c.w_s2);
  return after_launch("pack_s2_kernel");
}

int pack_conv_up(const ParamStore& ps, ConvW& c) {
  const int bn = conv_tc_bn_rows(c.Cout);
  if (c.ks != 3 || bn <= 0 || c.Cin % 64 || c.Cout % 64) return HSIDM_OK;
  const int64_t rows = round_up(c.Cout, bn), K = 16LL * c.Cin;
  if (c.w_up) cudaFree(c.w_up);
  HSIDM_CUDA(cudaMalloc(&c.w_up, sizeof(bf16) * rows * K));
  HSIDM_CUDA(cudaMemset(c.w_up, 0, sizeof(bf16) * rows * K));
  c.packed_bytes += sizeof(bf16) * rows * K;
  pack_up_kernel<<<(unsigned)std::min<int64_t>(ceil_div(c.Cout * K, 256), 4096), 256>>>(ps.dev(c.pw), c.Cout, c.Cin, c.w_up);
  return after_launch("pack_up_kernel");
}

int pack_conv(const ParamStore& ps, ConvW& c, bool bf16_too) {
  free_conv(c);
  const int taps = c.ks * c.ks;
  const int64_t K = (int64_t)taps * c.Cin;
  HSIDM_CUDA(cudaMalloc(&c.w_f32, sizeof(float) * K * c.Cout));
  c.packed_bytes = sizeof(float) * K * c.Cout;
  const int bn = conv_tc_bn_rows(c.Cout);
  if (bf16_too && bn > 0 && c.Cin % 64 == 0) {
    const int64_t rows = round_up(c.Cout, bn);
    HSIDM_CUDA(cudaMalloc(&c.w_bf16, sizeof(bf16) * rows * K));
    HSIDM_CUDA(cudaMemset(c.w_bf16, 0, sizeof(bf16) * rows * K));
    c.packed_bytes += sizeof(bf16) * rows * K;
  }
  if (bf16_too && bn > 0 && c.ks == 3 && 9 * c.Cin <= 64) {
    const int64_t rows = round_up(c.Cout, bn);
    HSIDM_CUDA(cudaMalloc(&c.w_col, sizeof(bf16) * rows * 64));
    HSIDM_CUDA(cudaMemset(c.w_col, 0, sizeof(bf16) * rows * 64));
    pack_col_kernel<<<(unsigned)ceil_div(c.Cout * 64, 256), 256>>>(ps.dev(c.pw), c.Cout, c.Cin, c.w_col);
    HSIDM_TRY(after_launch("pack_col_kernel"));
    c.packed_bytes += sizeof(bf16) * rows * 64;
  }
  const int64_t total = K * c.Cout;
  const int grid = (int)std::min<int64_t>(ceil_div(total, 256), 4096);
  pack_kernel<<<grid, 256>>>(ps.dev(c.pw), c.Cout, c.Cin, taps, c.w_f32, c.w_bf16);
  return after_launch("pack_kernel");
}

namespace {
// out[co][k]: k < 9*C -> conv2 weight (k = tap*C + c); k >= 9*C -> shortcut weight (res_conv, or identity when wr == null)
__global__ void pack_fused_kernel(const float* __restrict__ w2, const float* __restrict__ wr, const float* __restrict__ b2,
                                  const float* __restrict__ br, int Cout, int C, int Cr, bf16* __restrict__ out,
                                  float* __restrict__ bias) {
  const int K = 9 * C + Cr;
  const int64_t total = (int64_t)Cout * K;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int co = (int)(i / K), k = (int)(i % K);
    float v;
    if (k < 9 * C) {
      const int tap = k / C, c = k - tap * C;
      v = w2[((int64_t)co * C + c) * 9 + tap];
    } else {
      const int c = k - 9 * C;
      v = wr ? wr[(int64_t)co * Cr + c] : (c == co ? 1.0f : 0.0f);
    }
    out[i] = __float2bfloat16_rn(v);
    if (k == 0) bias[co] = b2[co] + (br ? br[co] : 0.0f);
  }
}
}  // namespace

namespace {
// Attention weight folding (n_head = 1, unet.py:124-143).  With Xn the normalised input [S][C]:
//   scores = (Xn Wq^T)(Xn Wk^T)^T = (Xn Mqk^T) Xn^T,  Mqk = Wk^T Wq          -> one C x C projection instead of q and k
//   out    = Wout (P (Xn Wv^T))^T + b = Wov (P Xn)^T + b,  Wov = Wout Wv     -> no v projection at all
// which: 0 -> o[co][ci] = sum_r a[r][co] * b[r][ci]  (a = Wk, b = Wq);  1 -> o[co][ci] = sum_r a[co][r] * b[r][ci]  (a = Wout, b = Wv)
__global__ void fold_attn_kernel(const float* __restrict__ a, const float* __restrict__ b, int C, int which, bf16* __restrict__ o) {
  const int co = blockIdx.x, ci = threadIdx.x + blockIdx.y * blockDim.x;
  if (ci >= C) return;
  float acc = 0.f;
  for (int r = 0; r < C; ++r) acc = fmaf(which == 0 ? a[(int64_t)r * C + co] : a[(int64_t)co * C + r], b[(int64_t)r * C + ci], acc);
  o[(int64_t)co * C + ci] = __float2bfloat16_rn(acc);
}
}  // namespace

void free_attn_fold(AttnFoldW& f) {
  if (f.w_qk) cudaFree(f.w_qk);
  if (f.w_ov) cudaFree(f.w_ov);
  f.w_qk = nullptr, f.w_ov = nullptr, f.bytes = 0;
}

int pack_attn_fold(const ParamStore& ps, const ConvW& qkv, const ConvW& aout, AttnFoldW& f) {
  free_attn_fold(f);
  const int C = aout.Cout;
  const int bn = conv_tc_bn_rows(C);
  if (bn <= 0 || C % 64 || qkv.Cout != 3 * C || qkv.Cin != C || aout.Cin != C) return HSIDM_OK;   // not a tensor-core shape: stay unfolded
  const int64_t rows = round_up(C, bn);
  HSIDM_CUDA(cudaMalloc(&f.w_qk, sizeof(bf16) * rows * C));
  HSIDM_CUDA(cudaMalloc(&f.w_ov, sizeof(bf16) * rows * C));
  HSIDM_CUDA(cudaMemset(f.w_qk, 0, sizeof(bf16) * rows * C));
  HSIDM_CUDA(cudaMemset(f.w_ov, 0, sizeof(bf16) * rows * C));
  f.C = C, f.bytes = 2 * (int64_t)sizeof(bf16) * rows * C;
  const float* w = ps.dev(qkv.pw);   // [3C][C]: q rows, k rows, v rows
  const dim3 grid(C, (unsigned)ceil_div(C, 256));
  fold_attn_kernel<<<grid, 256>>>(w + (int64_t)C * C, w, C, 0, f.w_qk);
  fold_attn_kernel<<<grid, 256>>>(ps.dev(aout.pw), w + 2LL * C * C, C, 1, f.w_ov);
  return after_launch("fold_attn_kernel");
}

void free_fused(FusedW& f) {
  if (f.w) cudaFree(f.w);
  if (f.bias) cudaFree(f.bias);
  f.w = nullptr, f.bias = nullptr, f.bytes = 0;
}

int pack_fused(const ParamStore& ps, const ConvW& c2, const ConvW* rc, FusedW& f) {
  free_fused(f);
  f.C = c2.Cin, f.Cout = c2.Cout, f.Cr = rc ? rc->Cin : c2.Cout;
  const int bn = conv_tc_bn_rows(f.Cout);
  if (bn <= 0 || f.C % 64 || f.Cr % 64 || f.Cout % 64) return HSIDM_OK;   // not a tensor-core shape: leave unfused
  const int64_t K = 9LL * f.C + f.Cr, rows = round_up(f.Cout, bn);
  HSIDM_CUDA(cudaMalloc(&f.w, sizeof(bf16) * rows * K));
  HSIDM_CUDA(cudaMemset(f.w, 0, sizeof(bf16) * rows * K));
  HSIDM_CUDA(cudaMalloc(&f.bias, sizeof(float) * f.Cout));
  f.bytes = sizeof(bf16) * rows * K + sizeof(float) * f.Cout;
  const int grid = (int)std::min<int64_t>(ceil_div((int64_t)f.Cout * K, 256), 4096);
  pack_fused_kernel<<<grid, 256>>>(ps.dev(c2.pw), rc ? ps.dev(rc->pw) : nullptr, ps.dev(c2.pb), rc ? ps.dev(rc->pb) : nullptr,
                                   f.Cout, f.C, f.Cr, f.w, f.bias);
  return after_launch("pack_fused_kernel");
}

// ---- dispatcher -----------------------------------------------------------------------------------------------
namespace {
enum Route { R_TC, R_DOWN, R_DOWN_HALO, R_UP, R_UP_SUBPIX, R_COL, R_SIMT };

// Decides how `op` runs and, for the lowered routes, fills `g` with the tensor-core op (its source pointer is patched
// once the temporary exists).
Route plan_conv(const Exec& ex, const ConvOp& op, const ConvW& w, ConvOp* g) {
  const int prec = ex.prec;
  *g = op;
  if (prec == HSIDM_BF16 && w.w_col && op.ksize == 3 && op.stride == 1 && !op.up && op.src[0].layout == L_NCHW_F32 &&
      !op.src[0].img_off && (op.src[1].C == 0 || (op.src[1].layout == L_NCHW_F32 && !op.src[1].img_off)) &&
      9 * (op.src[0].C + op.src[1].C) <= 64) {
    // tiny-Cin first conv (unet.py:196-197): im2col to K = 64, then a 1x1 tensor-core GEMM
    g->src[0] = ConvSrc();
    g->src[0].C = 64, g->src[1] = ConvSrc();
    g->ksize = 1, g->w_bf16 = w.w_col;
    if (conv_tc_supported(*g, prec)) return R_COL;
    *g = op;
  }
  if (prec != HSIDM_BF16 || !w.w_bf16) return R_SIMT;
  if (conv_tc_supported(op, prec)) return R_TC;
  const bool nhwc1 = op.src[0].layout == L_NHWC && op.src[1].C == 0 && op.src[0].C % 64 == 0;
  if (nhwc1 && op.stride == 2 && op.ksize == 3 && !op.up && op.Hin % 2 == 0 && op.Win % 2 == 0) {
    // Downsample (unet.py:68-74), preferred: the halo kernel over the four phase lattices of the input (no im2col)
    static const bool no_s2 = std::getenv("HSIDM_NO_S2HALO") != nullptr;   // A/B switch for profiling runs
    if (w.w_s2 && !no_s2) {
      g->Hin = op.Hout, g->Win = op.Wout, g->stride = 1, g->s2 = 1, g->w_bf16 = w.w_s2;
      if (conv_tc_supported(*g, prec) && conv_halo_ok(*g)) return R_DOWN_HALO;
      *g = op;
    }
    // else: gather the 9 taps once, then a 1x1 tensor-core GEMM with K = 9*C.
    g->src[0].C = 9 * op.src[0].C;
    g->Hin = op.Hout, g->Win = op.Wout, g->stride = 1, g->ksize = 1;
    if (conv_tc_supported(*g, prec)) return R_DOWN;
  } else if (nhwc1 && op.up && op.stride == 1) {
    // Upsample (unet.py:58-65), preferred: four sub-pixel 2x2-tap convs over the low-res source (no upsampled tensor)
    if (w.w_up && op.ksize == 3) {
      g->up = 0, g->up_parity = 0, g->w_bf16 = w.w_up;
      if (conv_tc_supported(*g, prec) && conv_halo_ok(*g)) return R_UP_SUBPIX;
      *g = op;
    }
    // else: materialise the nearest-2x tensor, then the ordinary 3x3 tensor-core conv.
    g->Hin = 2 * op.Hin, g->Win = 2 * op.Win, g->up = 0;
    if (conv_tc_supported(*g, prec)) return R_UP;
  }
  *g = op;
  return R_SIMT;
}

void fill_weights(ConvOp& op, const ConvW& w, const ParamStore& ps) {
  op.w_f32 = w.w_f32;
  op.w_bf16 = w.w_bf16;
  op.bias = w.bias_override ? w.bias_override : ps.dev(w.pb);
  op.ksize = w.ks;
  op.Cout = w.Cout;
}
}  // namespace

void run_conv(Exec& ex, ConvOp op, const ConvW& w, const ParamStore& ps) {
  fill_weights(op, w, ps);
  cudaStream_t st = ex.stream;
  const int prec = ex.prec;
  ConvOp g;
  const Route route = plan_conv(ex, op, w, &g);
  if (op.gn.on() && route != R_TC) {   // only the halo tensor-core kernel normalises its input on the fly
    if (ex.status == HSIDM_OK) {
      set_last_error("run_conv: fused input GroupNorm requested for an op the halo kernel does not take");
      ex.status = HSIDM_UNSUPPORTED_CFG;
    }
    return;
  }
  if (route == R_TC) {
    ex.run([&] { return conv_tc(op, st); });
  } else if (route == R_DOWN_HALO) {
    ex.run([&] { return conv_tc(g, st); });
  } else if (route == R_DOWN) {
    const int C = op.src[0].C;
    Act col = ex.alloc_act(op.N, op.Hout, op.Wout, 9 * C);
    const void* src = op.src[0].p;
    ex.run([&] { return im2col_s2(src, col.p, op.N, op.Hin, op.Win, C, st); });
    g.src[0].p = col.p;
    ex.run([&] { return conv_tc(g, st); });
    ex.release(col);
  } else if (route == R_UP_SUBPIX) {
    for (int parity = 0; parity < 4; ++parity) {
      g.up_parity = parity;
      ex.run([&] { return conv_tc(g, st); });
    }
  } else if (route == R_COL) {
    Act col = ex.alloc_act(op.N, op.Hin, op.Win, 64);
    const float* s0 = static_cast<const float*>(op.src[0].p);
    const float* s1 = static_cast<const float*>(op.src[1].p);
    ex.run([&] { return im2col_small(s0, op.src[0].C, s1, op.src[1].C, col.p, op.N, op.Hin, op.Win, st); });
    g.src[0].p = col.p;
    ex.run([&] { return conv_tc(g, st); });
    ex.release(col);
  } else if (route == R_UP) {
    const int C = op.src[0].C;
    Act big = ex.alloc_act(op.N, 2 * op.Hin, 2 * op.Win, C);
    const void* src = op.src[0].p;
    ex.run([&] { return upsample2x(src, big.p, op.N, op.Hin, op.Win, C, prec, st); });
    g.src[0].p = big.p;
    ex.run([&] { return conv_tc(g, st); });
    ex.release(big);
  } else {
    op.stats_out = nullptr, op.stats_slots = 0;   // the CUDA-core kernel has no fused statistics
    ex.run([&] { return conv_simt(op, prec, st); });
  }
}

Act alloc_conv_out(Exec& ex, ConvOp proto, const ConvW& w, const ParamStore& ps) {
  fill_weights(proto, w, ps);
  Act out = ex.alloc_act(proto.N, proto.Hout, proto.Wout, w.Cout);
  ConvOp g;
  const Route route = plan_conv(ex, proto, w, &g);
  static const bool no_stats = std::getenv("HSIDM_NO_STATS") != nullptr;   // A/B switch for profiling
  int slots = 0;
  if (!no_stats && route != R_SIMT) slots = conv_tc_stats_slots(route == R_TC ? proto : g);
  if (slots > 0) {
    out.stats = static_cast<float*>(ex.alloc_raw(sizeof(float) * 2 * (int64_t)proto.N * slots * w.Cout));
    out.slots = slots;
  }
  return out;
}

Act act_slice(const Act& a, int n0, int cnt, size_t esize) {
  Act s = a;
  s.N = cnt;
  s.p = static_cast<char*>(a.p) + (int64_t)n0 * a.H * a.W * a.C * (int64_t)esize;
  if (a.stats) s.stats = a.stats + (int64_t)n0 * a.slots * a.C * 2;
  return s;
}

void run_conv_stats(Exec& ex, ConvOp op, const ConvW& w, const ParamStore& ps, Act& out) {
  fill_weights(op, w, ps);
  ConvOp g;
  const Route route = plan_conv(ex, op, w, &g);
  int slots = 0;
  static const bool no_stats = std::getenv("HSIDM_NO_STATS") != nullptr;   // A/B switch for profiling
  if (!no_stats && route != R_SIMT && op.out_layout == L_NHWC) slots = conv_tc_stats_slots(route == R_TC ? op : g);
  if (slots > 0) {
    out.stats = static_cast<float*>(ex.alloc_raw(sizeof(float) * 2 * (int64_t)op.N * slots * op.Cout));
    out.slots = slots;
    op.stats_out = out.stats, op.stats_slots = slots;
  }
  run_conv(ex, op, w, ps);
}

}  // namespace hsidm
