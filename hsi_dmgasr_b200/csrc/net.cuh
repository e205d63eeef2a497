// Shared host-side machinery of the two network executors (UNet in unet.cu, GAE in gae.cu):
// a named fp32 parameter store fed by *_set_param, convolution weight packing, and the conv dispatcher that
// routes each ConvOp to the tensor-core kernel (BF16 mode, shapes permitting) or the CUDA-core kernel.
#pragma once
#include <string>
#include <vector>

#include "kernels.cuh"

namespace hsidm {

struct Param {
  std::string key;
  std::vector<int64_t> shape;
  float* dev = nullptr;
  bool set = false;
  int64_t numel() const {
    int64_t n = 1;
    for (auto d : shape) n *= d;
    return n;
  }
};

class ParamStore {
 public:
  ~ParamStore();
  int add(const std::string& key, std::vector<int64_t> shape);  // returns index
  int find(const std::string& key) const;
  int set(const char* key, const float* data, const int64_t* shape, int ndim);
  int check_all_set() const;
  int alloc_all();  // one device slab for every parameter
  const float* dev(int idx) const { return idx < 0 ? nullptr : params_[idx].dev; }
  int size() const { return (int)params_.size(); }
  const Param& at(int i) const { return params_[i]; }
  int64_t bytes() const { return slab_bytes_; }
  // Bitwise comparison of the stored fp32 master copies with the caller's current tensors: `table_dev` is a DEVICE array
  // of size() device pointers in parameter order.  *changed = 1 if any element differs.  Synchronises `stream`.
  int differs(const void* const* table_dev, int n, int* changed, cudaStream_t stream);

 private:
  std::vector<Param> params_;
  float* slab_ = nullptr;
  int64_t slab_bytes_ = 0;
  int64_t* cmp_meta_ = nullptr;   // device [size()][2] = (slab offset in floats, numel)
  int* cmp_flag_ = nullptr;       // device flag
};

// One convolution's parameters and packed copies.
struct ConvW {
  int pw = -1, pb = -1;  // indices into the ParamStore (weight [Cout,Cin,k,k], bias [Cout])
  int Cin = 0, Cout = 0, ks = 3;
  float* w_f32 = nullptr;  // [k*k*Cin][Cout]
  bf16* w_bf16 = nullptr;  // [rows >= Cout][k*k*Cin], zero padded rows
  bf16* w_up = nullptr;    // [rows >= Cout][4 parities][4 taps][Cin]: pre-summed weights of (nearest-2x upsample -> 3x3 conv)
  bf16* w_s2 = nullptr;    // [rows >= Cout][9*Cin] in the consumption order of the stride-2 (phase lattice) halo form
  bf16* w_col = nullptr;   // [rows >= Cout][64]: the same weights for the im2col route of tiny-Cin 3x3 convs (9*Cin <= 64)
  int64_t packed_bytes = 0;
  const float* bias_override = nullptr;  // used instead of the parameter bias when set (fused weights)
};

// conv2 of a res block with its shortcut folded in as extra K columns: W = [ W_conv2 (k = tap*C + c) | W_shortcut ],
// W_shortcut = res_conv weight, or the identity when the block has no res_conv; bias = b_conv2 (+ b_res_conv).
struct FusedW {
  bf16* w = nullptr;     // [rows >= Cout][9*C + Cr]
  float* bias = nullptr; // [Cout]
  int C = 0, Cr = 0, Cout = 0;
  int64_t bytes = 0;
};
int pack_fused(const ParamStore& ps, const ConvW& c2, const ConvW* rc, FusedW& f);
void free_fused(FusedW& f);

// SelfAttention with its projections folded (BF16 mode, n_head = 1): see pack_attn_fold in net.cu.
struct AttnFoldW {
  bf16* w_qk = nullptr;   // [rows >= C][C] = Wk^T Wq : scores = (Xn w_qk^T) Xn^T
  bf16* w_ov = nullptr;   // [rows >= C][C] = Wout Wv : out = (P Xn) w_ov^T + b_out
  int C = 0;
  int64_t bytes = 0;
};
int pack_attn_fold(const ParamStore& ps, const ConvW& qkv, const ConvW& aout, AttnFoldW& f);
void free_attn_fold(AttnFoldW& f);

// Registers "<prefix>.weight" / "<prefix>.bias" and returns the ConvW.
ConvW make_conv(ParamStore& ps, const std::string& prefix, int Cin, int Cout, int ks, bool bias = true);
// (Re)packs w_f32 always and w_bf16 when `bf16_too`. Idempotent; frees previous packs.
int pack_conv(const ParamStore& ps, ConvW& c, bool bf16_too);
int pack_conv_s2(const ParamStore& ps, ConvW& c);   // extra pack for Downsample convs (stride-2 halo form), after pack_conv
int pack_conv_up(const ParamStore& ps, ConvW& c);   // extra pack for Upsample convs (sub-pixel form), after pack_conv
void free_conv(ConvW& c);

// Execution context shared by both executors.
struct Exec {
  int prec = HSIDM_F32;
  Arena arena;
  bool dry = false;
  cudaStream_t stream = nullptr;
  int status = HSIDM_OK;  // first failure (sticky within one pass)
  unsigned* tickets = nullptr;  // GroupNorm ticket counters (zero between kernels)
  int tickets_cap = 0;
  int ensure_tickets(int n);  // grows the counter array (synchronises when it has to)
  ~Exec();

  size_t esize() const { return prec == HSIDM_BF16 ? 2 : 4; }
  Act alloc_act(int N, int H, int W, int C) {
    Act a;
    a.N = N, a.H = H, a.W = W, a.C = C;
    a.p = arena.alloc(a.numel() * (int64_t)esize());
    if (!a.p && status == HSIDM_OK) {
      set_last_error("workspace arena exhausted (%lld bytes requested, capacity %lld)",
                     (long long)(a.numel() * (int64_t)esize()), (long long)arena.capacity());
      status = HSIDM_OOM_WORKSPACE;
    }
    return a;
  }
  void* alloc_raw(int64_t bytes) {
    void* p = arena.alloc(bytes);
    if (!p && status == HSIDM_OK) {
      set_last_error("workspace arena exhausted (%lld bytes requested)", (long long)bytes);
      status = HSIDM_OOM_WORKSPACE;
    }
    return p;
  }
  void release(Act& a) {
    arena.free(a.p);
    arena.free(a.stats);
    a.p = nullptr, a.stats = nullptr, a.slots = 0;
  }
  void release_raw(void* p) { arena.free(p); }
  // run a launcher unless this is a dry (measuring) pass or an earlier step already failed
  template <typename F>
  void run(F&& f) {
    if (dry || status != HSIDM_OK) return;
    int s = f();
    if (s != HSIDM_OK) status = s;
  }
};

// Fills the weight/bias fields of `op` from `w` and dispatches it. In BF16 mode stride-2 and upsampled convs
// whose channels fit the tensor-core kernel are lowered to (im2col | upsample) + tensor-core GEMM.
void run_conv(Exec& ex, ConvOp op, const ConvW& w, const ParamStore& ps);
// Allocates the NHWC output of `proto` (shape fields and source channel counts/layouts filled, pointers irrelevant)
// together with the statistics buffer the kernel chosen for it will fill (none on the CUDA-core path).
Act alloc_conv_out(Exec& ex, ConvOp proto, const ConvW& w, const ParamStore& ps);
// View of images [n0, n0+cnt) of an activation (and of its statistics). Never release a view.
Act act_slice(const Act& a, int n0, int cnt, size_t esize);
// Same, additionally asking the kernel to leave GroupNorm partial statistics of its output in `out` (out.stats / out.slots
// stay null/0 when the chosen kernel cannot produce them, e.g. on the CUDA-core path).
void run_conv_stats(Exec& ex, ConvOp op, const ConvW& w, const ParamStore& ps, Act& out);

}  // namespace hsidm
