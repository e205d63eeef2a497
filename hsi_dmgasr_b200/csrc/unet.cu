// SR3 UNet executor + diffusion sampler behind the hsidm_ctx C ABI.
//
// The layer list is rebuilt from the constructor arguments exactly as UNet.__init__ does
// (model/sr3_modules/unet.py:190-236) and executed as straight-line host code that enqueues kernels on one
// stream; intermediate activations are NHWC in the context's element type and live in a stream-ordered arena.
// hsidm_sample captures one (UNet forward + posterior step) into a CUDA graph and replays it T times; the
// timestep lives in device memory so the graph is identical for every step.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "unet.cuh"

namespace {

constexpr float kGnEps = 1e-5f;

int build_layers(hsidm_ctx* c) {
  const hsidm_unet_cfg& g = c->cfg;
  ParamStore& ps = c->ps;
  const int ic = g.inner_channel;
  c->mlp1_w = ps.add("noise_level_mlp.1.weight", {4 * ic, ic});
  c->mlp1_b = ps.add("noise_level_mlp.1.bias", {4 * ic});
  c->mlp3_w = ps.add("noise_level_mlp.3.weight", {ic, 4 * ic});
  c->mlp3_b = ps.add("noise_level_mlp.3.bias", {ic});

  auto in_attn = [&](int res) {
    for (int i = 0; i < g.n_attn_res; ++i)
      if (g.attn_res[i] == res) return true;
    return false;
  };
  auto make_res = [&](const std::string& name, int cin, int cout, bool attn, int skip) {
    LayerW L;
    L.kind = LayerW::RES;
    ResW& r = L.rb;
    r.cin = cin, r.cout = cout, r.attn = attn, r.skip = skip, r.has_res = cin != cout;
    const std::string p = name + ".res_block";
    r.nf_w = ps.add(p + ".noise_func.noise_func.0.weight", {cout, ic});
    r.nf_b = ps.add(p + ".noise_func.noise_func.0.bias", {cout});
    r.noise_off = c->noise_total;
    c->noise_total += cout;
    r.gn1_w = ps.add(p + ".block1.block.0.weight", {cin});
    r.gn1_b = ps.add(p + ".block1.block.0.bias", {cin});
    r.c1 = make_conv(ps, p + ".block1.block.3", cin, cout, 3);
    r.gn2_w = ps.add(p + ".block2.block.0.weight", {cout});
    r.gn2_b = ps.add(p + ".block2.block.0.bias", {cout});
    r.c2 = make_conv(ps, p + ".block2.block.3", cout, cout, 3);
    if (r.has_res) r.rc = make_conv(ps, p + ".res_conv", cin, cout, 1);
    if (attn) {
      const std::string a = name + ".attn";
      r.an_w = ps.add(a + ".norm.weight", {cout});
      r.an_b = ps.add(a + ".norm.bias", {cout});
      r.qkv = make_conv(ps, a + ".qkv", cout, 3 * cout, 1, /*bias=*/false);
      r.aout = make_conv(ps, a + ".out", cout, cout, 1);
    }
    return L;
  };

  int pre = ic, now = g.image_size;
  std::vector<int> feat{pre};
  {
    LayerW L;
    L.kind = LayerW::CONV;
    L.conv = make_conv(ps, "downs.0", g.in_channel, ic, 3);
    c->downs.push_back(L);
  }
  for (int lev = 0; lev < g.n_mults; ++lev) {
    const int cm = ic * g.channel_mults[lev];
    for (int b = 0; b < g.res_blocks; ++b) {
      c->downs.push_back(make_res("downs." + std::to_string(c->downs.size()), pre, cm, in_attn(now), 0));
      pre = cm;
      feat.push_back(pre);
    }
    if (lev + 1 < g.n_mults) {
      LayerW L;
      L.kind = LayerW::DOWN;
      L.conv = make_conv(ps, "downs." + std::to_string(c->downs.size()) + ".conv", pre, pre, 3);
      c->downs.push_back(L);
      feat.push_back(pre);
      now /= 2;
    }
  }
  c->mid.push_back(make_res("mid.0", pre, pre, true, 0));
  c->mid.push_back(make_res("mid.1", pre, pre, false, 0));
  for (int lev = g.n_mults - 1; lev >= 0; --lev) {
    const int cm = ic * g.channel_mults[lev];
    for (int b = 0; b < g.res_blocks + 1; ++b) {
      const int sk = feat.back();
      feat.pop_back();
      c->ups.push_back(make_res("ups." + std::to_string(c->ups.size()), pre + sk, cm, in_attn(now), sk));
      pre = cm;
    }
    if (lev > 0) {
      LayerW L;
      L.kind = LayerW::UP;
      L.conv = make_conv(ps, "ups." + std::to_string(c->ups.size()) + ".conv", pre, pre, 3);
      c->ups.push_back(L);
      now *= 2;
    }
  }
  c->fin_gn_w = ps.add("final_conv.block.0.weight", {pre});
  c->fin_gn_b = ps.add("final_conv.block.0.bias", {pre});
  c->fin_conv = make_conv(ps, "final_conv.block.3", pre, g.out_channel, 3);
  return HSIDM_OK;
}

template <typename F>
void for_each_conv(hsidm_ctx* c, F&& f) {
  auto visit = [&](std::vector<LayerW>& v) {
    for (auto& L : v) {
      if (L.kind != LayerW::RES) {
        f(L.conv);
        continue;
      }
      f(L.rb.c1);
      f(L.rb.c2);
      if (L.rb.has_res) f(L.rb.rc);
      if (L.rb.attn) f(L.rb.qkv), f(L.rb.aout);
    }
  };
  visit(c->downs);
  visit(c->mid);
  visit(c->ups);
  f(c->fin_conv);
}

// ---- forward ------------------------------------------------------------------------------------------------
struct NoiseRef {
  const float* base;   // [.., noise_total]
  int64_t n_stride;    // per image
  const int* t_dev;    // optional device timestep
  int64_t t_stride;
};

// GroupNorm(weight gw, bias gb) over cat(x, skip) described for a consumer that evaluates the statistics itself from the
// per-slot partial sums the producers left (no kernel in between).  False when a tensor carries no partial sums or the
// grouping does not fit the consumers' fold (odd channels per group): the caller then computes statistics first.
bool gn_from_slots(hsidm_ctx* c, const Act& x, const Act* skip, int gw, int gb, bool swish, GnIn* gn) {
  const int C1 = skip ? skip->C : 0, groups = c->cfg.norm_groups, Ct = x.C + C1;
  *gn = GnIn();
  gn->C[0] = x.C, gn->C[1] = C1, gn->groups = groups, gn->eps = kGnEps, gn->swish = swish ? 1 : 0;
  gn->gamma = c->ps.dev(gw), gn->beta = c->ps.dev(gb);
  if ((conv_tc_variant() & 256) || !x.stats || (skip && !skip->stats) || x.C % 64 || C1 % 64 || Ct % groups || ((Ct / groups) & 1) || groups > 128)
    return false;
  gn->part[0] = x.stats, gn->slots[0] = x.slots;
  if (skip) gn->part[1] = skip->stats, gn->slots[1] = skip->slots;
  return true;
}

// GroupNorm(+Swish) over cat(a, b) -> new tensor with a.C + b.C channels.  Statistics come from the partial sums the
// producing convolutions left behind when both inputs carry them (folded inside the apply kernel); otherwise from a
// finalize launch or a pass over the tensors.
Act gn_act(hsidm_ctx* c, const Act& a, const Act* b, int gw, int gb, bool swish) {
  Exec& ex = c->ex;
  const int C1 = b ? b->C : 0;
  const int groups = c->cfg.norm_groups;
  const int HW = a.H * a.W;
  const void* p1 = b ? b->p : nullptr;
  GnIn gn;
  if (gn_from_slots(c, a, b, gw, gb, swish, &gn)) {
    Act out = ex.alloc_act(a.N, a.H, a.W, a.C + C1);
    ex.run([&] { return gn_apply_fused(a.p, p1, a.N, HW, gn, out.p, ex.prec, ex.stream); });
    return out;
  }
  float* stats = static_cast<float*>(ex.alloc_raw(sizeof(float) * 2 * a.N * groups));
  Act out = ex.alloc_act(a.N, a.H, a.W, a.C + C1);
  const bool fused = a.stats && (!b || b->stats) && a.C % 64 == 0 && C1 % 64 == 0;
  if (fused) {
    const float* s1 = b ? b->stats : nullptr;
    const int sl1 = b ? b->slots : 0;
    ex.run([&] { return gn_finalize(a.stats, a.slots, a.C, s1, sl1, C1, a.N, HW, groups, kGnEps, stats, ex.stream); });
  } else {
    void* scratch = ex.alloc_raw(gn_scratch_bytes(a.C, C1, a.N, HW, groups));
    ex.run([&] {
      return gn_stats(a.p, a.C, p1, C1, a.N, HW, groups, kGnEps, scratch, ex.tickets, stats, ex.prec, ex.stream);
    });
    ex.release_raw(scratch);
  }
  ex.run([&] {
    return gn_apply(a.p, a.C, p1, C1, a.N, HW, groups, stats, c->ps.dev(gw), c->ps.dev(gb), swish ? 1 : 0, out.p, ex.prec,
                    ex.stream);
  });
  ex.release_raw(stats);
  return out;
}

// Conv descriptor with NHWC sources; the output (and its statistics buffer, if it has one) is `out`.
ConvOp conv_op_nhwc(const Act& in, const Act* in2, const Act& out) {
  ConvOp op;
  op.src[0].p = in.p, op.src[0].C = in.C;
  if (in2) op.src[1].p = in2->p, op.src[1].C = in2->C;
  op.N = in.N, op.Hin = in.H, op.Win = in.W;
  op.Hout = out.H, op.Wout = out.W;
  op.out = out.p;
  op.stats_out = out.stats, op.stats_slots = out.slots;
  return op;
}

// Output tensor of a conv over a source of shape (N, Hin, Win, C0+C1) with the given geometry.
Act conv_out(hsidm_ctx* c, const ConvW& w, int N, int Hin, int Win, int C0, int C1, int stride = 1, int up = 0) {
  ConvOp proto;
  proto.src[0].C = C0, proto.src[1].C = C1;
  proto.N = N, proto.Hin = Hin, proto.Win = Win, proto.stride = stride, proto.up = up;
  const int He = up ? 2 * Hin : Hin, We = up ? 2 * Win : Win;
  proto.Hout = stride == 2 ? (He + 1) / 2 : He, proto.Wout = stride == 2 ? (We + 1) / 2 : We;
  return alloc_conv_out(c->ex, proto, w, c->ps);
}

// SelfAttention.forward (unet.py:124-143), n_head = 1, with the projections folded at commit (pack_attn_fold):
//   Xn = GroupNorm(x) written as [S][C] and as [C][S];  T = Xn Mqk^T;  P = softmax(T Xn^T / sqrt(C));  Y = P Xn;
//   out = Y Wov^T + b_out + x.     5 launches, 0.40 GFLOP per image instead of 0.67 (no q/k/v tensors).
// Needs the GroupNorm affine of x (from the producer's tail, or from gn_finalize when x carries slot statistics).
bool attention_folded(hsidm_ctx* c, const ResW& r, const Act& x, Act& out) {
  Exec& ex = c->ex;
  const int C = x.C, S = x.H * x.W, N = x.N;
  static const bool off = std::getenv("HSIDM_NO_ATTNFOLD") != nullptr;   // A/B switch
  if (off || (conv_tc_variant() & 512) || ex.prec != HSIDM_BF16 || !r.afold.w_qk || C % 64 || S % 64) return false;
  if (!(S == 64 || S == 128 || S == 256)) return false;   // the softmax runs in the scores GEMM's epilogue (one tile per row)
  ConvW wqk, wov;
  wqk.Cin = C, wqk.Cout = C, wqk.ks = 1, wqk.w_bf16 = r.afold.w_qk;
  wov = r.aout, wov.w_bf16 = r.afold.w_ov, wov.w_f32 = nullptr;
  {
    ConvOp t;
    t.src[0].C = C, t.N = N, t.Hin = t.Hout = x.H, t.Win = t.Wout = x.W, t.ksize = 1, t.Cout = C, t.w_bf16 = wqk.w_bf16;
    if (!conv_tc_supported(t, ex.prec)) return false;
    GemmTcOp probe;
    probe.M = S, probe.N = S, probe.K = C, probe.lda = C, probe.ldb = C, probe.ldc = S, probe.row_softmax = 1, probe.alpha = 1.f;
    probe.sA = probe.sB = (int64_t)S * C, probe.sC = (int64_t)S * S;
    if (!gemm_tc_supported(probe)) return false;
  }
  GnIn gn;
  float* stats = nullptr;
  if (!gn_from_slots(c, x, nullptr, r.an_w, r.an_b, false, &gn)) {
    if (!x.stats) return false;
    stats = static_cast<float*>(ex.alloc_raw(sizeof(float) * 2 * N * gn.groups));
    ex.run([&] { return gn_finalize(x.stats, x.slots, C, nullptr, 0, 0, N, S, gn.groups, kGnEps, stats, ex.stream); });
    gn.stats = stats;
  }
  Act nrm = ex.alloc_act(N, x.H, x.W, C);
  bf16* xt = static_cast<bf16*>(ex.alloc_raw(sizeof(bf16) * (int64_t)N * C * S));   // [N][C][S]
  ex.run([&] { return gn_apply_transposed(x.p, N, S, gn, nrm.p, xt, ex.stream); });
  if (stats) ex.release_raw(stats);
  Act t = ex.alloc_act(N, x.H, x.W, C);
  run_conv(ex, conv_op_nhwc(nrm, nullptr, t), wqk, c->ps);
  Act y = ex.alloc_act(N, x.H, x.W, C);
  AttnFlashOp fa;
  fa.Q = t.p, fa.K = nrm.p, fa.Vt = xt, fa.Y = y.p, fa.S = S, fa.Ck = C, fa.Cv = C, fa.batch = N;
  fa.ldq = fa.ldk = fa.ldy = C, fa.sQ = fa.sK = fa.sY = (int64_t)S * C, fa.sVt = (int64_t)C * S;
  fa.alpha = 1.0f / std::sqrt((float)C);
  if (!(conv_tc_variant() & 4096) && attn_flash_supported(fa)) {
    // scores, softmax and P.Xn in one kernel: the scores stay in tensor memory, the probabilities in shared memory
    ex.run([&] { return attn_flash(fa, ex.stream); });
  } else {
    bf16* prob = static_cast<bf16*>(ex.alloc_raw(sizeof(bf16) * (int64_t)N * S * S));
    GemmTcOp sc;
    sc.A = t.p, sc.B = nrm.p, sc.C = prob, sc.M = S, sc.N = S, sc.K = C, sc.batch = N;
    sc.lda = sc.ldb = C, sc.sA = sc.sB = (int64_t)S * C, sc.ldc = S, sc.sC = (int64_t)S * S;
    sc.alpha = fa.alpha, sc.c_f32 = 0, sc.row_softmax = 1;
    ex.run([&] { return gemm_tc(sc, ex.stream); });
    GemmTcOp pv;
    pv.A = prob, pv.B = xt, pv.C = y.p, pv.M = S, pv.N = C, pv.K = S, pv.batch = N;
    pv.lda = S, pv.sA = (int64_t)S * S, pv.ldb = S, pv.sB = (int64_t)C * S, pv.ldc = C, pv.sC = (int64_t)S * C, pv.c_f32 = 0;
    ex.run([&] { return gemm_tc(pv, ex.stream); });
    ex.release_raw(prob);
  }
  ex.release(t);
  ex.release(nrm);
  ex.release_raw(xt);
  ConvOp op = conv_op_nhwc(y, nullptr, out);
  op.resid = x.p;
  run_conv(ex, op, wov, c->ps);
  ex.release(y);
  return true;
}

// The same with q, k, v materialised (A/B reference for the folded form; shapes whose rows do not fit one softmax tile).
// Every contraction is a K-major
// "NT" GEMM: the V projection is computed transposed (V^T = W_v X^T) so that O = P V contracts over contiguous keys.
bool attention_tc(hsidm_ctx* c, const ResW& r, const Act& x, Act& out) {
  Exec& ex = c->ex;
  const int C = x.C, S = x.H * x.W, N = x.N;
  if (ex.prec != HSIDM_BF16 || !r.qkv.w_bf16) return false;
  GemmTcOp probe;
  probe.M = S, probe.N = S, probe.K = C, probe.lda = 2 * C, probe.ldb = 2 * C, probe.ldc = S, probe.c_f32 = 1;
  probe.sA = probe.sB = (int64_t)S * 2 * C, probe.sC = (int64_t)S * S;
  if (!gemm_tc_supported(probe) || C % 64 || S % 64) return false;
  {
    ConvOp t;
    t.src[0].C = C, t.N = N, t.Hin = t.Hout = x.H, t.Win = t.Wout = x.W, t.ksize = 1, t.Cout = 2 * C, t.w_bf16 = r.qkv.w_bf16;
    if (!conv_tc_supported(t, ex.prec)) return false;
  }
  Act nrm = gn_act(c, x, nullptr, r.an_w, r.an_b, false);
  // q, k = first 2C rows of the packed qkv weight
  Act qk = ex.alloc_act(N, x.H, x.W, 2 * C);
  {
    ConvW w2 = r.qkv;
    w2.Cout = 2 * C;
    run_conv(ex, conv_op_nhwc(nrm, nullptr, qk), w2, c->ps);
  }
  const bf16* wv = r.qkv.w_bf16 + (int64_t)2 * C * C;   // rows [2C, 3C) of the K-major [3C][C] matrix
  bf16* qkp = static_cast<bf16*>(qk.p);
  bf16* vt = static_cast<bf16*>(ex.alloc_raw(sizeof(bf16) * (int64_t)N * C * S));        // [N][C][S]
  GemmTcOp g;
  g.A = wv, g.B = nrm.p, g.C = vt, g.M = C, g.N = S, g.K = C, g.batch = N;
  g.lda = C, g.sA = 0, g.ldb = C, g.sB = (int64_t)S * C, g.ldc = S, g.sC = (int64_t)C * S, g.c_f32 = 0;
  ex.run([&] { return gemm_tc(g, ex.stream); });
  // probabilities = softmax(q k^T / sqrt(C)) as bf16: with S = 64, 128 or 256 a row is exactly one accumulator tile and the softmax
  // runs in the GEMM's epilogue; longer sequences go through fp32 scores and the row-softmax kernel
  bf16* prob = static_cast<bf16*>(ex.alloc_raw(sizeof(bf16) * (int64_t)N * S * S));
  GemmTcOp qkT;
  qkT.A = qkp, qkT.B = qkp + C, qkT.M = S, qkT.N = S, qkT.K = C, qkT.batch = N;
  qkT.lda = qkT.ldb = 2 * C, qkT.sA = qkT.sB = (int64_t)S * 2 * C, qkT.ldc = S, qkT.sC = (int64_t)S * S;
  qkT.alpha = 1.0f / std::sqrt((float)C);
  if (S == 64 || S == 128 || S == 256) {
    qkT.C = prob, qkT.c_f32 = 0, qkT.row_softmax = 1;
    ex.run([&] { return gemm_tc(qkT, ex.stream); });
  } else {
    float* scores = static_cast<float*>(ex.alloc_raw(sizeof(float) * (int64_t)N * S * S));  // [N][S][S]
    qkT.C = scores, qkT.c_f32 = 1;
    ex.run([&] { return gemm_tc(qkT, ex.stream); });
    ex.run([&] { return softmax_rows_bf16(scores, prob, (int64_t)N * S, S, ex.stream); });
    ex.release_raw(scores);
  }
  ex.release(nrm);
  ex.release(qk);
  Act av = ex.alloc_act(N, x.H, x.W, C);
  GemmTcOp pv;
  pv.A = prob, pv.B = vt, pv.C = av.p, pv.M = S, pv.N = C, pv.K = S, pv.batch = N;
  pv.lda = S, pv.sA = (int64_t)S * S, pv.ldb = S, pv.sB = (int64_t)C * S, pv.ldc = C, pv.sC = (int64_t)S * C, pv.c_f32 = 0;
  ex.run([&] { return gemm_tc(pv, ex.stream); });
  ex.release_raw(prob);
  ex.release_raw(vt);
  ConvOp op = conv_op_nhwc(av, nullptr, out);
  op.resid = x.p;
  run_conv(ex, op, r.aout, c->ps);
  ex.release(av);
  return true;
}

// Same on CUDA cores (F32 mode, or shapes the tensor-core GEMM does not take).  Writes into `out`.
void attention(hsidm_ctx* c, const ResW& r, const Act& x, Act& out) {
  Exec& ex = c->ex;
  if (attention_folded(c, r, x, out)) return;
  if (attention_tc(c, r, x, out)) return;
  const int C = x.C, S = x.H * x.W;
  Act nrm = gn_act(c, x, nullptr, r.an_w, r.an_b, false);
  Act qkv = ex.alloc_act(x.N, x.H, x.W, 3 * C);
  run_conv(ex, conv_op_nhwc(nrm, nullptr, qkv), r.qkv, c->ps);
  ex.release(nrm);
  float* scores = static_cast<float*>(ex.alloc_raw(sizeof(float) * (int64_t)x.N * S * S));
  Act av = ex.alloc_act(x.N, x.H, x.W, C);
  const size_t es = ex.esize();
  GemmOp qk;
  qk.A = qkv.p, qk.B = static_cast<const char*>(qkv.p) + es * C, qk.C = scores;
  qk.M = S, qk.N = S, qk.K = C, qk.lda = 3 * C, qk.ldb = 3 * C, qk.ldc = S;
  qk.sA = (int64_t)S * 3 * C, qk.sB = qk.sA, qk.sC = (int64_t)S * S;
  qk.batch = x.N, qk.transB = 1, qk.c_f32 = 1, qk.alpha = 1.0f / std::sqrt((float)C);
  ex.run([&] { return gemm_simt(qk, ex.prec, ex.stream); });
  ex.run([&] { return softmax_rows(scores, (int64_t)x.N * S, S, ex.stream); });
  GemmOp pv;
  pv.A = scores, pv.B = static_cast<const char*>(qkv.p) + es * 2 * C, pv.C = av.p;
  pv.M = S, pv.N = C, pv.K = S, pv.lda = S, pv.ldb = 3 * C, pv.ldc = C;
  pv.sA = (int64_t)S * S, pv.sB = (int64_t)S * 3 * C, pv.sC = (int64_t)S * C;
  pv.batch = x.N, pv.transB = 0, pv.a_f32 = 1, pv.c_f32 = 0;
  ex.run([&] { return gemm_simt(pv, ex.prec, ex.stream); });
  ex.release_raw(scores);
  ex.release(qkv);
  ConvOp op = conv_op_nhwc(av, nullptr, out);
  op.resid = x.p;
  run_conv(ex, op, r.aout, c->ps);
  ex.release(av);
}

// conv(w) over GroupNorm(+Swish)(cat(x, skip)) -> out; `fill` sets the op's epilogue fields (noise bias, residual,
// shortcut sources).  When the halo tensor-core kernel takes the op and both inputs carry the partial sums of their
// producers, the normalisation is fused into the conv's load path (the conv reads the raw tensors and applies the
// per-image per-channel affine + Swish to each halo tile in shared memory): the normalised tensor is never written or
// re-read.  Otherwise it goes through gn_act's normalised copy.
template <typename Fill>
void gn_conv(hsidm_ctx* c, const Act& x, const Act* skip, int gw, int gb, bool swish, Act& out, const ConvW& w, Fill&& fill) {
  Exec& ex = c->ex;
  const int C1 = skip ? skip->C : 0;
  static const bool no_fuse = std::getenv("HSIDM_NO_GNFUSE") != nullptr;   // A/B switch for profiling runs
  bool fuse = !no_fuse && !(conv_tc_variant() & 8) && ex.prec == HSIDM_BF16 && x.C % 64 == 0 && C1 % 64 == 0;
  if (fuse) {
    ConvOp probe = conv_op_nhwc(x, skip, out);
    fill(probe);
    probe.w_bf16 = w.w_bf16, probe.Cout = w.Cout, probe.ksize = w.ks;
    static const float kDummy = 0.f;
    probe.gn.gamma = &kDummy, probe.gn.stats = &kDummy;
    fuse = probe.w_bf16 && conv_tc_supported(probe, ex.prec) && conv_halo_ok(probe);
  }
  if (!fuse) {
    Act a = gn_act(c, x, skip, gw, gb, swish);
    ConvOp op = conv_op_nhwc(a, nullptr, out);
    fill(op);
    run_conv(ex, op, w, c->ps);
    ex.release(a);
    return;
  }
  // Statistics: folded by the conv's own transform warps from the producers' partial sums; where those are missing (or
  // variant 256 asks for it) from a gn_finalize launch / a statistics pass over the tensors - the conv then reads (mean, rstd).
  GnIn gn;
  float* stats = nullptr;
  if (!gn_from_slots(c, x, skip, gw, gb, swish, &gn)) {
    const int groups = gn.groups;
    stats = static_cast<float*>(ex.alloc_raw(sizeof(float) * 2 * x.N * groups));
    if (x.stats && (!skip || skip->stats)) {
      const float* s1 = skip ? skip->stats : nullptr;
      const int sl1 = skip ? skip->slots : 0;
      ex.run([&] { return gn_finalize(x.stats, x.slots, x.C, s1, sl1, C1, x.N, x.H * x.W, groups, kGnEps, stats, ex.stream); });
    } else {
      void* scratch = ex.alloc_raw(gn_scratch_bytes(x.C, C1, x.N, x.H * x.W, groups));
      const void* p1 = skip ? skip->p : nullptr;
      ex.run([&] { return gn_stats(x.p, x.C, p1, C1, x.N, x.H * x.W, groups, kGnEps, scratch, ex.tickets, stats, ex.prec, ex.stream); });
      ex.release_raw(scratch);
    }
    gn.stats = stats;
  }
  ConvOp op = conv_op_nhwc(x, skip, out);
  fill(op);
  op.gn = gn;
  run_conv(ex, op, w, c->ps);
  if (stats) ex.release_raw(stats);
}

// ResnetBlocWithAttn.forward (unet.py:105-111, 155-159) on cat(x, skip), written into `out`.  Inputs are not released.
void res_block(hsidm_ctx* c, const ResW& r, const Act& x, const Act* skip, const NoiseRef& nz, Act& out) {
  Exec& ex = c->ex;
  const int C1 = skip ? skip->C : 0;
  Act h = conv_out(c, r.c1, x.N, x.H, x.W, r.cin, 0);
  {
    ConvW w1 = r.c1;
    w1.pb = -1;   // conv1's bias is folded into the noise-embedding vector (see hsidm_unet_commit)
    gn_conv(c, x, skip, r.gn1_w, r.gn1_b, true, h, w1, [&](ConvOp& op) {
      op.nbias = nz.base + r.noise_off, op.nbias_stride = nz.n_stride;
      op.nbias_t = nz.t_dev, op.nbias_t_stride = nz.t_stride;
    });
  }
  // with attention the block output is an intermediate; otherwise conv2 writes straight into `out`
  Act mid;
  if (r.attn) mid = conv_out(c, r.c2, x.N, x.H, x.W, r.cout, 0);
  Act& y = r.attn ? mid : out;
  // Shortcut: res_conv is folded into conv2 as extra K columns on the halo tensor-core kernel; the identity shortcut is
  // the epilogue's residual add (folding it too - HSIDM_FUSE_ID_MAX=<max Cout> - measured no faster since the epilogue
  // stores through TMA: 19.5 vs 19.1 ms per step); without the halo kernel: a 1x1 conv / epilogue add.
  bool fused = false;
  ConvW wf;
  static const int fuse_id_max = std::getenv("HSIDM_FUSE_ID_MAX") ? atoi(std::getenv("HSIDM_FUSE_ID_MAX")) : 0;
  // A/B switch: res_conv of the blocks with at most this many output channels as its own 1x1 launch instead of extra K columns
  static const int no_rsfuse_max = std::getenv("HSIDM_NO_RSFUSE_MAX") ? atoi(std::getenv("HSIDM_NO_RSFUSE_MAX")) : 0;
  if (ex.prec == HSIDM_BF16 && r.fused.w && (r.has_res || r.cout <= fuse_id_max) && !(r.has_res && r.cout <= no_rsfuse_max)) {
    ConvOp probe = conv_op_nhwc(h, nullptr, y);
    probe.rsrc[0].p = x.p, probe.rsrc[0].C = x.C;
    if (skip) probe.rsrc[1].p = skip->p, probe.rsrc[1].C = skip->C;
    wf.Cin = r.cout, wf.Cout = r.cout, wf.ks = 3, wf.w_bf16 = r.fused.w, wf.bias_override = r.fused.bias;
    probe.w_bf16 = wf.w_bf16, probe.Cout = wf.Cout;
    fused = conv_tc_supported(probe, ex.prec) && conv_halo_ok(probe);
  }
  Act shortcut;
  const void* resid = x.p;
  if (!fused && r.has_res) {
    shortcut = ex.alloc_act(x.N, x.H, x.W, r.cout);
    run_conv(ex, conv_op_nhwc(x, skip, shortcut), r.rc, c->ps);
    resid = shortcut.p;
  }
  gn_conv(c, h, nullptr, r.gn2_w, r.gn2_b, true, y, fused ? wf : r.c2, [&](ConvOp& op) {
    if (fused) {
      op.rsrc[0].p = x.p, op.rsrc[0].C = x.C;
      if (skip) op.rsrc[1].p = skip->p, op.rsrc[1].C = skip->C;
    } else {
      op.resid = resid;
    }
  });
  if (!fused && r.has_res) ex.release(shortcut);
  ex.release(h);
  if (r.attn) {
    attention(c, r, mid, out);
    ex.release(mid);
  }
  (void)C1;
}

// Output tensor of a res layer: written by conv2, or by the attention out-projection.
Act res_out(hsidm_ctx* c, const ResW& r, int N, int H, int W) {
  return r.attn ? conv_out(c, r.aout, N, H, W, r.cout, 0) : conv_out(c, r.c2, N, H, W, r.cout, 0);
}

NoiseRef noise_slice(const NoiseRef& nz, int n0) {
  NoiseRef s = nz;
  if (s.base) s.base += (int64_t)n0 * s.n_stride;
  return s;
}

// Images per sub-batch at a level whose tensors have `per_image_bytes` bytes: consecutive layers of a level run
// sub-batch by sub-batch so that producer -> consumer traffic stays in the 126 MB L2 instead of round-tripping HBM.
int chunk_images(int N, int64_t per_image_bytes, int hw = 1 << 30) {
  static const int min_hw = [] {
    const char* e = std::getenv("HSIDM_CHUNK_MIN_HW");
    return e ? std::atoi(e) : 0;
  }();
  if (hw < min_hw) return N;
  static const int64_t target = [] {
    const char* e = std::getenv("HSIDM_CHUNK_MB");
    return (int64_t)(e ? std::atoi(e) : 0) << 20;   // off by default: measured slower on B200 (see profiles/README.md)
  }();
  if (target <= 0) return N;
  int chunk = (int)std::max<int64_t>(1, target / std::max<int64_t>(1, per_image_bytes));
  if (chunk >= N) return N;
  const int nchunks = (N + chunk - 1) / chunk;
  return (N + nchunks - 1) / nchunks;   // balanced
}

// UNet.forward (unet.py:239-263).  x0/x1: fp32 NCHW halves of the input; eps: fp32 NCHW.
// Layers are grouped into stages of equal resolution; each stage is executed sub-batch by sub-batch (see chunk_images).
void unet_forward_pass(hsidm_ctx* c, const float* x0, int c0, const float* x1, int c1, const NoiseRef& nz, float* eps,
                       int N, int H, int W) {
  Exec& ex = c->ex;
  const size_t es = ex.esize();
  std::vector<Act> feats;
  static const int pdl_max_px = std::getenv("HSIDM_PDL_MAX_PIXELS") ? atoi(std::getenv("HSIDM_PDL_MAX_PIXELS")) : 48 * 128 * 128;
  g_pdl_pass = (long long)N * H * W <= pdl_max_px;   // programmatic dependent launch only where kernels are short
  // ------------------------------------------------ down path ------------------------------------------------
  Act stage_in;   // full-N input of the current stage (empty for the first: raw NCHW halves)
  int Hc = H, Wc = W;
  size_t i = 0;
  while (i < c->downs.size()) {
    size_t j = i;
    while (j < c->downs.size() && c->downs[j].kind != LayerW::DOWN) ++j;
    if (j < c->downs.size()) ++j;   // the Downsample conv closes the stage
    std::vector<Act> outs;
    int cin = i == 0 ? 0 : stage_in.C;
    for (size_t k = i; k < j; ++k) {
      const LayerW& L = c->downs[k];
      if (L.kind == LayerW::CONV) {
        ConvOp proto;
        proto.src[0].C = c0, proto.src[0].layout = L_NCHW_F32, proto.src[1].C = c1, proto.src[1].layout = L_NCHW_F32;
        proto.N = N, proto.Hin = proto.Hout = Hc, proto.Win = proto.Wout = Wc;
        outs.push_back(alloc_conv_out(ex, proto, L.conv, c->ps));
      } else if (L.kind == LayerW::RES) {
        outs.push_back(res_out(c, L.rb, N, Hc, Wc));
      } else {
        outs.push_back(conv_out(c, L.conv, N, Hc, Wc, cin, 0, /*stride=*/2));
      }
      cin = outs.back().C;
    }
    const int chunk = chunk_images(N, (int64_t)Hc * Wc * outs[0].C * (int64_t)es, Hc * Wc);
    for (int n0 = 0; n0 < N; n0 += chunk) {
      const int cnt = std::min(chunk, N - n0);
      const NoiseRef nzs = noise_slice(nz, n0);
      Act xs = i == 0 ? Act() : act_slice(stage_in, n0, cnt, es);
      for (size_t k = i; k < j; ++k) {
        const LayerW& L = c->downs[k];
        Act ys = act_slice(outs[k - i], n0, cnt, es);
        if (L.kind == LayerW::CONV) {
          ConvOp op;
          op.src[0].p = x0 ? x0 + (int64_t)n0 * c0 * H * W : nullptr, op.src[0].C = c0, op.src[0].layout = L_NCHW_F32;
          if (c1) op.src[1].p = x1 ? x1 + (int64_t)n0 * c1 * H * W : nullptr, op.src[1].C = c1, op.src[1].layout = L_NCHW_F32;
          op.N = cnt, op.Hin = H, op.Win = W, op.Hout = H, op.Wout = W, op.out = ys.p;
          op.stats_out = ys.stats, op.stats_slots = ys.slots;
          run_conv(ex, op, L.conv, c->ps);
        } else if (L.kind == LayerW::RES) {
          res_block(c, L.rb, xs, nullptr, nzs, ys);
        } else {
          ConvOp op = conv_op_nhwc(xs, nullptr, ys);
          op.stride = 2;
          run_conv(ex, op, L.conv, c->ps);
        }
        xs = ys;
      }
    }
    for (auto& o : outs) feats.push_back(o);   // every downs output is a skip tensor (unet.py:243-249)
    stage_in = outs.back();
    Hc = stage_in.H, Wc = stage_in.W;
    i = j;
  }
  // ------------------------------------------------ middle ------------------------------------------------
  Act x = stage_in;       // aliases feats.back()
  bool x_owned = false;
  for (const LayerW& L : c->mid) {
    Act y = res_out(c, L.rb, N, x.H, x.W);
    res_block(c, L.rb, x, nullptr, nz, y);
    if (x_owned) ex.release(x);
    x = y, x_owned = true;
  }
  // ------------------------------------------------ up path ------------------------------------------------
  const int out_ch = c->cfg.out_channel;
  i = 0;
  while (i < c->ups.size()) {
    size_t j = i + 1;   // a stage = [optional leading Upsample conv] + res blocks up to the next Upsample
    while (j < c->ups.size() && c->ups[j].kind != LayerW::UP) ++j;
    const bool last_stage = j == c->ups.size();
    const bool lead_up = c->ups[i].kind == LayerW::UP;
    const int Hs = lead_up ? 2 * x.H : x.H, Ws = lead_up ? 2 * x.W : x.W;
    // skips consumed by this stage, in pop order
    std::vector<Act> skips;
    for (size_t k = i; k < j; ++k)
      if (c->ups[k].kind == LayerW::RES) skips.push_back(feats.back()), feats.pop_back();
    const ResW& tail = c->ups[j - 1].rb;   // a stage always ends with a res block
    Act stage_out;
    if (!last_stage) stage_out = res_out(c, tail, N, Hs, Ws);
    const int chunk = chunk_images(N, (int64_t)Hs * Ws * tail.cout * (int64_t)es, Hs * Ws);
    for (int n0 = 0; n0 < N; n0 += chunk) {
      const int cnt = std::min(chunk, N - n0);
      const NoiseRef nzs = noise_slice(nz, n0);
      Act xs = act_slice(x, n0, cnt, es);
      bool xs_owned = false;
      size_t sk = 0;
      for (size_t k = i; k < j; ++k) {
        const LayerW& L = c->ups[k];
        const bool is_tail = k + 1 == j;
        Act ys;
        if (L.kind == LayerW::UP) {
          ys = conv_out(c, L.conv, cnt, xs.H, xs.W, xs.C, 0, 1, /*up=*/1);
          ConvOp op = conv_op_nhwc(xs, nullptr, ys);
          op.up = 1;
          run_conv(ex, op, L.conv, c->ps);
        } else {
          const Act skip = act_slice(skips[sk++], n0, cnt, es);
          const bool into_stage_out = is_tail && !last_stage;
          ys = into_stage_out ? act_slice(stage_out, n0, cnt, es) : res_out(c, L.rb, cnt, xs.H, xs.W);
          res_block(c, L.rb, xs, &skip, nzs, ys);
          if (into_stage_out) {
            if (xs_owned) ex.release(xs);
            xs = ys, xs_owned = false;
            continue;
          }
        }
        if (xs_owned) ex.release(xs);
        xs = ys, xs_owned = true;
      }
      if (last_stage) {
        // final_conv = GroupNorm -> Swish -> conv to out_channel, fp32 NCHW (unet.py:236, 263)
        Act dst;   // the network output is not an arena tensor: only the shape fields matter here
        dst.N = cnt, dst.H = xs.H, dst.W = xs.W, dst.C = out_ch;
        float* eps_out = eps ? eps + (int64_t)n0 * out_ch * xs.H * xs.W : nullptr;
        gn_conv(c, xs, nullptr, c->fin_gn_w, c->fin_gn_b, true, dst, c->fin_conv, [&](ConvOp& op) {
          op.N = cnt;
          op.out = eps_out, op.out_layout = L_NCHW_F32;
          op.stats_out = nullptr, op.stats_slots = 0;
        });
        if (xs_owned) ex.release(xs);
      }
    }
    for (auto& s : skips) ex.release(s);
    if (x_owned) ex.release(x);
    x = stage_out, x_owned = true;
    i = j;
  }
}

int check_shape(hsidm_ctx* c, int c0, int c1, int N, int H, int W) {
  if (!c->committed) HSIDM_FAIL(HSIDM_BAD_STATE, "hsidm_unet_commit has not been called");
  if (N <= 0 || H <= 0 || W <= 0) HSIDM_FAIL(HSIDM_BAD_SHAPE, "non-positive shape N=%d H=%d W=%d", N, H, W);
  if (c0 + c1 != c->cfg.in_channel)
    HSIDM_FAIL(HSIDM_BAD_SHAPE, "input channels %d+%d != in_channel %d", c0, c1, c->cfg.in_channel);
  const int div = 1 << (c->cfg.n_mults - 1);
  if (H % div || W % div)
    HSIDM_FAIL(HSIDM_BAD_SHAPE, "H=%d, W=%d must be multiples of %d (one per down/up level, unet.py:68-74, 58-65)", H, W, div);
  return HSIDM_OK;
}

void drop_graph(hsidm_ctx* c) {
  if (c->graph_exec) {
    cudaDeviceSynchronize();
    cudaGraphExecDestroy(c->graph_exec);
  }
  if (c->graph) cudaGraphDestroy(c->graph);
  c->graph_exec = nullptr, c->graph = nullptr;
}

// Measure the arena peak for this shape with a dry pass and make sure the backing store is large enough.
int ensure_workspace(hsidm_ctx* c, int c0, int c1, int N, int H, int W) {
  if (c->ws_N == N && c->ws_H == H && c->ws_W == W) return HSIDM_OK;
  drop_graph(c);  // the arena may move
  Exec& ex = c->ex;
  HSIDM_TRY(ex.ensure_tickets(N));
  ex.dry = true, ex.status = HSIDM_OK;
  ex.arena.begin(true);
  NoiseRef nz{nullptr, 0, nullptr, 0};
  unet_forward_pass(c, nullptr, c0, nullptr, c1, nz, nullptr, N, H, W);
  ex.dry = false;
  if (ex.status != HSIDM_OK) return ex.status;
  HSIDM_TRY(ex.arena.reserve(ex.arena.peak()));
  c->ws_N = N, c->ws_H = H, c->ws_W = W;
  return HSIDM_OK;
}

int run_forward(hsidm_ctx* c, const float* x0, int c0, const float* x1, int c1, const NoiseRef& nz, float* eps, int N,
                int H, int W, cudaStream_t stream) {
  Exec& ex = c->ex;
  ex.stream = stream, ex.dry = false, ex.status = HSIDM_OK;
  ex.arena.begin(false);
  unet_forward_pass(c, x0, c0, x1, c1, nz, eps, N, H, W);
  return ex.status;
}

int ensure_table(hsidm_ctx* c, cudaStream_t stream) {
  if (!c->table_dirty) return HSIDM_OK;
  if (c->T <= 0) HSIDM_FAIL(HSIDM_BAD_STATE, "hsidm_set_schedule has not been called");
  drop_graph(c);   // a cached graph holds the old table pointer
  ++c->weight_gen;
  if (c->nbias_table) cudaFree(c->nbias_table);
  c->nbias_table = nullptr;
  HSIDM_CUDA(cudaMalloc(&c->nbias_table, sizeof(float) * (int64_t)c->T * c->noise_total));
  HSIDM_TRY(noise_embed(c->levels_dev, 1, c->T, c->cfg.inner_channel, c->ps.dev(c->mlp1_w), c->ps.dev(c->mlp1_b),
                        c->ps.dev(c->mlp3_w), c->ps.dev(c->mlp3_b), c->noise_layers_dev,
                        (int)c->noise_layers_host.size(), c->noise_total, c->nbias_table, stream));
  c->table_dirty = false;
  return HSIDM_OK;
}

}  // namespace

// =================================================================================================================
extern "C" {

int hsidm_ctx_create(const hsidm_unet_cfg* cfg, int device, hsidm_ctx** out) {
  if (!cfg || !out) HSIDM_FAIL(HSIDM_BAD_ARG, "hsidm_ctx_create: null argument");
  *out = nullptr;
  if (cfg->n_mults < 1 || cfg->n_mults > HSIDM_MAX_LEVELS || cfg->n_attn_res < 0 || cfg->n_attn_res > HSIDM_MAX_LEVELS)
    HSIDM_FAIL(HSIDM_UNSUPPORTED_CFG, "channel_multiplier / attn_res length out of range");
  if (cfg->precision != HSIDM_F32 && cfg->precision != HSIDM_BF16) HSIDM_FAIL(HSIDM_BAD_DTYPE, "unknown precision %d", cfg->precision);
  if (cfg->inner_channel <= 0 || cfg->inner_channel % 8 || cfg->norm_groups <= 0 || cfg->inner_channel % cfg->norm_groups)
    HSIDM_FAIL(HSIDM_UNSUPPORTED_CFG, "inner_channel=%d must be a positive multiple of 8 and of norm_groups=%d",
               cfg->inner_channel, cfg->norm_groups);
  if (cfg->in_channel <= 0 || cfg->out_channel <= 0 || cfg->res_blocks < 1)
    HSIDM_FAIL(HSIDM_UNSUPPORTED_CFG, "in_channel/out_channel/res_blocks must be positive");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    HSIDM_FAIL(HSIDM_CUDA_ERROR, "no CUDA device available (this library has no CPU path)");
  }
  HSIDM_DEVICE(device);
  hsidm_ctx* c = new hsidm_ctx();
  c->cfg = *cfg;
  c->device = device;
  c->ex.prec = cfg->precision;
  build_layers(c);
  int s = c->ps.alloc_all();
  if (s == HSIDM_OK && cudaMalloc(&c->t_dev, 8 * sizeof(int)) != cudaSuccess) s = HSIDM_CUDA_ERROR;
  if (s == HSIDM_OK && cfg->precision == HSIDM_BF16) s = conv_tc_init();
  if (s == HSIDM_OK && (cudaStreamCreateWithFlags(&c->side, cudaStreamNonBlocking) != cudaSuccess ||
                        cudaEventCreateWithFlags(&c->ev_in, cudaEventDisableTiming) != cudaSuccess ||
                        cudaEventCreateWithFlags(&c->ev_out, cudaEventDisableTiming) != cudaSuccess ||
                        cudaEventCreateWithFlags(&c->ev_arena, cudaEventDisableTiming) != cudaSuccess)) {
    set_last_error("could not create the sampler stream/events");
    s = HSIDM_CUDA_ERROR;
  }
  if (s != HSIDM_OK) {
    delete c;
    return s;
  }
  *out = c;
  return HSIDM_OK;
}

int hsidm_ctx_destroy(hsidm_ctx* c) {
  if (!c) return HSIDM_OK;
  DeviceGuard guard(c->device);
  cudaDeviceSynchronize();
  drop_graph(c);
  for_each_conv(c, [](ConvW& w) { free_conv(w); });
  for (auto* v : {&c->downs, &c->mid, &c->ups})
    for (auto& L : *v)
      if (L.kind == LayerW::RES) free_fused(L.rb.fused), free_attn_fold(L.rb.afold);
  if (c->noise_layers_dev) cudaFree(c->noise_layers_dev);
  if (c->coef_dev) cudaFree(c->coef_dev);
  if (c->levels_dev) cudaFree(c->levels_dev);
  if (c->nbias_table) cudaFree(c->nbias_table);
  if (c->t_dev) cudaFree(c->t_dev);
  if (c->nbias_buf) cudaFree(c->nbias_buf);
  if (c->samp_buf) cudaFree(c->samp_buf);
  if (c->side) cudaStreamDestroy(c->side);
  if (c->ev_in) cudaEventDestroy(c->ev_in);
  if (c->ev_out) cudaEventDestroy(c->ev_out);
  if (c->ev_arena) cudaEventDestroy(c->ev_arena);
  if (c->train && c->train_free) c->train_free(c->train);
  delete c;
  return HSIDM_OK;
}

int hsidm_unet_param_count(const hsidm_ctx* c) { return c ? c->ps.size() : 0; }
const char* hsidm_unet_param_name(const hsidm_ctx* c, int i) {
  return (c && i >= 0 && i < c->ps.size()) ? c->ps.at(i).key.c_str() : nullptr;
}

int hsidm_unet_set_param(hsidm_ctx* c, const char* key, const float* data, const int64_t* shape, int ndim) {
  if (!c) HSIDM_FAIL(HSIDM_BAD_ARG, "null context");
  HSIDM_DEVICE(c->device);
  c->committed = false;
  drop_graph(c);
  return c->ps.set(key, data, shape, ndim);
}

int hsidm_unet_commit(hsidm_ctx* c) {
  if (!c) HSIDM_FAIL(HSIDM_BAD_ARG, "null context");
  HSIDM_DEVICE(c->device);
  HSIDM_TRY(c->ps.check_all_set());
  drop_graph(c);   // every packed buffer is freed and reallocated below: a cached graph would replay dangling pointers
  ++c->weight_gen;
  int status = HSIDM_OK;
  c->packed_bytes = 0;
  for_each_conv(c, [&](ConvW& w) {
    if (status != HSIDM_OK) return;
    status = pack_conv(c->ps, w, c->cfg.precision == HSIDM_BF16);
    c->packed_bytes += w.packed_bytes;
  });
  HSIDM_TRY(status);
  if (c->cfg.precision == HSIDM_BF16)
    for (auto& L : c->downs)
      if (L.kind == LayerW::DOWN && status == HSIDM_OK) {
        const int64_t before = L.conv.packed_bytes;
        status = pack_conv_s2(c->ps, L.conv);
        c->packed_bytes += L.conv.packed_bytes - before;
      }
  HSIDM_TRY(status);
  if (c->cfg.precision == HSIDM_BF16)
    for (auto& L : c->ups)
      if (L.kind == LayerW::UP && status == HSIDM_OK) {
        const int64_t before = L.conv.packed_bytes;
        status = pack_conv_up(c->ps, L.conv);
        c->packed_bytes += L.conv.packed_bytes - before;
      }
  HSIDM_TRY(status);
  if (c->cfg.precision == HSIDM_BF16) {
    auto fuse = [&](std::vector<LayerW>& v) {
      for (auto& L : v)
        if (L.kind == LayerW::RES && status == HSIDM_OK) {
          status = pack_fused(c->ps, L.rb.c2, L.rb.has_res ? &L.rb.rc : nullptr, L.rb.fused);
          c->packed_bytes += L.rb.fused.bytes;
          if (status == HSIDM_OK && L.rb.attn) {
            status = pack_attn_fold(c->ps, L.rb.qkv, L.rb.aout, L.rb.afold);
            c->packed_bytes += L.rb.afold.bytes;
          }
        }
    };
    fuse(c->downs), fuse(c->mid), fuse(c->ups);
    HSIDM_TRY(status);
  }
  c->noise_layers_host.clear();
  auto visit = [&](std::vector<LayerW>& v) {
    for (auto& L : v)
      if (L.kind == LayerW::RES)
        c->noise_layers_host.push_back(
            NoiseLayer{c->ps.dev(L.rb.nf_w), c->ps.dev(L.rb.nf_b), c->ps.dev(L.rb.c1.pb), L.rb.cout, L.rb.noise_off});
  };
  visit(c->downs), visit(c->mid), visit(c->ups);
  if (c->noise_layers_dev) cudaFree(c->noise_layers_dev);
  HSIDM_CUDA(cudaMalloc(&c->noise_layers_dev, sizeof(NoiseLayer) * c->noise_layers_host.size()));
  HSIDM_CUDA(cudaMemcpy(c->noise_layers_dev, c->noise_layers_host.data(), sizeof(NoiseLayer) * c->noise_layers_host.size(),
                        cudaMemcpyHostToDevice));
  HSIDM_CUDA(cudaDeviceSynchronize());
  c->committed = true;
  c->table_dirty = true;
  return HSIDM_OK;
}

int hsidm_set_schedule(hsidm_ctx* c, const double* betas, int T) {
  if (!c || !betas || T <= 0) HSIDM_FAIL(HSIDM_BAD_ARG, "hsidm_set_schedule: bad argument");
  HSIDM_DEVICE(c->device);
  drop_graph(c);
  // float64 tables then fp32 casts, as set_new_noise_schedule does (diffusion.py:103-140)
  std::vector<float> coef((size_t)T * 5), levels(T);
  double ac = 1.0;
  for (int t = 0; t < T; ++t) {
    const double b = betas[t], alpha = 1.0 - b, ac_prev = ac;
    ac *= alpha;
    const double post_var = b * (1.0 - ac_prev) / (1.0 - ac);
    const float logvar = (float)std::log(std::max(post_var, 1e-20));
    coef[t * 5 + 0] = (float)std::sqrt(1.0 / ac);
    coef[t * 5 + 1] = (float)std::sqrt(1.0 / ac - 1.0);
    coef[t * 5 + 2] = (float)(b * std::sqrt(ac_prev) / (1.0 - ac));
    coef[t * 5 + 3] = (float)((1.0 - ac_prev) * std::sqrt(alpha) / (1.0 - ac));
    coef[t * 5 + 4] = std::exp(0.5f * logvar);  // (0.5 * log_variance).exp() in fp32, diffusion.py:175
    levels[t] = (float)std::sqrt(ac);           // sqrt_alphas_cumprod_prev[t+1], diffusion.py:154
  }
  if (c->coef_dev) cudaFree(c->coef_dev);
  if (c->levels_dev) cudaFree(c->levels_dev);
  c->coef_dev = nullptr, c->levels_dev = nullptr;
  HSIDM_CUDA(cudaMalloc(&c->coef_dev, sizeof(float) * coef.size()));
  HSIDM_CUDA(cudaMalloc(&c->levels_dev, sizeof(float) * T));
  HSIDM_CUDA(cudaMemcpy(c->coef_dev, coef.data(), sizeof(float) * coef.size(), cudaMemcpyHostToDevice));
  HSIDM_CUDA(cudaMemcpy(c->levels_dev, levels.data(), sizeof(float) * T, cudaMemcpyHostToDevice));
  c->coef_host = coef;
  c->T = T;
  c->table_dirty = true;
  return HSIDM_OK;
}

int hsidm_num_timesteps(const hsidm_ctx* c) { return c ? c->T : 0; }
int hsidm_snapshot_count(const hsidm_ctx* c) {
  if (!c || c->T <= 0) return 0;
  const int inter = 1 | (c->T / 10);
  return (c->T - 1) / inter + 1;
}
int64_t hsidm_ctx_bytes(const hsidm_ctx* c) {
  if (!c) return 0;
  return c->ex.arena.capacity() + c->packed_bytes + c->ps.bytes() + (int64_t)c->T * c->noise_total * 4 + c->samp_cap * 4;
}

int hsidm_unet_forward(hsidm_ctx* c, const float* x0, int c0, const float* x1, int c1, const float* noise_level,
                       int level_stride, float* eps_out, int N, int H, int W, hsidm_stream stream_) {
  if (!c || !x0 || !noise_level || !eps_out) HSIDM_FAIL(HSIDM_BAD_ARG, "hsidm_unet_forward: null argument");
  if (c1 > 0 && !x1) HSIDM_FAIL(HSIDM_BAD_ARG, "hsidm_unet_forward: c1 > 0 but x1 is null");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  HSIDM_DEVICE(c->device);
  HSIDM_TRY(check_shape(c, c0, c1, N, H, W));
  HSIDM_TRY(ensure_workspace(c, c0, c1, N, H, W));
  if (c->nbias_cap < N) {
    if (c->nbias_buf) {
      HSIDM_CUDA(cudaDeviceSynchronize());
      cudaFree(c->nbias_buf);
      c->nbias_buf = nullptr;
    }
    HSIDM_CUDA(cudaMalloc(&c->nbias_buf, sizeof(float) * (int64_t)N * c->noise_total));
    c->nbias_cap = N;
  }
  const int n_emb = level_stride == 0 ? 1 : N;
  HSIDM_CUDA(cudaStreamWaitEvent(stream, c->ev_arena, 0));
  HSIDM_TRY(noise_embed(noise_level, level_stride, n_emb, c->cfg.inner_channel, c->ps.dev(c->mlp1_w), c->ps.dev(c->mlp1_b),
                        c->ps.dev(c->mlp3_w), c->ps.dev(c->mlp3_b), c->noise_layers_dev, (int)c->noise_layers_host.size(),
                        c->noise_total, c->nbias_buf, stream));
  NoiseRef nz{c->nbias_buf, level_stride == 0 ? 0 : (int64_t)c->noise_total, nullptr, 0};
  HSIDM_CUDA(cudaStreamWaitEvent(stream, c->ev_arena, 0));   // the previous pass over the arena, whatever stream it ran on
  const int s = run_forward(c, x0, c0, x1, c1, nz, eps_out, N, H, W, stream);
  HSIDM_CUDA(cudaEventRecord(c->ev_arena, stream));
  return s;
}

int hsidm_posterior_step(hsidm_ctx* c, int t, const float* x_t, const float* eps, const float* noise, float* x_prev,
                         int64_t n, hsidm_stream stream_) {
  if (!c || !x_t || !eps || !x_prev) HSIDM_FAIL(HSIDM_BAD_ARG, "hsidm_posterior_step: null argument");
  if (c->T <= 0) HSIDM_FAIL(HSIDM_BAD_STATE, "hsidm_set_schedule has not been called");
  if (t < 0 || t >= c->T) HSIDM_FAIL(HSIDM_BAD_ARG, "timestep %d outside [0,%d)", t, c->T);
  HSIDM_DEVICE(c->device);
  PosteriorArgs a{};
  a.x_t = x_t, a.eps = eps, a.noise = noise, a.x_prev = x_prev, a.n = n;
  a.per_image = n, a.tape_image_stride = 0, a.tape_step_stride = 0;  // noise is given for this very step
  a.coef = c->coef_dev, a.T = c->T, a.t = t, a.t_dev = nullptr, a.seed = 0, a.use_philox = 0;
  a.snapshot_base = nullptr, a.inter = 1;
  // a.noise indexes with j = T-1-t: fold that into the pointer by using zero strides
  return posterior_step(a, static_cast<cudaStream_t>(stream_));
}

int hsidm_sample(hsidm_ctx* c, const float* cond, const float* x_T, const float* noise_tape, int64_t tape_image_stride,
                 int64_t tape_step_stride, uint64_t seed, float* out, float* snapshots, int N, int H, int W,
                 hsidm_stream stream_) {
  return hsidm_sample_at(c, cond, x_T, noise_tape, tape_image_stride, tape_step_stride, seed, 0, out, snapshots, N, H, W, stream_);
}

int hsidm_sample_at(hsidm_ctx* c, const float* cond, const float* x_T, const float* noise_tape, int64_t tape_image_stride,
                    int64_t tape_step_stride, uint64_t seed, int64_t first_image, float* out, float* snapshots, int N, int H,
                    int W, hsidm_stream stream_) {
  if (!c || !cond || !x_T || !out) HSIDM_FAIL(HSIDM_BAD_ARG, "hsidm_sample: null argument");
  if (first_image < 0) HSIDM_FAIL(HSIDM_BAD_ARG, "hsidm_sample_at: negative first_image");
  cudaStream_t caller = static_cast<cudaStream_t>(stream_);
  cudaStream_t stream = c->side;
  HSIDM_DEVICE(c->device);
  if (c->cfg.in_channel % 2 || c->cfg.out_channel * 2 != c->cfg.in_channel)
    HSIDM_FAIL(HSIDM_UNSUPPORTED_CFG, "conditional sampling needs in_channel == 2*out_channel (diffusion.py:158)");
  const int ch = c->cfg.out_channel;
  HSIDM_TRY(check_shape(c, ch, ch, N, H, W));
  HSIDM_TRY(ensure_workspace(c, ch, ch, N, H, W));
  HSIDM_TRY(ensure_table(c, stream));
  const int64_t per_image = (int64_t)ch * H * W, n = per_image * N;
  if (c->samp_cap < 3 * n) {
    drop_graph(c);
    if (c->samp_buf) {
      HSIDM_CUDA(cudaDeviceSynchronize());
      cudaFree(c->samp_buf);
      c->samp_buf = nullptr;
    }
    HSIDM_CUDA(cudaMalloc(&c->samp_buf, sizeof(float) * 3 * n));
    c->samp_cap = 3 * n;
  }
  float* cond_b = c->samp_buf;
  float* x_b = c->samp_buf + n;
  float* eps_b = c->samp_buf + 2 * n;
  // everything below is ordered after the caller's prior work ...
  HSIDM_CUDA(cudaEventRecord(c->ev_in, caller));
  HSIDM_CUDA(cudaStreamWaitEvent(stream, c->ev_in, 0));
  HSIDM_CUDA(cudaStreamWaitEvent(stream, c->ev_arena, 0));
  hsidm_ctx::GraphKey key;
  key.weight_gen = c->weight_gen, key.route_gen = conv_tc_route_gen();
  key.N = N, key.H = H, key.W = W, key.tape = noise_tape, key.s_img = tape_image_stride, key.s_step = tape_step_stride;
  key.snaps = snapshots;
  if (c->graph_exec && !(c->graph_key == key)) drop_graph(c);
  if (!c->graph_exec) {
    // ---- capture one denoising step: UNet forward + fused posterior + timestep decrement ----
    const int64_t launches_before = g_launches;
    HSIDM_CUDA(cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal));
    NoiseRef nz{c->nbias_table, 0, c->t_dev, (int64_t)c->noise_total};
    int s = run_forward(c, cond_b, ch, x_b, ch, nz, eps_b, N, H, W, stream);
    if (s == HSIDM_OK) {
      PosteriorArgs a{};
      a.x_t = x_b, a.eps = eps_b, a.noise = noise_tape, a.x_prev = x_b, a.n = n, a.per_image = per_image;
      a.tape_image_stride = tape_image_stride, a.tape_step_stride = tape_step_stride;
      a.coef = c->coef_dev, a.T = c->T, a.t = -1, a.t_dev = c->t_dev, a.seed = 0;
      a.seed_dev = reinterpret_cast<const unsigned long long*>(c->t_dev + 2), a.use_philox = noise_tape ? 0 : 1;
      a.snapshot_base = snapshots, a.inter = 1 | (c->T / 10);
      s = posterior_step(a, stream);
    }
    if (s == HSIDM_OK) s = step_counter_dec(c->t_dev, stream);
    cudaError_t ce = cudaStreamEndCapture(stream, &c->graph);
    if (s != HSIDM_OK) {
      drop_graph(c);
      return s;
    }
    if (ce != cudaSuccess) {
      drop_graph(c);
      HSIDM_FAIL(HSIDM_CUDA_ERROR, "cudaStreamEndCapture failed: %s", cudaGetErrorString(ce));
    }
    ce = cudaGraphInstantiate(&c->graph_exec, c->graph, 0);
    if (ce != cudaSuccess) {
      drop_graph(c);
      HSIDM_FAIL(HSIDM_CUDA_ERROR, "cudaGraphInstantiate failed: %s", cudaGetErrorString(ce));
    }
    c->graph_key = key;
    c->graph_nodes = g_launches - launches_before;  // captured, not executed: count them per replay instead
    g_launches = launches_before;
  }
  HSIDM_CUDA(cudaMemcpyAsync(cond_b, cond, sizeof(float) * n, cudaMemcpyDeviceToDevice, stream));
  HSIDM_CUDA(cudaMemcpyAsync(x_b, x_T, sizeof(float) * n, cudaMemcpyDeviceToDevice, stream));
  HSIDM_TRY(sampler_state_set(c->t_dev, c->T - 1, seed, (uint64_t)first_image * (uint64_t)(per_image / 4), stream));
  for (int i = 0; i < c->T; ++i) {
    HSIDM_CUDA(cudaGraphLaunch(c->graph_exec, stream));
    g_launches += c->graph_nodes;
  }
  HSIDM_CUDA(cudaMemcpyAsync(out, x_b, sizeof(float) * n, cudaMemcpyDeviceToDevice, stream));
  // ... and the caller's later work is ordered after the sampler
  HSIDM_CUDA(cudaEventRecord(c->ev_out, stream));
  HSIDM_CUDA(cudaEventRecord(c->ev_arena, stream));
  HSIDM_CUDA(cudaStreamWaitEvent(caller, c->ev_out, 0));
  return HSIDM_OK;
}

int hsidm_unet_params_changed(hsidm_ctx* c, const void* const* table_dev, int n, int* changed, hsidm_stream stream) {
  if (!c) HSIDM_FAIL(HSIDM_BAD_ARG, "null context");
  HSIDM_DEVICE(c->device);
  return c->ps.differs(table_dev, n, changed, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
