// Group autoencoder (GAE) executor behind the hsidm_gae C ABI.
//
// The reference encodes/decodes the G overlapping band groups one after another at batch 1 (AE.py:283-324).
// Here all B*G groups of a batch of cubes go through the shared Encoder / Decoder as ONE batch of images: the
// head conv gathers its n_subs bands straight out of the NCHW cube through a per-image offset table, and the
// decoder outputs are overlap-averaged with static 1/count weights before the residual trunk.
#include <cmath>

#include "net.cuh"

namespace hsidm {

struct SsbW {  // one SSB = ResBlock(k=3) -> ResAttentionBlock(k=1) + CALayer (AE.py:102-109)
  ConvW spa0, spa2, spc0, spc2;
  int du0_w = -1, du0_b = -1, du2_w = -1, du2_b = -1;
};

struct BranchW {  // BranchUnit(use_tail=False, up_scale=1) (AE.py:145-165)
  ConvW head;
  std::vector<SsbW> blocks;
  int feats = 0;
};

}  // namespace hsidm

using namespace hsidm;

struct hsidm_gae {
  hsidm_gae_cfg cfg;
  int device = 0;
  int G = 0;
  std::vector<int> start, end;
  ParamStore ps;
  BranchW enc, dec, trunk;
  ConvW enc_final, dec_final, fin;
  bool committed = false;
  Exec ex;
  int* start_dev = nullptr;
  float* inv_count_dev = nullptr;
  int64_t* img_off_dev = nullptr;  // [cap_B * G]
  int off_B = 0, off_H = 0, off_W = 0;
  int ws_kind = 0, ws_B = 0, ws_H = 0, ws_W = 0;
  cudaEvent_t ev_arena = nullptr;  // completion of the previous pass over the arena (calls may come from different streams)
};

namespace {

constexpr float kResScale = 0.1f;  // res_scale of every ResBlock / ResAttentionBlock (AE.py:192,225,268)

BranchW make_branch(ParamStore& ps, const std::string& p, int cin, int feats, int blocks) {
  BranchW b;
  b.feats = feats;
  b.head = make_conv(ps, p + ".head", cin, feats, 3);
  const int red = feats / 3;  // CALayer(n_feats, reduction=3), common.py:267
  for (int i = 0; i < blocks; ++i) {
    const std::string q = p + ".body.net." + std::to_string(i);
    SsbW s;
    s.spa0 = make_conv(ps, q + ".spa.body.0", feats, feats, 3);
    s.spa2 = make_conv(ps, q + ".spa.body.2", feats, feats, 3);
    s.spc0 = make_conv(ps, q + ".spc.body.0", feats, feats, 1);
    s.spc2 = make_conv(ps, q + ".spc.body.2", feats, feats, 1);
    s.du0_w = ps.add(q + ".spc.body.3.conv_du.0.weight", {red, feats, 1, 1});
    s.du0_b = ps.add(q + ".spc.body.3.conv_du.0.bias", {red});
    s.du2_w = ps.add(q + ".spc.body.3.conv_du.2.weight", {feats, red, 1, 1});
    s.du2_b = ps.add(q + ".spc.body.3.conv_du.2.bias", {feats});
    b.blocks.push_back(s);
  }
  return b;
}

template <typename F>
void for_each_conv(hsidm_gae* g, F&& f) {
  for (BranchW* b : {&g->enc, &g->dec, &g->trunk}) {
    f(b->head);
    for (auto& s : b->blocks) f(s.spa0), f(s.spa2), f(s.spc0), f(s.spc2);
  }
  f(g->enc_final), f(g->dec_final), f(g->fin);
}

ConvOp op_nhwc(const Act& in, const Act& out) {
  ConvOp op;
  op.src[0].p = in.p, op.src[0].C = in.C;
  op.N = in.N, op.Hin = in.H, op.Win = in.W, op.Hout = out.H, op.Wout = out.W;
  op.out = out.p;
  return op;
}

// body of BranchUnit after the head conv: SSPN = n x SSB + skip (AE.py:120-141). Consumes `y`.
Act branch_body(hsidm_gae* g, const BranchW& b, Act y) {
  Exec& ex = g->ex;
  const int N = y.N, H = y.H, W = y.W, F = b.feats, red = F / 3;
  Act h = y;  // h aliases y for the first block
  for (size_t i = 0; i < b.blocks.size(); ++i) {
    const SsbW& s = b.blocks[i];
    // ResBlock: h1 = h + 0.1 * conv(lrelu(conv(h)))   (common.py:163-182)
    Act t = ex.alloc_act(N, H, W, F);
    {
      ConvOp op = op_nhwc(h, t);
      op.act = ACT_LRELU;
      run_conv(ex, op, s.spa0, g->ps);
    }
    Act h1 = ex.alloc_act(N, H, W, F);
    {
      ConvOp op = op_nhwc(t, h1);
      op.scale = kResScale, op.resid = h.p;
      run_conv(ex, op, s.spa2, g->ps);
    }
    if (i > 0) ex.release(h);  // y itself is still needed for the SSPN skip
    // ResAttentionBlock: h2 = h1 + 0.1 * CA(conv1x1(lrelu(conv1x1(h1))))   (common.py:231-271)
    {
      ConvOp op = op_nhwc(h1, t);
      op.act = ACT_LRELU;
      run_conv(ex, op, s.spc0, g->ps);
    }
    Act r = ex.alloc_act(N, H, W, F);
    run_conv(ex, op_nhwc(t, r), s.spc2, g->ps);
    ex.release(t);
    float* mean = static_cast<float*>(ex.alloc_raw(sizeof(float) * 2 * N * F));
    float* gate = mean + (int64_t)N * F;
    ex.run([&] { return channel_mean(r.p, N, H * W, F, mean, ex.prec, ex.stream); });
    ex.run([&] {
      return ca_gate(mean, N, F, red, g->ps.dev(s.du0_w), g->ps.dev(s.du0_b), g->ps.dev(s.du2_w), g->ps.dev(s.du2_b), gate,
                     ex.stream);
    });
    Act h2 = ex.alloc_act(N, H, W, F);
    ex.run([&] { return scale_residual(r.p, gate, kResScale, h1.p, h2.p, N, H * W, F, ex.prec, ex.stream); });
    ex.release_raw(mean);
    ex.release(r);
    ex.release(h1);
    h = h2;
  }
  // SSPN skip: res + x
  Act out = ex.alloc_act(N, H, W, F);
  ex.run([&] { return scale_residual(h.p, nullptr, 1.0f, y.p, out.p, N, H * W, F, ex.prec, ex.stream); });
  if (!b.blocks.empty()) ex.release(h);
  ex.release(y);
  return out;
}

void encode_pass(hsidm_gae* g, const float* x, float* z, int B, int H, int W) {
  Exec& ex = g->ex;
  const int N = B * g->G;
  Act y = ex.alloc_act(N, H, W, g->cfg.n_feats);
  {
    ConvOp op;
    op.src[0].p = x, op.src[0].C = g->cfg.n_subs, op.src[0].layout = L_NCHW_F32, op.src[0].img_off = g->img_off_dev;
    op.N = N, op.Hin = H, op.Win = W, op.Hout = H, op.Wout = W, op.out = y.p;
    run_conv(ex, op, g->enc.head, g->ps);
  }
  Act f = branch_body(g, g->enc, y);
  {
    ConvOp op;
    op.src[0].p = f.p, op.src[0].C = f.C;
    op.N = N, op.Hin = H, op.Win = W, op.Hout = H, op.Wout = W, op.out = z, op.out_layout = L_NCHW_F32;
    run_conv(ex, op, g->enc_final, g->ps);
  }
  ex.release(f);
}

void decode_pass(hsidm_gae* g, const float* z, float* out, int B, int H, int W, int clamp01) {
  Exec& ex = g->ex;
  const int N = B * g->G;
  Act y = ex.alloc_act(N, H, W, g->cfg.n_feats);
  {
    ConvOp op;
    op.src[0].p = z, op.src[0].C = g->cfg.latent, op.src[0].layout = L_NCHW_F32;
    op.N = N, op.Hin = H, op.Win = W, op.Hout = H, op.Wout = W, op.out = y.p;
    run_conv(ex, op, g->dec.head, g->ps);
  }
  Act f = branch_body(g, g->dec, y);
  Act d = ex.alloc_act(N, H, W, g->cfg.n_subs);
  run_conv(ex, op_nhwc(f, d), g->dec_final, g->ps);
  ex.release(f);
  // y[:, s:e] += dec_g ; y /= count  (AE.py:288-297)
  Act avg = ex.alloc_act(B, H, W, g->cfg.n_colors);
  ex.run([&] {
    return overlap_average(d.p, B, g->G, H * W, g->cfg.n_subs, g->cfg.n_colors, g->start_dev, g->inv_count_dev, avg.p,
                           ex.prec, ex.stream);
  });
  ex.release(d);
  // y + final(trunk(y))  (AE.py:302-307)
  Act t = ex.alloc_act(B, H, W, g->cfg.trunk_feats);
  run_conv(ex, op_nhwc(avg, t), g->trunk.head, g->ps);
  Act tf = branch_body(g, g->trunk, t);
  {
    ConvOp op;
    op.src[0].p = tf.p, op.src[0].C = tf.C;
    op.N = B, op.Hin = H, op.Win = W, op.Hout = H, op.Wout = W, op.out = out, op.out_layout = L_NCHW_F32;
    op.resid = avg.p, op.clamp01 = clamp01;
    run_conv(ex, op, g->fin, g->ps);
  }
  ex.release(tf);
  ex.release(avg);
}

int prepare(hsidm_gae* g, int kind, int B, int H, int W) {
  if (!g->committed) HSIDM_FAIL(HSIDM_BAD_STATE, "hsidm_gae_commit has not been called");
  if (B <= 0 || H <= 0 || W <= 0) HSIDM_FAIL(HSIDM_BAD_SHAPE, "non-positive shape B=%d H=%d W=%d", B, H, W);
  if (g->off_B != B || g->off_H != H || g->off_W != W) {
    std::vector<int64_t> off((size_t)B * g->G);
    const int64_t hw = (int64_t)H * W;
    for (int b = 0; b < B; ++b)
      for (int k = 0; k < g->G; ++k) off[(size_t)b * g->G + k] = ((int64_t)b * g->cfg.n_colors + g->start[k]) * hw;
    HSIDM_CUDA(cudaDeviceSynchronize());
    if (g->img_off_dev) cudaFree(g->img_off_dev);
    HSIDM_CUDA(cudaMalloc(&g->img_off_dev, sizeof(int64_t) * off.size()));
    HSIDM_CUDA(cudaMemcpy(g->img_off_dev, off.data(), sizeof(int64_t) * off.size(), cudaMemcpyHostToDevice));
    g->off_B = B, g->off_H = H, g->off_W = W;
  }
  if (g->ws_kind != kind || g->ws_B != B || g->ws_H != H || g->ws_W != W) {
    Exec& ex = g->ex;
    ex.dry = true, ex.status = HSIDM_OK;
    ex.arena.begin(true);
    if (kind == 1)
      encode_pass(g, nullptr, nullptr, B, H, W);
    else
      decode_pass(g, nullptr, nullptr, B, H, W, 0);
    ex.dry = false;
    if (ex.status != HSIDM_OK) return ex.status;
    HSIDM_TRY(ex.arena.reserve(ex.arena.peak()));
    g->ws_kind = kind, g->ws_B = B, g->ws_H = H, g->ws_W = W;
  }
  return HSIDM_OK;
}

}  // namespace

extern "C" {

int hsidm_gae_create(const hsidm_gae_cfg* cfg, int device, hsidm_gae** out) {
  if (!cfg || !out) HSIDM_FAIL(HSIDM_BAD_ARG, "hsidm_gae_create: null argument");
  *out = nullptr;
  if (cfg->n_subs <= cfg->n_ovls || cfg->n_ovls < 0 || cfg->n_colors < cfg->n_subs)
    HSIDM_FAIL(HSIDM_UNSUPPORTED_CFG, "need n_colors >= n_subs > n_ovls >= 0 (got %d, %d, %d)", cfg->n_colors, cfg->n_subs, cfg->n_ovls);
  if (cfg->n_feats <= 0 || cfg->n_feats > 256 || cfg->trunk_feats <= 0 || cfg->trunk_feats > 256 || cfg->n_feats < 3 ||
      cfg->trunk_feats < 3 || cfg->latent <= 0 || cfg->n_blocks < 0 || cfg->trunk_blocks < 0)
    HSIDM_FAIL(HSIDM_UNSUPPORTED_CFG, "feature widths must be in [3,256]");
  if (cfg->precision != HSIDM_F32 && cfg->precision != HSIDM_BF16) HSIDM_FAIL(HSIDM_BAD_DTYPE, "unknown precision %d", cfg->precision);
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    HSIDM_FAIL(HSIDM_CUDA_ERROR, "no CUDA device available (this library has no CPU path)");
  }
  HSIDM_DEVICE(device);
  hsidm_gae* g = new hsidm_gae();
  g->cfg = *cfg;
  g->device = device;
  g->ex.prec = cfg->precision;
  // group layout, AE.py:264-280
  g->G = (int)std::ceil((double)(cfg->n_colors - cfg->n_ovls) / (double)(cfg->n_subs - cfg->n_ovls));
  std::vector<float> inv(cfg->n_colors, 0.f);
  for (int k = 0; k < g->G; ++k) {
    int s = (cfg->n_subs - cfg->n_ovls) * k, e = s + cfg->n_subs;
    if (e > cfg->n_colors) e = cfg->n_colors, s = cfg->n_colors - cfg->n_subs;
    g->start.push_back(s), g->end.push_back(e);
    for (int c = s; c < e; ++c) inv[c] += 1.f;
  }
  for (auto& v : inv) v = v > 0.f ? 1.0f / v : 0.f;
  g->enc = make_branch(g->ps, "Encoder.branch", cfg->n_subs, cfg->n_feats, cfg->n_blocks);
  g->enc_final = make_conv(g->ps, "Encoder.final", cfg->n_feats, cfg->latent, 3);
  g->dec = make_branch(g->ps, "Decoder.branch", cfg->latent, cfg->n_feats, cfg->n_blocks);
  g->dec_final = make_conv(g->ps, "Decoder.final", cfg->n_feats, cfg->n_subs, 3);
  g->trunk = make_branch(g->ps, "trunk", cfg->n_colors, cfg->trunk_feats, cfg->trunk_blocks);
  g->fin = make_conv(g->ps, "final", cfg->trunk_feats, cfg->n_colors, 3);
  int s = g->ps.alloc_all();
  if (s == HSIDM_OK && cudaMalloc(&g->start_dev, sizeof(int) * g->G) != cudaSuccess) s = HSIDM_CUDA_ERROR;
  if (s == HSIDM_OK && cudaMalloc(&g->inv_count_dev, sizeof(float) * cfg->n_colors) != cudaSuccess) s = HSIDM_CUDA_ERROR;
  if (s == HSIDM_OK) {
    cudaMemcpy(g->start_dev, g->start.data(), sizeof(int) * g->G, cudaMemcpyHostToDevice);
    cudaMemcpy(g->inv_count_dev, inv.data(), sizeof(float) * cfg->n_colors, cudaMemcpyHostToDevice);
    if (cfg->precision == HSIDM_BF16) s = conv_tc_init();
    if (s == HSIDM_OK && cudaEventCreateWithFlags(&g->ev_arena, cudaEventDisableTiming) != cudaSuccess) s = HSIDM_CUDA_ERROR;
  }
  if (s != HSIDM_OK) {
    if (s == HSIDM_CUDA_ERROR && g_last_error.empty()) set_last_error("device allocation failed in hsidm_gae_create");
    delete g;
    return s;
  }
  *out = g;
  return HSIDM_OK;
}

int hsidm_gae_destroy(hsidm_gae* g) {
  if (!g) return HSIDM_OK;
  DeviceGuard guard(g->device);
  cudaDeviceSynchronize();
  for_each_conv(g, [](ConvW& w) { free_conv(w); });
  if (g->start_dev) cudaFree(g->start_dev);
  if (g->inv_count_dev) cudaFree(g->inv_count_dev);
  if (g->img_off_dev) cudaFree(g->img_off_dev);
  if (g->ev_arena) cudaEventDestroy(g->ev_arena);
  delete g;
  return HSIDM_OK;
}

int hsidm_gae_param_count(const hsidm_gae* g) { return g ? g->ps.size() : 0; }
const char* hsidm_gae_param_name(const hsidm_gae* g, int i) {
  return (g && i >= 0 && i < g->ps.size()) ? g->ps.at(i).key.c_str() : nullptr;
}

int hsidm_gae_set_param(hsidm_gae* g, const char* key, const float* data, const int64_t* shape, int ndim) {
  if (!g) HSIDM_FAIL(HSIDM_BAD_ARG, "null gae handle");
  HSIDM_DEVICE(g->device);
  g->committed = false;
  return g->ps.set(key, data, shape, ndim);
}

int hsidm_gae_commit(hsidm_gae* g) {
  if (!g) HSIDM_FAIL(HSIDM_BAD_ARG, "null gae handle");
  HSIDM_DEVICE(g->device);
  HSIDM_TRY(g->ps.check_all_set());
  int status = HSIDM_OK;
  for_each_conv(g, [&](ConvW& w) {
    if (status == HSIDM_OK) status = pack_conv(g->ps, w, g->cfg.precision == HSIDM_BF16);
  });
  HSIDM_TRY(status);
  HSIDM_CUDA(cudaDeviceSynchronize());
  g->committed = true;
  g->ws_kind = 0;
  return HSIDM_OK;
}

int hsidm_gae_groups(const hsidm_gae* g, int32_t* start, int32_t* end) {
  if (!g) return 0;
  for (int k = 0; k < g->G; ++k) {
    if (start) start[k] = g->start[k];
    if (end) end[k] = g->end[k];
  }
  return g->G;
}

int hsidm_gae_encode(hsidm_gae* g, const float* x, float* z, int B, int H, int W, hsidm_stream stream) {
  if (!g || !x || !z) HSIDM_FAIL(HSIDM_BAD_ARG, "hsidm_gae_encode: null argument");
  HSIDM_DEVICE(g->device);
  HSIDM_TRY(prepare(g, 1, B, H, W));
  Exec& ex = g->ex;
  ex.stream = static_cast<cudaStream_t>(stream), ex.dry = false, ex.status = HSIDM_OK;
  ex.arena.begin(false);
  HSIDM_CUDA(cudaStreamWaitEvent(ex.stream, g->ev_arena, 0));
  encode_pass(g, x, z, B, H, W);
  HSIDM_CUDA(cudaEventRecord(g->ev_arena, ex.stream));
  return ex.status;
}

int hsidm_gae_decode(hsidm_gae* g, const float* z, float* y, int B, int H, int W, int clamp01, hsidm_stream stream) {
  if (!g || !z || !y) HSIDM_FAIL(HSIDM_BAD_ARG, "hsidm_gae_decode: null argument");
  HSIDM_DEVICE(g->device);
  HSIDM_TRY(prepare(g, 2, B, H, W));
  Exec& ex = g->ex;
  ex.stream = static_cast<cudaStream_t>(stream), ex.dry = false, ex.status = HSIDM_OK;
  ex.arena.begin(false);
  HSIDM_CUDA(cudaStreamWaitEvent(ex.stream, g->ev_arena, 0));
  decode_pass(g, z, y, B, H, W, clamp01);
  HSIDM_CUDA(cudaEventRecord(g->ev_arena, ex.stream));
  return ex.status;
}

int hsidm_gae_params_changed(hsidm_gae* g, const void* const* table_dev, int n, int* changed, hsidm_stream stream) {
  if (!g) HSIDM_FAIL(HSIDM_BAD_ARG, "null gae handle");
  HSIDM_DEVICE(g->device);
  return g->ps.differs(table_dev, n, changed, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
