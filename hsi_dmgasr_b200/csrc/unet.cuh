// Layer records and the context object behind the hsidm_ctx handle, shared by the inference executor (unet.cu) and the
// training step (train.cu).
#pragma once
#include <vector>

#include "net.cuh"

namespace hsidm {

struct ResW {
  int cin = 0, cout = 0, skip = 0;
  bool attn = false, has_res = false;
  int gn1_w = -1, gn1_b = -1, gn2_w = -1, gn2_b = -1;
  int nf_w = -1, nf_b = -1, noise_off = 0;
  ConvW c1, c2, rc;
  FusedW fused;   // conv2 + shortcut as one tensor-core GEMM (BF16 mode)
  int an_w = -1, an_b = -1;
  ConvW qkv, aout;
  AttnFoldW afold;   // folded projections (BF16 mode)
};

struct LayerW {
  enum Kind { CONV, RES, DOWN, UP } kind = CONV;
  ConvW conv;  // CONV / DOWN / UP
  ResW rb;     // RES
};

}  // namespace hsidm

using namespace hsidm;

struct hsidm_ctx {
  hsidm_unet_cfg cfg;
  int device = 0;
  ParamStore ps;
  std::vector<LayerW> downs, mid, ups;
  int fin_gn_w = -1, fin_gn_b = -1;
  ConvW fin_conv;
  int mlp1_w = -1, mlp1_b = -1, mlp3_w = -1, mlp3_b = -1;
  std::vector<NoiseLayer> noise_layers_host;  // filled at commit (device pointers)
  NoiseLayer* noise_layers_dev = nullptr;
  int noise_total = 0;
  bool committed = false;
  Exec ex;
  int64_t packed_bytes = 0;

  // schedule
  int T = 0;
  std::vector<float> coef_host;  // [T][5]
  float* coef_dev = nullptr;     // [T][5]
  float* levels_dev = nullptr;   // [T]
  float* nbias_table = nullptr;  // [T][noise_total]
  bool table_dirty = true;
  int* t_dev = nullptr;

  // persistent buffers
  float* nbias_buf = nullptr;  // [cap_n][noise_total] for hsidm_unet_forward
  int nbias_cap = 0;
  float* samp_buf = nullptr;   // cond | x | eps for hsidm_sample
  int64_t samp_cap = 0;

  // workspace bookkeeping: peak per (N,H,W) measured by a dry pass
  int ws_N = 0, ws_H = 0, ws_W = 0;

  // cached one-step CUDA graph of hsidm_sample and the arguments baked into it
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t graph_exec = nullptr;
  struct GraphKey {
    int N = 0, H = 0, W = 0;
    const float* tape = nullptr;
    int64_t s_img = 0, s_step = 0;
    float* snaps = nullptr;
    int weight_gen = 0, route_gen = 0;   // packed weights / tables and kernel routing the captured nodes point at
    bool operator==(const GraphKey& o) const {
      return N == o.N && H == o.H && W == o.W && tape == o.tape && s_img == o.s_img && s_step == o.s_step && snaps == o.snaps &&
             weight_gen == o.weight_gen && route_gen == o.route_gen;
    }
  } graph_key;
  int64_t graph_nodes = 0;  // kernels per replay (for hsidm_launch_count)
  int weight_gen = 0;       // bumped whenever packed weights or the noise-embedding table are (re)allocated
  // The arena, packed weights and tables are shared by every entry point, which may be called on different streams
  // (hsidm_unet_forward on the caller's, hsidm_sample on `side`): each pass waits for the previous one's completion.
  cudaEvent_t ev_arena = nullptr;
  // hsidm_sample runs on its own non-blocking stream (the caller's may be the legacy default stream, which cannot
  // be captured) and is stitched into the caller's stream with two events
  cudaStream_t side = nullptr;
  cudaEvent_t ev_in = nullptr, ev_out = nullptr;
  // training step (train.cu): saved activations of the last hsidm_train_forward, freed with the context
  void* train = nullptr;
  void (*train_free)(void*) = nullptr;
};

