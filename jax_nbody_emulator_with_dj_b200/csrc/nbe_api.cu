// libnbe_b200: C ABI (include/nbe.h) over the sm_100a kernels.  Host side: parameter upload,
// static description of the 15-block V-Net as 27 fused conv launches, tensor-core operand
// layout of the weights, per-shape plans (activation arena + TMA tensor maps) and the
// subbox loop with host<->device staging.
#include "../../include/nbe.h"

#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <thread>
#include <vector>

#include "aux_kernels.cuh"
#include "density.cuh"
#include "conv_mma.cuh"

using namespace nbe;

namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

enum ConvType { T_CONV3 = 0, T_SKIP1 = 1, T_DOWN = 2, T_UP = 3 };
enum Inst { I_128_128_2 = 0, I_256_256_1, I_FINAL, I_128_64_2, I_64_64_2, I_256_128_2, I_128_256_2, I_256_512_1, I_PAIR_128_256_2, I_PAIR_256_512_1, I_PAIR_128_128_2, I_PAIR_256_256_1, I_PAIR_128_256_1, I_EARLY_128_256_2, I_EARLY_256_512_1, I_F192_2, I_CHAIN_2, I_FOLD_2, I_FOLD128_1, I_FINAL_FOLD, I_COUNT };
struct InstInfo { int nrs, dc, tm; bool fin; bool pair = false; bool acc3 = false; bool early = false; bool f192 = false; bool chain = false; bool fold = false; };
const InstInfo kInst[I_COUNT] = {{128, 128, 2, false}, {256, 256, 1, false}, {32, 16, 2, true},
                                 {128, 64, 2, false},  {64, 64, 2, false},   {256, 128, 1, false},
                                 {128, 256, 2, false, false, true}, {256, 512, 1, false, false, true},
                                 // CTA-pair instances: nrs = rows of both CTAs = 2 x 1.5 x Cout
                                 {192, 256, 2, false, true, true}, {384, 512, 1, false, true, true},
                                 {192, 128, 2, false, true}, {384, 256, 1, false, true},
                                 // one tile per CTA, two accumulator stages: the acc3 epilogue overlaps the next item
                                 {192, 256, 1, false, true, true},
                                 // early-drain acc3 pair instances (fine-grained accumulator hand-over, conv_mma.cuh)
                                 {192, 256, 2, false, true, true, true}, {384, 512, 1, false, true, true, true},
                                 // F192: N = 192 fused [Wh|dW|Wl] main product, 2 x 128 weight rows per stage (conv_mma.cuh, EARLY == 2)
                                 {256, 256, 2, false, true, true, true, true},
                                 // F192 + accumulation chains (EARLY == 3): the two primal accumulators alternate every 3 taps
                                 {256, 256, 2, false, true, true, true, true, true},
                                 // FOLD (EARLY == 4): chains + folded tangent, N = 128 main product, 2 x 96 weight rows per stage
                                 {192, 256, 2, false, true, true, true, true, true, true},
                                 // folded tangent for the 128-output early-drain instance: 2 x 128 rows per stage
                                 // (main taps use the first 64: Wh halves; lo taps [Wl | Wh] halves)
                                 {256, 512, 1, false, true, true, true, false, false, true},
                                 // last layer with the folded tangent: columns [net | dnet | xh*Wl | -], 48 weight rows per stage
                                 // ([Wh | dW_res (skip only) | Wl | - ] for xh, [- | Wh] for dx'), three activation reads per tap
                                 {48, 32, 2, true, false, false, false, false, false, true}};

// activation tensors of the net
enum ActId {
  A_IN16 = 0, A_L00_0, A_L00, A_L01_0, A_Y0, A_D0, A_L1_0, A_Y1, A_D1, A_L2_0, A_Y2, A_D2, A_C_0, A_C,
  A_U2, A_R2_0, A_R2, A_U1, A_R1_0, A_R1, A_U0, A_R00_0, A_R00, A_R01_0, A_OUT, A_COUNT
};

struct Src { int act; int crop; int c0; bool kc16; };
struct ConvPart {
  int layer; int type; int off; std::vector<Src> src; int tile_base64 = 0, tile_base16 = 0;
  int fold_layer = -1;      // layer whose fold vector a is already in the stored tangent of this part's source tensor
  int fold_off = 0;         // ... input channel i of this part carries a[i + fold_off]
  int beta_layer = -1;      // layer whose beta the owning launch's epilogue applies (-1: none, or this part itself)
};
struct StaticLaunch {
  std::string name;
  int inst;
  std::vector<ConvPart> parts;
  int in_ref;      // activation whose frame defines the tile space
  int out_act;
  int cout;
  // per-sample packed weight regions (in halves, relative to the sample base)
  long long b64_off = 0, b16_off = 0;
  int n_tiles64 = 0, n_tiles16 = 0;
  long long bias_off = 0;   // floats, in the bias buffer
  // tangent folding (conv_mma.cuh, ConvLaunch::beta / anext)
  int beta_layer = -1;      // FOLD instance: layer whose beta the epilogue applies (the launch's 3^3 conv)
  int anext_layer = -1;     // layer whose fold vector a the epilogue adds to the stored tangent (fold consumer of out_act)
  int anext_off = 0;        // ... output channel j gets a[j + anext_off]
};

struct ActBuf { int c = 0, d = 0, h = 0, w = 0; size_t off_hi = 0, off_lo = 0, off_dx = 0; };

struct HostLaunch {
  ConvLaunch L;
  GroupTable G;
  int inst;
  int grid;
  double flops;
  std::string name;
};

struct Plan {
  int dims[3];
  int batch;
  ActBuf act[A_COUNT];
  size_t arena_bytes = 0;
  std::vector<HostLaunch> launches;   // [sample][launch] flattened
  int n_launch = 0;
};

struct Layer {
  std::string block, layer;
  int cout, cin, k;
  float *W = nullptr, *dW = nullptr, *SW = nullptr, *sb = nullptr;   // device
  std::vector<float> bias;
  long long w32_off = 0;    // floats into w32 / dw32 per sample
  std::vector<float> hSW, hsb;              // host copies of the style parameters (fold range check)
  float *pre_a = nullptr, *pre_beta = nullptr;   // premodulated weights: dW = W (.) (a_i + beta_o) factorisation (device)
  std::vector<float> h_pre_a;
  bool pre_ok = false;
};

}  // namespace

struct nbe_ctx {
  int device = 0;
  int num_sms = 148;
  std::string err;
  EncodeTiledFn encode = nullptr;
  cudaStream_t own_stream = nullptr, copy_stream = nullptr, up_stream = nullptr;
  cudaEvent_t mod_done = nullptr;   // recorded after the modulation kernel; every consumer stream waits on it
  bool mod_pending = false;

  bool have_params = false, premod = false, vel = true;
  float eps = 1e-8f;
  int precision = NBE_PREC_SPLIT;
  bool pair = true;         // CTA pairs (cta_group::2) for the 3^3 velocity launches (NBE_PAIR=0 disables)
  bool dbuf = false;        // 64-output acc3 pair launches: one tile per CTA + double-buffered TMEM (NBE_DBUF=0: two tiles)
  bool early = true;        // acc3 pair launches hand the per-kd accumulators over as they complete (NBE_EARLY=0: whole item)
  bool lo_box = true;       // lo-product weight stages are loaded with a box of only the rows they use (NBE_LOBOX=0)
  bool f192 = true;         // 64-output acc3 pair launches use the N = 192 fused instance (NBE_F192=0: EARLY / plain)
  bool trace = false;       // NBE_TRACE=1: upload timings on stderr (synchronises the upload stream: not for benchmarks)
  bool w_window = true;     // the first subbox's upload is windowed in W as well (NBE_WWIN=0: whole rows)
  bool chain = true;        // ... with accumulation chains of 3 taps (NBE_CHAIN=0: one chain per kd-plane)
  bool fold = true;         // ... and the folded tangent (4 instead of 5 products; NBE_FOLD=0 disables)
  bool fold_active = false; // what build_static used: fold && every fold vector in range (checked per modulation)
  bool fold_final = true;   // ... and in the last layer (NBE_FOLD_FINAL=0)
  bool fold128 = true;      // ... also in the 128-output launches (NBE_FOLD128=0)
  float fold_amax = 64.f;   // |a_i| above this (a modulation m_i close to zero) switches folding off (NBE_FOLD_AMAX)
  float* d_fold = nullptr; size_t fold_cap = 0;      // [sample][launch][beta 128 | anext 128] floats
  int band_h = 2;           // tile rows per h-band of the item order (NBE_BAND; 0: whole planes)
  bool wide16 = true;       // also for the 16-channel first layer (32-byte rows, SWIZZLE_32B row shifts; NBE_WIDE16)
  bool wide = true;         // w-halo'd activation blocks serving 9 taps per load (NBE_WIDE=0 disables)
  std::vector<Layer> layers;
  std::map<std::string, int> lidx;

  // static launch list + weight layout
  std::vector<StaticLaunch> sl;
  std::vector<LayerMeta> metas;
  long long packed_halves = 0;      // per sample
  long long w32_floats = 0;         // per sample
  LayerMeta* d_metas = nullptr;
  float* d_bias = nullptr;
  long long bias_floats = 0;

  // modulation state
  int mod_batch = 0;
  __half* d_packed = nullptr; size_t packed_cap = 0;
  float *d_w32 = nullptr, *d_dw32 = nullptr; size_t w32_cap = 0, dw32_cap = 0;
  float *d_s0 = nullptr, *d_s1 = nullptr; int s_cap = 0;

  // plans
  std::vector<Plan*> plans;
  uint8_t* arena = nullptr; size_t arena_cap = 0;
  int32_t* d_ident = nullptr; int ident_cap = 0;     // identity gather table for nbe_forward

  // process_box staging
  void* d_box = nullptr; size_t box_cap = 0;
  void* d_disp = nullptr; void* d_velo = nullptr; size_t out_cap = 0, velo_cap = 0;
  int32_t* d_idx = nullptr; size_t idx_cap = 0;

  // instrumentation
  int64_t launches = 0;
  bool profiling = false;
  std::vector<std::string> prof_names;      // per launch slot (pack_input + conv launches)
  std::vector<double> prof_flops;
  std::vector<cudaEvent_t> prof_events;     // (slots+1) events per recorded sample, resolved lazily
  std::vector<double> prof_sum_ms;
  long long prof_samples = 0;
};

namespace {

int fail(nbe_ctx* c, int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (c) c->err = buf;
  return code;
}

#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess)                                                                         \
      return fail(ctx, NBE_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
  } while (0)

size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Every entry point runs on its context's device and restores the caller's current device on exit
// (the host process may drive several GPUs, one context each, and PyTorch tracks the current device
// per thread).
struct DevGuard {
  int prev = -1;
  cudaError_t err;
  explicit DevGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    // always: cudaGetDevice reports device 0 in a thread that has no context bound yet, and the first CUDA call of
    // such a thread may be a DRIVER call (cuTensorMapEncodeTiled in build_plan, once every buffer is already
    // allocated) -- CUDA_ERROR_INVALID_CONTEXT on the second box of a two-GPU process_box_multi
    err = cudaSetDevice(dev);
    if (prev == dev) prev = -1;
  }
  ~DevGuard() { if (prev >= 0) cudaSetDevice(prev); }
};
#define ENTER_DEVICE(ctx)                                                                          \
  DevGuard dev_guard_((ctx)->device);                                                              \
  if (dev_guard_.err != cudaSuccess)                                                               \
    return fail((ctx), NBE_ERR_CUDA, "cudaSetDevice(%d) -> %s", (ctx)->device, cudaGetErrorString(dev_guard_.err))

int ensure(nbe_ctx* ctx, void** p, size_t* cap, size_t need) {
  if (*cap >= need && *p) return NBE_OK;
  if (*p) { CK(cudaDeviceSynchronize()); CK(cudaFree(*p)); *p = nullptr; *cap = 0; }
  CK(cudaMalloc(p, need));
  *cap = need;
  return NBE_OK;
}

int find_layer(nbe_ctx* ctx, const char* block, const char* layer) {
  auto it = ctx->lidx.find(std::string(block) + "/" + layer);
  return it == ctx->lidx.end() ? -1 : it->second;
}

// ----------------------------------------------------------------------------------------
// tangent folding: which layers can be the 3^3 conv of a FOLD launch, and whether their fold vectors are usable
// ----------------------------------------------------------------------------------------
bool fold_candidate(const Layer& l) { return l.k == 3 && l.cin >= 64; }

// Premodulated weights arrive as (W, dW) without the style parameters.  The reference's modulation gives
// dW[o,i,t] = W[o,i,t] * (a_i + beta_o) (style_layers_vel.py:86-93): recover a and beta from the per-(o,i) ratio
// R = <dW, W>_t / <W, W>_t = a_i + beta_o  (column means, row means minus the overall mean) and verify the
// structure element by element.  A tree that does not have it (hand-made dweight) keeps the 5-product kernels.
void factor_premod(Layer& L, const float* W, const float* dW, float amax, std::vector<float>& a, std::vector<float>& beta) {
  const int O = L.cout, I = L.cin, T = L.k * L.k * L.k;
  std::vector<double> R(static_cast<size_t>(O) * I);
  L.pre_ok = false;
  double dmax = 0;
  for (int o = 0; o < O; ++o)
    for (int i = 0; i < I; ++i) {
      double num = 0, den = 0;
      const size_t base = (static_cast<size_t>(o) * I + i) * T;
      for (int t = 0; t < T; ++t) { num += static_cast<double>(dW[base + t]) * W[base + t]; den += static_cast<double>(W[base + t]) * W[base + t];
                                    dmax = std::max(dmax, std::fabs(static_cast<double>(dW[base + t]))); }
      if (!(den > 0)) return;
      R[static_cast<size_t>(o) * I + i] = num / den;
    }
  std::vector<double> col(I, 0.0), row(O, 0.0);
  double mean = 0;
  for (int o = 0; o < O; ++o) for (int i = 0; i < I; ++i) { const double r = R[static_cast<size_t>(o) * I + i]; col[i] += r / O; row[o] += r / I; mean += r / (static_cast<double>(O) * I); }
  a.resize(I); beta.resize(O);
  for (int i = 0; i < I; ++i) { a[i] = static_cast<float>(col[i]); if (!(std::fabs(col[i]) <= amax)) return; }
  for (int o = 0; o < O; ++o) beta[o] = static_cast<float>(row[o] - mean);
  double worst = 0;
  for (int o = 0; o < O; ++o)
    for (int i = 0; i < I; ++i) {
      const size_t base = (static_cast<size_t>(o) * I + i) * T;
      const double f = static_cast<double>(a[i]) + beta[o];
      for (int t = 0; t < T; ++t) worst = std::max(worst, std::fabs(dW[base + t] - f * W[base + t]));
    }
  L.pre_ok = worst <= 1e-4 * dmax;
}

bool fold_static_ok(const nbe_ctx* ctx) {
  if (!ctx->fold || !ctx->vel || ctx->precision != NBE_PREC_SPLIT || !ctx->pair || !ctx->wide || !ctx->f192 ||
      !ctx->chain || ctx->dbuf)
    return false;
  if (ctx->premod)
    for (const auto& l : ctx->layers)
      if (fold_candidate(l) && !l.pre_ok) return false;
  return true;
}

// style models: every fold vector a_i = SW[i,1] / m_i of every candidate layer must be moderate for all samples
bool fold_range_ok(const nbe_ctx* ctx, const std::vector<float>& s0, const std::vector<float>& s1) {
  if (ctx->premod) return true;
  for (const auto& l : ctx->layers) {
    if (!fold_candidate(l)) continue;
    for (size_t b = 0; b < s0.size(); ++b)
      for (int i = 0; i < l.cin; ++i) {
        const float m = s0[b] * l.hSW[2 * i] + s1[b] * l.hSW[2 * i + 1] + l.hsb[i];
        const float a = l.hSW[2 * i + 1] / m;
        if (!(std::fabs(a) <= ctx->fold_amax)) return false;
      }
  }
  return true;
}

// ----------------------------------------------------------------------------------------
// static description of the net as fused launches (see DESIGN.md "launch list")
// ----------------------------------------------------------------------------------------
int build_static(nbe_ctx* ctx) {
  const bool vel = ctx->vel;
  const bool split = ctx->precision == NBE_PREC_SPLIT;
  ctx->sl.clear();
  auto L = [&](const char* b, const char* l) { return find_layer(ctx, b, l); };
  // vel + split precision: one primal accumulator per kd ("acc3")
  auto inst64 = [&]() { return vel ? (split ? I_128_256_2 : I_128_128_2) : (split ? I_128_64_2 : I_64_64_2); };
  auto inst128 = [&]() { return vel ? (split ? I_256_512_1 : I_256_256_1) : (split ? I_256_128_2 : I_128_128_2); };

  auto add = [&](const std::string& name, int inst, int in_ref, int out_act, int cout,
                 std::vector<ConvPart> parts) {
    StaticLaunch s;
    s.name = name; s.inst = inst; s.in_ref = in_ref; s.out_act = out_act; s.cout = cout; s.parts = parts;
    ctx->sl.push_back(s);
  };
  auto res_block = [&](const char* blk, std::vector<Src> in, int in_ref, int mid_act, int out_act, int mid,
                       bool fin) {
    add(std::string(blk) + ".conv_0", mid == 128 ? inst128() : inst64(), in_ref, mid_act, mid,
        {ConvPart{L(blk, "conv_0"), T_CONV3, 0, in}});
    std::vector<Src> midsrc;
    for (int q = 0; q < mid / 64; ++q) midsrc.push_back(Src{mid_act, 0, q * 64, false});
    add(std::string(blk) + ".conv_1+skip", fin ? I_FINAL : inst64(), mid_act, out_act, fin ? 3 : 64,
        {ConvPart{L(blk, "conv_1"), T_CONV3, 0, midsrc}, ConvPart{L(blk, "skip"), T_SKIP1, 2, in}});
  };
  auto S = [](int act, int crop = 0, int c0 = 0, bool kc16 = false) { return Src{act, crop, c0, kc16}; };

  res_block("conv_l00", {S(A_IN16, 0, 0, true)}, A_IN16, A_L00_0, A_L00, 64, false);
  res_block("conv_l01", {S(A_L00)}, A_L00, A_L01_0, A_Y0, 64, false);
  add("down_l0", inst64(), A_D0, A_D0, 64, {ConvPart{L("down_l0", "conv_0"), T_DOWN, 0, {S(A_Y0)}}});
  res_block("conv_l1", {S(A_D0)}, A_D0, A_L1_0, A_Y1, 64, false);
  add("down_l1", inst64(), A_D1, A_D1, 64, {ConvPart{L("down_l1", "conv_0"), T_DOWN, 0, {S(A_Y1)}}});
  res_block("conv_l2", {S(A_D1)}, A_D1, A_L2_0, A_Y2, 64, false);
  add("down_l2", inst64(), A_D2, A_D2, 64, {ConvPart{L("down_l2", "conv_0"), T_DOWN, 0, {S(A_Y2)}}});
  res_block("conv_c", {S(A_D2)}, A_D2, A_C_0, A_C, 64, false);
  add("up_r2", inst64(), A_C, A_U2, 64, {ConvPart{L("up_r2", "conv_0"), T_UP, 0, {S(A_C)}}});
  res_block("conv_r2", {S(A_Y2, 4), S(A_U2)}, A_U2, A_R2_0, A_R2, 128, false);
  add("up_r1", inst64(), A_R2, A_U1, 64, {ConvPart{L("up_r1", "conv_0"), T_UP, 0, {S(A_R2)}}});
  res_block("conv_r1", {S(A_Y1, 16), S(A_U1)}, A_U1, A_R1_0, A_R1, 128, false);
  add("up_r0", inst64(), A_R1, A_U0, 64, {ConvPart{L("up_r0", "conv_0"), T_UP, 0, {S(A_R1)}}});
  res_block("conv_r00", {S(A_Y0, 40), S(A_U0)}, A_U0, A_R00_0, A_R00, 128, false);
  res_block("conv_r01", {S(A_R00)}, A_R00, A_R01_0, A_OUT, 64, true);

  for (auto& s : ctx->sl)
    for (auto& p : s.parts)
      if (p.layer < 0) return fail(ctx, NBE_ERR_STATE, "layer missing for launch %s", s.name.c_str());
  // one primal accumulator per kd only where it matters (27-tap convs over 64-channel inputs); the
  // K = 16 first layer and the 1- / 8-tap resampling layers accumulate few terms and keep the
  // double-buffered single-accumulator instances (epilogue overlapped with the next item's MMAs)
  for (auto& s : ctx->sl) {
    bool deep = false;
    for (auto& p : s.parts) deep = deep || (p.type == T_CONV3 && !p.src[0].kc16);
    if (!deep && s.inst == I_128_256_2) s.inst = I_128_128_2;
    if (!deep && s.inst == I_256_512_1) s.inst = I_256_256_1;
  }
  if (ctx->pair)        // 3^3 (+ folded skip) launches of the split-precision velocity net run on CTA pairs
    for (auto& s : ctx->sl) {
      bool ok = ctx->wide && vel;
      for (auto& p : s.parts) ok = ok && (p.type == T_CONV3 || p.type == T_SKIP1);
      if (!ok) continue;
      if (s.inst == I_128_256_2)
        s.inst = ctx->dbuf ? I_PAIR_128_256_1
                 : (ctx->f192 ? (ctx->chain ? (ctx->fold_active ? I_FOLD_2 : I_CHAIN_2) : I_F192_2)
                              : (ctx->early ? I_EARLY_128_256_2 : I_PAIR_128_256_2));
      else if (s.inst == I_256_512_1)
        s.inst = ctx->early ? ((ctx->fold_active && ctx->fold128) ? I_FOLD128_1 : I_EARLY_256_512_1) : I_PAIR_256_512_1;
      else if (s.inst == I_128_128_2) s.inst = I_PAIR_128_128_2;
      else if (s.inst == I_256_256_1) s.inst = I_PAIR_256_256_1;
    }

  if (ctx->fold_active && ctx->fold_final)
    for (auto& s : ctx->sl)
      if (s.inst == I_FINAL) s.inst = I_FINAL_FOLD;

  // ---- tangent folding: the 3^3 conv of a FOLD launch defines the fold vector of the tensor(s) it reads; the
  // launch producing such a tensor adds a (.) y to the tangent it stores, every other reader subtracts it again
  // in its tangent weights (modulate_kernel)
  {
    // channel j of tensor X carries a[j + act_off[X]] of layer act_fold[X] (K-chunk q of that layer reads the 64
    // channels of its q-th source from c0 on)
    std::vector<int> act_fold(A_COUNT, -1), act_off(A_COUNT, 0);
    for (auto& s : ctx->sl) {
      if (!kInst[s.inst].fold) continue;
      for (auto& p : s.parts)
        if (p.type == T_CONV3) {
          s.beta_layer = p.layer;
          for (size_t q = 0; q < p.src.size(); ++q) {
            const Src& sc = p.src[q];
            const int off = 64 * static_cast<int>(q) - sc.c0;
            if (act_fold[sc.act] >= 0 && (act_fold[sc.act] != p.layer || act_off[sc.act] != off))
              return fail(ctx, NBE_ERR_STATE, "launch %s: source tensor already has a fold consumer", s.name.c_str());
            act_fold[sc.act] = p.layer; act_off[sc.act] = off;
          }
        }
    }
    for (auto& s : ctx->sl) {
      const bool has = vel && s.out_act >= 0 && s.out_act < A_COUNT;
      s.anext_layer = has ? act_fold[s.out_act] : -1;
      s.anext_off = has ? act_off[s.out_act] : 0;
      for (auto& p : s.parts) {
        // this part's input channel i is channel i - 64 q + c0 of its q-th source
        p.fold_layer = act_fold[p.src[0].act];
        p.fold_off = p.src[0].c0 + act_off[p.src[0].act];
        for (size_t q = 0; q < p.src.size(); ++q) {
          const Src& sc = p.src[q];
          if (act_fold[sc.act] != p.fold_layer || (p.fold_layer >= 0 && sc.c0 - 64 * static_cast<int>(q) + act_off[sc.act] != p.fold_off))
            return fail(ctx, NBE_ERR_STATE, "launch %s: sources with different fold vectors", s.name.c_str());
        }
        p.beta_layer = (s.beta_layer >= 0 && s.beta_layer != p.layer) ? s.beta_layer : -1;
      }
    }
  }

  // ---- weight layout: tiles, emit rules, LayerMeta
  const int nkind = (vel && split) ? 2 : 1;
  ctx->metas.assign(ctx->layers.size(), LayerMeta{});
  long long off = 0, boff = 0;
  int row0 = 0;
  for (size_t li = 0; li < ctx->layers.size(); ++li) {     // block order of the modulate grid
    ctx->metas[li].row0 = row0;
    row0 += ctx->layers[li].cout;
  }
  for (auto& s : ctx->sl) {
    const InstInfo ii = kInst[s.inst];
    int t64 = 0, t16 = 0;
    for (auto& p : s.parts) {
      const Layer& ly = ctx->layers[p.layer];
      const bool k16 = p.src[0].kc16;
      const int nkc = static_cast<int>(p.src.size());
      // FOLD: the folded 1^3 skip has a third tile kind, its residual tangent rows (x * dW_res -> dy)
      const int nk = k16 ? 1 : ((ii.fold && !ii.fin && p.type == T_SKIP1) ? 3 : nkind);
      LayerMeta& M = ctx->metas[p.layer];
      int& tb = k16 ? t16 : t64;
      (k16 ? p.tile_base16 : p.tile_base64) = tb;
      M.kc16 = k16; M.nrs = ii.nrs;
      const int k3 = ly.k * ly.k * ly.k;
      if (p.type == T_CONV3) {
        // stage order (kd, kc, kind, kw, kh): the 9 taps of a (kd, kc, kind) group are consecutive
        for (int kd = 0; kd < 3; ++kd) for (int kh = 0; kh < 3; ++kh) for (int kw = 0; kw < 3; ++kw)
          M.tap_tile[(kd * 3 + kh) * 3 + kw] = tb + (kd * nkc * nk) * 9 + kw * 3 + kh;
        M.kc_stride = nk * 9; M.kind_stride = 9;
        tb += 27 * nkc * nk;
      } else if (p.type == T_SKIP1) {
        M.tap_tile[0] = tb; M.kc_stride = nk; M.kind_stride = 1; tb += nkc * nk;
      } else if (p.type == T_DOWN) {
        for (int t = 0; t < 8; ++t) M.tap_tile[t] = tb + t * nkc * nk;
        M.kc_stride = nk; M.kind_stride = 1; tb += 8 * nkc * nk;
      } else {  // T_UP: parity p uses tap 7-p
        for (int t = 0; t < 8; ++t) M.tap_tile[t] = tb + (7 - t) * nkc * nk;
        M.kc_stride = nk; M.kind_stride = 1; tb += 8 * nkc * nk;
      }
      (void)k3;
      // emit rules
      int nr = 0;
      auto rule = [&](int what, int kind, int row, int kcol, int alt = 0) {
        M.rules[nr].what = static_cast<int8_t>(what); M.rules[nr].kind = static_cast<int8_t>(kind);
        M.rules[nr].row_base = static_cast<int16_t>(row); M.rules[nr].kcol = static_cast<int16_t>(kcol);
        M.rules[nr].alt_kd1 = static_cast<int16_t>(alt);
        ++nr;
      };
      const bool acc3 = ii.acc3;
      M.pair_rows = ii.nrs / 2;
      auto prule = [&](int what, int kind, int kdmask, int cta_base, int mod, int base, int kcol = 0) {
        EmitRule& R = M.rules[nr++];
        R.what = static_cast<int8_t>(what); R.kind = static_cast<int8_t>(kind); R.row_base = static_cast<int16_t>(base);
        R.kcol = static_cast<int16_t>(kcol); R.alt_kd1 = 0; R.mod = static_cast<int16_t>(mod);
        R.cta_base = static_cast<int8_t>(cta_base); R.kd_mask = static_cast<int8_t>(kdmask);
      };
      const int C = ly.cout;
      if (ii.fin && ii.fold) {
        // xh * rows [0, 32) = [Wh | dW_res | Wl | -] -> (net, dnet, ylo, -);  dx' * rows [32, 48) = [- | Wh] -> dnet;
        // lo tile: xl * rows [0, 16) = [Wh | -] -> net.  dW_res is zero for the 3^3 conv itself (not emitted).
        rule(EMIT_WH, 0, 0, 0); rule(EMIT_WL, 0, 16, 0); rule(EMIT_WH, 0, 40, 0);
        if (p.type == T_SKIP1) rule(EMIT_DW, 0, 8, 0);
        rule(EMIT_WH, 1, 0, 0);
      } else if (ii.fin) {
        if (vel) {
          rule(EMIT_WH, 0, 0, 0); rule(EMIT_DW, 0, 8, 0); rule(EMIT_WH, 0, 24, 0);
          if (split) { rule(EMIT_WL, 1, 0, 0); rule(EMIT_WH, 1, 16, 0); }
        } else {
          rule(EMIT_WH, 0, 0, 0);
          if (split) { rule(EMIT_WL, 0, 8, 0); rule(EMIT_WH, 0, 16, 0); }
        }
      } else if (ii.fold && k16) {   // 16-channel folded skip of a FOLD launch: N = 2C rows [dW | Wh..] -> (dy, y1)
        prule(EMIT_DW, 0, 0, 0, C, 0, 0);
        prule(EMIT_WH, 0, 0, 1, C, 0, 0); prule(EMIT_WH, 0, 0, 1, C, 0, 3); prule(EMIT_WL, 0, 0, 1, C, 0, 6);
      } else if (ii.fold && ii.tm == 1) {
        // 128-output folded instance: main taps stage this CTA's half of Wh (xh * Wh -> y_kd, dx' * Wh -> dy share the
        // rows), lo taps [Wl half | Wh half] (xh * Wl + xl * Wh -> y0)
        prule(EMIT_WH, 0, 0, 0, C / 2, 0);
        prule(EMIT_WL, 1, 0, 0, C / 2, 0);      prule(EMIT_WH, 1, 0, 0, C / 2, C / 2);
      } else if (ii.fold) {
        // per-CTA stage (1.5C rows): rows [0, C) = this CTA's half of the N = 2C operand ([Wl | Wh] -> (ylo, y0) for
        // blocks of phase 0, [Wh | Wl] -> (y1, ylo) for phase 1), rows [C, 1.5C) = its half of Wh for dx' * Wh.
        // lo stage: this CTA's half of Wh (xl * Wh -> y1); skip only: its half of the residual tangent rows.
        auto xrule = [&](int what, int kind, int kdmask, int cta, int base, int o0, int o1) {
          EmitRule& R = M.rules[nr++];
          R.what = static_cast<int8_t>(what); R.kind = static_cast<int8_t>(kind); R.row_base = static_cast<int16_t>(base);
          R.kcol = 0; R.alt_kd1 = 0; R.mod = 0; R.cta_base = static_cast<int8_t>(cta); R.kd_mask = static_cast<int8_t>(kdmask);
          R.o_min = static_cast<int16_t>(o0); R.o_max = static_cast<int16_t>(o1);
        };
        const int H2 = C / 2;
        xrule(EMIT_WL, 0, 0b01, 0, 0, 0, C);         xrule(EMIT_WH, 0, 0b01, 1, 0, 0, C);
        xrule(EMIT_WH, 0, 0b10, 0, 0, 0, C);         xrule(EMIT_WL, 0, 0b10, 1, 0, 0, C);
        xrule(EMIT_WH, 0, 0, 0, C, 0, H2);           xrule(EMIT_WH, 0, 0, 1, C, H2, C);           // dx' * Wh halves
        xrule(EMIT_WH, 1, 0, 0, 0, 0, H2);           xrule(EMIT_WH, 1, 0, 1, 0, H2, C);           // lo stage
        if (p.type == T_SKIP1) { xrule(EMIT_DW, 2, 0, 0, 0, 0, H2); xrule(EMIT_DW, 2, 0, 1, 0, H2, C); }
      } else if (ii.pair && k16) {   // N = 2C rows [Wh.. | dW]: CTA0 stages the primal rows, CTA1 the tangent rows
        prule(EMIT_WH, 0, 0, 0, C, 0, 0); prule(EMIT_WH, 0, 0, 0, C, 0, 3); prule(EMIT_WL, 0, 0, 0, C, 0, 6);
        prule(EMIT_DW, 0, 0, 1, C, 0, 0);
      } else if (ii.f192 && !k16) {
        // per-CTA stage (2C rows): rows [0, 1.5C) = this CTA's half of the N = 3C operand, rows [1.5C, 2C) = its half
        // of Wh for dx * Wh.  kd 0 / 2 (and the folded 1^3 skip): [dW | Wl | Wh] -> (dy, ylo, y0); kd 1: [Wh | dW | Wl]
        // -> (y1, dy, ylo).  lo stage: this CTA's half of Wh (xl * Wh -> y1).
        auto xrule = [&](int what, int kind, int kdmask, int cta, int base, int o0, int o1) {
          EmitRule& R = M.rules[nr++];
          R.what = static_cast<int8_t>(what); R.kind = static_cast<int8_t>(kind); R.row_base = static_cast<int16_t>(base);
          R.kcol = 0; R.alt_kd1 = 0; R.mod = 0; R.cta_base = static_cast<int8_t>(cta); R.kd_mask = static_cast<int8_t>(kdmask);
          R.o_min = static_cast<int16_t>(o0); R.o_max = static_cast<int16_t>(o1);
        };
        const int H2 = C / 2;
        // kd masks 0b101 / 0b010; with accumulation chains the same two row orders are selected by the phase of
        // the tap's 3-tap block instead (bit 0: y0 side, bit 1: y1 side), see LayerMeta::chain
        const int m0 = ii.chain ? 0b01 : 0b101;
        xrule(EMIT_DW, 0, m0, 0, 0, 0, C);           xrule(EMIT_WL, 0, m0, 0, C, 0, H2);
        xrule(EMIT_WL, 0, m0, 1, 0, H2, C);          xrule(EMIT_WH, 0, m0, 1, H2, 0, C);
        xrule(EMIT_WH, 0, 0b010, 0, 0, 0, C);        xrule(EMIT_DW, 0, 0b010, 0, C, 0, H2);
        xrule(EMIT_DW, 0, 0b010, 1, 0, H2, C);       xrule(EMIT_WL, 0, 0b010, 1, H2, 0, C);
        xrule(EMIT_WH, 0, 0, 0, C + H2, 0, H2);      xrule(EMIT_WH, 0, 0, 1, C + H2, H2, C);      // dx * Wh halves
        xrule(EMIT_WH, 1, 0, 0, 0, 0, H2);           xrule(EMIT_WH, 1, 0, 1, 0, H2, C);           // lo stage
      } else if (ii.pair && acc3) {
        // per-CTA stage rows: kd 0 / kd 1 [R0: C | R1: C/2], kd 2 [Wh half | dW half],
        // lo [Wl half | Wh half]; non-3^3 terms (folded skip) use the kd 0 form
        prule(EMIT_WH, 0, 0b001, 0, C, 0);      prule(EMIT_DW, 0, 0b001, 1, C, 0);         // kd0: xh*[Wh|dW] -> (y0, dy)
        prule(EMIT_DW, 0, 0b010, 0, C, 0);      prule(EMIT_WH, 0, 0b010, 1, C, 0);         // kd1: xh*[dW|Wh] -> (dy, y1)
        prule(EMIT_WH, 0, 0b011, 0, C / 2, C);                                             // kd0/1: dx*Wh -> dy
        prule(EMIT_WH, 0, 0b100, 0, C / 2, 0);  prule(EMIT_DW, 0, 0b100, 0, C / 2, C / 2); // kd2
        prule(EMIT_WL, 1, 0, 0, C / 2, 0);      prule(EMIT_WH, 1, 0, 0, C / 2, C / 2);     // lo products
      } else if (ii.pair) {          // single-product velocity: [R0: Wh|dW halves (C) | R1: Wh halves (C/2)]
        prule(EMIT_WH, 0, 0, 0, C, 0); prule(EMIT_DW, 0, 0, 1, C, 0); prule(EMIT_WH, 0, 0, 0, C / 2, C);
      } else if (k16) {
        rule(EMIT_WH, 0, 0, 0); rule(EMIT_WH, 0, 0, 3); rule(EMIT_WL, 0, 0, 6);
        if (vel) rule(EMIT_DW, 0, C, 0);
      } else if (vel) {
        // main stage [Wh | dW]; in acc3 mode the kd == 1 taps are packed [dW | Wh] so that one
        // N = 2C MMA lands on the adjacent accumulator blocks (dy, y1)
        rule(EMIT_WH, 0, 0, 0, acc3 ? C : 0); rule(EMIT_DW, 0, C, 0, acc3 ? -C : 0);
        if (split) { rule(EMIT_WL, 1, 0, 0); rule(EMIT_WH, 1, C, 0); }
      } else {
        rule(EMIT_WH, 0, 0, 0);
        if (split) rule(EMIT_WL, 0, C, 0);
      }
      M.n_rules = nr;
      M.chain = (ii.chain && !k16) ? 1 : 0;
      M.chain_nkc = nkc;
      M.chain_par0 = 0;
      if (M.chain && p.type == T_CONV3) {
        // blocks are numbered in launch order: the folded 64-channel skip's one-tap blocks first, then the
        // 3-tap blocks of (kd, kc, kw); a 16-channel skip belongs to the lo chain and has no block
        for (auto& p2 : s.parts)
          if (p2.type == T_SKIP1 && !p2.src[0].kc16) M.chain_par0 = static_cast<int>(p2.src.size()) & 1;
      }
    }
    s.n_tiles64 = t64; s.n_tiles16 = t16;
    s.b64_off = off; off += static_cast<long long>(t64) * ii.nrs * 64;
    s.b16_off = off; off += static_cast<long long>(t16) * ii.nrs * 16;
    off = static_cast<long long>(align_up(static_cast<size_t>(off), 512));
    s.bias_off = boff; boff += 128;
  }
  ctx->packed_halves = off;
  ctx->bias_floats = boff;
  // per-layer fp32 output offsets and destination pointers are filled in at modulate time
  long long w32 = 0;
  for (auto& ly : ctx->layers) { ly.w32_off = w32; w32 += static_cast<long long>(ly.cout) * ly.cin * ly.k * ly.k * ly.k; }
  ctx->w32_floats = w32;

  // bias buffers (conv bias + folded skip bias), fp32
  {
    std::vector<float> hb(static_cast<size_t>(boff), 0.f);
    for (auto& s : ctx->sl)
      for (auto& p : s.parts) {
        const Layer& ly = ctx->layers[p.layer];
        for (int o = 0; o < ly.cout; ++o) hb[s.bias_off + o] += ly.bias[o];
      }
    if (ctx->d_bias) { cudaFree(ctx->d_bias); ctx->d_bias = nullptr; }
    CK(cudaMalloc(&ctx->d_bias, hb.size() * sizeof(float)));
    CK(cudaMemcpy(ctx->d_bias, hb.data(), hb.size() * sizeof(float), cudaMemcpyHostToDevice));
  }
  // plans depend on the layout: drop them
  for (auto* p : ctx->plans) { delete p; }
  ctx->plans.clear();
  ctx->mod_batch = 0;
  return NBE_OK;
}

// ----------------------------------------------------------------------------------------
// tensor maps
// ----------------------------------------------------------------------------------------
int make_act_map(nbe_ctx* ctx, CUtensorMap* m, const __half* base, int C, int W, int H, int D, int kc, int box_h,
                 int par /* -1 or parity (a,b,c) of a stride-2 view */, int box_w = 8) {
  cuuint64_t dims[4] = {static_cast<cuuint64_t>(C), static_cast<cuuint64_t>(W), static_cast<cuuint64_t>(H),
                        static_cast<cuuint64_t>(D)};
  cuuint64_t str[3] = {static_cast<cuuint64_t>(C) * 2, static_cast<cuuint64_t>(C) * W * 2,
                       static_cast<cuuint64_t>(C) * W * H * 2};
  const __half* p = base;
  if (par >= 0) {
    const int a = (par >> 2) & 1, b = (par >> 1) & 1, c = par & 1;
    p = base + ((static_cast<size_t>(a) * H + b) * W + c) * C;
    dims[1] = W / 2; dims[2] = H / 2; dims[3] = D / 2;
    str[0] *= 2; str[1] *= 2; str[2] *= 2;
  }
  cuuint32_t box[4] = {static_cast<cuuint32_t>(kc), static_cast<cuuint32_t>(box_w), static_cast<cuuint32_t>(box_h), 1u};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = ctx->encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<__half*>(p), dims, str, box, es,
                           CU_TENSOR_MAP_INTERLEAVE_NONE,
                           kc == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_32B,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(ctx, NBE_ERR_CUDA, "cuTensorMapEncodeTiled(act C=%d W=%d H=%d D=%d) -> %d", C, W, H, D, (int)r);
  return NBE_OK;
}

int make_b_map(nbe_ctx* ctx, CUtensorMap* m, const __half* base, int kc, long long rows, int nrs) {
  if (rows <= 0) { memset(m, 0, sizeof(*m)); return NBE_OK; }
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(kc), static_cast<cuuint64_t>(rows)};
  cuuint64_t str[1] = {static_cast<cuuint64_t>(kc) * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(kc), static_cast<cuuint32_t>(nrs)};
  cuuint32_t es[2] = {1, 1};
  CUresult r = ctx->encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half*>(base), dims, str, box, es,
                           CU_TENSOR_MAP_INTERLEAVE_NONE,
                           kc == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_32B,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(ctx, NBE_ERR_CUDA, "cuTensorMapEncodeTiled(B kc=%d rows=%lld) -> %d", kc, rows, (int)r);
  return NBE_OK;
}

// 16-channel weight tiles as {16 ch, nrs rows, tiles}: one box = this CTA's rows of three consecutive taps
int make_b_map16(nbe_ctx* ctx, CUtensorMap* m, const __half* base, int n_tiles, int nrs, int box_rows) {
  if (n_tiles <= 0) { memset(m, 0, sizeof(*m)); return NBE_OK; }
  cuuint64_t dims[3] = {16, static_cast<cuuint64_t>(nrs), static_cast<cuuint64_t>(n_tiles)};
  cuuint64_t str[2] = {32, static_cast<cuuint64_t>(nrs) * 32};
  cuuint32_t box[3] = {16, static_cast<cuuint32_t>(box_rows), 3};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = ctx->encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<__half*>(base), dims, str, box, es,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(ctx, NBE_ERR_CUDA, "cuTensorMapEncodeTiled(B16 tiles=%d) -> %d", n_tiles, (int)r);
  return NBE_OK;
}

// 64-channel weight tiles as {64 ch, nrs rows, tiles}: one box = `box_rows` rows of three consecutive taps
int make_b_map_lo3(nbe_ctx* ctx, CUtensorMap* m, const __half* base, int n_tiles, int nrs, int box_rows) {
  if (n_tiles <= 0) { memset(m, 0, sizeof(*m)); return NBE_OK; }
  cuuint64_t dims[3] = {64, static_cast<cuuint64_t>(nrs), static_cast<cuuint64_t>(n_tiles)};
  cuuint64_t str[2] = {128, static_cast<cuuint64_t>(nrs) * 128};
  cuuint32_t box[3] = {64, static_cast<cuuint32_t>(box_rows), 3};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = ctx->encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<__half*>(base), dims, str, box, es,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(ctx, NBE_ERR_CUDA, "cuTensorMapEncodeTiled(B lo3 tiles=%d) -> %d", n_tiles, (int)r);
  return NBE_OK;
}

// ----------------------------------------------------------------------------------------
// per-shape plan
// ----------------------------------------------------------------------------------------
int chain(nbe_ctx* ctx, int n, int out[A_COUNT]) {
  if (n % 8 != 0 || n < 104) return fail(ctx, NBE_ERR_ARG, "spatial size %d must be a multiple of 8 and >= 104", n);
  out[A_IN16] = n; out[A_L00_0] = n - 2; out[A_L00] = n - 4; out[A_L01_0] = n - 6; out[A_Y0] = n - 8;
  out[A_D0] = out[A_Y0] / 2; out[A_L1_0] = out[A_D0] - 2; out[A_Y1] = out[A_D0] - 4;
  out[A_D1] = out[A_Y1] / 2; out[A_L2_0] = out[A_D1] - 2; out[A_Y2] = out[A_D1] - 4;
  out[A_D2] = out[A_Y2] / 2; out[A_C_0] = out[A_D2] - 2; out[A_C] = out[A_D2] - 4;
  out[A_U2] = 2 * out[A_C]; out[A_R2_0] = out[A_U2] - 2; out[A_R2] = out[A_U2] - 4;
  out[A_U1] = 2 * out[A_R2]; out[A_R1_0] = out[A_U1] - 2; out[A_R1] = out[A_U1] - 4;
  out[A_U0] = 2 * out[A_R1]; out[A_R00_0] = out[A_U0] - 2; out[A_R00] = out[A_U0] - 4;
  out[A_R01_0] = out[A_R00] - 2; out[A_OUT] = out[A_R00] - 4;
  if (out[A_Y2] - out[A_U2] != 8 || out[A_Y1] - out[A_U1] != 32 || out[A_Y0] - out[A_U0] != 80 || out[A_OUT] != n - 96)
    return fail(ctx, NBE_ERR_ARG, "size %d does not produce consistent skip crops", n);
  return NBE_OK;
}

int act_channels(int a) {
  if (a == A_IN16) return 16;
  if (a == A_R2_0 || a == A_R1_0 || a == A_R00_0) return 128;
  if (a == A_OUT) return 0;
  return 64;
}

int build_plan(nbe_ctx* ctx, const int32_t dims[3], int batch, Plan** out) {
  for (auto* p : ctx->plans)
    if (p->dims[0] == dims[0] && p->dims[1] == dims[1] && p->dims[2] == dims[2] && p->batch == batch) { *out = p; return NBE_OK; }
  const bool vel = ctx->vel;
  const bool split = ctx->precision == NBE_PREC_SPLIT;
  Plan* P = new Plan();
  P->dims[0] = dims[0]; P->dims[1] = dims[1]; P->dims[2] = dims[2]; P->batch = batch;
  int cd[A_COUNT], ch[A_COUNT], cw[A_COUNT];
  int rc;
  if ((rc = chain(ctx, dims[0], cd)) || (rc = chain(ctx, dims[1], ch)) || (rc = chain(ctx, dims[2], cw))) { delete P; return rc; }
  size_t off = 0;
  for (int a = 0; a < A_COUNT; ++a) {
    ActBuf& B = P->act[a];
    B.c = act_channels(a); B.d = cd[a]; B.h = ch[a]; B.w = cw[a];
    if (a == A_OUT) continue;
    const size_t bytes = align_up(static_cast<size_t>(B.c) * B.d * B.h * B.w * 2, 1024);
    B.off_hi = off; off += bytes;
    if (a != A_IN16) {
      if (split) { B.off_lo = off; off += bytes; }
      if (vel) { B.off_dx = off; off += bytes; }
    }
  }
  P->arena_bytes = off;
  if (ctx->arena_cap < off) {
    rc = ensure(ctx, reinterpret_cast<void**>(&ctx->arena), &ctx->arena_cap, off);
    if (rc) { delete P; return rc; }
    // the arena moved: every cached plan holds stale pointers
    for (auto* p : ctx->plans) { delete p; }
    ctx->plans.clear();
  }
  uint8_t* arena = ctx->arena;
  auto hi = [&](int a) { return reinterpret_cast<__half*>(arena + P->act[a].off_hi); };
  auto lo = [&](int a) { return reinterpret_cast<__half*>(arena + P->act[a].off_lo); };
  auto dx = [&](int a) { return reinterpret_cast<__half*>(arena + P->act[a].off_dx); };

  const int nl = static_cast<int>(ctx->sl.size());
  P->n_launch = nl;
  P->launches.resize(static_cast<size_t>(nl) * batch);
  for (int b = 0; b < batch; ++b) {
    const int wb = (ctx->mod_batch == 1) ? 0 : b;       // shared weights when modulated for one sample
    const __half* packed = ctx->d_packed + static_cast<size_t>(wb) * ctx->packed_halves;
    for (int li = 0; li < nl; ++li) {
      const StaticLaunch& s = ctx->sl[li];
      const InstInfo ii = kInst[s.inst];
      HostLaunch& H = P->launches[static_cast<size_t>(b) * nl + li];
      H.inst = s.inst; H.name = s.name; H.flops = 0;
      ConvLaunch& Lc = H.L;
      memset(&Lc, 0, sizeof Lc);
      memset(&H.G, 0, sizeof H.G);
      const int box_h = 16 * ii.tm + 2;
      std::map<std::pair<const void*, int>, int> amap;
      int n_amap = 0;
      int cur_box_w = 8;
      auto get_map = [&](const __half* ptr, int act, int par) -> int {
        auto key = std::make_pair(static_cast<const void*>(ptr), par * 16 + cur_box_w);
        auto it = amap.find(key);
        if (it != amap.end()) return it->second;
        if (n_amap >= kMaxAMaps) return -1;
        const ActBuf& B = P->act[act];
        if (make_act_map(ctx, &Lc.amap[n_amap], ptr, B.c, B.w, B.h, B.d, B.c == 16 ? 16 : 64, box_h, par, cur_box_w)) return -1;
        amap[key] = n_amap;
        return n_amap++;
      };
      const int box_rows = ii.pair ? ii.nrs / 2 : ii.nrs;
      const bool fold128 = ii.fold && ii.tm == 1;
      Lc.main_rows = fold128 ? s.cout / 2 : 0;
      if ((rc = make_b_map(ctx, &Lc.bmap64, packed + s.b64_off, 64, static_cast<long long>(s.n_tiles64) * ii.nrs, fold128 ? s.cout / 2 : box_rows)) ||
          (rc = make_b_map16(ctx, &Lc.bmap16, packed + s.b16_off, s.n_tiles16, ii.nrs, box_rows))) { delete P; return rc; }
      // pair acc3 layout: a lo stage [Wl half | Wh half] fills only the first `cout` of the 1.5 x cout rows
      // each CTA owns per stage; loading just those saves a sixth of the weight traffic
      Lc.lo_rows = 0;
      memset(&Lc.bmap64_lo, 0, sizeof Lc.bmap64_lo);
      if (ctx->lo_box && ii.pair && ii.acc3 && s.n_tiles64 > 0) {
        Lc.lo_rows = ii.f192 ? s.cout / 2 : s.cout;
        Lc.lo_taps = ii.f192 ? 3 : 1;
        if (ii.f192) rc = make_b_map_lo3(ctx, &Lc.bmap64_lo, packed + s.b64_off, s.n_tiles64, ii.nrs, Lc.lo_rows);
        else rc = make_b_map(ctx, &Lc.bmap64_lo, packed + s.b64_off, 64, static_cast<long long>(s.n_tiles64) * ii.nrs, Lc.lo_rows);
        if (rc) { delete P; return rc; }
      }

      int ng = 0;
      bool bad = false;
      const bool lo3 = ctx->lo_box && ii.f192;     // F192: a lo tap is 32 rows per CTA; three of them share a stage
      // lo-product groups are issued before the main groups (see conv_mma.cuh: truncating
      // accumulation), so collect them separately
      // main groups are further ordered by the accumulator they complete: kd 0 and the folded skip
      // (y0, dy) | kd 1 (dy, y1) | kd 2 (y2, dy), which is what lets the EARLY instances drain y0 / y1
      // while later kd-planes are still running
      std::vector<GroupDesc> g_lo, g_main[3], g_skip16, g_skip64, g_dwres;   // the skip lists are used by the chain instances only
      int cur_kind = 0, cur_kd = 0, cur_skip = 0;
      auto push = [&](const GroupDesc& G) {
        if (cur_kind == 1) g_lo.push_back(G);
        else if (cur_kind == 2) g_dwres.push_back(G);
        else if (ii.chain && cur_skip == 1) g_skip16.push_back(G);
        else if (ii.chain && cur_skip == 2) g_skip64.push_back(G);
        else g_main[cur_kd].push_back(G);
      };
      const int nkind = (vel && split) ? 2 : 1;
      const ActBuf& OB = P->act[s.out_act];
      // tile space = output voxels, except for the up-sampling conv (input voxels)
      bool is_up = false;
      for (const auto& p : s.parts) {
        const Layer& ly = ctx->layers[p.layer];
        const int nkc = static_cast<int>(p.src.size());
        const bool k16 = p.src[0].kc16;
        const int nk = k16 ? 1 : ((ii.fold && !ii.fin && p.type == T_SKIP1) ? 3 : nkind);
        const int C = ly.cout;
        const int tb = k16 ? p.tile_base16 : p.tile_base64;
        // OP(a, n8, b_row, d_col) -> device layout with byte offsets >> 4 precomputed
        auto OP = [&](int a, int n8, int b_row, int d_col) {
          MmaOp o;
          o.a_off = static_cast<uint16_t>(a);
          o.b_off = static_cast<uint16_t>(b_row * (k16 ? 32 : 128) / 16);
          o.d_col = static_cast<uint16_t>(d_col);
          o.n8 = static_cast<uint8_t>(n8);
          o.pad_ = 0;
          return o;
        };
        auto fill_ops = [&](GroupDesc& G, int kind, const Src& sc, int par, int kd) {
          const bool acc3 = ii.acc3 && !k16;
          const __half* ph = hi(sc.act);
          if (ii.f192 && !k16) {
            if (kind == 1) {      // lo phase: xl * Wh -> y1 (column 0; FOLD layout: column C)
              G.n_a = 1; G.a_map[0] = static_cast<int16_t>(get_map(lo(sc.act), sc.act, par)); G.a_map[1] = -1; G.n_ops = 1;
              G.ops[0] = OP(0, C / 8, 0, ii.fold ? C : 0);
            } else if (ii.fold && kind == 2) {     // folded skip: xh * dW_res -> dy (column 0)
              G.n_a = 1; G.a_map[0] = static_cast<int16_t>(get_map(ph, sc.act, par)); G.a_map[1] = -1; G.n_ops = 1;
              G.ops[0] = OP(0, C / 8, 0, 0);
            } else if (ii.fold) {   // xh * [..2C rows..] -> (ylo, y0) or (y1, ylo), picked per 3-tap block;  dx' * Wh -> dy
              G.n_a = 2; G.a_map[0] = static_cast<int16_t>(get_map(ph, sc.act, par));
              G.a_map[1] = static_cast<int16_t>(get_map(dx(sc.act), sc.act, par)); G.n_ops = 2;
              G.ops[0] = OP(0, 2 * C / 8, 0, 2 * C);
              G.ops[1] = OP(1, C / 8, C, 0);
            } else {              // xh * [..3C rows..] -> (dy, ylo, y0) or (y1, dy, ylo);  dx * Wh -> dy
              G.n_a = 2; G.a_map[0] = static_cast<int16_t>(get_map(ph, sc.act, par));
              G.a_map[1] = static_cast<int16_t>(get_map(dx(sc.act), sc.act, par)); G.n_ops = 2;
              G.ops[0] = OP(0, 3 * C / 8, 0, (kd == 1 && !ii.chain) ? 0 : C);      // chains: the kernel picks 0 / C per 3-tap block
              G.ops[1] = OP(1, C / 8, C + C / 2, C);
            }
            return;
          }
          if (ii.pair) {      // CTA-pair operand layout (see build_static): b_row is a row of the per-CTA stage
            G.a_map[0] = static_cast<int16_t>(get_map(ph, sc.act, par));
            if (k16) {
              G.n_a = 1; G.a_map[1] = -1; G.n_ops = 1;
              G.ops[0] = OP(0, 2 * C / 8, 0, 0);
            } else if (kind == 1) {
              G.n_a = 2; G.a_map[1] = static_cast<int16_t>(get_map(lo(sc.act), sc.act, par)); G.n_ops = 2;
              G.ops[0] = OP(0, C / 8, 0, 0);
              G.ops[1] = OP(1, C / 8, C / 2, 0);
            } else {
              G.n_a = 2; G.a_map[1] = static_cast<int16_t>(get_map(dx(sc.act), sc.act, par));
              if (ii.fold) {            // xh * Wh -> y_kd (columns 0 / 2C / 3C);  dx' * Wh -> dy (same weight rows)
                G.n_ops = 2;
                G.ops[0] = OP(0, C / 8, 0, kd <= 0 ? 0 : (kd == 1 ? 2 * C : 3 * C));
                G.ops[1] = OP(1, C / 8, 0, C);
              } else if (acc3 && kd == 1) {
                G.n_ops = 2;
                G.ops[0] = OP(0, 2 * C / 8, 0, C);
                G.ops[1] = OP(1, C / 8, C, C);
              } else if (acc3 && kd == 2) {
                G.n_ops = 3;
                G.ops[0] = OP(0, C / 8, 0, 3 * C);
                G.ops[1] = OP(0, C / 8, C / 2, C);
                G.ops[2] = OP(1, C / 8, 0, C);
              } else {
                G.n_ops = 2;
                G.ops[0] = OP(0, 2 * C / 8, 0, 0);
                G.ops[1] = OP(1, C / 8, C, C);
              }
            }
            return;
          }
          if (k16) {
            G.n_a = 1; G.a_map[0] = static_cast<int16_t>(get_map(ph, sc.act, par)); G.a_map[1] = -1;
            G.n_ops = 1;
            G.ops[0] = OP(0, (vel ? 2 * C : C) / 8, 0, 0);
            return;
          }
          const __half* pl = split ? lo(sc.act) : nullptr;
          const __half* pd = vel ? dx(sc.act) : nullptr;
          if (ii.fin && ii.fold) {
            if (kind == 0) {
              G.n_a = 2; G.a_map[0] = static_cast<int16_t>(get_map(ph, sc.act, par));
              G.a_map[1] = static_cast<int16_t>(get_map(pd, sc.act, par));
              G.n_ops = 2;
              G.ops[0] = OP(0, 4, 0, 0);
              G.ops[1] = OP(1, 2, 32, 0);
            } else {
              G.n_a = 1; G.a_map[0] = static_cast<int16_t>(get_map(pl, sc.act, par)); G.a_map[1] = -1;
              G.n_ops = 1;
              G.ops[0] = OP(0, 2, 0, 0);
            }
            return;
          }
          if (ii.fin) {
            if (vel) {
              G.n_a = 2; G.a_map[0] = static_cast<int16_t>(get_map(ph, sc.act, par));
              G.a_map[1] = static_cast<int16_t>(get_map(kind == 0 ? pd : pl, sc.act, par));
              G.n_ops = 2;
              G.ops[0] = OP(0, 2, 0, 0);
              G.ops[1] = OP(1, 2, 16, 0);
            } else if (split) {
              G.n_a = 2; G.a_map[0] = static_cast<int16_t>(get_map(ph, sc.act, par));
              G.a_map[1] = static_cast<int16_t>(get_map(pl, sc.act, par));
              G.n_ops = 2;
              G.ops[0] = OP(0, 2, 0, 0);
              G.ops[1] = OP(1, 2, 16, 0);
            } else {
              G.n_a = 1; G.a_map[0] = static_cast<int16_t>(get_map(ph, sc.act, par)); G.a_map[1] = -1;
              G.n_ops = 1;
              G.ops[0] = OP(0, 2, 0, 0);
            }
            return;
          }
          if (vel) {
            G.n_a = 2; G.a_map[0] = static_cast<int16_t>(get_map(ph, sc.act, par));
            if (kind == 0 && acc3 && kd == 1) {   // [dW | Wh]:  D[C:3C] += xh*[dW|Wh] -> (dy, y1);  dy += dx*Wh
              G.a_map[1] = static_cast<int16_t>(get_map(pd, sc.act, par));
              G.n_ops = 2;
              G.ops[0] = OP(0, 2 * C / 8, 0, C);
              G.ops[1] = OP(1, C / 8, C, C);
            } else if (kind == 0 && acc3 && kd == 2) {   // [Wh | dW]:  y2 += xh*Wh;  dy += xh*dW + dx*Wh
              G.a_map[1] = static_cast<int16_t>(get_map(pd, sc.act, par));
              G.n_ops = 3;
              G.ops[0] = OP(0, C / 8, 0, 3 * C);
              G.ops[1] = OP(0, C / 8, C, C);
              G.ops[2] = OP(1, C / 8, 0, C);
            } else if (kind == 0) {  // [Wh | dW]:  D[0:2C] = xh*[Wh|dW];  D[C:2C] += dx*Wh
              G.a_map[1] = static_cast<int16_t>(get_map(pd, sc.act, par));
              G.n_ops = 2;
              G.ops[0] = OP(0, 2 * C / 8, 0, 0);
              G.ops[1] = OP(1, C / 8, 0, C);
            } else {                 // [Wl | Wh]:  D[0:C] += xh*Wl + xl*Wh
              G.a_map[1] = static_cast<int16_t>(get_map(pl, sc.act, par));
              G.n_ops = 2;
              G.ops[0] = OP(0, C / 8, 0, 0);
              G.ops[1] = OP(1, C / 8, C, 0);
            }
          } else if (split) {        // [Wh | Wl]:  D = xh*Wh + xh*Wl + xl*Wh
            G.n_a = 2; G.a_map[0] = static_cast<int16_t>(get_map(ph, sc.act, par));
            G.a_map[1] = static_cast<int16_t>(get_map(pl, sc.act, par));
            G.n_ops = 3;
            G.ops[0] = OP(0, C / 8, 0, 0);
            G.ops[1] = OP(0, C / 8, C, 0);
            G.ops[2] = OP(1, C / 8, 0, 0);
          } else {
            G.n_a = 1; G.a_map[0] = static_cast<int16_t>(get_map(ph, sc.act, par)); G.a_map[1] = -1;
            G.n_ops = 1;
            G.ops[0] = OP(0, C / 8, 0, 0);
          }
        };
        auto mk = [&](const Src& sc, int par, int dw_, int dh_, int dd_, int ntaps, int tile0, int kind, int kd = -1) {
          GroupDesc G;
          memset(&G, 0, sizeof G);
          G.ntaps = static_cast<int8_t>(ntaps); G.kc16 = k16 ? 1 : 0; G.c0 = static_cast<int16_t>(sc.c0);
          G.dw = static_cast<int8_t>(sc.crop + p.off + dw_); G.dh = static_cast<int8_t>(sc.crop + p.off + dh_);
          G.dd = static_cast<int8_t>(sc.crop + p.off + dd_);
          G.brow0 = tile0 * ii.nrs; G.brow_step = static_cast<int16_t>(ii.nrs);
          G.tap_rows = static_cast<int16_t>(ii.pair ? ii.nrs / 2 : ii.nrs);
          G.pitch = static_cast<int8_t>(ntaps == 9 ? 10 : 8);
          G.tps = static_cast<int8_t>((k16 && ntaps >= 3) ? 3 : 1);
          if (!k16 && kind >= 1 && lo3) {                              // three lo taps per weight stage
            if (ntaps == 9) G.tps = 3;
            G.tap_rows = static_cast<int16_t>(s.cout / 2);
          }
          cur_box_w = G.pitch;
          cur_kind = k16 ? 0 : kind;
          cur_kd = (ii.acc3 && !k16 && kd > 0) ? kd : 0;
          cur_skip = p.type == T_SKIP1 ? (k16 ? 1 : 2) : 0;
          G.lo_stage = (cur_kind >= 1) ? 1 : 0;
          fill_ops(G, kind, sc, par, kd);
          if (G.a_map[0] < 0 || (G.n_a == 2 && G.a_map[1] < 0)) bad = true;
          push(G);
        };
        const double m = vel ? ((k16 || false) ? 2.0 : 3.0) : 1.0;
        double vout = static_cast<double>(OB.d) * OB.h * OB.w;
        if (p.type == T_CONV3) {
          const bool wide = ctx->wide && (!k16 || ctx->wide16);
          for (int kd = 0; kd < 3; ++kd)
            for (int q = 0; q < nkc; ++q) for (int kind = 0; kind < nk; ++kind) {
              const int t0 = tb + ((kd * nkc + q) * nk + kind) * 9;
              if (wide) mk(p.src[q], -1, 0, 0, kd, 9, t0, kind, kd);
              else for (int kw = 0; kw < 3; ++kw) mk(p.src[q], -1, kw, 0, kd, 3, t0 + kw * 3, kind, kd);
            }
          H.flops += 2.0 * ly.cout * ly.cin * 27 * vout * m;
        } else if (p.type == T_SKIP1) {
          for (int q = 0; q < nkc; ++q) for (int kind = 0; kind < nk; ++kind)
            mk(p.src[q], -1, 0, 0, 0, 1, tb + q * nk + kind, kind);
          H.flops += 2.0 * ly.cout * ly.cin * vout * m;
        } else if (p.type == T_DOWN) {
          for (int t = 0; t < 8; ++t) for (int q = 0; q < nkc; ++q) for (int kind = 0; kind < nk; ++kind)
            mk(p.src[q], t, 0, 0, 0, 1, tb + (t * nkc + q) * nk + kind, kind);
          H.flops += 2.0 * ly.cout * ly.cin * 8 * vout * m;
        } else {
          is_up = true;
          for (int q = 0; q < nkc; ++q) for (int kind = 0; kind < nk; ++kind)
            mk(p.src[q], -1, 0, 0, 0, 1, tb + q * nk + kind, kind);
          Lc.par_brow_step = nkc * nk * ii.nrs;
          H.flops += 2.0 * ly.cout * ly.cin * vout * m;
        }
      }
      Lc.n_chains = 0;
      if (ii.chain) {
        // accumulation chains (conv_mma.cuh, EARLY == 3).  Chain 0 = the lo phase (+ a 16-channel folded skip,
        // whose fixed operand layout writes (y1, dy)) into y1; then the 64-channel skip's one-tap blocks and the
        // 3-tap blocks of the conv groups, alternating y0 / y1 starting with y0.
        if (g_lo.empty() || g_main[0].empty() || g_main[1].empty() || g_main[2].empty()) bad = true;
        else {
          g_lo.front().pre_wait = 3;
          if (!g_skip16.empty()) { g_skip16.front().pre_wait = 2; g_skip16.back().post_sig = 2; }
          else g_lo.back().post_sig = 2;
          g_lo.insert(g_lo.end(), g_skip16.begin(), g_skip16.end());
          int blk = 0;
          std::vector<GroupDesc> mains;
          mains.insert(mains.end(), g_skip64.begin(), g_skip64.end());
          for (int kq = 0; kq < 3; ++kq) { mains.insert(mains.end(), g_main[kq].begin(), g_main[kq].end()); g_main[kq].clear(); }
          mains.front().pre_wait = 2;             // first MMA touching dy / ylo / y0 of this item
          for (auto& G : mains) {
            G.chain = 1; G.phase0 = static_cast<int8_t>(blk & 1);
            blk += (G.ntaps + 2) / 3;
          }
          mains.insert(mains.end(), g_dwres.begin(), g_dwres.end());   // FOLD: residual skip tangent, accumulates into dy only
          g_main[0] = mains;
          Lc.n_chains = 1 + blk;
        }
      } else if (ii.f192) {
        if (g_lo.empty() || g_main[0].empty() || g_main[1].empty() || g_main[2].empty()) bad = true;
        else {
          g_lo.front().pre_wait = 3;            // lo phase accumulates into y1 (drained during the previous item's kd 2)
          g_main[0].front().pre_wait = 2;       // first MMA touching (dy, ylo, y0)
          g_main[0].back().post_sig = 1;        // y0 complete
          g_main[1].back().post_sig = 2;        // y1 complete
          g_main[2].front().pre_wait = 1;       // kd 2 re-uses the y0 columns: drained and re-zeroed first
        }
      } else if (ii.early) {
        if (g_lo.empty() || g_main[0].empty() || g_main[1].empty() || g_main[2].empty()) bad = true;
        else {
          g_lo.front().pre_wait = 1;            // lo products accumulate into y0
          g_main[0].front().pre_wait = 2;       // first MMA touching dy
          g_main[0].back().post_sig = 1;        // y0 complete
          g_main[1].back().post_sig = 2;        // y1 complete
        }
      }
      for (const auto& G : g_lo) { if (ng < kMaxGroups) H.G.g[ng++] = G; else bad = true; }
      for (int kq = 0; kq < 3; ++kq)
        for (const auto& G : g_main[kq]) { if (ng < kMaxGroups) H.G.g[ng++] = G; else bad = true; }
      if (bad) { delete P; return fail(ctx, NBE_ERR_STATE, "launch %s: too many groups / tensor maps", s.name.c_str()); }
      H.G.n_groups = ng;
      Lc.cout = s.cout; Lc.vel = vel ? 1 : 0; Lc.act = 1;
      Lc.acc3 = ii.acc3 ? 1 : 0;
      Lc.bias = ctx->d_bias + s.bias_off;
      {
        const float* fold = ctx->d_fold + (static_cast<size_t>(wb) * nl + li) * 256;   // anext is stored with its offset applied
        Lc.beta = (vel && s.beta_layer >= 0) ? fold : nullptr;
        Lc.anext = (vel && s.anext_layer >= 0) ? fold + 128 : nullptr;
      }
      if (!ii.fin) {
        Lc.out_h_ptr = hi(s.out_act);
        Lc.out_l_ptr = split ? lo(s.out_act) : nullptr;
        Lc.out_d_ptr = vel ? dx(s.out_act) : nullptr;
      }
      const int64_t C = OB.c;
      if (is_up) {
        const ActBuf& IB = P->act[s.in_ref];
        Lc.n_par = 8; Lc.out_w = IB.w; Lc.out_h = IB.h; Lc.out_d = IB.d;
        Lc.out_sw = 2 * C; Lc.out_sh = 2 * C * OB.w; Lc.out_sd = 2 * C * OB.w * OB.h;
        Lc.par_ow = C; Lc.par_oh = C * OB.w; Lc.par_od = C * OB.w * OB.h;
      } else {
        Lc.n_par = 1; Lc.out_w = OB.w; Lc.out_h = OB.h; Lc.out_d = OB.d;
        Lc.out_sw = C; Lc.out_sh = C * OB.w; Lc.out_sd = C * OB.w * OB.h;
      }
      Lc.band_h = ctx->band_h;
      const long long tiles = 1ll * Lc.n_par * Lc.out_d * ((Lc.out_h + 16 * ii.tm - 1) / (16 * ii.tm)) * ((Lc.out_w + 7) / 8);
      H.grid = static_cast<int>(std::min<long long>(tiles, ctx->num_sms));
      if (ii.pair) H.grid = 2 * static_cast<int>(std::min<long long>((tiles + 1) / 2, ctx->num_sms / 2));
    }
  }
  ctx->plans.push_back(P);
  *out = P;
  return NBE_OK;
}

// The opt-in to > 48 KB of dynamic shared memory is per function AND per device: one process may
// drive several GPUs (nbe_process_box_multi), each from its own host thread.
template <class K>
cudaError_t opt_in_smem(K kern, int bytes, int device, std::atomic<uint64_t>& done) {
  const uint64_t bit = 1ull << (device & 63);
  if (done.load(std::memory_order_acquire) & bit) return cudaSuccess;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess) done.fetch_or(bit, std::memory_order_release);
  return e;
}

template <int NRS, int DC, int TM, bool FIN>
cudaError_t launch_inst(int device, const ConvLaunch& dl, const GroupTable& gt, const FinalArgs& fa, int grid, cudaStream_t st) {
  using Cfg = ConvCfg<NRS, DC, TM>;
  static std::atomic<uint64_t> done{0};
  auto kern = conv_mma_kernel<NRS, DC, TM, FIN>;
  cudaError_t e = opt_in_smem(kern, Cfg::kSmemBytes, device, done);
  if (e != cudaSuccess) return e;
  kern<<<grid, kConvThreads, Cfg::kSmemBytes, st>>>(dl, gt, fa);
  return cudaGetLastError();
}

template <int NRS, int DC, int TM, int EARLY = 0>
cudaError_t launch_pair(int device, const ConvLaunch& dl, const GroupTable& gt, const FinalArgs& fa, int grid, cudaStream_t st) {
  using Cfg = ConvCfg<NRS, DC, TM, true>;
  static std::atomic<uint64_t> done{0};
  auto kern = conv_mma_kernel<NRS, DC, TM, false, true, EARLY>;
  cudaError_t e = opt_in_smem(kern, Cfg::kSmemBytes, device, done);
  if (e != cudaSuccess) return e;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kConvThreads); cfg.dynamicSmemBytes = Cfg::kSmemBytes; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, dl, gt, fa);
}

cudaError_t launch_conv(int device, int inst, const ConvLaunch& dl, const GroupTable& gt, const FinalArgs& fa, int grid, cudaStream_t st) {
  switch (inst) {
    case I_128_128_2: return launch_inst<128, 128, 2, false>(device, dl, gt, fa, grid, st);
    case I_256_256_1: return launch_inst<256, 256, 1, false>(device, dl, gt, fa, grid, st);
    case I_FINAL: return launch_inst<32, 16, 2, true>(device, dl, gt, fa, grid, st);
    case I_128_64_2: return launch_inst<128, 64, 2, false>(device, dl, gt, fa, grid, st);
    case I_64_64_2: return launch_inst<64, 64, 2, false>(device, dl, gt, fa, grid, st);
    case I_256_128_2: return launch_inst<256, 128, 1, false>(device, dl, gt, fa, grid, st);
    case I_128_256_2: return launch_inst<128, 256, 2, false>(device, dl, gt, fa, grid, st);
    case I_256_512_1: return launch_inst<256, 512, 1, false>(device, dl, gt, fa, grid, st);
    case I_PAIR_128_256_2: return launch_pair<192, 256, 2>(device, dl, gt, fa, grid, st);
    case I_PAIR_256_512_1: return launch_pair<384, 512, 1>(device, dl, gt, fa, grid, st);
    case I_PAIR_128_128_2: return launch_pair<192, 128, 2>(device, dl, gt, fa, grid, st);
    case I_PAIR_256_256_1: return launch_pair<384, 256, 1>(device, dl, gt, fa, grid, st);
    case I_PAIR_128_256_1: return launch_pair<192, 256, 1>(device, dl, gt, fa, grid, st);
    case I_EARLY_128_256_2: return launch_pair<192, 256, 2, 1>(device, dl, gt, fa, grid, st);
    case I_EARLY_256_512_1: return launch_pair<384, 512, 1, 1>(device, dl, gt, fa, grid, st);
    case I_F192_2: return launch_pair<256, 256, 2, 2>(device, dl, gt, fa, grid, st);
    case I_CHAIN_2: return launch_pair<256, 256, 2, 3>(device, dl, gt, fa, grid, st);
    case I_FOLD_2: return launch_pair<192, 256, 2, 4>(device, dl, gt, fa, grid, st);
    case I_FOLD128_1: return launch_pair<256, 512, 1, 1>(device, dl, gt, fa, grid, st);
    case I_FINAL_FOLD: return launch_inst<48, 32, 2, true>(device, dl, gt, fa, grid, st);
  }
  return cudaErrorInvalidValue;
}

// one sample through the net: pack + 27 conv launches.  `pk` and `fa` describe where the
// padded input comes from and where the (n-96)^3 outputs go.
int run_sample(nbe_ctx* ctx, Plan* P, int sample, PackArgs pk, FinalArgs fa, cudaStream_t st) {
  pk.out = reinterpret_cast<__half*>(ctx->arena + P->act[A_IN16].off_hi);
  pk.n0 = P->dims[0]; pk.n1 = P->dims[1]; pk.n2 = P->dims[2];
  const long long nvox = 1ll * pk.n0 * pk.n1 * pk.n2;
  const int pgrid = static_cast<int>(std::min<long long>((nvox + 255) / 256, 16ll * ctx->num_sms));
  const int slots = P->n_launch + 1;
  const bool prof = ctx->profiling && ctx->prof_events.size() < 400000;
  if (prof && static_cast<int>(ctx->prof_names.size()) != slots) {
    ctx->prof_names.assign(1, "pack_input");
    ctx->prof_flops.assign(1, 0.0);
    for (int li = 0; li < P->n_launch; ++li) {
      ctx->prof_names.push_back(P->launches[li].name);
      ctx->prof_flops.push_back(P->launches[li].flops);
    }
    ctx->prof_sum_ms.assign(slots, 0.0);
    ctx->prof_samples = 0;
  }
  auto mark = [&]() {
    if (prof) { cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, st); ctx->prof_events.push_back(e); }
  };
  if (ctx->mod_pending) CK(cudaStreamWaitEvent(st, ctx->mod_done, 0));
  mark();
  pack_input_kernel<<<pgrid, 256, 0, st>>>(pk);
  CK(cudaGetLastError());
  ctx->launches++;
  mark();
  for (int li = 0; li < P->n_launch; ++li) {
    const HostLaunch& H = P->launches[static_cast<size_t>(sample) * P->n_launch + li];
    CK(launch_conv(ctx->device, H.inst, H.L, H.G, fa, H.grid, st));
    ctx->launches++;
    mark();
  }
  return NBE_OK;
}

// fold the recorded events into per-slot sums (synchronises the device)
void resolve_profile(nbe_ctx* ctx) {
  const int slots = static_cast<int>(ctx->prof_names.size());
  if (slots == 0 || ctx->prof_events.empty()) return;
  cudaDeviceSynchronize();
  const size_t per = static_cast<size_t>(slots) + 1;
  for (size_t s0 = 0; s0 + per <= ctx->prof_events.size(); s0 += per) {
    for (int i = 0; i < slots; ++i) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, ctx->prof_events[s0 + i], ctx->prof_events[s0 + i + 1]) == cudaSuccess)
        ctx->prof_sum_ms[i] += ms;
    }
    ctx->prof_samples++;
  }
  for (auto e : ctx->prof_events) cudaEventDestroy(e);
  ctx->prof_events.clear();
}

int ensure_ident(nbe_ctx* ctx, int n) {
  if (ctx->ident_cap >= n) return NBE_OK;
  if (ctx->d_ident) { CK(cudaDeviceSynchronize()); CK(cudaFree(ctx->d_ident)); ctx->d_ident = nullptr; }
  std::vector<int32_t> h(static_cast<size_t>(n));
  for (int i = 0; i < n; ++i) h[i] = i;
  CK(cudaMalloc(&ctx->d_ident, sizeof(int32_t) * n));
  CK(cudaMemcpy(ctx->d_ident, h.data(), sizeof(int32_t) * n, cudaMemcpyHostToDevice));
  ctx->ident_cap = n;
  return NBE_OK;
}

size_t dtype_size(int dt) { return dt == NBE_F32 ? 4 : 2; }

// Device-resident core of process_box: box_dev (3,size) in, disp_dev / vel_dev (3,size) out,
// asynchronous on `stream`.  After subbox s has been enqueued `after(s)` is called (used by the
// host wrapper to overlap the copy-back).
template <class After>
int process_box_core(nbe_ctx* ctx, const void* box_dev, int in_dtype, const int32_t size[3], const int32_t size_out[3],
                     const int32_t crop[3], const int32_t plen[3], const int32_t* crop_idx, const int32_t* add_idx0, int sub_first,
                     int sub_count, float Dz, float vel_fac, void* disp_dev, void* vel_dev, int out_dtype,
                     cudaStream_t st, After after, bool blocks = false, cudaEvent_t first_wait = nullptr) {
  const int64_t S0 = size[0], S1 = size[1], S2 = size[2];
  const int64_t O0 = size_out[0], O1 = size_out[1], O2 = size_out[2];
  int rc;
  const int per = plen[0] + plen[1] + plen[2];
  const size_t idx_bytes = static_cast<size_t>(sub_count) * per * sizeof(int32_t);
  if ((rc = ensure(ctx, reinterpret_cast<void**>(&ctx->d_idx), &ctx->idx_cap, idx_bytes))) return rc;
  CK(cudaMemcpyAsync(ctx->d_idx, crop_idx + static_cast<size_t>(sub_first) * per, idx_bytes, cudaMemcpyHostToDevice, st));
  Plan* P = nullptr;
  if ((rc = build_plan(ctx, plen, 1, &P))) return rc;
  if (first_wait) CK(cudaStreamWaitEvent(st, first_wait, 0));
  const size_t es = dtype_size(out_dtype);
  for (int s = 0; s < sub_count; ++s) {
    const int32_t* ai = add_idx0 + static_cast<size_t>(sub_first + s) * 3;
    const int32_t* di = ctx->d_idx + static_cast<size_t>(s) * per;
    PackArgs pk{};
    pk.src = box_dev; pk.src_dtype = in_dtype; pk.src_sc = S0 * S1 * S2; pk.src_sd = S1 * S2; pk.src_sh = S2;
    pk.idx_d = di; pk.idx_h = di + plen[0]; pk.idx_w = di + plen[0] + plen[1];
    pk.in_norm = Dz / 6.0f;
    FinalArgs fa{};
    fa.src = box_dev; fa.src_dtype = in_dtype; fa.src_sc = pk.src_sc; fa.src_sd = pk.src_sd; fa.src_sh = pk.src_sh;
    fa.idx_d = pk.idx_d + 48; fa.idx_h = pk.idx_h + 48; fa.idx_w = pk.idx_w + 48;
    if (blocks) {
      // block layout [sub_count][3][c0][c1][c2]: each subbox's output is one contiguous record (the
      // unit of the NCCL all-gather of a sharded box)
      const int64_t c0 = crop[0], c1 = crop[1], c2 = crop[2];
      const size_t base = static_cast<size_t>(s) * 3 * c0 * c1 * c2;
      fa.disp = static_cast<uint8_t*>(disp_dev) + base * es;
      fa.vel = ctx->vel ? static_cast<uint8_t*>(vel_dev) + base * es : nullptr;
      fa.o_sc = c0 * c1 * c2; fa.o_sd = c1 * c2; fa.o_sh = c2;
    } else {
      const size_t base = (static_cast<size_t>(ai[0]) * O1 + ai[1]) * O2 + ai[2];
      fa.disp = static_cast<uint8_t*>(disp_dev) + base * es;
      fa.vel = ctx->vel ? static_cast<uint8_t*>(vel_dev) + base * es : nullptr;
      fa.o_sc = O0 * O1 * O2; fa.o_sd = O1 * O2; fa.o_sh = O2;
    }
    fa.out_dtype = out_dtype; fa.mid_dtype = in_dtype;
    fa.in_norm = pk.in_norm; fa.six = 6.0f; fa.dx_norm = vel_fac * 6.0f; fa.x0_norm = vel_fac * 6.0f / Dz;
    if ((rc = run_sample(ctx, P, 0, pk, fa, st))) return rc;
    if ((rc = after(s))) return rc;
  }
  return NBE_OK;
}

int check_box_args(nbe_ctx* ctx, const void* in, const int32_t* size, const int32_t* crop, const int32_t* plen,
                          const int32_t* crop_idx, const int32_t* add_idx0, const void* disp, const void* vel,
                          int sub_count, int in_dtype, int out_dtype) {
  if (!ctx->have_params) return fail(ctx, NBE_ERR_STATE, "process_box before nbe_set_params");
  if (ctx->mod_batch != 1) return fail(ctx, NBE_ERR_STATE, "process_box needs weights modulated for one sample");
  if (!in || !size || !crop || !plen || !crop_idx || !add_idx0 || !disp || sub_count < 0)
    return fail(ctx, NBE_ERR_ARG, "null argument");
  if (ctx->vel && !vel) return fail(ctx, NBE_ERR_ARG, "velocity model: velocity output required");
  if (in_dtype < 0 || in_dtype > 2 || out_dtype < 0 || out_dtype > 2) return fail(ctx, NBE_ERR_ARG, "bad dtype");
  for (int d = 0; d < 3; ++d)
    if (plen[d] - crop[d] != 96) return fail(ctx, NBE_ERR_ARG, "padding must be 48 per side (the models hard-code the 48-voxel crop)");
  // validate the padded subbox shape before any copy is enqueued
  int tmp[A_COUNT];
  for (int d = 0; d < 3; ++d) { int rc = chain(ctx, plen[d], tmp); if (rc) return rc; }
  return NBE_OK;
}

// The gather tables and paste anchors index host and device memory: reject anything out of range
// before a copy or a launch is enqueued (O(n_sub * plen), negligible).
int check_tables(nbe_ctx* ctx, const int32_t* size, const int32_t* crop, const int32_t* plen, const int32_t* crop_idx,
                 const int32_t* add_idx0, int sub_first, int sub_count, int64_t out_S0) {
  if (sub_first < 0) return fail(ctx, NBE_ERR_ARG, "sub_first < 0");
  const int per = plen[0] + plen[1] + plen[2];
  for (int s = sub_first; s < sub_first + sub_count; ++s) {
    const int32_t* t = crop_idx + static_cast<size_t>(s) * per;
    for (int d = 0; d < 3; ++d) {
      for (int i = 0; i < plen[d]; ++i)
        if (t[i] < 0 || t[i] >= size[d]) return fail(ctx, NBE_ERR_ARG, "crop_idx of subbox %d, dim %d out of range: %d", s, d, t[i]);
      t += plen[d];
      const int32_t a = add_idx0[static_cast<size_t>(s) * 3 + d];
      if (a < 0 || a + crop[d] > (d == 0 ? out_S0 : size[d])) return fail(ctx, NBE_ERR_ARG, "add_idx0 of subbox %d, dim %d out of range: %d", s, d, a);
    }
  }
  return NBE_OK;
}


}  // namespace

// ========================================================================================
// C ABI
// ========================================================================================
extern "C" {

const char* nbe_version(void) { return "nbe_b200 0.1 (sm_100a, tcgen05/TMA)"; }

int nbe_create(nbe_ctx** out, int device) {
  if (!out) return NBE_ERR_ARG;
  *out = nullptr;
  nbe_ctx* ctx = new nbe_ctx();
  ctx->device = device;
  DevGuard dev_guard_(device);
  if (dev_guard_.err != cudaSuccess) { delete ctx; return NBE_ERR_CUDA; }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete ctx; return NBE_ERR_CUDA; }
  if (prop.major != 10) { delete ctx; return NBE_ERR_UNSUPPORTED; }
  ctx->num_sms = prop.multiProcessorCount;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) {
    delete ctx;
    return NBE_ERR_CUDA;
  }
  ctx->encode = reinterpret_cast<EncodeTiledFn>(fn);
  if (const char* e = getenv("NBE_WIDE")) ctx->wide = atoi(e) != 0;
  if (const char* e = getenv("NBE_PAIR")) ctx->pair = atoi(e) != 0;
  if (const char* e = getenv("NBE_DBUF")) ctx->dbuf = atoi(e) != 0;
  if (const char* e = getenv("NBE_EARLY")) ctx->early = atoi(e) != 0;
  if (const char* e = getenv("NBE_LOBOX")) ctx->lo_box = atoi(e) != 0;
  if (const char* e = getenv("NBE_F192")) ctx->f192 = atoi(e) != 0;
  if (const char* e = getenv("NBE_CHAIN")) ctx->chain = atoi(e) != 0;
  if (const char* e = getenv("NBE_FOLD")) ctx->fold = atoi(e) != 0;
  if (const char* e = getenv("NBE_FOLD128")) ctx->fold128 = atoi(e) != 0;
  if (const char* e = getenv("NBE_FOLD_FINAL")) ctx->fold_final = atoi(e) != 0;
  if (const char* e = getenv("NBE_FOLD_AMAX")) ctx->fold_amax = static_cast<float>(atof(e));
  if (const char* e = getenv("NBE_WWIN")) ctx->w_window = atoi(e) != 0;
  if (const char* e = getenv("NBE_TRACE")) ctx->trace = atoi(e) != 0;
  if (const char* e = getenv("NBE_WIDE16")) ctx->wide16 = atoi(e) != 0;
  if (const char* e = getenv("NBE_BAND")) ctx->band_h = atoi(e);
  cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking);
  cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking);
  cudaStreamCreateWithFlags(&ctx->up_stream, cudaStreamNonBlocking);
  cudaEventCreateWithFlags(&ctx->mod_done, cudaEventDisableTiming);
  *out = ctx;
  return NBE_OK;
}

void nbe_destroy(nbe_ctx* ctx) {
  if (!ctx) return;
  DevGuard dev_guard_(ctx->device);
  cudaDeviceSynchronize();
  for (auto* p : ctx->plans) { delete p; }
  for (auto& l : ctx->layers) { cudaFree(l.W); cudaFree(l.dW); cudaFree(l.SW); cudaFree(l.sb); cudaFree(l.pre_a); cudaFree(l.pre_beta); }
  cudaFree(ctx->d_fold);
  cudaFree(ctx->d_metas); cudaFree(ctx->d_bias); cudaFree(ctx->d_packed); cudaFree(ctx->d_w32); cudaFree(ctx->d_dw32);
  cudaFree(ctx->d_s0); cudaFree(ctx->d_s1); cudaFree(ctx->arena); cudaFree(ctx->d_ident); cudaFree(ctx->d_box);
  cudaFree(ctx->d_disp); cudaFree(ctx->d_velo); cudaFree(ctx->d_idx);
  if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  if (ctx->up_stream) cudaStreamDestroy(ctx->up_stream);
  if (ctx->mod_done) cudaEventDestroy(ctx->mod_done);
  delete ctx;
}

const char* nbe_last_error(const nbe_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int nbe_set_precision(nbe_ctx* ctx, int precision) {
  if (!ctx) return NBE_ERR_ARG;
  if (precision != NBE_PREC_SPLIT && precision != NBE_PREC_FP16) return fail(ctx, NBE_ERR_ARG, "unknown precision %d", precision);
  ENTER_DEVICE(ctx);
  if (precision == ctx->precision) return NBE_OK;
  ctx->precision = precision;
  ctx->fold_active = fold_static_ok(ctx);
  if (ctx->have_params) { CK(cudaDeviceSynchronize()); return build_static(ctx); }
  return NBE_OK;
}

int nbe_set_params(nbe_ctx* ctx, const nbe_layer_params* layers, int n_layers, int premodulated, int compute_vel,
                   float eps) {
  if (!ctx || !layers) return NBE_ERR_ARG;
  ENTER_DEVICE(ctx);
  CK(cudaDeviceSynchronize());
  if (n_layers != 33) return fail(ctx, NBE_ERR_ARG, "expected 33 conv layers, got %d", n_layers);
  for (auto& l : ctx->layers) { cudaFree(l.W); cudaFree(l.dW); cudaFree(l.SW); cudaFree(l.sb); cudaFree(l.pre_a); cudaFree(l.pre_beta); }
  ctx->layers.clear(); ctx->lidx.clear();
  ctx->premod = premodulated != 0; ctx->vel = compute_vel != 0; ctx->eps = eps; ctx->have_params = false;
  for (int i = 0; i < n_layers; ++i) {
    const nbe_layer_params& p = layers[i];
    if (!p.block || !p.layer || !p.weight || !p.bias) return fail(ctx, NBE_ERR_ARG, "layer %d: null field", i);
    if (!ctx->premod && (!p.style_weight || !p.style_bias)) return fail(ctx, NBE_ERR_ARG, "layer %s/%s: style_weight/style_bias required", p.block, p.layer);
    if (ctx->premod && ctx->vel && !p.dweight) return fail(ctx, NBE_ERR_ARG, "layer %s/%s: dweight required", p.block, p.layer);
    if (!((p.cin == 3 || p.cin == 64 || p.cin == 128) && (p.cout == 3 || p.cout == 64 || p.cout == 128) && p.k >= 1 && p.k <= 3))
      return fail(ctx, NBE_ERR_UNSUPPORTED, "layer %s/%s: unsupported shape cout=%d cin=%d k=%d (mid_chan must be 64)", p.block, p.layer, p.cout, p.cin, p.k);
    Layer L;
    L.block = p.block; L.layer = p.layer; L.cout = p.cout; L.cin = p.cin; L.k = p.k;
    const size_t nw = static_cast<size_t>(p.cout) * p.cin * p.k * p.k * p.k;
    CK(cudaMalloc(&L.W, nw * 4)); CK(cudaMemcpy(L.W, p.weight, nw * 4, cudaMemcpyHostToDevice));
    if (p.dweight && ctx->premod) { CK(cudaMalloc(&L.dW, nw * 4)); CK(cudaMemcpy(L.dW, p.dweight, nw * 4, cudaMemcpyHostToDevice)); }
    if (!ctx->premod) {
      CK(cudaMalloc(&L.SW, p.cin * 2 * 4)); CK(cudaMemcpy(L.SW, p.style_weight, p.cin * 2 * 4, cudaMemcpyHostToDevice));
      CK(cudaMalloc(&L.sb, p.cin * 4)); CK(cudaMemcpy(L.sb, p.style_bias, p.cin * 4, cudaMemcpyHostToDevice));
      L.hSW.assign(p.style_weight, p.style_weight + 2 * p.cin);
      L.hsb.assign(p.style_bias, p.style_bias + p.cin);
    } else if (ctx->vel && p.dweight && fold_candidate(L)) {
      std::vector<float> a, beta;
      factor_premod(L, p.weight, p.dweight, ctx->fold_amax, a, beta);
      if (L.pre_ok) {
        CK(cudaMalloc(&L.pre_a, a.size() * 4)); CK(cudaMemcpy(L.pre_a, a.data(), a.size() * 4, cudaMemcpyHostToDevice));
        CK(cudaMalloc(&L.pre_beta, beta.size() * 4)); CK(cudaMemcpy(L.pre_beta, beta.data(), beta.size() * 4, cudaMemcpyHostToDevice));
      }
    }
    L.bias.assign(p.bias, p.bias + p.cout);
    ctx->lidx[L.block + "/" + L.layer] = static_cast<int>(ctx->layers.size());
    ctx->layers.push_back(L);
  }
  ctx->fold_active = fold_static_ok(ctx);
  int rc = build_static(ctx);
  if (rc) return rc;
  ctx->have_params = true;
  return NBE_OK;
}

int nbe_modulate(nbe_ctx* ctx, const float* Om, const float* Dz, int batch, void* stream) {
  if (!ctx) return NBE_ERR_ARG;
  if (!ctx->have_params) return fail(ctx, NBE_ERR_STATE, "nbe_modulate before nbe_set_params");
  if (batch < 1 || !Dz) return fail(ctx, NBE_ERR_ARG, "batch >= 1 and Dz required");
  if (!ctx->premod && !Om) return fail(ctx, NBE_ERR_ARG, "Om required for style models");
  ENTER_DEVICE(ctx);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // style vector in fp32 as in the reference (style_nbody_emulator_vel_core.py:126-128)
  std::vector<float> s0(batch), s1(batch);
  for (int b = 0; b < batch; ++b) {
    s0[b] = Om ? (Om[b] - 0.3f) * 5.0f : 0.f;
    s1[b] = Dz[b] - 1.0f;
  }
  {
    // tangent folding divides by the modulation m_i: a cosmology that drives some m_i towards zero falls back to
    // the 5-product kernels (different weight layout: rebuild the static description)
    const bool want = fold_static_ok(ctx) && fold_range_ok(ctx, s0, s1);
    if (want != ctx->fold_active) {
      CK(cudaDeviceSynchronize());
      ctx->fold_active = want;
      int rcs = build_static(ctx);
      if (rcs) return rcs;
    }
  }
  const size_t need_p = static_cast<size_t>(batch) * ctx->packed_halves * 2;
  const size_t need_f = static_cast<size_t>(batch) * ctx->sl.size() * 256 * 4;
  const size_t need_w = static_cast<size_t>(batch) * ctx->w32_floats * 4;
  const bool moved = ctx->packed_cap < need_p || ctx->fold_cap < need_f;
  int rc;
  if ((rc = ensure(ctx, reinterpret_cast<void**>(&ctx->d_fold), &ctx->fold_cap, need_f))) return rc;
  if ((rc = ensure(ctx, reinterpret_cast<void**>(&ctx->d_packed), &ctx->packed_cap, need_p))) return rc;
  if ((rc = ensure(ctx, reinterpret_cast<void**>(&ctx->d_w32), &ctx->w32_cap, need_w))) return rc;
  if (ctx->vel) { if ((rc = ensure(ctx, reinterpret_cast<void**>(&ctx->d_dw32), &ctx->dw32_cap, need_w))) return rc; }
  if (ctx->s_cap < batch) {
    if (ctx->d_s0) { CK(cudaDeviceSynchronize()); cudaFree(ctx->d_s0); cudaFree(ctx->d_s1); }
    CK(cudaMalloc(&ctx->d_s0, batch * 4)); CK(cudaMalloc(&ctx->d_s1, batch * 4)); ctx->s_cap = batch;
  }
  if (moved || batch != ctx->mod_batch) {
    // packed buffer moved or sample count changed: tensor maps are stale
    CK(cudaDeviceSynchronize());
    for (auto* p : ctx->plans) { delete p; }
    ctx->plans.clear();
  }
  CK(cudaMemsetAsync(ctx->d_fold, 0, need_f, st));
  CK(cudaMemcpyAsync(ctx->d_s0, s0.data(), batch * 4, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(ctx->d_s1, s1.data(), batch * 4, cudaMemcpyHostToDevice, st));
  CK(cudaMemsetAsync(ctx->d_packed, 0, need_p, st));
  // finish LayerMeta
  std::vector<LayerMeta> hm = ctx->metas;
  for (auto& s : ctx->sl)
    for (auto& p : s.parts) {
      LayerMeta& M = hm[p.layer];
      const Layer& ly = ctx->layers[p.layer];
      M.W = ly.W; M.dW = ly.dW; M.SW = ly.SW; M.sb = ly.sb;
      M.w32 = ctx->d_w32 + ly.w32_off * batch;
      M.dw32 = ctx->vel ? ctx->d_dw32 + ly.w32_off * batch : nullptr;
      M.dst = ctx->d_packed + (M.kc16 ? s.b16_off : s.b64_off);
      M.dst_sample_stride = ctx->packed_halves;
      M.cout = ly.cout; M.cin = ly.cin; M.k3 = ly.k * ly.k * ly.k;
      M.first = (ly.block == "conv_l00" && (ly.layer == "conv_0" || ly.layer == "skip")) ? 1 : 0;
      M.premod = ctx->premod ? 1 : 0; M.vel = ctx->vel ? 1 : 0;
      M.fold_SW = nullptr; M.fold_sb = nullptr; M.fold_a = nullptr; M.beta_out = nullptr;
      M.pre_a = ly.pre_a; M.pre_beta = ly.pre_beta;
      M.beta_layer = p.beta_layer;
      M.fold_stride = static_cast<int>(ctx->sl.size()) * 256;
      M.n_a_out = 0;
      if (p.fold_layer >= 0) {
        const Layer& fl = ctx->layers[p.fold_layer];
        if (ctx->premod) M.fold_a = fl.pre_a + p.fold_off;
        else { M.fold_SW = fl.SW + 2 * p.fold_off; M.fold_sb = fl.sb + p.fold_off; }
      }
      const size_t li = static_cast<size_t>(&s - ctx->sl.data());
      if (s.beta_layer == p.layer) {
        M.beta_out = ctx->d_fold + li * 256;
        // its a is the anext of the launch(es) producing the tensor(s) it reads
        for (size_t lj = 0; lj < ctx->sl.size(); ++lj)
          if (ctx->sl[lj].anext_layer == p.layer && M.n_a_out < 2) {
            M.a_out[M.n_a_out] = ctx->d_fold + lj * 256 + 128;
            M.a_out_off[M.n_a_out] = ctx->sl[lj].anext_off;
            M.a_out_n[M.n_a_out] = act_channels(ctx->sl[lj].out_act);
            ++M.n_a_out;
          }
      }
    }
  if (!ctx->d_metas) CK(cudaMalloc(&ctx->d_metas, hm.size() * sizeof(LayerMeta)));
  CK(cudaMemcpyAsync(ctx->d_metas, hm.data(), hm.size() * sizeof(LayerMeta), cudaMemcpyHostToDevice, st));
  int rows = 0;
  for (auto& l : ctx->layers) rows += l.cout;
  CK(cudaStreamSynchronize(st));     // hm / s0 / s1 are stack-owned
  modulate_kernel<<<dim3(rows, batch), 128, 0, st>>>(ctx->d_metas, static_cast<int>(hm.size()), ctx->d_s0, ctx->d_s1, ctx->eps);
  CK(cudaGetLastError());
  // the packed weights are consumed on other streams (the caller's in nbe_forward, the context's own
  // in nbe_process_box): they wait on this event instead of relying on stream coincidence
  CK(cudaEventRecord(ctx->mod_done, st));
  ctx->mod_pending = true;
  ctx->launches++;
  ctx->mod_batch = batch;
  return NBE_OK;
}

int nbe_get_modulated(nbe_ctx* ctx, int layer_index, int sample, float* w_host, float* dw_host) {
  if (!ctx || !w_host) return NBE_ERR_ARG;
  if (ctx->mod_batch < 1) return fail(ctx, NBE_ERR_STATE, "nbe_get_modulated before nbe_modulate");
  if (layer_index < 0 || layer_index >= static_cast<int>(ctx->layers.size()) || sample < 0 || sample >= ctx->mod_batch)
    return fail(ctx, NBE_ERR_ARG, "layer/sample out of range");
  ENTER_DEVICE(ctx);
  CK(cudaDeviceSynchronize());
  const Layer& ly = ctx->layers[layer_index];
  const size_t n = static_cast<size_t>(ly.cout) * ly.cin * ly.k * ly.k * ly.k;
  const size_t off = static_cast<size_t>(ly.w32_off) * ctx->mod_batch + static_cast<size_t>(sample) * n;
  CK(cudaMemcpy(w_host, ctx->d_w32 + off, n * 4, cudaMemcpyDeviceToHost));
  if (dw_host) {
    if (!ctx->vel) return fail(ctx, NBE_ERR_STATE, "no dweight: parameters were set with compute_vel = 0");
    CK(cudaMemcpy(dw_host, ctx->d_dw32 + off, n * 4, cudaMemcpyDeviceToHost));
  }
  return NBE_OK;
}

size_t nbe_workspace_bytes(nbe_ctx* ctx, const int32_t dims[3]) {
  if (!ctx || !dims) return 0;
  int cd[A_COUNT], ch[A_COUNT], cw[A_COUNT];
  if (chain(ctx, dims[0], cd) || chain(ctx, dims[1], ch) || chain(ctx, dims[2], cw)) return 0;
  const bool split = ctx->precision == NBE_PREC_SPLIT;
  size_t off = 0;
  for (int a = 0; a < A_OUT; ++a) {
    const size_t bytes = align_up(static_cast<size_t>(act_channels(a)) * cd[a] * ch[a] * cw[a] * 2, 1024);
    off += bytes;
    if (a != A_IN16) off += bytes * ((split ? 1 : 0) + (ctx->vel ? 1 : 0));
  }
  return off;
}

int nbe_forward(nbe_ctx* ctx, const void* x_dev, int in_dtype, int batch, const int32_t dims[3], const float* Dz,
                const float* vel_fac, void* disp_dev, void* vel_dev, int out_dtype, void* stream) {
  if (!ctx) return NBE_ERR_ARG;
  if (!ctx->have_params) return fail(ctx, NBE_ERR_STATE, "nbe_forward before nbe_set_params");
  if (ctx->mod_batch < 1) return fail(ctx, NBE_ERR_STATE, "nbe_forward before nbe_modulate");
  if (!x_dev || !dims || !Dz || !disp_dev || batch < 1) return fail(ctx, NBE_ERR_ARG, "null argument");
  if (ctx->vel && (!vel_dev || !vel_fac)) return fail(ctx, NBE_ERR_ARG, "velocity model: vel_dev and vel_fac required");
  if (!ctx->vel && vel_dev) return fail(ctx, NBE_ERR_ARG, "displacement-only model: vel_dev must be NULL");
  if (ctx->mod_batch != 1 && ctx->mod_batch != batch) return fail(ctx, NBE_ERR_STATE, "weights modulated for %d samples, batch is %d", ctx->mod_batch, batch);
  if (in_dtype < 0 || in_dtype > 2 || out_dtype < 0 || out_dtype > 2) return fail(ctx, NBE_ERR_ARG, "bad dtype");
  ENTER_DEVICE(ctx);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Plan* P = nullptr;
  int rc = build_plan(ctx, dims, batch, &P);
  if (rc) return rc;
  const int nmax = std::max(dims[0], std::max(dims[1], dims[2]));
  if ((rc = ensure_ident(ctx, nmax))) return rc;
  const int64_t n0 = dims[0], n1 = dims[1], n2 = dims[2];
  const int64_t o0 = n0 - 96, o1 = n1 - 96, o2 = n2 - 96;
  for (int b = 0; b < batch; ++b) {
    PackArgs pk{};
    pk.src = static_cast<const uint8_t*>(x_dev) + static_cast<size_t>(b) * 3 * n0 * n1 * n2 * dtype_size(in_dtype);
    pk.src_dtype = in_dtype; pk.src_sc = n0 * n1 * n2; pk.src_sd = n1 * n2; pk.src_sh = n2;
    pk.idx_d = ctx->d_ident; pk.idx_h = ctx->d_ident; pk.idx_w = ctx->d_ident;
    pk.in_norm = Dz[b] / 6.0f;
    FinalArgs fa{};
    fa.src = pk.src; fa.src_dtype = in_dtype; fa.src_sc = pk.src_sc; fa.src_sd = pk.src_sd; fa.src_sh = pk.src_sh;
    fa.idx_d = ctx->d_ident + 48; fa.idx_h = ctx->d_ident + 48; fa.idx_w = ctx->d_ident + 48;
    fa.disp = static_cast<uint8_t*>(disp_dev) + static_cast<size_t>(b) * 3 * o0 * o1 * o2 * dtype_size(out_dtype);
    fa.vel = vel_dev ? static_cast<uint8_t*>(vel_dev) + static_cast<size_t>(b) * 3 * o0 * o1 * o2 * dtype_size(out_dtype) : nullptr;
    fa.out_dtype = out_dtype; fa.mid_dtype = in_dtype; fa.o_sc = o0 * o1 * o2; fa.o_sd = o1 * o2; fa.o_sh = o2;
    fa.in_norm = pk.in_norm; fa.six = 6.0f;
    fa.dx_norm = vel_fac ? vel_fac[b] * 6.0f : 0.f;
    fa.x0_norm = vel_fac ? vel_fac[b] * 6.0f / Dz[b] : 0.f;
    if ((rc = run_sample(ctx, P, b, pk, fa, st))) return rc;
  }
  return NBE_OK;
}

int nbe_process_box_dev(nbe_ctx* ctx, const void* box_dev, int in_dtype, const int32_t size[3], const int32_t crop[3],
                        const int32_t plen[3], const int32_t* crop_idx, const int32_t* add_idx0, int sub_first,
                        int sub_count, float Dz, float vel_fac, void* disp_dev, void* vel_dev, int out_dtype,
                        void* stream) {
  if (!ctx) return NBE_ERR_ARG;
  int rc = check_box_args(ctx, box_dev, size, crop, plen, crop_idx, add_idx0, disp_dev, vel_dev, sub_count, in_dtype, out_dtype);
  if (rc) return rc;
  ENTER_DEVICE(ctx);
  return process_box_core(ctx, box_dev, in_dtype, size, size, crop, plen, crop_idx, add_idx0, sub_first, sub_count, Dz,
                          vel_fac, disp_dev, vel_dev, out_dtype, static_cast<cudaStream_t>(stream),
                          [](int) { return NBE_OK; });
}

// SubboxProcessor.process_box on host buffers for one context (see nbe_process_box in nbe.h).  With
// disp_blk / vel_blk set the results stay on the device in block layout instead of going back to the host.
static int process_box_host(nbe_ctx* ctx, const void* in_host, int in_dtype, const int32_t size[3], const int32_t crop[3],
                            const int32_t plen[3], const int32_t* crop_idx, const int32_t* add_idx0, int sub_first,
                            int sub_count, float Dz, float vel_fac, void* disp_host, void* vel_host, int out_dtype,
                            void* disp_blk, void* vel_blk, int64_t out_S0 = 0) {
  const bool blocks = disp_blk != nullptr;
  if (out_S0 <= 0) out_S0 = size[0];          // D extent of the host output arrays (a streamed D-slab holds fewer planes)
  int rc = check_box_args(ctx, in_host, size, crop, plen, crop_idx, add_idx0, blocks ? disp_blk : disp_host,
                          blocks ? vel_blk : vel_host, sub_count, in_dtype, out_dtype);
  if (rc) return rc;
  if (sub_count == 0) return NBE_OK;
  if ((rc = check_tables(ctx, size, crop, plen, crop_idx, add_idx0, sub_first, sub_count, out_S0))) return rc;
  ENTER_DEVICE(ctx);
  cudaStream_t st = ctx->own_stream, cs = ctx->copy_stream, us = ctx->up_stream;
  const int64_t S0 = size[0], S1 = size[1], S2 = size[2];
  const int per = plen[0] + plen[1] + plen[2];
  const size_t ies = dtype_size(in_dtype), es = dtype_size(out_dtype);
  const size_t plane_out = static_cast<size_t>(S1) * S2 * es;

  // ---- the device holds only the (D, H) window this range of subboxes reads: the D-planes and
  // H-rows that occur in its gather tables (owned blocks + 48-voxel halo, periodic), compacted into
  // slots in increasing index order; the tables are remapped index -> slot.  On 8 ranks of the 512^3
  // box that is 224 x 352 of 512 x 512 rows (30 % of the box); it is also what lets a box larger than
  // one GPU's HBM be processed when sharded (BASELINE config 5).
  std::vector<int32_t> slotD(static_cast<size_t>(S0), -1), slotH(static_cast<size_t>(S1), -1);
  std::vector<int32_t> tabs(crop_idx + static_cast<size_t>(sub_first) * per,
                            crop_idx + static_cast<size_t>(sub_first + sub_count) * per);
  for (int s = 0; s < sub_count; ++s) {
    const int32_t* t = &tabs[static_cast<size_t>(s) * per];
    for (int i = 0; i < plen[0]; ++i) slotD[t[i]] = 0;
    for (int i = 0; i < plen[1]; ++i) slotH[t[plen[0] + i]] = 0;
  }
  std::vector<int32_t> srcD, srcH;            // slot -> index
  for (int64_t d = 0; d < S0; ++d) if (slotD[d] == 0) { slotD[d] = static_cast<int32_t>(srcD.size()); srcD.push_back(static_cast<int32_t>(d)); }
  for (int64_t h = 0; h < S1; ++h) if (slotH[h] == 0) { slotH[h] = static_cast<int32_t>(srcH.size()); srcH.push_back(static_cast<int32_t>(h)); }
  const int64_t nD = static_cast<int64_t>(srcD.size()), nH = static_cast<int64_t>(srcH.size());
  const std::vector<int32_t> tabs_src = tabs;                 // un-remapped copy: upload units compare these
  for (int s = 0; s < sub_count; ++s) {
    int32_t* t = &tabs[static_cast<size_t>(s) * per];
    for (int i = 0; i < plen[0]; ++i) t[i] = slotD[t[i]];
    for (int i = 0; i < plen[1]; ++i) t[plen[0] + i] = slotH[t[plen[0] + i]];
  }
  // output: the D-range [dlo, dhi) spanned by the owned blocks
  int64_t dlo = S0, dhi = 0;
  for (int s = 0; s < sub_count; ++s) {
    const int32_t a0 = add_idx0[static_cast<size_t>(sub_first + s) * 3];
    dlo = std::min<int64_t>(dlo, a0); dhi = std::max<int64_t>(dhi, a0 + crop[0]);
  }
  const int64_t ND = dhi - dlo;
  std::vector<int32_t> anchors(add_idx0 + static_cast<size_t>(sub_first) * 3, add_idx0 + static_cast<size_t>(sub_first + sub_count) * 3);
  for (int s = 0; s < sub_count; ++s) anchors[static_cast<size_t>(s) * 3] -= static_cast<int32_t>(dlo);

  const size_t row_in = static_cast<size_t>(S2) * ies;
  const size_t in_bytes = static_cast<size_t>(3) * nD * nH * row_in;
  const size_t out_bytes = static_cast<size_t>(3) * ND * plane_out;
  if ((rc = ensure(ctx, &ctx->d_box, &ctx->box_cap, in_bytes))) return rc;
  if (!blocks) {
    if ((rc = ensure(ctx, &ctx->d_disp, &ctx->out_cap, out_bytes))) return rc;
    if (ctx->vel && (rc = ensure(ctx, &ctx->d_velo, &ctx->velo_cap, out_bytes))) return rc;
  }
  // every allocation of this call happens before the first copy is enqueued: the plan (arena, tensor
  // maps) and the index tables are sized here, not between uploads
  {
    Plan* P = nullptr;
    if ((rc = build_plan(ctx, plen, 1, &P))) return rc;
    if ((rc = ensure(ctx, reinterpret_cast<void**>(&ctx->d_idx), &ctx->idx_cap, static_cast<size_t>(sub_count) * per * sizeof(int32_t)))) return rc;
  }

  // ---- upload, in the order the subboxes need it, on its own stream.  Consecutive subboxes with the
  // same D and H tables form a unit; a unit uploads the (plane, row) rectangles no earlier unit has
  // brought in -- one strided 2-D copy per channel and rectangle -- and records an event the compute
  // stream waits on before the unit's first subbox.  Only the first unit's window (224 x 224 rows of
  // the 512^3 box: 5 ms) is exposed; the rest arrives behind the running subboxes.
  std::vector<cudaEvent_t> up_ev(static_cast<size_t>(sub_count), nullptr);
  auto destroy_events = [&]() { for (auto e : up_ev) if (e) cudaEventDestroy(e); };
  {
    // W is windowed as well, in blocks of WB columns (<= 64 blocks, one bit each): the first subbox then
    // waits for 224 x 224 x 256 instead of 224 x 224 x 512 voxels per channel.  The device rows keep the
    // full width (the W tables are not remapped), only the transfers are narrower.
    const int64_t WB = ctx->w_window ? std::max<int64_t>(32, (S2 + 63) / 64) : S2;
    const int nWB = static_cast<int>((S2 + WB - 1) / WB);
    const uint64_t all_w = nWB >= 64 ? ~0ull : ((1ull << nWB) - 1ull);
    std::vector<uint64_t> have(static_cast<size_t>(nD * nH), 0ull);
    struct Run { int32_t h0, n; uint64_t mask; bool operator==(const Run& o) const { return h0 == o.h0 && n == o.n && mask == o.mask; } };
    typedef std::vector<Run> Runs;
    // Only the FIRST subbox of the call is W-windowed (it is the one transfer nothing hides); every later
    // unit brings whole rows, or the rest of a partly present row, in large strided copies.
    const size_t cmp_n = static_cast<size_t>(plen[0] + plen[1]);
    cudaEvent_t tr0 = nullptr;
    if (ctx->trace) { cudaEventCreate(&tr0); cudaEventRecord(tr0, us); }
    for (int u0 = 0; u0 < sub_count;) {
      int u1 = u0;
      const bool first_win = ctx->w_window && u0 == 0;
      while (!first_win && u1 + 1 < sub_count &&
             memcmp(&tabs_src[static_cast<size_t>(u1 + 1) * per], &tabs_src[static_cast<size_t>(u0) * per], sizeof(int32_t) * cmp_n) == 0) ++u1;
      const int32_t* t = &tabs[static_cast<size_t>(u0) * per];
      std::vector<int32_t> ud(t, t + plen[0]), uh(t + plen[0], t + plen[0] + plen[1]);
      std::sort(ud.begin(), ud.end()); ud.erase(std::unique(ud.begin(), ud.end()), ud.end());
      std::sort(uh.begin(), uh.end()); uh.erase(std::unique(uh.begin(), uh.end()), uh.end());
      uint64_t need_w = 0;
      if (first_win) for (int i = 0; i < plen[2]; ++i) need_w |= 1ull << (t[plen[0] + plen[1] + i] / WB);
      else need_w = all_w;
      auto missing = [&](int32_t d) {
        Runs r;
        for (size_t i = 0; i < uh.size();) {
          const uint64_t m = need_w & ~have[static_cast<size_t>(d) * nH + uh[i]];
          if (!m) { ++i; continue; }
          size_t j = i;
          while (j + 1 < uh.size() && uh[j + 1] == uh[j] + 1 && srcH[uh[j + 1]] == srcH[uh[j]] + 1 &&
                 (need_w & ~have[static_cast<size_t>(d) * nH + uh[j + 1]]) == m) ++j;
          r.push_back(Run{uh[i], static_cast<int32_t>(j - i + 1), m});
          i = j + 1;
        }
        return r;
      };
      bool any = false;
      for (size_t i = 0; i < ud.size();) {
        const Runs r0 = missing(ud[i]);
        size_t j = i;
        while (j + 1 < ud.size() && ud[j + 1] == ud[j] + 1 && srcD[ud[j + 1]] == srcD[ud[j]] + 1 && missing(ud[j + 1]) == r0) ++j;
        const size_t n_planes = j - i + 1;
        for (const auto& run : r0) {
          for (int b0 = 0; b0 < nWB;) {                     // maximal runs of missing W blocks
            if (!((run.mask >> b0) & 1ull)) { ++b0; continue; }
            int b1 = b0;
            while (b1 + 1 < nWB && ((run.mask >> (b1 + 1)) & 1ull)) ++b1;
            const int64_t wlo = b0 * WB, whi = std::min<int64_t>(S2, (b1 + 1) * WB);
            const size_t w_bytes = static_cast<size_t>(whi - wlo) * ies;
            // own bounds check of every box (compute-sanitizer is not available on this pool)
            if (ud[i] + static_cast<int64_t>(n_planes) > nD || run.h0 + run.n > nH || whi > S2 ||
                srcD[ud[i]] + static_cast<int64_t>(n_planes) > S0 || srcH[run.h0] + run.n > S1) {
              cudaStreamSynchronize(us); destroy_events();
              return fail(ctx, NBE_ERR_STATE, "process_box upload: box out of bounds (internal error)");
            }
            for (int c = 0; c < 3; ++c) {
              cudaError_t e;
              if (w_bytes == row_in) {                      // full rows: a 2-D copy (planes x contiguous row run)
                uint8_t* dst = static_cast<uint8_t*>(ctx->d_box) + ((static_cast<size_t>(c) * nD + ud[i]) * nH + run.h0) * row_in;
                const uint8_t* src = static_cast<const uint8_t*>(in_host) +
                                     ((static_cast<size_t>(c) * S0 + srcD[ud[i]]) * S1 + srcH[run.h0]) * row_in;
                e = cudaMemcpy2DAsync(dst, static_cast<size_t>(nH) * row_in, src, static_cast<size_t>(S1) * row_in,
                                      static_cast<size_t>(run.n) * row_in, n_planes, cudaMemcpyHostToDevice, us);
              } else {
                cudaMemcpy3DParms p3 = {};
                p3.srcPtr = make_cudaPitchedPtr(const_cast<uint8_t*>(static_cast<const uint8_t*>(in_host)) + static_cast<size_t>(c) * S0 * S1 * row_in,
                                                row_in, static_cast<size_t>(S2), static_cast<size_t>(S1));
                p3.dstPtr = make_cudaPitchedPtr(static_cast<uint8_t*>(ctx->d_box) + static_cast<size_t>(c) * nD * nH * row_in,
                                                row_in, static_cast<size_t>(S2), static_cast<size_t>(nH));
                p3.srcPos = make_cudaPos(static_cast<size_t>(wlo) * ies, static_cast<size_t>(srcH[run.h0]), static_cast<size_t>(srcD[ud[i]]));
                p3.dstPos = make_cudaPos(static_cast<size_t>(wlo) * ies, static_cast<size_t>(run.h0), static_cast<size_t>(ud[i]));
                p3.extent = make_cudaExtent(w_bytes, static_cast<size_t>(run.n), n_planes);
                p3.kind = cudaMemcpyHostToDevice;
                e = cudaMemcpy3DAsync(&p3, us);
              }
              if (e != cudaSuccess) {
                cudaStreamSynchronize(us); destroy_events();
                return fail(ctx, NBE_ERR_CUDA, "process_box upload: %s", cudaGetErrorString(e));
              }
            }
            b0 = b1 + 1;
          }
          for (size_t q = i; q <= j; ++q)
            for (int32_t h = run.h0; h < run.h0 + run.n; ++h) have[static_cast<size_t>(ud[q]) * nH + h] |= run.mask;
          any = true;
        }
        i = j + 1;
      }
      if (any) {
        if (ctx->trace) cudaEventCreate(&up_ev[u0]); else cudaEventCreateWithFlags(&up_ev[u0], cudaEventDisableTiming);
        cudaEventRecord(up_ev[u0], us);
      }
      u0 = u1 + 1;
    }
    if (ctx->trace) {           // NBE_TRACE=1: when did the first unit's window land, when the last one
      cudaStreamSynchronize(us);
      float t_first = 0.f, t_last = 0.f;
      cudaEvent_t last = nullptr;
      for (auto e : up_ev) if (e) last = e;
      if (up_ev[0]) cudaEventElapsedTime(&t_first, tr0, up_ev[0]);
      if (last) cudaEventElapsedTime(&t_last, tr0, last);
      fprintf(stderr, "[nbe trace] gpu %d: upload window %lld x %lld rows, first unit after %.2f ms, all after %.2f ms (%.1f MB)\n",
              ctx->device, (long long)nD, (long long)nH, t_first, t_last, in_bytes / 1e6);
      cudaEventDestroy(tr0);
    }
  }
  cudaEvent_t done;
  if (cudaEventCreateWithFlags(&done, cudaEventDisableTiming) != cudaSuccess) {
    cudaStreamSynchronize(us); destroy_events();
    return fail(ctx, NBE_ERR_CUDA, "cudaEventCreate failed");
  }
  // Copy-back policy: consecutive subboxes sharing a D anchor form a run; when a run tiles the
  // whole (H, W) plane the finished D-slab is contiguous per channel and goes back as one large
  // copy per channel, otherwise each owned block is pasted with a strided 3-D copy.
  const bool tiles_hw = crop[1] > 0 && crop[2] > 0 && S1 % crop[1] == 0 && S2 % crop[2] == 0;
  const int run_full = tiles_hw ? static_cast<int>((S1 / crop[1]) * (S2 / crop[2])) : -1;
  int run_start = 0;
  auto flush_run = [&](int s_end) -> int {      // subboxes [run_start, s_end] are enqueued on st
    CK(cudaEventRecord(done, st));
    CK(cudaStreamWaitEvent(cs, done, 0));
    const int n_run = s_end - run_start + 1;
    const int32_t* a0 = &anchors[static_cast<size_t>(run_start) * 3];
    for (int f = 0; f < (ctx->vel ? 2 : 1); ++f) {
      uint8_t* dsrc = static_cast<uint8_t*>(f == 0 ? ctx->d_disp : ctx->d_velo);
      uint8_t* hdst = static_cast<uint8_t*>(f == 0 ? disp_host : vel_host);
      for (int c = 0; c < 3; ++c) {
        uint8_t* dch = dsrc + static_cast<size_t>(c) * ND * plane_out;            // device slab channel
        uint8_t* hch = hdst + (static_cast<size_t>(c) * out_S0 + dlo) * plane_out;    // same planes in the host box
        if (n_run == run_full) {
          const size_t off = static_cast<size_t>(a0[0]) * plane_out;
          CK(cudaMemcpyAsync(hch + off, dch + off, static_cast<size_t>(crop[0]) * plane_out, cudaMemcpyDeviceToHost, cs));
        } else {
          for (int s = run_start; s <= s_end; ++s) {
            const int32_t* ai = &anchors[static_cast<size_t>(s) * 3];
            cudaMemcpy3DParms p3 = {};
            p3.srcPtr = make_cudaPitchedPtr(dch, S2 * es, S2, S1);
            p3.dstPtr = make_cudaPitchedPtr(hch, S2 * es, S2, S1);
            p3.srcPos = make_cudaPos(static_cast<size_t>(ai[2]) * es, ai[1], ai[0]);
            p3.dstPos = p3.srcPos;
            p3.extent = make_cudaExtent(static_cast<size_t>(crop[2]) * es, crop[1], crop[0]);
            p3.kind = cudaMemcpyDeviceToHost;
            CK(cudaMemcpy3DAsync(&p3, cs));
          }
        }
      }
    }
    run_start = s_end + 1;
    return NBE_OK;
  };
  std::vector<int> group_end(static_cast<size_t>(sub_count), 0);
  std::vector<uint8_t> bulk(static_cast<size_t>(sub_count), 0);
  for (int g0 = 0; g0 < sub_count;) {
    int g1 = g0;
    while (g1 + 1 < sub_count && g1 + 1 - g0 < std::max(run_full, 1) &&
           anchors[static_cast<size_t>(g1 + 1) * 3] == anchors[static_cast<size_t>(g0) * 3]) ++g1;
    const bool is_bulk = (g1 - g0 + 1) == run_full && g1 + 1 < sub_count;
    for (int q = g0; q <= g1; ++q) { group_end[q] = g1; bulk[q] = is_bulk ? 1 : 0; }
    g0 = g1 + 1;
  }
  const int32_t size_in[3] = {static_cast<int32_t>(nD), static_cast<int32_t>(nH), size[2]};
  const int32_t size_out[3] = {static_cast<int32_t>(ND), size[1], size[2]};
  rc = process_box_core(ctx, ctx->d_box, in_dtype, size_in, size_out, crop, plen, tabs.data(), anchors.data(), 0, sub_count,
                        Dz, vel_fac, blocks ? disp_blk : ctx->d_disp, blocks ? vel_blk : ctx->d_velo, out_dtype, st,
                        [&](int s) -> int {
                          // before subbox s + 1 may start, its unit's upload must have landed
                          if (s + 1 < sub_count && up_ev[s + 1]) CK(cudaStreamWaitEvent(st, up_ev[s + 1], 0));
                          if (blocks) return NBE_OK;
                          // full (H, W)-tiling runs go back as one contiguous copy per channel when they end;
                          // everything else -- partial runs (sharded ranges) and the last run, whose copy
                          // nothing would hide -- is pasted subbox by subbox behind the next subbox's compute
                          return (!bulk[s] || s == group_end[s]) ? flush_run(s) : NBE_OK;
                        }, blocks, up_ev[0]);
  cudaError_t e0 = cudaStreamSynchronize(us);      // also on error paths: the caller's buffers must be quiescent
  cudaError_t e1 = cudaStreamSynchronize(st);
  cudaError_t e2 = cudaStreamSynchronize(cs);
  cudaEventDestroy(done);
  destroy_events();
  if (rc) return rc;
  if (e0 != cudaSuccess) return fail(ctx, NBE_ERR_CUDA, "process_box upload: %s", cudaGetErrorString(e0));
  if (e1 != cudaSuccess) return fail(ctx, NBE_ERR_CUDA, "process_box: %s", cudaGetErrorString(e1));
  if (e2 != cudaSuccess) return fail(ctx, NBE_ERR_CUDA, "process_box copy: %s", cudaGetErrorString(e2));
  return NBE_OK;
}

int nbe_process_box(nbe_ctx* ctx, const void* in_host, int in_dtype, const int32_t size[3], const int32_t crop[3],
                    const int32_t plen[3], const int32_t* crop_idx, const int32_t* add_idx0, int sub_first,
                    int sub_count, float Dz, float vel_fac, void* disp_host, void* vel_host, int out_dtype) {
  if (!ctx) return NBE_ERR_ARG;
  return process_box_host(ctx, in_host, in_dtype, size, crop, plen, crop_idx, add_idx0, sub_first, sub_count, Dz, vel_fac,
                          disp_host, vel_host, out_dtype, nullptr, nullptr);
}

int nbe_process_box_blocks(nbe_ctx* ctx, const void* in_host, int in_dtype, const int32_t size[3], const int32_t crop[3],
                           const int32_t plen[3], const int32_t* crop_idx, const int32_t* add_idx0, int sub_first,
                           int sub_count, float Dz, float vel_fac, void* disp_blocks_dev, void* vel_blocks_dev, int out_dtype) {
  if (!ctx) return NBE_ERR_ARG;
  if (!disp_blocks_dev) return fail(ctx, NBE_ERR_ARG, "null argument");
  return process_box_host(ctx, in_host, in_dtype, size, crop, plen, crop_idx, add_idx0, sub_first, sub_count, Dz, vel_fac,
                          nullptr, nullptr, out_dtype, disp_blocks_dev, vel_blocks_dev);
}

// One call, all GPUs: the subbox range is cut into contiguous shares (the first n % ngpu contexts get
// one more), one host thread per context runs nbe_process_box on its share.  Every GPU reads its
// window from the SAME page-locked input and DMAs its finished blocks into the SAME output arrays:
// the shares are disjoint, so there is no exchange step and no per-GPU copy of the box on the host.
int nbe_process_box_multi(nbe_ctx** ctxs, int ngpu, const void* in_host, int in_dtype, const int32_t size[3],
                          const int32_t crop[3], const int32_t plen[3], const int32_t* crop_idx, const int32_t* add_idx0,
                          int sub_first, int sub_count, float Dz, float vel_fac, void* disp_host, void* vel_host,
                          int out_dtype, int32_t out_size0) {
  if (!ctxs || ngpu < 1) return NBE_ERR_ARG;
  for (int g = 0; g < ngpu; ++g) if (!ctxs[g]) return NBE_ERR_ARG;
  if (sub_count < 0) return fail(ctxs[0], NBE_ERR_ARG, "sub_count < 0");
  std::vector<int> rcs(static_cast<size_t>(ngpu), NBE_OK);
  std::vector<std::thread> th;
  const int base = sub_count / ngpu, rem = sub_count % ngpu;
  for (int g = 0; g < ngpu; ++g) {
    const int lo = g * base + std::min(g, rem), cnt = base + (g < rem ? 1 : 0);
    th.emplace_back([=, &rcs]() {
      rcs[g] = process_box_host(ctxs[g], in_host, in_dtype, size, crop, plen, crop_idx, add_idx0, sub_first + lo, cnt, Dz,
                                vel_fac, disp_host, vel_host, out_dtype, nullptr, nullptr, out_size0);
    });
  }
  for (auto& t : th) t.join();
  for (int g = 0; g < ngpu; ++g)
    if (rcs[g]) {
      if (g != 0) ctxs[0]->err = "gpu " + std::to_string(ctxs[g]->device) + ": " + ctxs[g]->err;
      return rcs[g];
    }
  return NBE_OK;
}

// Page-lock / unlock a caller-owned host buffer so that nbe_process_box's copies are single
// asynchronous DMAs.  Returns 1 if this call registered the buffer, 0 if it was already
// page-locked (nothing to undo), negative on error.  Never leaves a sticky CUDA error behind.
int nbe_host_register(nbe_ctx* ctx, void* ptr, size_t bytes) {
  if (!ctx || !ptr) return NBE_ERR_ARG;
  ENTER_DEVICE(ctx);
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, ptr) == cudaSuccess && at.type == cudaMemoryTypeHost) return 0;
  cudaGetLastError();
  // portable: every GPU of the process DMAs from / into the same buffer (nbe_process_box_multi)
  cudaError_t e = cudaHostRegister(ptr, bytes, cudaHostRegisterPortable);
  if (e == cudaSuccess) return 1;
  cudaGetLastError();
  if (e == cudaErrorHostMemoryAlreadyRegistered) return 0;
  return fail(ctx, NBE_ERR_CUDA, "cudaHostRegister: %s", cudaGetErrorString(e));
}

int nbe_host_unregister(nbe_ctx* ctx, void* ptr) {
  if (!ctx || !ptr) return NBE_ERR_ARG;
  ENTER_DEVICE(ctx);
  cudaError_t e = cudaHostUnregister(ptr);
  cudaGetLastError();
  return e == cudaSuccess ? NBE_OK : NBE_ERR_CUDA;
}

int nbe_density_from_psi(nbe_ctx* ctx, const float* psi_dev, const int32_t n[3], float boxsize, int32_t res,
                         int32_t worder, float* delta_dev, void* stream) {
  if (!ctx) return NBE_ERR_ARG;
  if (!psi_dev || !delta_dev || !n || n[0] < 1 || n[1] < 1 || n[2] < 1 || res < 1 || !(boxsize > 0.f))
    return fail(ctx, NBE_ERR_ARG, "nbe_density_from_psi: bad argument");
  if (worder < 1 || worder > 4) return fail(ctx, NBE_ERR_ARG, "Unsupported mass-assignment order: %d", worder);
  ENTER_DEVICE(ctx);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long cells = 1ll * res * res * res, np = 1ll * n[0] * n[1] * n[2];
  CK(cudaMemsetAsync(delta_dev, 0, cells * sizeof(float), st));
  const int grid = static_cast<int>(std::min<long long>((np + 255) / 256, 32ll * ctx->num_sms));
  const float scale = static_cast<float>(res) / boxsize;
  switch (worder) {
    case 1: paint_kernel<1><<<grid, 256, 0, st>>>(psi_dev, n[0], n[1], n[2], res, scale, delta_dev); break;
    case 2: paint_kernel<2><<<grid, 256, 0, st>>>(psi_dev, n[0], n[1], n[2], res, scale, delta_dev); break;
    case 3: paint_kernel<3><<<grid, 256, 0, st>>>(psi_dev, n[0], n[1], n[2], res, scale, delta_dev); break;
    default: paint_kernel<4><<<grid, 256, 0, st>>>(psi_dev, n[0], n[1], n[2], res, scale, delta_dev); break;
  }
  const int g2 = static_cast<int>(std::min<long long>((cells + 255) / 256, 32ll * ctx->num_sms));
  rho_to_delta_kernel<<<g2, 256, 0, st>>>(delta_dev, cells, static_cast<float>(static_cast<double>(cells) / np));
  ctx->launches += 2;
  CK(cudaGetLastError());
  return NBE_OK;
}

int nbe_mas_deconvolve(nbe_ctx* ctx, void* delta_k_dev, int32_t res, int32_t worder, void* stream) {
  if (!ctx) return NBE_ERR_ARG;
  if (!delta_k_dev || res < 1 || worder < 1 || worder > 4) return fail(ctx, NBE_ERR_ARG, "nbe_mas_deconvolve: bad argument");
  ENTER_DEVICE(ctx);
  const long long n = 1ll * res * res * (res / 2 + 1);
  const int grid = static_cast<int>(std::min<long long>((n + 255) / 256, 32ll * ctx->num_sms));
  mas_deconvolve_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<float2*>(delta_k_dev), res, worder);
  ctx->launches += 1;
  CK(cudaGetLastError());
  return NBE_OK;
}

int nbe_pk_bins(nbe_ctx* ctx, const void* delta_k_dev, int32_t res, int32_t mas_order, int32_t nbins,
                double* out_dev, void* stream) {
  if (!ctx) return NBE_ERR_ARG;
  if (!delta_k_dev || !out_dev || res < 1 || mas_order < 0 || mas_order > 4 || nbins < 1 || nbins > kPkMaxBins)
    return fail(ctx, NBE_ERR_ARG, "nbe_pk_bins: bad argument");
  ENTER_DEVICE(ctx);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CK(cudaMemsetAsync(out_dev, 0, 3 * sizeof(double) * nbins, st));
  const long long n = 1ll * res * res * (res / 2 + 1);
  const int grid = static_cast<int>(std::min<long long>((n + 255) / 256, 4ll * ctx->num_sms));
  const size_t smem = 3 * sizeof(double) * nbins;
  if (smem > 48 * 1024) CK(cudaFuncSetAttribute(pk_bins_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  pk_bins_kernel<<<grid, 256, smem, st>>>(static_cast<const float2*>(delta_k_dev), res, mas_order, nbins, out_dev);
  ctx->launches += 1;
  CK(cudaGetLastError());
  return NBE_OK;
}

int nbe_za_psi_k(nbe_ctx* ctx, const void* delta_k_dev, int32_t res, float boxsize, void* psi_k_dev, void* stream) {
  if (!ctx) return NBE_ERR_ARG;
  if (!delta_k_dev || !psi_k_dev || res < 1 || !(boxsize > 0.f)) return fail(ctx, NBE_ERR_ARG, "nbe_za_psi_k: bad argument");
  ENTER_DEVICE(ctx);
  const long long n = 1ll * res * res * (res / 2 + 1);
  const int grid = static_cast<int>(std::min<long long>((n + 255) / 256, 32ll * ctx->num_sms));
  za_psi_k_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const float2*>(delta_k_dev), res, boxsize / 6.283185307179586f, static_cast<float2*>(psi_k_dev));
  ctx->launches += 1;
  CK(cudaGetLastError());
  return NBE_OK;
}

// Give the activation arena, the cached plans and the box staging buffers back to the driver (they are
// re-created on demand).  Parameters and packed weights stay.
int nbe_release_workspace(nbe_ctx* ctx) {
  if (!ctx) return NBE_ERR_ARG;
  ENTER_DEVICE(ctx);
  CK(cudaDeviceSynchronize());
  for (auto* p : ctx->plans) { delete p; }
  ctx->plans.clear();
  cudaFree(ctx->arena); ctx->arena = nullptr; ctx->arena_cap = 0;
  cudaFree(ctx->d_box); ctx->d_box = nullptr; ctx->box_cap = 0;
  cudaFree(ctx->d_disp); ctx->d_disp = nullptr; ctx->out_cap = 0;
  cudaFree(ctx->d_velo); ctx->d_velo = nullptr; ctx->velo_cap = 0;
  return NBE_OK;
}

int64_t nbe_launch_count(nbe_ctx* ctx, int reset) {
  if (!ctx) return -1;
  const int64_t n = ctx->launches;
  if (reset) ctx->launches = 0;
  return n;
}

int nbe_set_profiling(nbe_ctx* ctx, int enable) {
  if (!ctx) return NBE_ERR_ARG;
  ENTER_DEVICE(ctx);
  resolve_profile(ctx);
  ctx->profiling = enable != 0;
  if (enable) {     // start a fresh accumulation
    std::fill(ctx->prof_sum_ms.begin(), ctx->prof_sum_ms.end(), 0.0);
    ctx->prof_samples = 0;
  }
  return NBE_OK;
}

int nbe_get_profile(nbe_ctx* ctx, int cap, const char** names, float* ms, double* flops) {
  if (!ctx) return NBE_ERR_ARG;
  ENTER_DEVICE(ctx);
  resolve_profile(ctx);
  const int n = static_cast<int>(ctx->prof_names.size());
  for (int i = 0; i < n && i < cap; ++i) {
    if (names) names[i] = ctx->prof_names[i].c_str();
    if (ms) ms[i] = ctx->prof_samples ? static_cast<float>(ctx->prof_sum_ms[i] / ctx->prof_samples) : 0.f;
    if (flops) flops[i] = ctx->prof_flops[i];
  }
  return n;
}

// Debug: copy one activation tensor (NDHWC fp16) of the most recent plan to the host.
// which: 0 = hi, 1 = lo, 2 = tangent.  shape_out = {d, h, w, c}.  Returns bytes copied or <0.
long long nbe_debug_read_act(nbe_ctx* ctx, int act, int which, void* host, size_t cap, int32_t shape_out[4]) {
  if (!ctx || ctx->plans.empty() || act < 0 || act >= A_OUT) return NBE_ERR_ARG;
  DevGuard dev_guard_(ctx->device);
  cudaDeviceSynchronize();
  Plan* P = ctx->plans.back();
  const ActBuf& B = P->act[act];
  const size_t bytes = static_cast<size_t>(B.c) * B.d * B.h * B.w * 2;
  if (shape_out) { shape_out[0] = B.d; shape_out[1] = B.h; shape_out[2] = B.w; shape_out[3] = B.c; }
  if (!host) return static_cast<long long>(bytes);
  if (cap < bytes) return NBE_ERR_ARG;
  const size_t off = which == 0 ? B.off_hi : (which == 1 ? B.off_lo : B.off_dx);
  if (cudaMemcpy(host, ctx->arena + off, bytes, cudaMemcpyDeviceToHost) != cudaSuccess) return NBE_ERR_CUDA;
  return static_cast<long long>(bytes);
}


// Debug: the fold vector a that the producer of activation `act` added to its stored tangent (dx' = dx + a (.) x);
// zeros when the tensor has none.  Returns the channel count or <0.
int nbe_debug_act_fold(nbe_ctx* ctx, int act, int sample, float* a_host, int cap) {
  if (!ctx || !a_host || act < 0 || act >= A_OUT || ctx->plans.empty()) return NBE_ERR_ARG;
  DevGuard dev_guard_(ctx->device);
  cudaDeviceSynchronize();
  const int c = act_channels(act);
  if (cap < c) return NBE_ERR_ARG;
  for (int i = 0; i < c; ++i) a_host[i] = 0.f;
  if (sample < 0 || sample >= ctx->mod_batch) return NBE_ERR_ARG;
  for (size_t li = 0; li < ctx->sl.size(); ++li)
    if (ctx->sl[li].out_act == act && ctx->sl[li].anext_layer >= 0 && ctx->vel) {
      const float* src = ctx->d_fold + (static_cast<size_t>(sample) * ctx->sl.size() + li) * 256 + 128;
      if (cudaMemcpy(a_host, src, c * sizeof(float), cudaMemcpyDeviceToHost) != cudaSuccess) return NBE_ERR_CUDA;
    }
  return c;
}

// 1 when the velocity launches run with the folded tangent (FOLD instances), else 0
int nbe_fold_active(nbe_ctx* ctx) { return (ctx && ctx->fold_active) ? 1 : 0; }

}  // extern "C"
