// Implicit-GEMM 3D convolution for sm_100a: TMA-fed tcgen05.mma with fp32 accumulators in
// TMEM, warp-specialised (A-producer / B-producer / MMA issuer / 4 epilogue warps),
// persistent over output tiles.
//
//   GEMM view:  M = output voxels (tile 8w x 16h, TM tiles stacked in h per work item),
//               N = output channels of [y | dy],  K = taps x input channels.
//
// One launch evaluates a *sum of conv terms* ("groups") into the same accumulators: the
// 27 taps of a 3^3 conv, the K-chunks of 128-channel inputs (un-materialised concat), the
// folded 1^3 skip conv of a ResNet block (style_blocks_vel.py:112-159), the Dz-tangent
// terms x*dW + dx*W (style_layers_vel.py:129-141) and, in split precision, the hi/lo
// products.  A group loads one halo'd activation block {C, 8w, 16*TM+2 h, 1 d} per operand
// with a single TMA box; its up-to-3 `kh` taps are served from that block by advancing the
// UMMA descriptor start address by whole 8-row swizzle atoms (1024 B), so every activation
// byte is fetched once per (kd,kw) instead of once per tap.
//
// Epilogue (fused): + bias, LeakyReLU with the tangent rule dy = y>0 ? dy : 0.01 dy
// (layers_vel.py:178-186), fp16 hi (+lo) / tangent stores in NDHWC; or, for the last layer,
// the model tail  disp = (net + x0)*6, vel = dnet*6vf + x0*6vf/Dz
// (style_nbody_emulator_vel_core.py:187-193) written as NCDHW into the caller's box.
#pragma once
#include "ptx.cuh"

namespace nbe {

constexpr int kMaxGroups = 64;
constexpr int kMaxAMaps = 32;
constexpr int kConvThreads = 384;          // 4 pipeline warps + 8 epilogue warps
constexpr int kSmemLimit = 227 * 1024;
constexpr float kWeightScale = 256.0f;          // packed weights carry this factor (see modulate_kernel)
constexpr float kInvWeightScale = 1.0f / 256.0f;

struct MmaOp {
  uint16_t a_off;   // index (0 / 1) of the group's activation block this op multiplies
  uint16_t b_off;   // byte offset >> 4 of the first B row inside the weight stage
  uint16_t d_col;   // first accumulator column
  uint8_t n8;       // N / 8
  uint8_t pad_;
};

struct GroupDesc {
  int16_t a_map[2];   // tensor-map index of operand 0 / 1 (-1: absent)
  int8_t n_a;         // operand blocks to load (1 or 2)
  int8_t ntaps;       // taps served from the block: 1, 3 (kh) or 9 (kw-major: j = kw*3 + kh)
  int8_t kc16;        // 1: 16-channel rows (32 B, SWIZZLE_32B); 0: 64-channel rows (128 B)
  int8_t n_ops;
  int16_t c0;         // channel coordinate of the box
  int8_t dw, dh, dd;  // block origin relative to the tile origin
  int8_t pitch;       // rows per h-line of the block: 8, or 10 (w-halo'd "wide" block, kw taps by row shift)
  int8_t tps;         // taps per weight stage: 1, or 3 for 16-channel rows (three small tiles share a stage)
  // early-drain protocol of the per-kd accumulator layout (EARLY instances, see the epilogue):
  int8_t pre_wait;    // before this group's MMAs wait for  1: y0_empty  2: acc_empty (dy | y1 | y2)
  int8_t post_sig;    // after this group commit           1: y0_full   2: y1_full
  int8_t lo_stage;    // weight stages of this group use the short box of bmap64_lo (lo products)
  int32_t brow0;      // B row of tap 0
  int16_t brow_step;  // B rows between consecutive taps
  int16_t tap_rows;   // rows between consecutive taps INSIDE a multi-tap weight stage (this CTA's rows per tap)
  // EARLY == 3 (accumulation chains): the group's taps are cut into blocks of 3; block b accumulates into
  // primal accumulator (phase0 + b) & 1 (0: y0, 1: y1), which is handed to the epilogue after the block
  int8_t chain;
  int8_t phase0;
  int8_t pad_[2];
  MmaOp ops[3];
};

// Passed by value as a __grid_constant__ kernel parameter: lives in the constant bank, so the
// pipeline warps read it with uniform loads straight into uniform registers.
struct GroupTable {
  int32_t n_groups;
  int32_t pad_[3];
  GroupDesc g[kMaxGroups];
};

struct alignas(128) ConvLaunch {
  CUtensorMap amap[kMaxAMaps];
  CUtensorMap bmap64;
  CUtensorMap bmap16;
  CUtensorMap bmap64_lo;    // same tensor as bmap64 with a box of lo_rows rows (lo-product stages use fewer rows);
                            // lo_taps == 3: a 3-D map {64, rows, tiles} whose box carries three consecutive taps
  int32_t lo_rows;          // rows per CTA of a lo stage (0: lo stages use bmap64)
  int32_t lo_taps;          // taps per lo-stage box: 1 or 3
  int32_t main_rows;        // rows per CTA of a main 64-channel stage (0: the whole stage); the folded 128-output
                            // instance stages 64 Wh rows per tap and 128 rows [Wl | Wh] per lo tap
  int32_t n_chains;         // EARLY == 3: accumulation chains per item (1 lo chain + the 3-tap blocks of the main groups)
  int32_t n_par;            // 1, or 8 for the x2 up-sampling conv (one parity per item)
  int32_t par_brow_step;    // B rows between parities
  int32_t out_w, out_h, out_d;         // extent of the tile space
  int64_t out_sw, out_sh, out_sd;      // output strides (elements) in tile space
  int64_t par_ow, par_oh, par_od;      // output offsets (elements) of parity bits c, b, a
  __half* out_h_ptr;        // primal hi  [.., cout_stride]
  __half* out_l_ptr;        // primal lo or nullptr
  __half* out_d_ptr;        // tangent or nullptr
  const float* bias;        // [cout]
  // Tangent folding.  dW = W (.) (a_i + beta_o) for a style-modulated layer, so
  //   x * dW + dx * W  =  (dx + a (.) x) * W  +  beta (.) (x * W):
  // `anext` (or nullptr) is the fold vector a of the layer that consumes this launch's output -- the epilogue
  // stores dx' = dy + anext (.) y instead of dy; `beta` (FOLD instance only) is this launch's own per-output
  // factor, applied to the primal sum.
  const float* beta;        // [cout] or nullptr
  const float* anext;       // [cout] or nullptr
  int32_t cout;             // output channels (primal)
  int32_t vel;              // accumulator holds [y | dy]
  int32_t act;              // LeakyReLU
  int32_t acc3;             // accumulator columns are [y0 | dy | y1 | y2] (one primal accumulator per kd)
  int32_t band_h;           // item order: bands of band_h tile rows, d swept inside a band (0: whole planes)
};

// Per-call arguments of the last layer's fused model tail.
struct FinalArgs {
  const void* src;          // raw input box (3, S0, S1, S2), in_dtype
  int32_t src_dtype;
  int64_t src_sc, src_sd, src_sh;      // element strides (w contiguous)
  const int32_t* idx_d;     // gather tables of the centre crop (already offset by the pad)
  const int32_t* idx_h;
  const int32_t* idx_w;
  void* disp;               // (3, ...) out_dtype, base at the paste anchor
  void* vel;                // or nullptr
  int32_t out_dtype;
  int32_t mid_dtype;        // compute dtype of the model: results are rounded to it first
  int64_t o_sc, o_sd, o_sh;
  float in_norm;            // Dz/6
  float six;                // 6
  float dx_norm;            // 6*vel_fac
  float x0_norm;            // 6*vel_fac/Dz
};

__device__ __forceinline__ float load_as_f32(const void* p, int64_t i, int dtype) {
  if (dtype == 0) return reinterpret_cast<const float*>(p)[i];
  if (dtype == 1) return __half2float(reinterpret_cast<const __half*>(p)[i]);
  return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i]);
}
__device__ __forceinline__ float round_to_dtype(float v, int dtype) {
  if (dtype == 1) return __half2float(__float2half_rn(v));
  if (dtype == 2) return __bfloat162float(__float2bfloat16_rn(v));
  return v;
}
// x * (Dz/6) as the reference computes it, i.e. in the compute dtype with one rounding
// (style_nbody_emulator_vel_core.py:132-134).
__device__ __forceinline__ float scale_in_dtype(float x, float in_norm, int dtype) {
  return round_to_dtype(x * round_to_dtype(in_norm, dtype), dtype);
}
__device__ __forceinline__ void store_from_f32(void* p, int64_t i, int dtype, float v) {
  if (dtype == 0) reinterpret_cast<float*>(p)[i] = v;
  else if (dtype == 1) reinterpret_cast<__half*>(p)[i] = __float2half_rn(v);
  else reinterpret_cast<__nv_bfloat16*>(p)[i] = __float2bfloat16_rn(v);
}


template <int NRS, int DC, int TM, bool PAIR = false>
struct ConvCfg {
  static constexpr int kHRows = 16 * TM + 2;               // h-lines of one activation block
  static constexpr int kABlk = (kHRows * 10 * 128 + 1023) / 1024 * 1024;   // bytes, wide 64-channel block
  static constexpr int kAStage = 2 * kABlk;
  // pair mode: NRS counts the rows of BOTH CTAs; each stages its halves of every op's rows
  static constexpr int kBRows = PAIR ? NRS / 2 : NRS;       // rows this CTA stages per tap
  static constexpr int kBStage = kBRows * 128;
  static constexpr int kNA = (TM == 1) ? 3 : 2;
  // The activation ring is a ring of BLOCKS: a group takes as many consecutive slots as it has operand
  // blocks (1 or 2), so the one-block groups of a lo phase leave room to prefetch further ahead.
  static constexpr int kNAB = 2 * kNA;
  static constexpr int kCtrl = 2048;                       // barriers, TMEM slot, bias
  static constexpr int kBudget = kSmemLimit - 1024 /*align*/ - kCtrl - kNA * kAStage;
  static constexpr int kNBraw = kBudget / kBStage;
  static constexpr int kNB = kNBraw > 12 ? 12 : kNBraw;
  static constexpr int kNBuf = (2 * TM * DC <= 512) ? 2 : 1;
  static constexpr int kColsRaw = kNBuf * TM * DC;
  static constexpr int kTmemCols = kColsRaw <= 32 ? 32 : kColsRaw <= 64 ? 64 : kColsRaw <= 128 ? 128
                                   : kColsRaw <= 256 ? 256 : 512;
  static constexpr int kSmemBytes = 1024 + kNA * kAStage + kNB * kBStage + kCtrl;
  static_assert(kNB >= 2, "not enough shared memory for the B ring");
  static_assert(3 * kBRows * 32 <= kBStage, "a weight stage holds three 16-channel taps");
  static_assert(kColsRaw <= 512, "TMEM overflow");
};

// 0: base_offset field left 0 for row-shifted operand starts; 1: base_offset = (addr >> 7) & 7
#ifndef NBE_BASE_OFFSET
#define NBE_BASE_OFFSET 0
#endif
constexpr bool kBaseOffsetMode = NBE_BASE_OFFSET != 0;

// 1: 9-tap 64-channel groups are issued one kw-block (three taps) per elected section; 0: tap by tap
#ifndef NBE_BLOCK_ISSUE
#define NBE_BLOCK_ISSUE 1
#endif
constexpr bool kBlockIssue = NBE_BLOCK_ISSUE != 0 && !kBaseOffsetMode;

// EARLY (only with the per-kd accumulator layout [y0 | dy | y1 | y2], one TMEM stage): fine-grained
// accumulator hand-over.  The issuers order an item as  lo products (-> y0) | kd 0 (+ folded skip)
// -> (y0, dy) | kd 1 -> (dy, y1) | kd 2 -> (y2, dy)  and commit y0_full / y1_full as soon as the
// last MMA writing y0 / y1 has been issued; the epilogue drains y0 and y1 into registers while
// the later kd-planes are still being multiplied, and only (y2, dy) are left when the item ends.
// The next item's lo products need nothing but y0 (already drained and re-zeroed), so they run
// under the epilogue's math and stores: the single-buffered TMEM no longer idles the tensor pipe.
// Same MMAs per accumulator in the same order and the same (y0 + y1) + y2 sum: bit-identical to the
// plain acc3 path.
//
// EARLY == 2 ("F192", 64-output pair instance): columns [y1 | dy | ylo | y0].  The three products that
// share the activation operand xh are ONE instruction per k-step, N = 192: weight rows [dW | Wl | Wh]
// land on (dy, ylo, y0) for kd 0 / 2 and rows [Wh | dW | Wl] on (y1, dy, ylo) for kd 1; dx * Wh -> dy
// (N = 64) is the second; the lo phase is the single product xl * Wh (N = 64) accumulated into y1 BEFORE
// kd 1's hi products.  Per k-step that is 3 MMAs and 3 activation-operand reads instead of 4.33, and
// xh is fetched by TMA once per kd-plane, not twice.  kd 2 re-uses the y0 columns: the epilogue drains
// y0 (then y1) while kd 1 (kd 2) runs, the issuers wait for y0_empty before kd 2, and the next item's
// lo phase needs only y1 (drained during kd 2), so the end-of-item drain of (dy, ylo, y0') hides under it.
// Sum order ((y0 + y1) + y0') + ylo.
//
// EARLY == 3 ("chains", same instance and column layout as F192): the tensor core TRUNCATES when it adds
// into the fp32 accumulator, a systematic relative shrink that grows with the number of accumulations
// into one column.  Here the two primal accumulators are used in turn: the lo
// phase (xl * Wh, all taps) is one chain into y1, then every block of three taps of the main groups
// (12 accumulations) goes to y0 / y1 alternately, with the weight rows of a tap ordered for the
// accumulator its block uses ([dW | Wl | Wh] -> (dy, ylo, y0), [Wh | dW | Wl] -> (y1, dy, ylo)).  After each
// chain the issuers commit y?_full, the epilogue (idle 80 % of the time) adds the 64 columns to its
// register sum with round-to-nearest adds, re-zeroes them and arrives on y?_empty, which the issuers
// wait on before the next chain into the same accumulator -- one whole block later.  dy and ylo
// accumulate through the item as before.
//
// EARLY == 4 ("FOLD", chains + folded tangent): the consumer-side product x * dW is gone (see ConvLaunch::beta /
// anext), so a tap is  xh * [Wl | Wh] -> (ylo, y0)  or  xh * [Wh | Wl] -> (y1, ylo)  (ONE instruction, N = 128) and
// dx' * Wh -> dy (N = 64): 4 MMA column-blocks per k-step instead of 5.  Columns [dy | y1 | ylo | y0], 2 x 96
// weight rows per stage.  A folded 1^3 skip conv reads a tensor with a different fold vector: its residual
// x * dW_res -> dy is a one-tap group of its own with a short weight stage.
template <int NRS, int DC, int TM, bool FINAL, bool PAIR = false, int EARLY = 0>
__global__ void __launch_bounds__(kConvThreads, 1)
conv_mma_kernel(const __grid_constant__ ConvLaunch launch, const __grid_constant__ GroupTable gt,
                const __grid_constant__ FinalArgs fa) {
  // The launch record (tensor maps included) is a kernel parameter: TMA reads the maps from the constant bank, which
  // the driver keeps coherent per launch.  (Maps in global memory written by cudaMemcpy would need a
  // fence.proxy.tensormap::generic.acquire.sys per map and producer thread -- tried: 25 serial fences cost the short
  // resampling launches 0.1 ms each.)
  const ConvLaunch* const L = &launch;
  using Cfg = ConvCfg<NRS, DC, TM, PAIR>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);

  uint8_t* a_smem = smem;
  uint8_t* b_smem = a_smem + Cfg::kNA * Cfg::kAStage;
  uint64_t* bars = reinterpret_cast<uint64_t*>(b_smem + Cfg::kNB * Cfg::kBStage);
  uint64_t* a_full = bars;
  uint64_t* a_empty = a_full + Cfg::kNAB;
  uint64_t* b_full = a_empty + Cfg::kNAB;
  uint64_t* b_empty = b_full + Cfg::kNB;
  uint64_t* acc_full = b_empty + Cfg::kNB;
  uint64_t* acc_empty = acc_full + 2;
  uint64_t* y0_full = acc_empty + 2;       // EARLY only
  uint64_t* y1_full = y0_full + 1;
  uint64_t* y0_empty = y1_full + 1;
  uint64_t* y1_empty = y0_empty + 1;       // EARLY == 2 only
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(y1_empty + 1);
  static_assert(EARLY == 0 || (Cfg::kNBuf == 1 && !FINAL && TM * DC == 512), "EARLY: single-stage acc3 instances only");
  static_assert(EARLY < 2 || (PAIR && TM == 2 && DC == 256), "F192: 64-output pair instance only");
  constexpr bool kChain = EARLY >= 3;      // accumulation chains
  constexpr bool kFold = EARLY == 4;       // ... with the folded tangent and the [dy | y1 | ylo | y0] layout
  float* bias_s = reinterpret_cast<float*>(tmem_slot + 2);   // up to 128 floats
  float* beta_s = bias_s + 128;
  float* anext_s = beta_s + 128;
  static_assert(8 * (2 * Cfg::kNAB + 2 * Cfg::kNB + 8) + 8 + 3 * 512 <= Cfg::kCtrl, "control block overflow");

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr int kIssuers = TM == 2 ? 2 : 1;
  // CTA pair (cta_group::2): rank 0 is the leader; its barriers collect the TMA bytes of both
  // CTAs and it alone issues the M = 256 MMAs, each CTA contributing its own 128 voxel rows and
  // its half of every weight operand's rows
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;

  const int n_groups = gt.n_groups;
  const int out_w = L->out_w, out_h = L->out_h, out_d = L->out_d;
  const int tiles_w = (out_w + 7) >> 3;
  const int tiles_h = (out_h + 16 * TM - 1) / (16 * TM);
  const int n_par = L->n_par;
  const long long n_items = 1ll * n_par * out_d * tiles_h * tiles_w;

  if (threadIdx.x == 0) {
    // kIssuers MMA-issuing warps (one per M-tile when TM == 2): each commits to the empty / full
    // barriers, so those expect kIssuers arrivals
    for (int i = 0; i < Cfg::kNAB; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], kIssuers); }
    for (int i = 0; i < Cfg::kNB; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], kIssuers); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], kIssuers); mbar_init(&acc_empty[i], PAIR ? 512 : 256); }
    mbar_init(y0_full, kIssuers); mbar_init(y1_full, kIssuers); mbar_init(y0_empty, PAIR ? 512 : 256);
    mbar_init(y1_empty, PAIR ? 512 : 256);
    fence_mbar_init();
  }
  if (warp == 2) {
    if constexpr (PAIR) tmem_alloc_2sm<Cfg::kTmemCols>(tmem_slot);
    else tmem_alloc<Cfg::kTmemCols>(tmem_slot);
  }
  if (warp == 3) {
    const int nb = FINAL ? 16 : L->cout;
    for (int i = lane; i < nb; i += 32) bias_s[i] = L->bias[i];
    const float* bp = L->beta;
    const float* ap = L->anext;
    for (int i = lane; i < 128; i += 32) {
      beta_s[i] = (bp != nullptr && i < nb) ? bp[i] : 0.f;
      anext_s[i] = (ap != nullptr && i < nb) ? ap[i] : 0.f;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // Accumulators start at zero and are re-zeroed by the epilogue after each read, so every MMA
  // accumulates and the launch may order its terms freely (small lo-products first: the tensor
  // core truncates on accumulation, and truncation error scales with the running sum).
  if (warp >= 4 && warp < 8) {
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    if constexpr (DC >= 32) {
      for (int c = 0; c < Cfg::kNBuf * TM * DC; c += 32) tmem_st32_zero(lane_base + c);
    } else {
      for (int c = 0; c < Cfg::kNBuf * TM * DC; c += 16) tmem_st16_zero(lane_base + c);
    }
    tmem_st_wait();
  }
  tc_fence_before();
  if constexpr (PAIR) cluster_sync_all(); else __syncthreads();
  tc_fence_after();

  // a pair walks the item list two at a time (leader takes the even one); both CTAs of a pair
  // therefore run the same number of iterations
  const long long item_first = PAIR ? static_cast<long long>(blockIdx.x & ~1u) : static_cast<long long>(blockIdx.x);
  // Item order (slowest to fastest): h-band, d, tile row inside the band, tile column, parity.
  // The CTAs in flight then cover a narrow band over several d planes, so the three input planes
  // a 3^3 tap window re-reads stay L2-resident between their uses.
  const int band_h = (L->band_h > 0 && L->band_h < tiles_h) ? L->band_h : tiles_h;
  const long long band_items = 1ll * out_d * band_h * tiles_w;
  const int n_bands = (tiles_h + band_h - 1) / band_h;
  auto decode = [&](long long item, int& par, int& w0, int& h0, int& d0) {
    // x2 up-sampling: the 8 output parities of one input tile are CONSECUTIVE items (they run on 8 CTAs at the same
    // time), so the input tile is read from DRAM once and the interleaved output lines are completed together
    par = n_par > 1 ? static_cast<int>(item % n_par) : 0;
    long long r = n_par > 1 ? item / n_par : item;
    int band = static_cast<int>(r / band_items);
    if (band > n_bands - 1) band = n_bands - 1;
    r -= band * band_items;
    const int bh = min(band_h, tiles_h - band * band_h);
    const int plane = bh * tiles_w;
    d0 = static_cast<int>(r / plane);
    const int r3 = static_cast<int>(r - static_cast<long long>(d0) * plane);
    const int th = band * band_h + r3 / tiles_w;
    w0 = (r3 % tiles_w) * 8;
    h0 = th * 16 * TM;
  };

  // With two h-stacked tiles per item, the upper tile of the last tile row can lie wholly outside
  // the output (e.g. out_h = 130: rows 144..159).  Its MMAs are skipped (the barrier protocol is
  // kept), which saves the tensor work and, the board being power-limited, its energy.  In a pair
  // the M = 256 instruction covers both CTAs' items, so the tile must be dead in both.
  auto tile_dead = [&](long long it0, int t) -> bool {
    if (TM != 2 || t == 0) return false;
    bool dead = true;
    for (int c = 0; c < (PAIR ? 2 : 1); ++c) {
      const long long item = it0 + c;
      if (item >= n_items) continue;
      int par, w0, h0, d0;
      decode(item, par, w0, h0, d0);
      dead = dead && (h0 + 16 * t >= out_h);
    }
    return dead;
  };

  if (warp == 0) {
    // ------------------------------------------------ A producer (activation blocks)
    if (lane == 0) {
      uint32_t s = 0, ph = 0;
      for (long long it0 = item_first; it0 < n_items; it0 += gridDim.x) {
        const long long item = it0 + rank;
        int par, w0, h0, d0;
        decode(item < n_items ? item : 0, par, w0, h0, d0);
        if (item >= n_items) d0 = out_d + 1024;        // idle half of the last pair: every load is out of bounds (zeros)
        for (int g = 0; g < n_groups; ++g) {
          const GroupDesc& G = gt.g[g];
          const uint32_t blk = static_cast<uint32_t>(Cfg::kHRows) * G.pitch * (G.kc16 ? 32u : 128u);
          for (int q = 0; q < G.n_a; ++q) {
            mbar_wait(&a_empty[s], ph ^ 1);
            if (!PAIR || rank == 0) mbar_expect_tx(&a_full[s], blk * (PAIR ? 2 : 1));
            if constexpr (PAIR)
              tma_load_4d_2sm(a_smem + s * Cfg::kABlk, &L->amap[G.a_map[q]], &a_full[s],
                              G.c0, w0 + G.dw, h0 + G.dh, d0 + G.dd);
            else
              tma_load_4d(a_smem + s * Cfg::kABlk, &L->amap[G.a_map[q]], &a_full[s],
                          G.c0, w0 + G.dw, h0 + G.dh, d0 + G.dd);
            if (++s == Cfg::kNAB) { s = 0; ph ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------ B producer (weight tiles)
    if (lane == 0) {
      uint32_t s = 0, ph = 0;
      const int par_brow_step = L->par_brow_step;
      const int lo_rows = L->lo_rows, lo_taps = L->lo_taps;
      const uint32_t main_rows = L->main_rows > 0 ? static_cast<uint32_t>(L->main_rows) : static_cast<uint32_t>(Cfg::kBRows);
      for (long long it0 = item_first; it0 < n_items; it0 += gridDim.x) {
        int par, w0, h0, d0;
        decode(it0, par, w0, h0, d0);
        for (int g = 0; g < n_groups; ++g) {
          const GroupDesc& G = gt.g[g];
          const bool lo_box = G.lo_stage != 0 && lo_rows > 0;
          const CUtensorMap* bm = G.kc16 ? &L->bmap16 : (lo_box ? &L->bmap64_lo : &L->bmap64);
          const uint32_t bytes = lo_box ? static_cast<uint32_t>(lo_rows) * 128u : (G.kc16 ? Cfg::kBRows * 32u : main_rows * 128u);
          const int row0 = G.brow0 + par * par_brow_step + (PAIR ? static_cast<int>(rank) * Cfg::kBRows : 0);
          const int tps = G.tps;
          for (int j = 0; j < G.ntaps; j += tps) {
            mbar_wait(&b_empty[s], ph ^ 1);
            uint8_t* dst = b_smem + s * Cfg::kBStage;
            if (G.kc16) {
              // 16-channel tiles: one 3-D box {16 ch, this CTA's rows, 3 consecutive taps} per stage
              if (!PAIR || rank == 0) mbar_expect_tx(&b_full[s], bytes * 3 * (PAIR ? 2 : 1));
              const int tile = (G.brow0 + par * par_brow_step) / NRS + j;
              const int r0 = PAIR ? static_cast<int>(rank) * Cfg::kBRows : 0;
              if constexpr (PAIR) tma_load_3d_2sm(dst, bm, &b_full[s], 0, r0, tile);
              else tma_load_3d(dst, bm, &b_full[s], 0, r0, tile);
            } else if (lo_box && lo_taps == 3) {
              // three taps of lo rows per stage (the box always carries three tiles; a one-tap group uses the first)
              if (!PAIR || rank == 0) mbar_expect_tx(&b_full[s], bytes * 3 * (PAIR ? 2 : 1));
              const int tile = (G.brow0 + par * par_brow_step) / NRS + j;
              const int r0 = PAIR ? static_cast<int>(rank) * Cfg::kBRows : 0;
              if constexpr (PAIR) tma_load_3d_2sm(dst, bm, &b_full[s], 0, r0, tile);
              else tma_load_3d(dst, bm, &b_full[s], 0, r0, tile);
            } else {
              if (!PAIR || rank == 0) mbar_expect_tx(&b_full[s], bytes * (PAIR ? 2 : 1));
              if constexpr (PAIR) tma_load_2d_2sm(dst, bm, &b_full[s], 0, row0 + j * G.brow_step);
              else tma_load_2d(dst, bm, &b_full[s], 0, row0 + j * G.brow_step);
            }
            if (++s == Cfg::kNB) { s = 0; ph ^= 1; }
          }
        }
      }
    }
  } else if (warp == 2 || (warp == 3 && kIssuers == 2)) {
    // ------------------------------------------------ MMA issuer(s)
    // The whole warp runs this loop with warp-uniform control flow and values (so they stay in
    // uniform registers); one elected lane issues tcgen05.mma / tcgen05.commit.  With TM == 2 two
    // warps issue, one per M-tile: the issue path (not the tensor pipe) was the limiter, and
    // tile-disjoint accumulators keep the result independent of the interleaving.
    const int my_tile = warp - 2;
    auto mma = [](uint32_t d, uint64_t ad, uint64_t bd, uint32_t id, uint32_t acc) {
      if constexpr (PAIR) umma_f16_2sm(d, ad, bd, id, acc); else umma_f16(d, ad, bd, id, acc);
    };
    auto commit = [](uint64_t* bar) {
      if constexpr (PAIR) umma_commit_2sm(bar); else umma_commit(bar);
    };
    constexpr uint32_t idesc_base = umma_idesc_f16(PAIR ? 256 : 128, 0, false);
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    uint32_t sa = 0, pa = 0, sb = 0, pb = 0, buf = 0, pacc = 0;
    uint32_t pe = 0;      // chains: bit a = parity of the next wait on y<a>_empty
    for (long long item = item_first; item < n_items && (!PAIR || rank == 0); item += gridDim.x) {
      if constexpr (EARLY == 0) {
        mbar_wait(&acc_empty[buf], pacc ^ 1);
        tc_fence_after();
      }
      const bool dead = tile_dead(item, my_tile);
      for (int g = 0; g < n_groups; ++g) {
        const GroupDesc& G = gt.g[g];
        const int ntaps = G.ntaps, n_ops = G.n_ops;
        if constexpr (EARLY != 0) {
          // pre_wait 1: y0_empty, 2: acc_empty, 3: y1_empty.  All are signalled by the PREVIOUS item's
          // epilogue, except y0_empty in the F192 layout, which kd 2 needs from the CURRENT item
          const int pw = G.pre_wait;
          if (kChain && pw == 3) {                // start of the lo chain (into y1)
            mbar_wait(y1_empty, ((pe >> 1) & 1u) ^ 1u);
            pe ^= 2u;
            tc_fence_after();
          } else if (pw != 0) {
            uint64_t* bar = pw == 1 ? y0_empty : (pw == 2 ? &acc_empty[0] : y1_empty);
            mbar_wait(bar, (EARLY == 2 && pw == 1) ? pacc : (pacc ^ 1));
            tc_fence_after();
          }
        }
        const bool k16 = G.kc16 != 0;
        const uint32_t rowb = k16 ? 32u : 128u;
        const uint32_t row16 = rowb >> 4;                       // one row in 16-byte units
        const uint32_t pitch = static_cast<uint32_t>(G.pitch);
        const uint32_t sbo16 = pitch * row16;                   // h-line stride = 8-row group stride
        const uint32_t bsbo16 = 8u * row16;                     // B stages are dense
        // upper descriptor word: SBO | version | layout
        const uint32_t desc_hi = sbo16 | (1u << 14) | ((k16 ? 6u : 2u) << 29);
        const uint32_t bdesc_hi = bsbo16 | (1u << 14) | ((k16 ? 6u : 2u) << 29);
        // the group's operand blocks sit in n_a consecutive ring slots (with wrap-around).  Slot
        // arithmetic is kept branch-free: everything the MMA descriptors are built from must stay in
        // uniform registers (an R2UR per operand word in front of every UTCHMMA -- which is what a
        // conditional wait loop over the slots, or a __shfl_sync broadcast, produced -- costs a third
        // of the whole net).
        uint32_t a_blk1;
        {
          const uint32_t s1 = (sa + 1 == Cfg::kNAB) ? 0u : sa + 1;
          mbar_wait(&a_full[sa], pa);
          if (G.n_a == 2) mbar_wait(&a_full[s1], (sa + 1 == Cfg::kNAB) ? (pa ^ 1u) : pa);
          a_blk1 = ((smem_u32(a_smem + s1 * Cfg::kABlk) & 0x3FFFFu) >> 4) | (1u << 16);
        }
        tc_fence_after();
        const uint32_t a_blk0 = ((smem_u32(a_smem + sa * Cfg::kABlk) & 0x3FFFFu) >> 4) | (1u << 16);
        uint32_t op_a[3], op_b[3], op_d[3], op_i[3];
#pragma unroll
        for (int o = 0; o < 3; ++o) {
          op_a[o] = (G.ops[o].a_off & 1) ? a_blk1 : a_blk0;
          op_b[o] = G.ops[o].b_off;
          op_d[o] = G.ops[o].d_col;
          op_i[o] = idesc_base | (static_cast<uint32_t>(G.ops[o].n8) << 17);
        }
        const int tps = G.tps;
        // one tap's rows inside a multi-tap stage, in 16-byte units (lo stages carry lo_rows per tap)
        const uint32_t tap16 = static_cast<uint32_t>(G.tap_rows) * row16;
        if (k16 && tps == 3) {
          // 16-channel rows: one MMA per tap, so the per-tap wait / elect / commit round trip (about 400
          // cycles of dependent uniform instructions) was the limiter of the first layer; issue the three
          // taps of a stage from one elected section
          for (int j = 0; j < ntaps; j += 3) {
            mbar_wait(&b_full[sb], pb);
            tc_fence_after();
            const uint32_t b_base = ((smem_u32(b_smem + sb * Cfg::kBStage) & 0x3FFFFu) >> 4) | (1u << 16);
            if (elect_one()) {
              if (!dead) {
                const uint32_t d_tile = tmem_u + (buf * TM + my_tile) * DC;
#pragma unroll
                for (int jt = 0; jt < 3; ++jt) {
                  const uint32_t jj = static_cast<uint32_t>(j + jt);
                  const uint32_t kw = ntaps == 9 ? jj / 3u : 0u;
                  const uint32_t kh = ntaps == 9 ? jj % 3u : jj;
                  const uint32_t a_row = (static_cast<uint32_t>(my_tile * 16) + kh) * sbo16 + kw * row16;
                  const uint32_t a_hi = desc_hi | (kBaseOffsetMode ? (((a_blk0 + a_row) >> 3) & 7u) << 17 : 0u);
#pragma unroll
                  for (int o = 0; o < 3; ++o)
                    if (o < n_ops)
                      mma(d_tile + op_d[o], (static_cast<uint64_t>(a_hi) << 32) | (op_a[o] + a_row),
                          (static_cast<uint64_t>(bdesc_hi) << 32) | (b_base + jt * tap16 + op_b[o]), op_i[o], 1u);
                }
              }
              commit(&b_empty[sb]);
            }
            __syncwarp();
            if (++sb == Cfg::kNB) { sb = 0; pb ^= 1; }
          }
        } else if (!k16 && ntaps == 9 && kBlockIssue && EARLY != 0) {     // the big 3^3 velocity instances only
          // 64-channel rows, 9 taps (kw-major): one elected section per kw-block of three kh taps.  The per-tap
          // round trip of the generic loop below (wait, elect, reconverge, ring arithmetic: ~400 cycles of dependent
          // uniform instructions) is as long as the 8 MMAs it brackets take on the tensor pipe once the folded
          // tangent has removed a fifth of them: the issuers, not the pipe, set the pace.  Here the elected lane waits
          // for each tap's weight stage itself and hands the stages back one by one, so the ring behaves as before.
          // Same MMAs into every accumulator in the same order: bit-identical.
          const uint64_t a_hi64 = static_cast<uint64_t>(desc_hi) << 32;
          const uint64_t b_hi64 = static_cast<uint64_t>(bdesc_hi) << 32;
          for (int jb = 0; jb < 3; ++jb) {
            uint32_t chain_p = 0u, d_sel = op_d[0];
            bool chain_end = false;
            if constexpr (kChain) {
              if (G.chain != 0) {
                chain_p = (static_cast<uint32_t>(G.phase0) + static_cast<uint32_t>(jb)) & 1u;
                mbar_wait(chain_p ? y1_empty : y0_empty, ((pe >> chain_p) & 1u) ^ 1u);
                pe ^= (1u << chain_p);
                tc_fence_after();
                d_sel = chain_p ? (kFold ? static_cast<uint32_t>(DC / 4) : 0u) : op_d[0];
                chain_end = true;
              }
            }
            if (elect_one()) {
              uint32_t s = sb, p = pb;
#pragma unroll
              for (int kh = 0; kh < 3; ++kh) {
                if (tps == 1 || kh == 0) {
                  mbar_wait(&b_full[s], p);
                  tc_fence_after();
                }
                const uint32_t b_lo = (((smem_u32(b_smem + s * Cfg::kBStage) & 0x3FFFFu) >> 4) +
                                       (tps == 3 ? static_cast<uint32_t>(kh) * tap16 : 0u)) | (1u << 16);
#pragma unroll
                for (int t = 0; t < TM; ++t) {
                  if ((kIssuers == 2 && t != my_tile) || dead) continue;
                  const uint32_t d_tile = tmem_u + (buf * TM + t) * DC;
                  const uint32_t a_row = (static_cast<uint32_t>(t * 16 + kh)) * sbo16 + static_cast<uint32_t>(jb) * row16;
                  const uint64_t ad0 = a_hi64 | (op_a[0] + a_row), bd0 = b_hi64 | (b_lo + op_b[0]);
                  const uint64_t ad1 = a_hi64 | (op_a[1] + a_row), bd1 = b_hi64 | (b_lo + op_b[1]);
                  const uint64_t ad2 = a_hi64 | (op_a[2] + a_row), bd2 = b_hi64 | (b_lo + op_b[2]);
                  if (n_ops == 2) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                      mma(d_tile + d_sel, ad0 + 2u * k, bd0 + 2u * k, op_i[0], 1u);
                      mma(d_tile + op_d[1], ad1 + 2u * k, bd1 + 2u * k, op_i[1], 1u);
                    }
                  } else if (n_ops == 1) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) mma(d_tile + d_sel, ad0 + 2u * k, bd0 + 2u * k, op_i[0], 1u);
                  } else {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                      mma(d_tile + d_sel, ad0 + 2u * k, bd0 + 2u * k, op_i[0], 1u);
                      mma(d_tile + op_d[1], ad1 + 2u * k, bd1 + 2u * k, op_i[1], 1u);
                      mma(d_tile + op_d[2], ad2 + 2u * k, bd2 + 2u * k, op_i[2], 1u);
                    }
                  }
                }
                if (tps == 1 || kh == 2) {
                  commit(&b_empty[s]);
                  if (++s == Cfg::kNB) { s = 0; p ^= 1u; }
                }
              }
              if constexpr (kChain) {
                if (chain_end) commit(chain_p ? y1_full : y0_full);
              }
            }
            __syncwarp();
            {   // advance the ring by the stages consumed (the ring may be shorter than three stages)
              uint32_t adv = sb + (tps == 3 ? 1u : 3u);
              if constexpr (Cfg::kNB < 3) {
                if (adv >= Cfg::kNB) { adv -= Cfg::kNB; pb ^= 1u; }
              }
              pb ^= (adv >= Cfg::kNB) ? 1u : 0u;
              sb = (adv >= Cfg::kNB) ? adv - Cfg::kNB : adv;
            }
          }
        } else
        for (int j = 0, jt = 0; j < ntaps; ++j) {
          // EARLY == 3: block j / 3 of a chain group accumulates into y0 (phase 0) or y1 (phase 1); the
          // accumulator must have been drained since its previous chain
          uint32_t chain_p = 0u, d_sel = op_d[0];
          bool chain_end = false;
          if constexpr (kChain) {
            if (G.chain != 0) {
              chain_p = (static_cast<uint32_t>(G.phase0) + static_cast<uint32_t>(j) / 3u) & 1u;
              if (j % 3 == 0) {
                mbar_wait(chain_p ? y1_empty : y0_empty, ((pe >> chain_p) & 1u) ^ 1u);
                pe ^= (1u << chain_p);
                tc_fence_after();
              }
              d_sel = chain_p ? (kFold ? static_cast<uint32_t>(DC / 4) : 0u) : op_d[0];
              chain_end = (j % 3 == 2) || (j == ntaps - 1);
            }
          }
          if (jt == 0) {
            mbar_wait(&b_full[sb], pb);
            tc_fence_after();
          }
          const uint32_t b_lo =
              (((smem_u32(b_smem + sb * Cfg::kBStage) & 0x3FFFFu) >> 4) + static_cast<uint32_t>(jt) * tap16) | (1u << 16);
          if (elect_one()) {
#pragma unroll
            for (int t = 0; t < TM; ++t) {
              if ((kIssuers == 2 && t != my_tile) || dead) continue;
              const uint32_t d_tile = tmem_u + (buf * TM + t) * DC;
              const uint32_t kw = ntaps == 9 ? static_cast<uint32_t>(j) / 3u : 0u;
              const uint32_t kh = ntaps == 9 ? static_cast<uint32_t>(j) % 3u : static_cast<uint32_t>(j);
              const uint32_t a_row = (static_cast<uint32_t>(t * 16) + kh) * sbo16 + kw * row16;
              // kw taps start at a row that is not a multiple of the 8-row swizzle atom.  Measured on
              // B200: the MMA unit swizzles on absolute smem address bits (like TMA), so the plain
              // shifted start with base_offset = 0 is correct; setting base_offset = (addr>>7)&7 is
              // WRONG (NBE_BASE_OFFSET=1 reproduces that experiment).
              const uint32_t a_hi = desc_hi | (kBaseOffsetMode ? (((a_blk0 + a_row) >> 3) & 7u) << 17 : 0u);
              if (k16) {
#pragma unroll
                for (int o = 0; o < 3; ++o)
                  if (o < n_ops)
                    mma(d_tile + op_d[o], (static_cast<uint64_t>(a_hi) << 32) | (op_a[o] + a_row),
                        (static_cast<uint64_t>(bdesc_hi) << 32) | (b_lo + op_b[o]), op_i[o], 1u);
              } else {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
#pragma unroll
                  for (int o = 0; o < 3; ++o)
                    if (o < n_ops)
                      mma(d_tile + (o == 0 ? d_sel : op_d[o]),
                          (static_cast<uint64_t>(a_hi) << 32) | (op_a[o] + a_row + 2u * k),
                          (static_cast<uint64_t>(bdesc_hi) << 32) | (b_lo + op_b[o] + 2u * k), op_i[o], 1u);
                }
              }
            }
            if (jt == tps - 1) commit(&b_empty[sb]);
            if constexpr (kChain) {
              if (chain_end) commit(chain_p ? y1_full : y0_full);
            }
          }
          __syncwarp();
          if (++jt == tps) {
            jt = 0;
            if (++sb == Cfg::kNB) { sb = 0; pb ^= 1; }
          }
        }
        const int ps = EARLY != 0 ? G.post_sig : 0;
        if (elect_one()) {
          commit(&a_empty[sa]);
          if (G.n_a == 2) commit(&a_empty[sa + 1 == Cfg::kNAB ? 0u : sa + 1]);
          if (ps == 1) commit(y0_full);
          else if (ps == 2) commit(y1_full);
        }
        __syncwarp();
        {
          const uint32_t adv = sa + static_cast<uint32_t>(G.n_a);
          pa ^= (adv >= Cfg::kNAB) ? 1u : 0u;
          sa = (adv >= Cfg::kNAB) ? adv - Cfg::kNAB : adv;
        }
      }
      if (elect_one()) commit(&acc_full[buf]);
      __syncwarp();
      if (++buf == Cfg::kNBuf) { buf = 0; pacc ^= 1; }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------ epilogue (2 x 128 threads, thread <-> row)
    // two warp sets share the work of an item: with TM == 2 each set owns one tile, with
    // TM == 1 they take alternate 32-channel chunks
    const int q = warp & 3;                       // TMEM lane quarter of this warp
    const int r = q * 32 + lane;                  // accumulator row
    const int eg = (warp - 4) >> 2;               // epilogue warp set 0 / 1
    // launch fields are read once: the asm memory clobbers below would otherwise make every item
    // re-load them from global memory (a ~1 us dependent stall per item in the ncu source view)
    const int cout = L->cout;
    const bool vel = L->vel != 0;
    const bool act = L->act != 0;
    const bool acc3 = L->acc3 != 0;
    const bool fold_out = L->anext != nullptr;
    const int64_t out_sw = L->out_sw, out_sh = L->out_sh, out_sd = L->out_sd;
    const int64_t par_ow = L->par_ow, par_oh = L->par_oh, par_od = L->par_od;
    __half* const out_h_ptr = L->out_h_ptr;
    __half* const out_l_ptr = L->out_l_ptr;
    __half* const out_d_ptr = L->out_d_ptr;
    uint32_t buf = 0, pacc = 0;
    if constexpr (EARLY != 0) {
      // EARLY == 1: [y0 | dy | y1 | y2],  EARLY == 2: [y1 | dy | ylo | y0]  (x COUT columns per tile); this
      // thread owns row r of tile t and the two 32-channel chunks ch(0), ch(1) (TM == 1: the warp sets
      // take alternate chunks of the 128)
      constexpr int COUT = DC / 4;
      constexpr int cY0 = EARLY >= 2 ? 3 * COUT : 0;          // first primal accumulator to complete
      constexpr int cY1 = kFold ? COUT : (EARLY >= 2 ? 0 : 2 * COUT);   // second
      uint32_t pf = 0;                                        // chains: bit a = parity of the next wait on y<a>_full
      const int n_chains = L->n_chains;
      constexpr int cYZ = 3 * COUT;                           // last: y2, or y0 re-used by kd 2
      constexpr int cDY = kFold ? 0 : COUT;
      constexpr int cLO = 2 * COUT;                           // F192 / FOLD only: xh * Wl
      const int t = TM == 2 ? eg : 0;
      auto ch = [&](int i) { return TM == 1 ? eg * 32 + 64 * i : 32 * i; };
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + t * DC;
      auto arrive = [](uint64_t* bar) {
        if constexpr (PAIR) mbar_arrive_leader(bar); else mbar_arrive(bar);
      };
      for (long long it0 = item_first; it0 < n_items; it0 += gridDim.x) {
        const long long item = it0 + rank;
        const bool item_ok = item < n_items;
        int par, w0, h0, d0;
        decode(item_ok ? item : 0, par, w0, h0, d0);
        const bool dead = tile_dead(it0, eg);     // nothing is accumulated into a dead tile: its columns stay zero
        const int w = w0 + (r & 7);
        const int h = h0 + t * 16 + (r >> 3);
        const bool valid = item_ok && (w < out_w) && (h < out_h);
        uint32_t ps[64];
        if constexpr (kChain) {
          // accumulation chains: chain 0 (lo phase) comes back in y1, chain c >= 1 in y0 / y1 alternately
#pragma unroll
          for (int k = 0; k < 64; ++k) ps[k] = 0u;
          for (int cn = 0; cn < n_chains; ++cn) {
            const uint32_t a = cn == 0 ? 1u : (static_cast<uint32_t>(cn - 1) & 1u);
            mbar_wait(a ? y1_full : y0_full, (pf >> a) & 1u);
            pf ^= (1u << a);
            tc_fence_after();
            if (!dead) {
              // both chunks in flight at once: the drain is latency-bound and sits on the issuers' critical
              // path (the next chain into this accumulator waits for it)
              const uint32_t col = taddr + (a ? cY1 : cY0);
              uint32_t yv[64];
              tmem_ld32(col + ch(0), yv);
              tmem_ld32(col + ch(1), yv + 32);
              tmem_ld_wait();
              tmem_st32_zero(col + ch(0));
              tmem_st32_zero(col + ch(1));
              tmem_st_wait();
              tc_fence_before();
              arrive(a ? y1_empty : y0_empty);      // hand the columns back before the register adds
#pragma unroll
              for (int j = 0; j < 64; ++j) ps[j] = __float_as_uint(__uint_as_float(ps[j]) + __uint_as_float(yv[j]));
            } else {
              tc_fence_before();
              arrive(a ? y1_empty : y0_empty);
            }
          }
        } else {
        // ---- y0: complete once the kd = 0 groups are done (two more kd-planes of MMAs still to come)
        mbar_wait(y0_full, pacc);
        tc_fence_after();
        if (!dead) {
          tmem_ld32(taddr + cY0 + ch(0), ps);
          tmem_ld32(taddr + cY0 + ch(1), ps + 32);
          tmem_ld_wait();
          tmem_st32_zero(taddr + cY0 + ch(0));
          tmem_st32_zero(taddr + cY0 + ch(1));
          tmem_st_wait();
        }
        tc_fence_before();
        arrive(y0_empty);                         // EARLY 1: the next item's lo products may start; 2: kd 2 may start
        // ---- y1
        mbar_wait(y1_full, pacc);
        tc_fence_after();
        if (!dead) {
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            uint32_t y1[32];
            tmem_ld32(taddr + cY1 + ch(i), y1);
            tmem_ld_wait();
            tmem_st32_zero(taddr + cY1 + ch(i));
#pragma unroll
            for (int j = 0; j < 32; ++j)
              ps[32 * i + j] = __float_as_uint(__uint_as_float(ps[32 * i + j]) + __uint_as_float(y1[j]));
          }
        }
        if constexpr (EARLY == 2) {
          tmem_st_wait();
          tc_fence_before();
          arrive(y1_empty);                       // the next item's lo phase (xl * Wh -> y1) may start
        }
        }
        // ---- last primal accumulator, dy (and ylo): end of the item
        mbar_wait(&acc_full[0], pacc);
        tc_fence_after();
        if (!dead) {
          const int64_t voff = static_cast<int64_t>(d0) * out_sd + static_cast<int64_t>(h) * out_sh +
                               static_cast<int64_t>(w) * out_sw;
#pragma unroll
          for (int i = 0; i < 2; ++i) {
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
              const int c = ch(i) + 16 * hf;
              uint32_t y2[16], dy[16], yl[16];
              if constexpr (!kChain) tmem_ld16(taddr + cYZ + c, y2);
              tmem_ld16(taddr + cDY + c, dy);
              if constexpr (EARLY >= 2) tmem_ld16(taddr + cLO + c, yl);
              tmem_ld_wait();
              if constexpr (!kChain) tmem_st16_zero(taddr + cYZ + c);
              tmem_st16_zero(taddr + cDY + c);
              if constexpr (EARLY >= 2) tmem_st16_zero(taddr + cLO + c);
              uint32_t ph[8], pl[8], pd[8];
#pragma unroll
              for (int j = 0; j < 16; j += 2) {
                float s0 = __uint_as_float(ps[32 * i + 16 * hf + j]);
                float s1 = __uint_as_float(ps[32 * i + 16 * hf + j + 1]);
                if constexpr (!kChain) {
                  s0 += __uint_as_float(y2[j]);
                  s1 += __uint_as_float(y2[j + 1]);
                }
                if constexpr (EARLY >= 2) {
                  s0 += __uint_as_float(yl[j]);
                  s1 += __uint_as_float(yl[j + 1]);
                }
                float y0v = s0 * kInvWeightScale + bias_s[c + j];
                float y1v = s1 * kInvWeightScale + bias_s[c + j + 1];
                float d0v = __uint_as_float(dy[j]) * kInvWeightScale;
                float d1v = __uint_as_float(dy[j + 1]) * kInvWeightScale;
                if constexpr (kFold || EARLY == 1) {      // + beta_o * (x * W)_o: the demodulation part of x * dW (beta_s is zero unless folded)
                  d0v = fmaf(beta_s[c + j], s0 * kInvWeightScale, d0v);
                  d1v = fmaf(beta_s[c + j + 1], s1 * kInvWeightScale, d1v);
                }
                if (act) {
                  d0v = y0v > 0.f ? d0v : 0.01f * d0v;
                  d1v = y1v > 0.f ? d1v : 0.01f * d1v;
                  y0v = y0v >= 0.f ? y0v : 0.01f * y0v;
                  y1v = y1v >= 0.f ? y1v : 0.01f * y1v;
                }
                // the consumer's fold vector: it multiplies dx' = dy + a (.) y by W (anext_s is zero without one)
                d0v = fmaf(anext_s[c + j], y0v, d0v);
                d1v = fmaf(anext_s[c + j + 1], y1v, d1v);
                const __half2 hh = __floats2half2_rn(y0v, y1v);
                const float2 hf2 = __half22float2(hh);
                const __half2 ll = __floats2half2_rn(y0v - hf2.x, y1v - hf2.y);
                const __half2 dd = __floats2half2_rn(d0v, d1v);
                ph[j >> 1] = *reinterpret_cast<const uint32_t*>(&hh);
                pl[j >> 1] = *reinterpret_cast<const uint32_t*>(&ll);
                pd[j >> 1] = *reinterpret_cast<const uint32_t*>(&dd);
              }
              if (valid) {
                st_global_v8(out_h_ptr + voff + c, make_uint4(ph[0], ph[1], ph[2], ph[3]), make_uint4(ph[4], ph[5], ph[6], ph[7]));
                st_global_v8(out_l_ptr + voff + c, make_uint4(pl[0], pl[1], pl[2], pl[3]), make_uint4(pl[4], pl[5], pl[6], pl[7]));
                st_global_v8(out_d_ptr + voff + c, make_uint4(pd[0], pd[1], pd[2], pd[3]), make_uint4(pd[4], pd[5], pd[6], pd[7]));
              }
            }
          }
        }
        tmem_st_wait();
        tc_fence_before();
        arrive(&acc_empty[0]);
        pacc ^= 1;
      }
    } else
    for (long long it0 = item_first; it0 < n_items; it0 += gridDim.x) {
      const long long item = it0 + rank;
      const bool item_ok = item < n_items;
      int par, w0, h0, d0;
      decode(item_ok ? item : 0, par, w0, h0, d0);
      mbar_wait(&acc_full[buf], pacc);
      tc_fence_after();
      const bool dead = tile_dead(it0, eg);       // nothing was accumulated: the columns are still zero
#pragma unroll
      for (int t = 0; t < TM; ++t) {
        if ((TM == 2 && t != eg) || dead) continue;
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + (buf * TM + t) * DC;
        const int w = w0 + (r & 7);
        const int h = h0 + t * 16 + (r >> 3);
        const bool valid = item_ok && (w < out_w) && (h < out_h);
        if constexpr (FINAL) {
          if (TM == 1 && eg != 0) continue;
          // DC == 16: columns [net | dnet] (displacement-only: [xh*Wh + xl*Wh | xh*Wl]);
          // DC == 32 (folded tangent): [net | dnet | xh*Wl | -], dnet = (dx + a x)*Wh (+ the skip's residual x*dW_res)
          uint32_t v[DC >= 32 ? 32 : 16];
          if constexpr (DC >= 32) tmem_ld32(taddr, v); else tmem_ld16(taddr, v);
          tmem_ld_wait();
          if constexpr (DC >= 32) tmem_st32_zero(taddr); else tmem_st16_zero(taddr);
          if (valid) {
            const int64_t sidx = static_cast<int64_t>(fa.idx_d[d0]) * fa.src_sd +
                                 static_cast<int64_t>(fa.idx_h[h]) * fa.src_sh + fa.idx_w[w];
            const int64_t oidx = static_cast<int64_t>(d0) * fa.o_sd + static_cast<int64_t>(h) * fa.o_sh + w;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              const float x0 = scale_in_dtype(load_as_f32(fa.src, sidx + c * fa.src_sc, fa.src_dtype),
                                              fa.in_norm, fa.src_dtype);
              float net = __uint_as_float(v[c]) * kInvWeightScale + bias_s[c];
              if constexpr (DC >= 32) {
                const float prim = (__uint_as_float(v[c]) + __uint_as_float(v[16 + c])) * kInvWeightScale;
                net = prim + bias_s[c];
                const float dnet = fmaf(beta_s[c], prim, __uint_as_float(v[8 + c]) * kInvWeightScale);
                store_from_f32(fa.vel, oidx + c * fa.o_sc, fa.out_dtype,
                               round_to_dtype(dnet * fa.dx_norm + x0 * fa.x0_norm, fa.mid_dtype));
              } else if (fa.vel != nullptr) {
                const float dnet = __uint_as_float(v[8 + c]) * kInvWeightScale;
                store_from_f32(fa.vel, oidx + c * fa.o_sc, fa.out_dtype,
                               round_to_dtype(dnet * fa.dx_norm + x0 * fa.x0_norm, fa.mid_dtype));
              } else {
                net += __uint_as_float(v[8 + c]) * kInvWeightScale;     // displacement-only: columns 8..10 hold xh*Wl
              }
              store_from_f32(fa.disp, oidx + c * fa.o_sc, fa.out_dtype,
                             round_to_dtype((net + x0) * fa.six, fa.mid_dtype));
            }
          }
        } else {
          const int pc = par & 1, pb2 = (par >> 1) & 1, pa2 = (par >> 2) & 1;
          const int64_t voff = static_cast<int64_t>(d0) * out_sd + static_cast<int64_t>(h) * out_sh +
                               static_cast<int64_t>(w) * out_sw + pc * par_ow + pb2 * par_oh + pa2 * par_od;
          __half* oh = out_h_ptr + voff;
          __half* ol = out_l_ptr ? out_l_ptr + voff : nullptr;
          __half* od = out_d_ptr ? out_d_ptr + voff : nullptr;
          for (int c = (TM == 1 ? eg * 32 : 0); c < cout; c += (TM == 1 ? 64 : 32)) {
            uint32_t y[32], dy[32];
            tmem_ld32(taddr + c, y);
            if (vel) tmem_ld32(taddr + cout + c, dy);
            tmem_ld_wait();
            tmem_st32_zero(taddr + c);
            if (vel) tmem_st32_zero(taddr + cout + c);
            if (acc3) {
              // per-kd primal accumulators: each saw a third of the truncating accumulations at
              // ~1/sqrt(3) of the magnitude; combine them with round-to-nearest fp32 adds
              uint32_t y1[32];
              tmem_ld32(taddr + 2 * cout + c, y1);
              tmem_ld_wait();
              tmem_st32_zero(taddr + 2 * cout + c);
#pragma unroll
              for (int i = 0; i < 32; ++i) y[i] = __float_as_uint(__uint_as_float(y[i]) + __uint_as_float(y1[i]));
              tmem_ld32(taddr + 3 * cout + c, y1);
              tmem_ld_wait();
              tmem_st32_zero(taddr + 3 * cout + c);
#pragma unroll
              for (int i = 0; i < 32; ++i) y[i] = __float_as_uint(__uint_as_float(y[i]) + __uint_as_float(y1[i]));
            }
            uint32_t ph[16], pl[16], pd[16];
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              float y0 = __uint_as_float(y[i]) * kInvWeightScale + bias_s[c + i];
              float y1 = __uint_as_float(y[i + 1]) * kInvWeightScale + bias_s[c + i + 1];
              float d0v = vel ? __uint_as_float(dy[i]) * kInvWeightScale : 0.f;
              float d1v = vel ? __uint_as_float(dy[i + 1]) * kInvWeightScale : 0.f;
              if (act) {
                d0v = y0 > 0.f ? d0v : 0.01f * d0v;
                d1v = y1 > 0.f ? d1v : 0.01f * d1v;
                y0 = y0 >= 0.f ? y0 : 0.01f * y0;
                y1 = y1 >= 0.f ? y1 : 0.01f * y1;
              }
              if (fold_out) {                               // the consumer's fold vector
                d0v = fmaf(anext_s[c + i], y0, d0v);
                d1v = fmaf(anext_s[c + i + 1], y1, d1v);
              }
              const __half2 hh = __floats2half2_rn(y0, y1);
              const float2 hf = __half22float2(hh);
              const __half2 ll = __floats2half2_rn(y0 - hf.x, y1 - hf.y);
              const __half2 dd = __floats2half2_rn(d0v, d1v);
              ph[i >> 1] = *reinterpret_cast<const uint32_t*>(&hh);
              pl[i >> 1] = *reinterpret_cast<const uint32_t*>(&ll);
              pd[i >> 1] = *reinterpret_cast<const uint32_t*>(&dd);
            }
            if (valid) {
#pragma unroll
              for (int v8 = 0; v8 < 2; ++v8) {
                const int j = 8 * v8;
                st_global_v8(oh + c + v8 * 16, make_uint4(ph[j], ph[j + 1], ph[j + 2], ph[j + 3]),
                             make_uint4(ph[j + 4], ph[j + 5], ph[j + 6], ph[j + 7]));
                if (ol) st_global_v8(ol + c + v8 * 16, make_uint4(pl[j], pl[j + 1], pl[j + 2], pl[j + 3]),
                                     make_uint4(pl[j + 4], pl[j + 5], pl[j + 6], pl[j + 7]));
                if (od) st_global_v8(od + c + v8 * 16, make_uint4(pd[j], pd[j + 1], pd[j + 2], pd[j + 3]),
                                     make_uint4(pd[j + 4], pd[j + 5], pd[j + 6], pd[j + 7]));
              }
            }
          }
        }
      }
      tmem_st_wait();
      tc_fence_before();
      if constexpr (PAIR) mbar_arrive_leader(&acc_empty[buf]); else mbar_arrive(&acc_empty[buf]);
      if (++buf == Cfg::kNBuf) { buf = 0; pacc ^= 1; }
    }
  }

  tc_fence_before();
  if constexpr (PAIR) cluster_sync_all(); else __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    if constexpr (PAIR) tmem_dealloc_2sm<Cfg::kTmemCols>(tmem_base);
    else tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
}

}  // namespace nbe
