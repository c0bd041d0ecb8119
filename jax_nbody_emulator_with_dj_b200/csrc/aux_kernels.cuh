// HBM-bound helper kernels: fused weight modulation / demodulation / Dz-tangent / operand
// packing, and the periodic gather + scale + channel-pack of the input subbox.
#pragma once
#include "conv_mma.cuh"

namespace nbe {

// ---------------------------------------------------------------------------------------
// Weight modulation  (style_layers_vel.py:62-105; nbody_emulator.py:131-148, :189-219)
//
//   m_i   = s0*SW[i,0] + s1*SW[i,1] + sb[i]            w = W*m        dws = W*SW[i,1]
//   n_o   = sqrt(sum_{i,t} w^2 + eps)                  dn_o = -sum(w*dws)/n^3
//   Wn    = w/n                                        dWn  = dws/n + w*dn  (+ Wn/Dz on the
//                                                      layers fed by the raw Dz-scaled field)
//
// One block per (layer, output channel, sample).  Results are written twice: fp32 OIDHW
// (for nbe_get_modulated) and as fp16 tensor-core operand rows (hi, lo = fp16(Wn - hi), and
// tangent), placed by per-layer emit rules into the B tensors of the conv launches.
// ---------------------------------------------------------------------------------------
enum { EMIT_WH = 0, EMIT_WL = 1, EMIT_DW = 2 };

struct EmitRule {
  int8_t what;        // EMIT_*
  int8_t kind;        // tile-kind offset (x kind_stride tiles)
  int16_t row_base;   // row of output channel 0 inside the B stage
  int16_t kcol;       // column offset (16-channel rows only)
  int16_t alt_kd1;    // added to row_base for the taps with kd == 1 of a 3^3 conv (see acc3 layout)
  // CTA-pair layout: a stage holds [CTA0 rows | CTA1 rows] (pair_rows each); output channel o of
  // this rule lands in CTA (o / mod + cta_base) at row row_base + o % mod.  mod == 0: no split.
  int16_t mod;
  int8_t cta_base;
  int8_t kd_mask;     // bit kd set: the rule applies to taps of that kd-plane (0: all)
  // explicit form (o_max > 0): output channels [o_min, o_max) go to CTA `cta_base`, rows row_base + (o - o_min)
  int16_t o_min, o_max;
};
constexpr int kMaxRules = 16;

struct LayerMeta {
  const float* W;        // (cout, cin, k3) raw weight, or premodulated weight
  const float* dW;       // premodulated dweight or nullptr
  const float* SW;       // (cin, 2) or nullptr
  const float* sb;       // (cin)
  float* w32;            // [B][cout*cin*k3] modulated fp32 out
  float* dw32;           // idem (vel) or nullptr
  __half* dst;           // packed operand buffer of the owning launch (sample 0)
  long long dst_sample_stride;   // halves between samples
  int cout, cin, k3;
  int first;             // add Wn/Dz to the tangent
  int premod;
  int vel;
  int kc16;              // destination rows are 16 channels wide
  int nrs;               // rows per B stage
  int kc_stride, kind_stride;    // tiles
  int n_rules;
  int row0;              // first block index (prefix sum of cout)
  int tap_tile[27];
  EmitRule rules[kMaxRules];
  int pair_rows;         // rows per CTA in a pair-layout stage
  // accumulation chains (conv_mma.cuh, EARLY == 3): kd_mask of a rule selects by the PHASE of the 3-tap
  // block a tap belongs to, phase = (chain_par0 + kd * chain_nkc + kc + kw) & 1
  int chain, chain_par0, chain_nkc;
  // Tangent folding (conv_mma.cuh, ConvLaunch::beta / anext).  With m_i the modulation of input channel i,
  //   dWn[o,i,t] = Wn[o,i,t] * (a_i + beta_o),   a_i = SW[i,1] / m_i,   beta_o = n_o * dn_o = -sum(w*dws) / n_o^2.
  // The tangent rows this layer EMITS are  dWn - (f_i + bl_o) * Wn,  where f is the fold vector already added
  // to the stored tangent of the tensor it reads (the a of that tensor's fold consumer; zero if none) and bl
  // the beta the owning launch applies in its epilogue (the beta of the launch's main conv; zero for launches
  // that are not FOLD instances).  For the main conv of a FOLD launch that is identically zero: no tangent rows.
  const float* fold_SW;     // style params of the layer that defines f of this layer's source tensor (or nullptr)
  const float* fold_sb;
  const float* fold_a;      // premodulated weights: f itself (device, [cin]) or nullptr
  int beta_layer;           // meta index of the layer whose beta the owning launch applies (-1: none)
  const float* pre_a;       // premodulated weights: this layer's own a (cin) / beta (cout) from the host factorisation
  const float* pre_beta;
  float* beta_out;          // main conv of a FOLD launch: beta goes here ([sample][fold_stride])
  float* a_out[2];          // ... and a[a_out_off + j], j < a_out_n, here (the anext of the launch producing a source tensor)
  int a_out_off[2], a_out_n[2], n_a_out;
  int fold_stride;          // floats between samples in beta_out / a_out
};

// sum over (i, t) of w^2 and w * dws for output row o of layer M: every thread returns the block-wide sums
__device__ __forceinline__ void modulate_row_sums(const LayerMeta& M, int o, float s0, float s1, float* red1,
                                                  float* red2, float& a1, float& a2) {
  const int ne = M.cin * M.k3;
  const float* Wrow = M.W + static_cast<long long>(o) * ne;
  a1 = 0.f; a2 = 0.f;
  for (int e = threadIdx.x; e < ne; e += blockDim.x) {
    const int i = e / M.k3;
    const float m = s0 * M.SW[2 * i] + s1 * M.SW[2 * i + 1] + M.sb[i];
    const float w = Wrow[e] * m;
    const float dws = Wrow[e] * M.SW[2 * i + 1];
    a1 += w * w;
    a2 += w * dws;
  }
  for (int off = 16; off > 0; off >>= 1) {
    a1 += __shfl_xor_sync(0xffffffffu, a1, off);
    a2 += __shfl_xor_sync(0xffffffffu, a2, off);
  }
  __syncthreads();          // red1 / red2 may still be read from a previous call
  if ((threadIdx.x & 31) == 0) { red1[threadIdx.x >> 5] = a1; red2[threadIdx.x >> 5] = a2; }
  __syncthreads();
  a1 = red1[0] + red1[1] + red1[2] + red1[3];
  a2 = red2[0] + red2[1] + red2[2] + red2[3];
}

__global__ void __launch_bounds__(128)
modulate_kernel(const LayerMeta* __restrict__ metas, int n_layers, const float* __restrict__ s0a,
                const float* __restrict__ s1a, float eps) {
  __shared__ float red1[4], red2[4];
  __shared__ int s_layer;
  if (threadIdx.x == 0) {
    int l = 0;
    while (l + 1 < n_layers && metas[l + 1].row0 <= static_cast<int>(blockIdx.x)) ++l;
    s_layer = l;
  }
  __syncthreads();
  const LayerMeta& M = metas[s_layer];
  const int o = blockIdx.x - M.row0;
  const int b = blockIdx.y;
  const int ne = M.cin * M.k3;
  const float s0 = s0a[b], s1 = s1a[b];
  const float* Wrow = M.W + static_cast<long long>(o) * ne;

  float inv_n = 1.f, dn = 0.f;
  float beta_own = 0.f;       // this layer's own beta_o
  if (!M.premod) {
    float a1, a2;
    modulate_row_sums(M, o, s0, s1, red1, red2, a1, a2);
    const float n = sqrtf(a1 + eps);
    inv_n = 1.f / n;
    dn = -a2 / (n * n * n);
    beta_own = -a2 / (a1 + eps);
  } else if (M.pre_beta != nullptr) {
    beta_own = M.pre_beta[o];
  }
  // beta applied by the owning launch's epilogue (row o of its main conv)
  float bl = 0.f;
  if (M.vel && M.beta_layer >= 0) {
    const LayerMeta& MB = metas[M.beta_layer];
    if (!M.premod) {
      float a1, a2;
      modulate_row_sums(MB, o, s0, s1, red1, red2, a1, a2);
      bl = -a2 / (a1 + eps);
    } else {
      bl = MB.pre_beta[o];
    }
  }
  if (M.vel && M.beta_out != nullptr && threadIdx.x == 0) M.beta_out[static_cast<long long>(b) * M.fold_stride + o] = beta_own;
  if (M.vel && o == 0)
    for (int t = 0; t < M.n_a_out; ++t)
      for (int j = threadIdx.x; j < M.a_out_n[t]; j += blockDim.x) {
        const int i = j + M.a_out_off[t];
        M.a_out[t][static_cast<long long>(b) * M.fold_stride + j] =
            M.premod ? M.pre_a[i] : M.SW[2 * i + 1] / (s0 * M.SW[2 * i] + s1 * M.SW[2 * i + 1] + M.sb[i]);
      }
  const float inv_Dz = 1.f / (s1 + 1.f);
  const long long sample_off = static_cast<long long>(b) * M.dst_sample_stride;
  const int rowlen = M.kc16 ? 16 : 64;
  // emission loop: input channel fastest, so that consecutive threads write consecutive
  // columns of one operand row (coalesced 2-byte stores); the weight row itself is L1-resident
  for (int e2 = threadIdx.x; e2 < ne; e2 += blockDim.x) {
    const int tap = e2 / M.cin;
    const int i = e2 - tap * M.cin;
    const int e = i * M.k3 + tap;
    float wn, dwn = 0.f;
    if (M.premod) {
      wn = Wrow[e];
      if (M.vel) dwn = M.dW[static_cast<long long>(o) * ne + e];
    } else {
      const float m = s0 * M.SW[2 * i] + s1 * M.SW[2 * i + 1] + M.sb[i];
      const float w = Wrow[e] * m;
      wn = w * inv_n;
      if (M.vel) {
        const float dws = Wrow[e] * M.SW[2 * i + 1];
        dwn = dws * inv_n + w * dn;
        if (M.first) dwn += wn * inv_Dz;
      }
    }
    const long long oidx = static_cast<long long>(b) * M.cout * ne + static_cast<long long>(o) * ne + e;
    M.w32[oidx] = wn;
    if (M.dw32) M.dw32[oidx] = dwn;
    if (M.vel) {              // what the stored tangent of the source / the launch's epilogue already carry
      float f = bl;
      if (M.fold_SW != nullptr) f += M.fold_SW[2 * i + 1] / (s0 * M.fold_SW[2 * i] + s1 * M.fold_SW[2 * i + 1] + M.fold_sb[i]);
      else if (M.fold_a != nullptr) f += M.fold_a[i];
      dwn = fmaf(-f, wn, dwn);
    }
    // operands are packed scaled by kWeightScale (a power of two, undone exactly in the conv
    // epilogue) so that lo = W - hi(W) ~ 2^-12 |W| stays a NORMAL fp16 number: unscaled, the lo
    // part of a typical demodulated weight (~0.02) is subnormal and keeps only ~5 bits.
    const float ws = wn * kWeightScale, dws_ = dwn * kWeightScale;
    const __half wh = __float2half_rn(ws);
    const __half wl = __float2half_rn(ws - __half2float(wh));
    const __half dh = __float2half_rn(dws_);
    const int kc = M.kc16 ? 0 : (i >> 6);
    for (int r = 0; r < M.n_rules; ++r) {
      const EmitRule R = M.rules[r];
      const __half v = R.what == EMIT_WH ? wh : (R.what == EMIT_WL ? wl : dh);
      const long long tile = M.tap_tile[tap] + kc * M.kc_stride + R.kind * M.kind_stride;
      const int col = M.kc16 ? (R.kcol + i) : (i & 63);
      const int kd = M.k3 == 27 ? tap / 9 : 0;
      const int sel = M.chain ? ((M.chain_par0 + kd * M.chain_nkc + kc + (M.k3 == 27 ? tap % 3 : 0)) & 1) : kd;
      if (R.kd_mask && !((R.kd_mask >> sel) & 1)) continue;
      if (R.o_max > 0 && (o < R.o_min || o >= R.o_max)) continue;
      const int row = R.o_max > 0 ? (R.cta_base * M.pair_rows + R.row_base + (o - R.o_min))
                      : R.mod ? ((o / R.mod + R.cta_base) * M.pair_rows + R.row_base + o % R.mod)
                              : (R.row_base + (kd == 1 ? R.alt_kd1 : 0) + o);
      M.dst[sample_off + (tile * M.nrs + row) * rowlen + col] = v;
    }
  }
}

// ---------------------------------------------------------------------------------------
// Input pack: periodic gather of the padded subbox with the reference's integer tables
// (subbox.py:81-97), scale by Dz/6 (style_nbody_emulator_vel_core.py:132-134), split into
// fp16 hi/lo and write 16-channel NDHWC records
//     [xh0 xh1 xh2 | xl0 xl1 xl2 | xh0 xh1 xh2 | 0 x 7]
// so that a single K=16 MMA against rows [Wh | Wh | Wl | 0] yields the 3-product primal.
// ---------------------------------------------------------------------------------------
struct PackArgs {
  const void* src;
  int32_t src_dtype;
  int64_t src_sc, src_sd, src_sh;
  const int32_t* idx_d;
  const int32_t* idx_h;
  const int32_t* idx_w;
  int32_t n0, n1, n2;
  float in_norm;
  __half* out;
};

__global__ void __launch_bounds__(256) pack_input_kernel(const PackArgs a) {
  const int64_t nvox = static_cast<int64_t>(a.n0) * a.n1 * a.n2;
  for (int64_t v = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; v < nvox;
       v += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int w = static_cast<int>(v % a.n2);
    const int64_t t = v / a.n2;
    const int h = static_cast<int>(t % a.n1);
    const int d = static_cast<int>(t / a.n1);
    const int64_t sidx = static_cast<int64_t>(a.idx_d[d]) * a.src_sd + static_cast<int64_t>(a.idx_h[h]) * a.src_sh +
                         a.idx_w[w];
    __half xh[3], xl[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float x;
      if (a.src_dtype == 0) x = reinterpret_cast<const float*>(a.src)[sidx + c * a.src_sc];
      else if (a.src_dtype == 1) x = __half2float(reinterpret_cast<const __half*>(a.src)[sidx + c * a.src_sc]);
      else x = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(a.src)[sidx + c * a.src_sc]);
      x = scale_in_dtype(x, a.in_norm, a.src_dtype);
      xh[c] = __float2half_rn(x);
      xl[c] = __float2half_rn(x - __half2float(xh[c]));
    }
    const __half z = __float2half_rn(0.f);
    __half rec[16] = {xh[0], xh[1], xh[2], xl[0], xl[1], xl[2], xh[0], xh[1], xh[2], z, z, z, z, z, z, z};
    uint4* dst = reinterpret_cast<uint4*>(a.out + v * 16);
    dst[0] = *reinterpret_cast<const uint4*>(&rec[0]);
    dst[1] = *reinterpret_cast<const uint4*>(&rec[8]);
  }
}

}  // namespace nbe
