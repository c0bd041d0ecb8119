// Displacement -> density contrast -> P(k) bins: the step right after the emulator
// (scripts/core.py:396-409, 446-458 call DISCO-DJ's get_delta_from_psi; scripts/utils.py:1083-1090
// call Pylians' PKL.Pk).  Scatter / reduction kernels only: HBM- and L2-atomic-bound byte work,
// grids sized in multiples of the SM count; the FFT between them is the caller's (cuFFT).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace nbe {

// B-spline weights of order P at position x (mesh units); i0 = first cell
template <int P>
__device__ __forceinline__ void mas_weights(float x, int& i0, float (&w)[P]) {
  if constexpr (P == 1) {
    i0 = __float2int_rd(x + 0.5f);
    w[0] = 1.f;
  } else if constexpr (P == 2) {
    i0 = __float2int_rd(x);
    const float d = x - static_cast<float>(i0);
    w[0] = 1.f - d; w[1] = d;
  } else if constexpr (P == 3) {
    const int ic = __float2int_rd(x + 0.5f);
    const float d = x - static_cast<float>(ic);
    i0 = ic - 1;
    w[0] = 0.5f * (0.5f - d) * (0.5f - d); w[1] = 0.75f - d * d; w[2] = 0.5f * (0.5f + d) * (0.5f + d);
  } else {
    const int ib = __float2int_rd(x);
    const float d = x - static_cast<float>(ib), e = 1.f - d;
    i0 = ib - 1;
    w[0] = e * e * e * (1.f / 6.f);
    w[1] = (4.f - 6.f * d * d + 3.f * d * d * d) * (1.f / 6.f);
    w[2] = (4.f - 6.f * e * e + 3.f * e * e * e) * (1.f / 6.f);
    w[3] = d * d * d * (1.f / 6.f);
  }
}

__device__ __forceinline__ int wrap(int i, int n) {
  i %= n;
  return i < 0 ? i + n : i;
}

// One thread per particle of the (n0, n1, n2) lattice, w fastest: neighbouring threads hit
// neighbouring cells, so the fp32 reductions mostly merge in L2.
template <int P>
__global__ void __launch_bounds__(256) paint_kernel(const float* __restrict__ psi, int n0, int n1, int n2, int res,
                                                    float scale /* res / boxsize */, float* __restrict__ rho) {
  const long long np = 1ll * n0 * n1 * n2;
  const float s0 = static_cast<float>(res) / n0, s1 = static_cast<float>(res) / n1, s2 = static_cast<float>(res) / n2;
  for (long long t = blockIdx.x * 256ll + threadIdx.x; t < np; t += 256ll * gridDim.x) {
    const int k = static_cast<int>(t % n2);
    const int j = static_cast<int>((t / n2) % n1);
    const int i = static_cast<int>(t / (1ll * n1 * n2));
    int a0, b0, c0;
    float wa[P], wb[P], wc[P];
    mas_weights<P>(i * s0 + psi[t] * scale, a0, wa);
    mas_weights<P>(j * s1 + psi[np + t] * scale, b0, wb);
    mas_weights<P>(k * s2 + psi[2 * np + t] * scale, c0, wc);
#pragma unroll
    for (int a = 0; a < P; ++a) {
      const long long ia = 1ll * wrap(a0 + a, res) * res;
#pragma unroll
      for (int b = 0; b < P; ++b) {
        const long long ib = (ia + wrap(b0 + b, res)) * res;
        const float wab = wa[a] * wb[b];
#pragma unroll
        for (int c = 0; c < P; ++c) atomicAdd(rho + ib + wrap(c0 + c, res), wab * wc[c]);
      }
    }
  }
}

__global__ void __launch_bounds__(256) rho_to_delta_kernel(float* __restrict__ rho, long long n, float norm) {
  for (long long t = blockIdx.x * 256ll + threadIdx.x; t < n; t += 256ll * gridDim.x) rho[t] = rho[t] * norm - 1.f;
}

__device__ __forceinline__ float sinc_pi(float x) {        // sin(pi x) / (pi x)
  return x == 0.f ? 1.f : sinpif(x) / (3.14159265358979f * x);
}
__device__ __forceinline__ float mas_window(int kx, int ky, int kz, int res, int order) {
  const float inv = 1.f / res;
  const float w = sinc_pi(kx * inv) * sinc_pi(ky * inv) * sinc_pi(kz * inv);
  float r = 1.f;
  for (int p = 0; p < order; ++p) r *= w;
  return r;
}

// delta_k /= W(k)^order on the half-complex (res, res, res/2+1) cube (float2 per mode)
__global__ void __launch_bounds__(256) mas_deconvolve_kernel(float2* __restrict__ dk, int res, int order) {
  const int nz = res / 2 + 1;
  const long long n = 1ll * res * res * nz;
  for (long long t = blockIdx.x * 256ll + threadIdx.x; t < n; t += 256ll * gridDim.x) {
    const int kz = static_cast<int>(t % nz);
    int ky = static_cast<int>((t / nz) % res), kx = static_cast<int>(t / (1ll * nz * res));
    if (kx > res / 2) kx -= res;
    if (ky > res / 2) ky -= res;
    const float inv = 1.f / mas_window(kx, ky, kz, res, order);
    float2 v = dk[t];
    v.x *= inv; v.y *= inv;
    dk[t] = v;
  }
}

// Shell sums of |delta_k / W^order|^2, |k| and the mode count over the independent modes of the
// half-complex cube; bins = floor(|k|) in units of k_F.  Per-block shared histograms in double,
// one global atomic per bin and block.
constexpr int kPkMaxBins = 2048;
__global__ void __launch_bounds__(256) pk_bins_kernel(const float2* __restrict__ dk, int res, int order, int nbins,
                                                      double* __restrict__ out /* [3][nbins]: P, k, N */) {
  extern __shared__ double sh[];                 // 3 * nbins
  for (int i = threadIdx.x; i < 3 * nbins; i += 256) sh[i] = 0.0;
  __syncthreads();
  const int nz = res / 2 + 1, mid = res / 2;
  const bool even = (res % 2) == 0;
  const long long n = 1ll * res * res * nz;
  for (long long t = blockIdx.x * 256ll + threadIdx.x; t < n; t += 256ll * gridDim.x) {
    const int kz = static_cast<int>(t % nz);
    int ky = static_cast<int>((t / nz) % res), kx = static_cast<int>(t / (1ll * nz * res));
    if (kx > mid) kx -= res;
    if (ky > mid) ky -= res;
    // the kz = 0 and kz = Nyquist planes hold every mode and its conjugate: keep one of each pair
    const bool plane = kz == 0 || (even && kz == mid);
    const int sx = (even && kx == mid) ? 0 : kx, sy = (even && ky == mid) ? 0 : ky;   // self-conjugate coordinates
    if (plane && (sx < 0 || (sx == 0 && sy < 0))) continue;
    const float k2 = static_cast<float>(kx * kx + ky * ky + kz * kz);
    if (k2 == 0.f) continue;
    const double kmod = sqrt(static_cast<double>(k2));
    const int bin = static_cast<int>(kmod);
    if (bin >= nbins) continue;
    const float2 v = dk[t];
    double p = static_cast<double>(v.x) * v.x + static_cast<double>(v.y) * v.y;
    if (order > 0) {
      const double w = mas_window(kx, ky, kz, res, order);
      p /= w * w;
    }
    atomicAdd(&sh[bin], p);
    atomicAdd(&sh[nbins + bin], kmod);
    atomicAdd(&sh[2 * nbins + bin], 1.0);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 3 * nbins; i += 256)
    if (sh[i] != 0.0) atomicAdd(&out[i], sh[i]);
}

// 1LPT (Zel'dovich) displacement in k space: psi_j(k) = i k_j / k^2 * delta(k), k in units of
// k_F = 2 pi / boxsize; the k = 0 mode and the Nyquist component of each derivative are zeroed so
// that the inverse real FFT is exact (scripts/core.py:396-397: dj.with_lpt(n_order=1) +
// evaluate_lpt_psi_at_a).  out: three half-complex cubes, back to back.
__global__ void __launch_bounds__(256) za_psi_k_kernel(const float2* __restrict__ dk, int res, float inv_kf,
                                                        float2* __restrict__ out) {
  const int nz = res / 2 + 1, mid = res / 2;
  const bool even = (res % 2) == 0;
  const long long n = 1ll * res * res * nz;
  for (long long t = blockIdx.x * 256ll + threadIdx.x; t < n; t += 256ll * gridDim.x) {
    const int kz = static_cast<int>(t % nz);
    int ky = static_cast<int>((t / nz) % res), kx = static_cast<int>(t / (1ll * nz * res));
    if (kx > mid) kx -= res;
    if (ky > mid) ky -= res;
    const float k2 = static_cast<float>(kx * kx + ky * ky + kz * kz);
    const float2 v = dk[t];
    const float s = k2 > 0.f ? inv_kf / k2 : 0.f;
    const int kk[3] = {(even && kx == mid) ? 0 : kx, (even && ky == mid) ? 0 : ky, (even && kz == mid) ? 0 : kz};
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const float f = s * kk[j];
      out[j * n + t] = make_float2(-f * v.y, f * v.x);       // i * f * (x + i y)
    }
  }
}

}  // namespace nbe
