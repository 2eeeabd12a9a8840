// pbk_blue.cuh -- arbitrary transform lengths through Bluestein's chirp-z identity on top of the
// power-of-two tile-FFT passes.  The reference accepts any length (scipy.fft / pocketfft,
// fft.py:34; its tests use 4224, 4233 and nperseg = 33), so lengths that are not powers of two are
// handled here instead of being refused:
//     X[k] = w[k] * sum_j (x[j] w[j]) conj(w)[k - j],        w[j] = exp(-i pi j^2 / n)
// i.e. multiply by w, convolve with conj(w) (circular convolution of length M >= 2n-1, M = 2^m,
// done with two M-point FFTs and the precomputed spectrum of the filter), multiply by w.  The
// inverse uses conj(w) everywhere.  j^2 mod 2n is formed in 64-bit integers, so the chirp phases
// are exact.  In coherent dedispersion the post-multiply of the forward transform and the
// pre-multiply of the inverse cancel (w conj(w) = 1): the middle step is just "times H[k], zero
// the padding".  These are plain elementwise kernels: this path is for generality, the
// power-of-two path is the optimised one.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "pbk_fft.cuh"

namespace pbk {

// element offset of (o, index, lane) in a user array: o*eo + index*ei + (lane / P)*ec + (lane % P)*ep
struct BlueMap {
  long long eo, ei, ec, ep;
};

__device__ __forceinline__ float2 cmulf(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// w[j] = exp(-i pi j^2 / n) for j < n
__global__ void __launch_bounds__(256) blue_chirp_kernel(float2* __restrict__ w, long long n) {
  for (long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x; j < n;
       j += (long long)gridDim.x * blockDim.x) {
    const unsigned long long r = ((unsigned long long)j * (unsigned long long)j) %
                                 (2ull * (unsigned long long)n);
    const double x = (double)r / (double)n;   // in [0, 2): phase in half-cycles
    double s, c;
    sincospi(x, &s, &c);
    w[j] = make_float2((float)c, (float)(-s));
  }
}

// filter b[j] = conj(w[|j|]) for |j| < n (indices mod M), 0 elsewhere
__global__ void __launch_bounds__(256) blue_filter_kernel(const float2* __restrict__ w,
                                                          float2* __restrict__ b, long long n,
                                                          long long M) {
  for (long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x; j < M;
       j += (long long)gridDim.x * blockDim.x) {
    float2 v = make_float2(0.f, 0.f);
    if (j < n) v = make_float2(w[j].x, -w[j].y);
    else if (M - j < n) v = make_float2(w[M - j].x, -w[M - j].y);
    b[j] = v;
  }
}

struct BlueIO {
  BlueMap map;
  long long O, n, M, I;
  int P;
  int kind;          // input: LOAD_C64 / LOAD_I8X2 / LOAD_F32 ; output: EPI_C64 only
  int conj_w;        // use conj(w)
  long long shift;   // index rotation: user index = (j + shift) mod n   (fftshift bookkeeping)
  float scale;
  long long lo, hi;  // output: keep lo <= k < hi (crop); rows are written at k - lo
};

// A[o, j, i] = x[o, j', i] * w[j] (j < n), 0 for the padding
__global__ void __launch_bounds__(256) blue_pre_kernel(const void* __restrict__ in,
                                                       float2* __restrict__ A,
                                                       const float2* __restrict__ w, BlueIO a) {
  const long long total = a.O * a.M * a.I;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const long long o = t / (a.M * a.I), r = t - o * a.M * a.I;
    const long long j = r / a.I;
    const int i = (int)(r - j * a.I);
    float2 v = make_float2(0.f, 0.f);
    if (j < a.n) {
      long long ju = j + a.shift;
      if (ju >= a.n) ju -= a.n;
      const long long e = o * a.map.eo + ju * a.map.ei + (long long)(i / a.P) * a.map.ec +
                          (long long)(i % a.P) * a.map.ep;
      if (a.kind == LOAD_C64) {
        v = __ldg(reinterpret_cast<const float2*>(in) + e);
      } else if (a.kind == LOAD_I8X2) {
        const char2 c = __ldg(reinterpret_cast<const char2*>(in) + e);
        v = make_float2((float)c.x, (float)c.y);
      } else if (a.kind == LOAD_U4X2 || a.kind == LOAD_U2X2) {
        v = ld_packed(in, e, a.kind);
      } else {
        v = make_float2(__ldg(reinterpret_cast<const float*>(in) + e), 0.f);
      }
      float2 ww = w[j];
      if (a.conj_w) ww.y = -ww.y;
      v = cmulf(v, ww);
    }
    A[t] = v;
  }
}

// A[o, k, i] *= bhat[k]  (or its conjugate)
__global__ void __launch_bounds__(256) blue_mul_kernel(float2* __restrict__ A,
                                                       const float2* __restrict__ bhat,
                                                       long long O, long long M, long long I,
                                                       int conj) {
  const long long total = O * M * I;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const long long k = (t / I) % M;
    float2 b = __ldg(bhat + k);
    if (conj) b.y = -b.y;
    A[t] = cmulf(A[t], b);
  }
}

// middle step of dedispersion: A[k] *= H[k, chan] for k < n, 0 for the padding.
// p carries the chirp description (same fields as the pass kernels use).
__global__ void __launch_bounds__(256) blue_midH_kernel(float2* __restrict__ A, long long n,
                                                        long long M, long long I, int P,
                                                        const __grid_constant__ PassArgs p) {
  const long long total = M * I;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const long long k = t / I;
    const int i = (int)(t - k * I);
    float2 v = make_float2(0.f, 0.f);
    if (k < n) {
      const int chan = i / P;
      float2 h;
      if (p.chirp_kind == CHIRP_ARRAY) {
        h = __ldg(p.chirp_arr + k * p.chirp_sk + (long long)chan * p.chirp_sc);
        h.x *= p.scale;
        h.y *= p.scale;
      } else if (p.chirp_kind == CHIRP_RAMP) {
        h = ramp_value(p, i, k);   // ramp plans have one column per lane (P = 1)
      } else {
        h = chirp_value(p, p.chan_freq[chan], k);
      }
      v = cmulf(A[t], h);
    }
    A[t] = v;
  }
}

// out[o, k', i] = A[o, k, i] * w[k] * scale for lo <= k < hi, written at row k - lo
__global__ void __launch_bounds__(256) blue_post_kernel(const float2* __restrict__ A,
                                                        float2* __restrict__ out,
                                                        const float2* __restrict__ w, BlueIO a) {
  const long long rows = a.hi - a.lo;
  const long long total = a.O * rows * a.I;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const long long o = t / (rows * a.I), r = t - o * rows * a.I;
    const long long k = r / a.I + a.lo;
    const int i = (int)(r % a.I);
    float2 ww = w[k];
    if (a.conj_w) ww.y = -ww.y;
    float2 v = cmulf(A[(o * a.M + k) * a.I + i], ww);
    v.x *= a.scale;
    v.y *= a.scale;
    long long ku = k - a.lo + a.shift;
    if (a.shift && ku >= a.n) ku -= a.n;
    out[o * a.map.eo + ku * a.map.ei + (long long)(i / a.P) * a.map.ec +
        (long long)(i % a.P) * a.map.ep] = v;
  }
}

}  // namespace pbk
