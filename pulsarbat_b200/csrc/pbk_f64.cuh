// pbk_f64.cuh -- complex128 transforms in FP64.
//
// The reference keeps complex128 through the transform (dedispersion.py:125, fft.py:34; its own
// tests/test_dedispersion.py:73-98 asserts +-DM reversibility to atol 3e-8), so a complex128 input
// must not be narrowed to complex64 arithmetic.  complex128 is not the throughput path (BASELINE
// is complex64 / int8), so this is the simple, obviously correct formulation: a Stockham autosort
// radix-4 FFT along the slow axis of an (outer, n, inner) array, one full read + write of the array
// per stage (32 B per sample and stage; B200's FP64 rate is far above what that needs), lanes
// (`inner`) innermost so that every access is coalesced, twiddles from sincospi in FP64.
// Other lengths (the reference accepts any: a dedispersed, cropped signal is dedispersed again in
// its own reversibility test) go through Bluestein's chirp-z identity on top of the same stages, with
// the chirp phases pi j^2 / n reduced exactly in integers (f64_fft_any).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pbk {

__device__ __forceinline__ double2 zadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 zsub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double2 zmul(double2 a, double2 b) {
  return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
// exp(sign * 2 pi i * num / den), num < den, den a power of two
__device__ __forceinline__ double2 zroot(long long num, long long den, int sign) {
  double s, c;
  sincospi(2.0 * (double)num / (double)den, &s, &c);
  return make_double2(c, sign < 0 ? -s : s);
}

// one Stockham stage: sub-length `len` (multiple of RADIX), stride `s`; x -> y
//   element (o, j, i) of the (outer, n, inner) array lives at ((o n + j) inner + i)
template <int RADIX>
__global__ void __launch_bounds__(256) f64_stage_kernel(const double2* __restrict__ x,
                                                        double2* __restrict__ y, long long outer,
                                                        long long n, long long inner,
                                                        long long len, long long s, int sign) {
  const long long m = len / RADIX;
  const long long per = (n / RADIX) * inner;          // butterflies x lanes per outer index
  const long long total = outer * per;
  for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < total;
       g += (long long)gridDim.x * blockDim.x) {
    const long long o = g / per, r = g - o * per;
    const long long t = r / inner, i = r - t * inner;
    // t enumerates (group, p, q): within a group of length len*s... the autosort recursion keeps
    // n/len independent sub-transforms interleaved with stride s: index j = q + s * (p + m * k)
    const long long q = t % s, p = t / s;             // p in [0, n / (RADIX s)) = [0, m)
    const double2* xs = x + (o * n) * inner + i;
    double2* ys = y + (o * n) * inner + i;
    if (RADIX == 2) {
      const double2 a = xs[(q + s * p) * inner];
      const double2 b = xs[(q + s * (p + m)) * inner];
      const double2 w = zroot(p, len, sign);
      ys[(q + s * (2 * p)) * inner] = zadd(a, b);
      ys[(q + s * (2 * p + 1)) * inner] = zmul(zsub(a, b), w);
    } else {
      const double2 a = xs[(q + s * p) * inner];
      const double2 b = xs[(q + s * (p + m)) * inner];
      const double2 c = xs[(q + s * (p + 2 * m)) * inner];
      const double2 d = xs[(q + s * (p + 3 * m)) * inner];
      const double2 apc = zadd(a, c), amc = zsub(a, c), bpd = zadd(b, d), bmd = zsub(b, d);
      // forward (sign -1): y1 = (a - c) - i (b - d), y3 = (a - c) + i (b - d); inverse swaps them
      const double2 jb = sign < 0 ? make_double2(bmd.y, -bmd.x) : make_double2(-bmd.y, bmd.x);
      const double2 w1 = zroot(p, len, sign);
      const double2 w2 = zmul(w1, w1), w3 = zmul(w2, w1);
      ys[(q + s * (4 * p)) * inner] = zadd(apc, bpd);
      ys[(q + s * (4 * p + 1)) * inner] = zmul(zadd(amc, jb), w1);
      ys[(q + s * (4 * p + 2)) * inner] = zmul(zsub(apc, bpd), w2);
      ys[(q + s * (4 * p + 3)) * inner] = zmul(zsub(amc, jb), w3);
    }
  }
}

// multiply the spectrum by the dedispersion chirp, rounded to complex64 as the reference's
// (dedispersion.py:19-23: `.astype(np.complex64)`), or by an explicit (n, nchan) complex64 array
struct F64Chirp {
  long long N, nchan, npol;
  double df, fr_sub, inv_fr, a0, D;
  const double* chan_freq;
  const float2* chirp_arr;     // explicit chirp, or nullptr
};
__global__ void __launch_bounds__(256) f64_chirp_kernel(double2* __restrict__ x, const F64Chirp c) {
  const long long I = c.nchan * c.npol, total = c.N * I;
  for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < total;
       g += (long long)gridDim.x * blockDim.x) {
    const long long k = g / I, col = g - k * I, ch = col / c.npol;
    float2 h;
    if (c.chirp_arr) {
      h = c.chirp_arr[k * c.nchan + ch];
    } else {
      const long long ks = (k >= ((c.N + 1) >> 1)) ? k - c.N : k;      // numpy fftfreq
      const double f = fma((double)ks, c.df, c.chan_freq[ch]);
      const double a = fma(f - c.fr_sub, c.inv_fr, c.a0);
      const double phi = (c.D * a) * a / f;                             // cycles
      const double fr = phi - rint(phi);
      double sn, cs;
      sincospi(2.0 * fr, &sn, &cs);
      h = make_float2((float)cs, (float)(-sn));
    }
    x[g] = zmul(x[g], make_double2((double)h.x, (double)h.y));
  }
}

// rows [lo, hi) of (n, I) scaled -> out (complex128), or detected (float64: per element |z|^2, or
// summed over pol pairs when stokes)
__global__ void __launch_bounds__(256) f64_store_kernel(const double2* __restrict__ x, void* out,
                                                        long long lo, long long hi, long long I,
                                                        double scale, int kind /*0 c128, 1 |z|^2,
                                                        2 Stokes I*/) {
  const long long E = kind == 2 ? I / 2 : I;
  const long long total = (hi - lo) * E;
  for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < total;
       g += (long long)gridDim.x * blockDim.x) {
    const long long r = g / E, e = g - r * E;
    if (kind == 2) {
      const double2 a = x[(lo + r) * I + 2 * e], b = x[(lo + r) * I + 2 * e + 1];
      reinterpret_cast<double*>(out)[g] =
          ((a.x * a.x + a.y * a.y) + (b.x * b.x + b.y * b.y)) * (scale * scale);
    } else {
      const double2 v = x[(lo + r) * I + e];
      if (kind == 0) reinterpret_cast<double2*>(out)[g] = make_double2(v.x * scale, v.y * scale);
      else reinterpret_cast<double*>(out)[g] = (v.x * v.x + v.y * v.y) * (scale * scale);
    }
  }
}

// STFT layout change (misc.py:43-52 / 81-91) around the per-segment transforms:
//   forward: y[s, k, c, p] (the (nseg, n, C P) transform output) -> out[s, c n + ((k + n/2) mod n), p] / n
//   inverse: in[s, c n + k', p] -> x[s, (k' + n/2) mod n... ifftshift, c, p] (then transformed)
__global__ void __launch_bounds__(256) f64_stft_permute_kernel(const double2* __restrict__ in,
                                                               double2* __restrict__ out,
                                                               long long nseg, long long n,
                                                               long long C, long long P,
                                                               int inverse, double scale) {
  const long long total = nseg * n * C * P;
  for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < total;
       g += (long long)gridDim.x * blockDim.x) {
    // g enumerates the (s, k, c, p) side
    long long r = g;
    const long long p = r % P; r /= P;
    const long long c = r % C; r /= C;
    const long long k = r % n, s = r / n;
    const long long sh = (k + n / 2) % n;                       // fftshift position of bin k
    const long long chan_side = ((s * C + c) * n + sh) * P + p; // (s, c n + shifted k, p)
    if (!inverse) {
      const double2 v = in[g];
      out[chan_side] = make_double2(v.x * scale, v.y * scale);
    } else {
      // np.fft.ifftshift then ifft: bin k of the transform input is the channel at shifted position
      out[g] = in[chan_side];   // ifftshift(y)[k] = y[(k + n // 2) mod n]
    }
  }
}

// ---- Bluestein: X[k] = w[k] sum_j (x[j] w[j]) conj(w)[k - j],  w[j] = exp(sign i pi j^2 / n) ----
__device__ __forceinline__ double2 f64_blue_w(long long j, long long n, int sign) {
  const long long e = (long long)(((unsigned long long)j * (unsigned long long)j) % (unsigned long long)(2 * n));
  double s, c;
  sincospi((double)e / (double)n, &s, &c);       // e / n in [0, 2): exact argument reduction
  return make_double2(c, sign < 0 ? -s : s);
}
// a[o, j, i] = x[o, j, i] w[j] for j < n, 0 for n <= j < M
__global__ void __launch_bounds__(256) f64_blue_pre_kernel(const double2* __restrict__ x,
                                                           double2* __restrict__ a, long long outer,
                                                           long long n, long long M, long long inner,
                                                           int sign) {
  const long long total = outer * M * inner;
  for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < total;
       g += (long long)gridDim.x * blockDim.x) {
    const long long i = g % inner, r = g / inner, j = r % M, o = r / M;
    a[g] = j < n ? zmul(x[(o * n + j) * inner + i], f64_blue_w(j, n, sign)) : make_double2(0.0, 0.0);
  }
}
// b[j] = conj(w[|j|]) on the circle of length M (j = -(n-1) .. n-1), 0 elsewhere
__global__ void __launch_bounds__(256) f64_blue_filter_kernel(double2* __restrict__ b, long long n,
                                                              long long M, int sign) {
  for (long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x; j < M;
       j += (long long)gridDim.x * blockDim.x) {
    const long long d = j < n ? j : (M - j < n ? M - j : -1);
    double2 v = make_double2(0.0, 0.0);
    if (d >= 0) { v = f64_blue_w(d, n, sign); v.y = -v.y; }
    b[j] = v;
  }
}
__global__ void __launch_bounds__(256) f64_blue_mul_kernel(double2* __restrict__ a,
                                                           const double2* __restrict__ bhat,
                                                           long long outer, long long M,
                                                           long long inner) {
  const long long total = outer * M * inner;
  for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < total;
       g += (long long)gridDim.x * blockDim.x)
    a[g] = zmul(a[g], bhat[(g / inner) % M]);
}
// X[o, k, i] = w[k] c[o, k, i] / M for k < n
__global__ void __launch_bounds__(256) f64_blue_post_kernel(const double2* __restrict__ c,
                                                            double2* __restrict__ X, long long outer,
                                                            long long n, long long M, long long inner,
                                                            int sign) {
  const long long total = outer * n * inner;
  const double sc = 1.0 / (double)M;
  for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < total;
       g += (long long)gridDim.x * blockDim.x) {
    const long long i = g % inner, r = g / inner, k = r % n, o = r / n;
    const double2 v = zmul(c[(o * M + k) * inner + i], f64_blue_w(k, n, sign));
    X[g] = make_double2(v.x * sc, v.y * sc);
  }
}

static inline unsigned f64_blocks(long long total) {
  long long b = (total + 255) / 256;
  return (unsigned)(b > 148ll * 16 ? 148ll * 16 : (b < 1 ? 1 : b));
}

// n-point transform (n = 2^ln) along axis 1 of (outer, n, inner); `src` is left untouched, the
// result ends in the returned buffer (one of a / b).  sign -1 forward, +1 inverse (unscaled).
static inline cudaError_t f64_fft(const double2* src, double2* a, double2* b, long long outer,
                                  long long n, long long inner, int sign, cudaStream_t st,
                                  double2** result) {
  const double2* x = src;
  double2* bufs[2] = {a, b};
  int w = 0;
  long long len = n, s = 1;
  if (n == 1) {
    cudaError_t e = cudaMemcpyAsync(a, src, (size_t)outer * inner * sizeof(double2),
                                    cudaMemcpyDeviceToDevice, st);
    *result = a;
    return e;
  }
  while (len > 1) {
    double2* y = bufs[w];
    const long long total = outer * (n / (len % 4 == 0 ? 4 : 2)) * inner;
    if (len % 4 == 0) {
      f64_stage_kernel<4><<<f64_blocks(total), 256, 0, st>>>(x, y, outer, n, inner, len, s, sign);
      len /= 4;
      s *= 4;
    } else {
      f64_stage_kernel<2><<<f64_blocks(total), 256, 0, st>>>(x, y, outer, n, inner, len, s, sign);
      len /= 2;
      s *= 2;
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    x = y;
    w ^= 1;
  }
  *result = const_cast<double2*>(x);
  return cudaSuccess;
}

// Any length.  `alloc(bytes)` hands out device arrays that live until the caller's work is queued
// (stream-ordered scratch); the unscaled transform of `src` ends in *result.
template <class Alloc>
static inline cudaError_t f64_fft_any(const double2* src, Alloc&& alloc, long long outer,
                                      long long n, long long inner, int sign, cudaStream_t st,
                                      double2** result) {
  cudaError_t e;
  double2 *a = nullptr, *b = nullptr;
  bool pow2 = n > 0 && (n & (n - 1)) == 0;
  if (pow2) {
    const size_t bytes = (size_t)outer * n * inner * sizeof(double2);
    if ((e = alloc((void**)&a, bytes)) != cudaSuccess) return e;
    if ((e = alloc((void**)&b, bytes)) != cudaSuccess) return e;
    return f64_fft(src, a, b, outer, n, inner, sign, st, result);
  }
  long long M = 16;
  while (M < 2 * n - 1) M <<= 1;
  const size_t bytes = (size_t)outer * M * inner * sizeof(double2);
  double2 *pad, *filt, *fa, *fb, *bhat, *A, *c;
  if ((e = alloc((void**)&pad, bytes)) != cudaSuccess) return e;
  if ((e = alloc((void**)&a, bytes)) != cudaSuccess) return e;
  if ((e = alloc((void**)&b, bytes)) != cudaSuccess) return e;
  if ((e = alloc((void**)&filt, (size_t)M * sizeof(double2))) != cudaSuccess) return e;
  if ((e = alloc((void**)&fa, (size_t)M * sizeof(double2))) != cudaSuccess) return e;
  if ((e = alloc((void**)&fb, (size_t)M * sizeof(double2))) != cudaSuccess) return e;
  f64_blue_filter_kernel<<<f64_blocks(M), 256, 0, st>>>(filt, n, M, sign);
  if ((e = f64_fft(filt, fa, fb, 1, M, 1, -1, st, &bhat)) != cudaSuccess) return e;
  f64_blue_pre_kernel<<<f64_blocks(outer * M * inner), 256, 0, st>>>(src, pad, outer, n, M, inner,
                                                                      sign);
  if ((e = f64_fft(pad, a, b, outer, M, inner, -1, st, &A)) != cudaSuccess) return e;
  f64_blue_mul_kernel<<<f64_blocks(outer * M * inner), 256, 0, st>>>(A, bhat, outer, M, inner);
  // inverse M-point transform: A is the source; results alternate between the other two arrays
  double2* o1 = A == a ? b : a;
  if ((e = f64_fft(A, o1, pad, outer, M, inner, +1, st, &c)) != cudaSuccess) return e;
  double2* X = c == pad ? o1 : pad;       // n <= M: the first outer*n*inner entries of a free array
  f64_blue_post_kernel<<<f64_blocks(outer * n * inner), 256, 0, st>>>(c, X, outer, n, M, inner, sign);
  *result = X;
  return cudaGetLastError();
}

}  // namespace pbk
