// pbk_fast.cuh -- second-generation tile-FFT pass kernel: everything about the tile is a
// compile-time constant (radix list, lane pairs per tile, thread count), CTAs are persistent
// (grid = resident CTAs, each loops over tiles), stage twiddles live in shared memory, and the
// inter-level twiddle W_M^(nrest*k) is factored into a per-task root times a per-tile table so a
// task needs one sincospif instead of seven.
//
// These kernels are specialised at compile time to the coherent-dedispersion passes: FWD =
// forward level with its level twiddle (complex64 or int8 input), MID = fft * generated chirp *
// ifft, INV = inverse level + epilogue.  Run-time switches around code that touches the 64-register
// data array make the compiler copy the whole array at every merge point, so there are none.
// Preconditions checked by the host (pbk_api.cu: setup_fast): uniform lane pairs (I and P even,
// pair adjacent and 16-byte aligned in both maps), a tile never straddles a row (I % W == 0),
// no fftshift bookkeeping, forward sign, unit scale, generated chirp.  Everything else runs on
// the generic kernel in pbk_fft.cuh, which is the same algorithm with runtime tile parameters.
#pragma once
#include "pbk_fft.cuh"

namespace pbk {

constexpr int ilog2c(int v) { return v <= 1 ? 0 : 1 + ilog2c(v >> 1); }

template <int R0_, int R1_, int R2_, int R3_, int LOG2PW_, int NT_, int MINB_>
struct FastCfg {
  static constexpr int NS = (R0_ > 1) + (R1_ > 1) + (R2_ > 1) + (R3_ > 1);
  static constexpr int L = R0_ * R1_ * R2_ * R3_;
  static constexpr int LOG2L = ilog2c(L);
  static constexpr int LOG2PW = LOG2PW_;
  static constexpr int PW = 1 << LOG2PW_;
  static constexpr int W = 2 * PW;
  static constexpr int NT = NT_;
  static constexpr int MINB = MINB_;
  __host__ __device__ static constexpr int radix(int j) {
    return j == 0 ? R0_ : j == 1 ? R1_ : j == 2 ? R2_ : R3_;
  }
  __host__ __device__ static constexpr int stride(int j) {  // S_j = L / (R_0 .. R_j)
    int s = L;
    for (int i = 0; i <= j; ++i) s /= radix(i);
    return s;
  }
  static constexpr int RL = radix(NS - 1);       // radix of the last (stride-1) stage
  static constexpr int LOG2RL = ilog2c(RL);
  static constexpr int KS = L / RL;              // k spacing between last-stage registers
  static constexpr int PPB = (8 >> LOG2PW_) > 1 ? (8 >> LOG2PW_) : 1;  // points per 128 bytes
  __host__ __device__ static constexpr int tw_off(int j) {  // float2 offset of stage j's table
    int o = 0;
    for (int i = 0; i < j; ++i) o += (radix(i) - 1) * stride(i);
    return o;
  }
  static constexpr int TW_TOTAL = tw_off(NS - 1) > 0 ? tw_off(NS - 1) : 1;
  static constexpr size_t TILE_BYTES = (size_t)L * PW * sizeof(float4);
  static constexpr size_t SMEM_BYTES = TILE_BYTES + (size_t)TW_TOTAL * sizeof(float2) +
                                       (size_t)RL * sizeof(float2) + 16;
};

// host-side builder of the stage tables in the layout the kernel expects:
// stage j, [m-1][q] -> exp(-2 pi i q m / (R_j S_j)), m = 1..R_j-1, q < S_j
template <class C>
static inline void fast_build_tables(float2* dst) {
  for (int j = 0; j + 1 < C::NS; ++j) {
    const int R = C::radix(j), S = C::stride(j);
    const long long M = (long long)R * S;
    float2* t = dst + C::tw_off(j);
    for (int m = 1; m < R; ++m)
      for (int q = 0; q < S; ++q) {
        const long long e = ((long long)q * m) % M;
        const double ang = -2.0 * 3.14159265358979323846 * (double)e / (double)M;
        t[(m - 1) * S + q] = make_float2((float)cos(ang), (float)sin(ang));
      }
  }
}

// ------------------------------------------------------------------------------------------
// shared-memory addressing (float4 index) with the XOR swizzle on the low point bits
// ------------------------------------------------------------------------------------------
template <class C, int S>
__device__ __forceinline__ int phys_pt(int base, int i) {
  if constexpr (C::PPB == 1) {
    return base + i * S;
  } else if constexpr (S % (C::RL * C::PPB) == 0) {
    const int sw = (base >> C::LOG2RL) & (C::PPB - 1);
    return (base ^ sw) + i * S;
  } else {
    static_assert(S >= C::RL, "non-final stage strides are multiples of the last radix");
    const int sw = ((base >> C::LOG2RL) + i * (S / C::RL)) & (C::PPB - 1);
    return (base + i * S) ^ sw;
  }
}

__device__ __forceinline__ c2 lds_c2(const float4* s, int idx) {
  const float4 t = s[idx];
  c2 v;
  v.re = make_float2(t.x, t.y);
  v.im = make_float2(t.z, t.w);
  return v;
}
__device__ __forceinline__ void sts_c2(float4* s, int idx, c2 v) {
  s[idx] = make_float4(v.re.x, v.re.y, v.im.x, v.im.y);
}

template <int R, int S, bool CONJ>
__device__ __forceinline__ void stage_twiddle(c2* v, const float2* tw, int q) {
#pragma unroll
  for (int m = 1; m < R; ++m) {
    const float2 w = tw[(m - 1) * S + q];
    v[m] = cmul(v[m], p_bc(w.x), p_bc(CONJ ? -w.y : w.y));
  }
}

// digit reversal: last-stage block index -> low part of the tile-level frequency index
template <class C>
__device__ __forceinline__ int klo_of(int b) {
  if constexpr (C::NS <= 1) {
    return 0;
  } else if constexpr (C::NS == 2) {
    return b;
  } else if constexpr (C::NS == 3) {
    return (b / C::radix(1)) + C::radix(0) * (b % C::radix(1));
  } else {
    const int m2 = b % C::radix(2);
    const int t = b / C::radix(2);
    const int m1 = t % C::radix(1);
    const int m0 = t / C::radix(1);
    return m0 + C::radix(0) * (m1 + C::radix(1) * m2);
  }
}

// per-tile context (uniform across the CTA) and per-thread lane-pair pointers
struct FastTile {
  const char* gin;    // byte pointer to this thread's lane pair at row 0 of the tile (input)
  char* gout;         // same for the output map
  unsigned nrest;     // inner time offset shared by the whole tile
  unsigned klow;      // low part of the full frequency index (MID chirp)
  int chan;           // channel of this thread's pair
};

template <class C>
__device__ __forceinline__ void fast_tile_setup(const PassArgs& p, long long tile, int pr,
                                                FastTile& T, int in_elem_bytes,
                                                int out_elem_bytes) {
  const long long q0 = tile * C::W;              // first lane of the tile
  const long long o = q0 / p.RI;
  const long long r0 = q0 - o * p.RI;
  const long long nrest = r0 / p.I;
  const int col = (int)(r0 - nrest * p.I) + 2 * pr;
  const long long o_orig = o >> p.log2Kprev;
  const long long kprev = o & ((1ll << p.log2Kprev) - 1);
  const long long klow = (kprev >> p.kl_sa) + ((kprev & p.kl_mb) << p.kl_sb);
  const long long bi = map_base(p.min, o_orig, kprev, klow, nrest, col, p.P);
  const long long bo = map_base(p.mout, o_orig, kprev, klow, nrest, col, p.P);
  T.gin = reinterpret_cast<const char*>(p.in) + bi * in_elem_bytes;
  T.gout = reinterpret_cast<char*>(p.out) + bo * out_elem_bytes;
  T.nrest = (unsigned)nrest;
  T.klow = (unsigned)klow;
  T.chan = col / p.P;
}

// streaming accesses: every element is touched once per pass, so mark it evict-first
__device__ __forceinline__ float4 ldg_stream_f4(const void* ptr) {
  return __ldcs(reinterpret_cast<const float4*>(ptr));
}

template <bool I8>
__device__ __forceinline__ c2 fast_load(const FastTile& T, unsigned row, unsigned rowbytes) {
  const char* a = T.gin + (unsigned long long)row * rowbytes;
  c2 v;
  if (!I8) {
    const float4 t = ldg_stream_f4(a);
    v.re = make_float2(t.x, t.z);
    v.im = make_float2(t.y, t.w);
  } else {
    const char4 t = __ldcs(reinterpret_cast<const char4*>(a));
    v.re = make_float2((float)t.x, (float)t.z);
    v.im = make_float2((float)t.y, (float)t.w);
  }
  return v;
}

__device__ __forceinline__ void fast_store_c64(const FastTile& T, unsigned row, unsigned rowbytes,
                                               c2 v) {
  __stcs(reinterpret_cast<float4*>(T.gout + (unsigned long long)row * rowbytes),
         make_float4(v.re.x, v.im.x, v.re.y, v.im.y));
}

// final epilogue (same semantics as store_row_epi in pbk_fft.cuh)
__device__ __forceinline__ void fast_store_epi(const PassArgs& p, const FastTile& T, unsigned row,
                                               unsigned rowbytes, c2 v) {
  const long long n = (long long)row * p.n_mul + T.nrest;
  if (n < p.crop_start || n >= p.crop_stop) return;
  char* a = T.gout + (unsigned long long)row * rowbytes;
  if (p.epi_kind == EPI_C64) {
    a -= p.crop_start * p.mout.a_n * 8;
    *reinterpret_cast<float4*>(a) = make_float4(v.re.x, v.im.x, v.re.y, v.im.y);
  } else {
    a -= p.crop_start * p.mout.a_n * 4;
    const float2 pw = p_fma(v.re, v.re, p_mul(v.im, v.im));
    if (p.epi_kind == EPI_INTENSITY) *reinterpret_cast<float2*>(a) = pw;
    else *reinterpret_cast<float*>(a) = pw.x + pw.y;
  }
}

// level twiddle for the R registers of a last-stage group: W_M^(nrest*(klo + KS*m))
//   = E * G[m],  E = W_M^(nrest*klo) (one exact root per task), G[m] = W_M^(nrest*KS*m) (per tile)
template <int R, bool CONJ>
__device__ __forceinline__ void level_twiddle(const PassArgs& p, c2* v, unsigned nrest,
                                              unsigned klo, const float2* G) {
  const float2 E = unit_root((unsigned long long)nrest * klo, p.log2M);
#pragma unroll
  for (int m = 0; m < R; ++m) {
    float2 w = E;
    if (m > 0) w = cmul1(E, G[m]);
    v[m] = cmul(v[m], p_bc(w.x), p_bc(CONJ ? -w.y : w.y));
  }
}

// chirp for the R registers of a last-stage group (uniform pair: one value serves both lanes).
// Only the generated chirp runs here; an explicit chirp array goes through the generic kernel.
template <int R, class C>
__device__ __forceinline__ void fast_chirp(const PassArgs& p, const FastTile& T, c2* v, int klo,
                                           double fchan) {
#pragma unroll
  for (int m = 0; m < R; ++m) {
    const long long kf = (long long)T.klow + ((long long)(klo + m * C::KS) << p.log2Kmul);
    const float2 h = chirp_value(p, fchan, kf);
    v[m] = cmul(v[m], p_bc(h.x), p_bc(h.y));
  }
}

// ------------------------------------------------------------------------------------------
// stages
// ------------------------------------------------------------------------------------------
// first forward stage: global -> registers -> smem (DIF, stage table 0)
template <class C, bool I8>
__device__ __forceinline__ void fwd_first(const FastTile& T, float4* tile, const float2* tws,
                                          int tid, unsigned rowbytes_in) {
  constexpr bool SIGNINV = false;
  constexpr int R = C::radix(0), S = C::stride(0);
  constexpr int TASKS = S * C::PW;
  constexpr int ITERS = (TASKS + C::NT - 1) / C::NT;
  const int pr = tid & (C::PW - 1);
#pragma unroll
  for (int it = 0; it < ITERS; ++it) {
    const int tau = tid + it * C::NT;
    if (TASKS % C::NT != 0 && tau >= TASKS) break;
    const int b = tau >> C::LOG2PW;
    c2 v[R];
#pragma unroll
    for (int i = 0; i < R; ++i) v[i] = fast_load<I8>(T, (unsigned)(b + i * S), rowbytes_in);
    Butterfly<R, SIGNINV>::run(v);
    stage_twiddle<R, S, SIGNINV>(v, tws + C::tw_off(0), b);
#pragma unroll
    for (int i = 0; i < R; ++i) sts_c2(tile, (phys_pt<C, S>(b, i) << C::LOG2PW) + pr, v[i]);
  }
}

// generic middle stage smem -> smem
template <class C, int J, bool DIT, bool SIGNINV>
__device__ __forceinline__ void mid_stage(float4* tile, const float2* tws, int tid) {
  constexpr int R = C::radix(J), S = C::stride(J);
  constexpr int TASKS = (C::L / R) * C::PW;
  constexpr int ITERS = (TASKS + C::NT - 1) / C::NT;
  const int pr = tid & (C::PW - 1);
#pragma unroll
  for (int it = 0; it < ITERS; ++it) {
    const int tau = tid + it * C::NT;
    if (TASKS % C::NT != 0 && tau >= TASKS) break;
    const int b = tau >> C::LOG2PW;
    const int q = b & (S - 1);
    const int base = (b / S) * (R * S) + q;
    c2 v[R];
#pragma unroll
    for (int i = 0; i < R; ++i) v[i] = lds_c2(tile, (phys_pt<C, S>(base, i) << C::LOG2PW) + pr);
    if (DIT) {
      stage_twiddle<R, S, true>(v, tws + C::tw_off(J), q);
      Butterfly<R, true>::run(v);
    } else {
      Butterfly<R, SIGNINV>::run(v);
      stage_twiddle<R, S, SIGNINV>(v, tws + C::tw_off(J), q);
    }
#pragma unroll
    for (int i = 0; i < R; ++i) sts_c2(tile, (phys_pt<C, S>(base, i) << C::LOG2PW) + pr, v[i]);
  }
}

template <class C, bool DIT, bool SIGNINV>
__device__ __forceinline__ void mid_stages(float4* tile, const float2* tws, int tid) {
  // forward order 1..NS-2, inverse order NS-2..1
  if constexpr (!DIT) {
    if constexpr (C::NS > 2) { mid_stage<C, 1, false, SIGNINV>(tile, tws, tid); __syncthreads(); }
    if constexpr (C::NS > 3) { mid_stage<C, 2, false, SIGNINV>(tile, tws, tid); __syncthreads(); }
  } else {
    if constexpr (C::NS > 3) { mid_stage<C, 2, true, false>(tile, tws, tid); __syncthreads(); }
    if constexpr (C::NS > 2) { mid_stage<C, 1, true, false>(tile, tws, tid); __syncthreads(); }
  }
}

// index of last-stage register i of block b (float4 index)
template <class C>
__device__ __forceinline__ int last_idx(int b, int i, int pr) {
  const int sw = C::PPB > 1 ? (b & (C::PPB - 1)) : 0;
  return ((b * C::RL + (i ^ sw)) << C::LOG2PW) + pr;
}

// last inverse stage: smem -> registers -> epilogue (mirror of fwd_first)
template <class C>
__device__ __forceinline__ void inv_last(const PassArgs& p, const FastTile& T, float4* tile,
                                         const float2* tws, int tid, unsigned rowbytes_out) {
  constexpr int R = C::radix(0), S = C::stride(0);
  constexpr int TASKS = S * C::PW;
  constexpr int ITERS = (TASKS + C::NT - 1) / C::NT;
  const int pr = tid & (C::PW - 1);
#pragma unroll
  for (int it = 0; it < ITERS; ++it) {
    const int tau = tid + it * C::NT;
    if (TASKS % C::NT != 0 && tau >= TASKS) break;
    const int b = tau >> C::LOG2PW;
    c2 v[R];
#pragma unroll
    for (int i = 0; i < R; ++i) v[i] = lds_c2(tile, (phys_pt<C, S>(b, i) << C::LOG2PW) + pr);
    stage_twiddle<R, S, true>(v, tws + C::tw_off(0), b);
    Butterfly<R, true>::run(v);
#pragma unroll
    for (int i = 0; i < R; ++i) fast_store_epi(p, T, (unsigned)(b + i * S), rowbytes_out, v[i]);
  }
}

// ------------------------------------------------------------------------------------------
// kernel
// ------------------------------------------------------------------------------------------
template <int MODE, class C, bool I8>
__global__ void __launch_bounds__(C::NT, C::MINB)
fast_pass_kernel(const __grid_constant__ PassArgs p, const float2* __restrict__ tables,
                 long long ntiles) {
  static_assert(C::NS >= 2, "fast kernel needs at least two stages");
  extern __shared__ float4 smem_dyn[];
  float4* tile = smem_dyn;
  float2* tws = reinterpret_cast<float2*>(tile + (size_t)C::L * C::PW);
  float2* G = tws + C::TW_TOTAL;
  const int tid = threadIdx.x;
  const int pr = tid & (C::PW - 1);

  for (int i = tid; i < C::TW_TOTAL; i += C::NT) tws[i] = tables[i];
  __syncthreads();

  constexpr bool SIGNINV = false;
  const int in_eb = I8 ? 2 : 8;
  const int out_eb = (MODE != MODE_FWD && p.epi_kind != EPI_C64) ? 4 : 8;
  const unsigned rb_in = (unsigned)(p.min.a_row * in_eb);
  const unsigned rb_out = (unsigned)(p.mout.a_row * out_eb);

  constexpr int RL = C::RL;
  constexpr int LTASKS = (C::L / RL) * C::PW;
  constexpr int LITERS = (LTASKS + C::NT - 1) / C::NT;

  for (long long t = p.tile0 + blockIdx.x; t < ntiles; t += gridDim.x) {
    FastTile T;
    fast_tile_setup<C>(p, t, pr, T, in_eb, out_eb);
    if (MODE != MODE_MID && tid < RL)
      G[tid] = unit_root((unsigned long long)T.nrest * (unsigned)(C::KS * tid), p.log2M);
    if (MODE == MODE_INV) __syncthreads();  // INV consumes G in its first phase

    if (MODE == MODE_FWD) {
      fwd_first<C, I8>(T, tile, tws, tid, rb_in);
      __syncthreads();
      mid_stages<C, false, SIGNINV>(tile, tws, tid);
#pragma unroll
      for (int it = 0; it < LITERS; ++it) {
        const int tau = tid + it * C::NT;
        if (LTASKS % C::NT != 0 && tau >= LTASKS) break;
        const int b = tau >> C::LOG2PW;
        c2 v[RL];
#pragma unroll
        for (int i = 0; i < RL; ++i) v[i] = lds_c2(tile, last_idx<C>(b, i, pr));
        Butterfly<RL, SIGNINV>::run(v);
        const int klo = klo_of<C>(b);
        level_twiddle<RL, SIGNINV>(p, v, T.nrest, (unsigned)klo, G);
#pragma unroll
        for (int i = 0; i < RL; ++i) fast_store_c64(T, (unsigned)(klo + i * C::KS), rb_out, v[i]);
      }
    } else if (MODE == MODE_MID) {
      fwd_first<C, false>(T, tile, tws, tid, rb_in);
      __syncthreads();
      mid_stages<C, false, false>(tile, tws, tid);
      const double fchan = p.chan_freq[T.chan];
#pragma unroll
      for (int it = 0; it < LITERS; ++it) {
        const int tau = tid + it * C::NT;
        if (LTASKS % C::NT != 0 && tau >= LTASKS) break;
        const int b = tau >> C::LOG2PW;
        c2 v[RL];
#pragma unroll
        for (int i = 0; i < RL; ++i) v[i] = lds_c2(tile, last_idx<C>(b, i, pr));
        Butterfly<RL, false>::run(v);
        fast_chirp<RL, C>(p, T, v, klo_of<C>(b), fchan);
        Butterfly<RL, true>::run(v);
#pragma unroll
        for (int i = 0; i < RL; ++i) sts_c2(tile, last_idx<C>(b, i, pr), v[i]);
      }
      __syncthreads();
      mid_stages<C, true, false>(tile, tws, tid);
      inv_last<C>(p, T, tile, tws, tid, rb_out);
    } else {  // MODE_INV
#pragma unroll
      for (int it = 0; it < LITERS; ++it) {
        const int tau = tid + it * C::NT;
        if (LTASKS % C::NT != 0 && tau >= LTASKS) break;
        const int b = tau >> C::LOG2PW;
        const int klo = klo_of<C>(b);
        c2 v[RL];
#pragma unroll
        for (int i = 0; i < RL; ++i) v[i] = fast_load<false>(T, (unsigned)(klo + i * C::KS), rb_in);
        level_twiddle<RL, true>(p, v, T.nrest, (unsigned)klo, G);
        Butterfly<RL, true>::run(v);
#pragma unroll
        for (int i = 0; i < RL; ++i) sts_c2(tile, last_idx<C>(b, i, pr), v[i]);
      }
      __syncthreads();
      mid_stages<C, true, false>(tile, tws, tid);
      inv_last<C>(p, T, tile, tws, tid, rb_out);
    }
    __syncthreads();  // tile buffer and G are reused by the next tile
  }
}

}  // namespace pbk
