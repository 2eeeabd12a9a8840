// pbk_l2pipe.cuh -- the three middle passes of a 3-level plan as ONE persistent kernel that keeps
// their intermediate results in the 126 MB L2.
//
// After the first forward level a 2^n-point column is 2^l1 independent contiguous blocks (one per
// first-level bin k1, N / 2^l1 rows each: 16 MB for BASELINE config 2).  The next three passes --
// FWD level 2, MID level 3 (fft * chirp * ifft), INV level 2 -- work inside a block.  Run as three
// launches over the whole array they cost three HBM round trips; here their tiles are handed out
// by a ticket counter in a software-pipelined order
//     A(b0) | A(b1) B(b0) | A(b2) B(b1) C(b0) | A(b3) B(b2) C(b1) | ...
// (A, B, C = the three passes, interleaved tile by tile inside a slot) so that what pass A writes
// for block b is read by pass B, and what B writes by C, while it is still in L2: HBM sees one
// read of the block (A) and one write (C).  Dependencies are per (block, pass) completion counters
// in global memory: a tile of pass B waits until every tile of pass A of its block has been
// written (release: stores, __threadfence, atomicAdd; acquire: volatile poll, __threadfence,
// ld.global.cg).  Tickets are taken in order and every ticket only depends on SMALLER tickets,
// which are held by CTAs that are resident (the grid is sized to the resident capacity), so the
// waits cannot deadlock; a poll that exceeds a generous bound sets an error flag instead of
// hanging the GPU.
//
// The per-tile code is that of fast_pass_kernel (wide tiles, pair-planar scratch in and out).  A
// CTA is CA::NT threads with one CA tile of shared memory; a pass-B work item is TWO level-3 tiles,
// one per half of the CTA (thread groups with named barriers, each in its half of the buffer).
#pragma once
#include "pbk_fast.cuh"
#include "pbk_l2pipe_launch.h"

namespace pbk {


template <int NT>
struct NamedSync {   // named barrier `id` over NT threads (id 0 is __syncthreads)
  int id;
  __device__ __forceinline__ void operator()() const {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(NT) : "memory");
  }
};

// ---- one tile of each pass kind (wide tiles, LK_PLANAR in, EPI_SCRATCH out) -------------------
template <class C, class Sync>
__device__ __forceinline__ void l2p_tile_fwd(const PassArgs& p, const FastTile& T, float4* tile,
                                             const float2* tws, float4* G4, int tid, Sync sync) {
  constexpr int RL = C::RL;
  constexpr int LTASKS = (C::L / RL) * C::PW;
  constexpr int LITERS = (LTASKS + C::NT - 1) / C::NT;
  const int pr = tid & (C::PW - 1);
  const unsigned rb_in = (unsigned)(p.min.a_row * 8), rb_out = (unsigned)(p.mout.a_row * 8);
  if (tid < RL) {
    const float2 g = unit_root((unsigned long long)T.nrest * (unsigned)(C::KS * tid), p.log2M);
    G4[tid] = make_float4(g.x, g.y, g.y, g.x);
  }
  fwd_first<C, LK_PLANAR, false>(T, tile, tws, tid, rb_in);
  sync();
  mid_stages<C, false, false>(tile, tws, tid, sync);
#pragma unroll
  for (int it = 0; it < LITERS; ++it) {
    const int tau = tid + it * C::NT;
    if (LTASKS % C::NT != 0 && tau >= LTASKS) break;
    const int b = tau >> C::LOG2PW;
    c2 v[RL];
#pragma unroll
    for (int i = 0; i < RL; ++i) v[i] = lds_c2(tile, last_idx<C>(b, i, pr));
    Butterfly<RL, false>::run(v);
    const int klo = klo_of<C>(b);
    level_twiddle<RL, false>(p, v, T.nrest, (unsigned)klo, G4);
#pragma unroll
    for (int i = 0; i < RL; ++i) fast_store_c64(T, (unsigned)(klo + i * C::KS), rb_out, v[i]);
  }
}

template <class C, bool TWOCH, class Sync>
__device__ __forceinline__ void l2p_tile_mid(const PassArgs& p, const FastTile& T, float4* tile,
                                             const float2* tws, int tid, Sync sync) {
  constexpr int RL = C::RL;
  constexpr int LTASKS = (C::L / RL) * C::PW;
  constexpr int LITERS = (LTASKS + C::NT - 1) / C::NT;
  const int pr = tid & (C::PW - 1);
  const unsigned rb_in = (unsigned)(p.min.a_row * 8), rb_out = (unsigned)(p.mout.a_row * 8);
  fwd_first<C, LK_PLANAR>(T, tile, tws, tid, rb_in);
  sync();
  mid_stages<C, false, false>(tile, tws, tid, sync);
#pragma unroll
  for (int it = 0; it < LITERS; ++it) {
    const int tau = tid + it * C::NT;
    if (LTASKS % C::NT != 0 && tau >= LTASKS) break;
    const int b = tau >> C::LOG2PW;
    c2 v[RL];
#pragma unroll
    for (int i = 0; i < RL; ++i) v[i] = lds_c2(tile, last_idx<C>(b, i, pr));
    Butterfly<RL, false>::run(v);
    fast_chirp<RL, C, TWOCH>(p, T, v, klo_of<C>(b));
    Butterfly<RL, true>::run(v);
#pragma unroll
    for (int i = 0; i < RL; ++i) sts_c2(tile, last_idx<C>(b, i, pr), v[i]);
  }
  sync();
  mid_stages<C, true, false>(tile, tws, tid, sync);
  inv_last<C, EPI_SCRATCH>(T, tile, tws, tid, rb_out);
}

template <class C, class Sync>
__device__ __forceinline__ void l2p_tile_inv(const PassArgs& p, const FastTile& T, float4* tile,
                                             const float2* tws, float4* G4, int tid, Sync sync) {
  constexpr int RL = C::RL;
  constexpr int LTASKS = (C::L / RL) * C::PW;
  constexpr int LITERS = (LTASKS + C::NT - 1) / C::NT;
  const int pr = tid & (C::PW - 1);
  const unsigned rb_in = (unsigned)(p.min.a_row * 8), rb_out = (unsigned)(p.mout.a_row * 8);
  if (tid < RL) {
    const float2 g = unit_root((unsigned long long)T.nrest * (unsigned)(C::KS * tid), p.log2M);
    G4[tid] = make_float4(g.x, g.y, g.y, g.x);
  }
  sync();
#pragma unroll
  for (int it = 0; it < LITERS; ++it) {
    const int tau = tid + it * C::NT;
    if (LTASKS % C::NT != 0 && tau >= LTASKS) break;
    const int b = tau >> C::LOG2PW;
    const int klo = klo_of<C>(b);
    c2 v[RL];
#pragma unroll
    for (int i = 0; i < RL; ++i) v[i] = fast_load<LK_PLANAR>(T, (unsigned)(klo + i * C::KS), rb_in);
    level_twiddle<RL, true>(p, v, T.nrest, (unsigned)klo, G4);
    Butterfly<RL, true>::run(v);
#pragma unroll
    for (int i = 0; i < RL; ++i) sts_c2(tile, last_idx<C>(b, i, pr), v[i]);
  }
  sync();
  mid_stages<C, true, false>(tile, tws, tid, sync);
  inv_last<C, EPI_SCRATCH>(T, tile, tws, tid, rb_out);
}

// per-thread tile context from the tile record (wide tiles)
template <class C>
__device__ __forceinline__ FastTile l2p_context(const PassArgs& p, const TileInfo& ti, int pr) {
  const int colt = 2 * pr;
  const long long off_in = ((long long)(colt / p.P) * p.min.a_c + (colt % p.P) * p.min.a_p) * 8;
  const long long off_out = ((long long)(colt / p.P) * p.mout.a_c + (colt % p.P) * p.mout.a_p) * 8;
  FastTile T;
  T.gin = reinterpret_cast<const char*>(p.in) + ti.bi + off_in;
  T.gout = reinterpret_cast<char*>(p.out) + ti.bo + off_out;
  T.nrest = ti.nrest;
  T.klow = ti.klow;
  T.chan = ti.chan0 + colt / p.P;
  T.row_lo = 0;
  T.row_cnt = C::L;
  return T;
}

template <class CA, class CB, bool TWOCH>
__global__ void __launch_bounds__(CA::NT, 2)
l2pipe_kernel(const __grid_constant__ PassArgs pa, const __grid_constant__ PassArgs pb,
              const __grid_constant__ PassArgs pc, const float2* __restrict__ tab_a,
              const float2* __restrict__ tab_b, const L2PipeArgs q) {
  static_assert(CA::NT == 2 * CB::NT, "a level-3 work item is two tiles, one per half CTA");
  static_assert(2 * CB::TILE_BYTES <= CA::TILE_BYTES, "both level-3 tiles fit the level-2 buffer");
  extern __shared__ float4 smem_dyn[];
  float4* tile = smem_dyn;
  float2* tws_a = reinterpret_cast<float2*>(tile + (size_t)CA::L * CA::PW);
  float2* tws_b = tws_a + CA::TW_PAD;
  float4* G4 = reinterpret_cast<float4*>(tws_b + CB::TW_PAD);
  TileInfo* sinfo = reinterpret_cast<TileInfo*>(G4 + CA::RL);     // [2]
  __shared__ long long s_ticket;
  const int tid = threadIdx.x;
  for (int i = tid; i < CA::TW_TOTAL; i += CA::NT) tws_a[i] = tab_a[i];
  for (int i = tid; i < CB::TW_TOTAL; i += CA::NT) tws_b[i] = tab_b[i];

  const long long items_b = q.tiles_b / 2;
  const long long per_phase = q.tiles_a > items_b ? q.tiles_a : items_b;
  const long long per_slot = 3 * per_phase;
  const long long total = (long long)(q.nblocks + 2) * per_slot;
  const int g = tid / CB::NT, gtid = tid - g * CB::NT;       // half-CTA groups of pass B

  for (;;) {
    __syncthreads();                       // tile buffer, G, records and s_ticket are free again
    if (tid == 0) s_ticket = atomicAdd(q.ticket, 1u);
    __syncthreads();
    const long long t = s_ticket;
    if (t >= total) break;
    const int slot = (int)(t / per_slot);
    const long long r = t - (long long)slot * per_slot;
    const int phase = (int)(r % 3);
    const long long idx = r / 3;
    const int blk = slot - phase;
    if (blk < 0 || blk >= q.nblocks) continue;
    if (idx >= (phase == 1 ? items_b : q.tiles_a)) continue;
    if (tid == 0) {
      if (phase > 0) {     // every tile of the previous pass of this block has been written
        const volatile unsigned* d = q.done + 2 * blk + (phase - 1);
        const unsigned need = (unsigned)(phase == 1 ? q.tiles_a : q.tiles_b);
        unsigned spins = 0;
        while (*d < need) {
          __nanosleep(64);
          if (++spins > (1u << 24)) { atomicExch(q.err, 1u); break; }
        }
        __threadfence();
      }
      if (phase == 0)
        fast_tile_info<CA, EPI_SCRATCH>(pa, (long long)blk * q.tiles_a + idx, sinfo[0], 64, 8);
      else if (phase == 2)
        fast_tile_info<CA, EPI_SCRATCH>(pc, (long long)blk * q.tiles_a + idx, sinfo[0], 64, 8);
    }
    if (phase == 1 && gtid == 0)
      fast_tile_info<CB, EPI_SCRATCH>(pb, (long long)blk * q.tiles_b + 2 * idx + g, sinfo[g], 64, 8);
    __syncthreads();
    if (phase == 0) {
      const FastTile T = l2p_context<CA>(pa, sinfo[0], tid & (CA::PW - 1));
      l2p_tile_fwd<CA>(pa, T, tile, tws_a, G4, tid, CtaSync());
    } else if (phase == 2) {
      const FastTile T = l2p_context<CA>(pc, sinfo[0], tid & (CA::PW - 1));
      l2p_tile_inv<CA>(pc, T, tile, tws_a, G4, tid, CtaSync());
    } else {
      const FastTile T = l2p_context<CB>(pb, sinfo[g], gtid & (CB::PW - 1));
      l2p_tile_mid<CB, TWOCH>(pb, T, tile + (size_t)g * CB::L * CB::PW, tws_b, gtid,
                              NamedSync<CB::NT>{1 + g});
    }
    if (phase < 2) {       // publish: this thread's stores, then the block's counter
      __threadfence();
      __syncthreads();
      if (tid == 0) atomicAdd(q.done + 2 * blk + phase, phase == 1 ? 2u : 1u);
    }
  }
}

template <class CA, class CB>
static cudaError_t l2pipe_launch(const PassArgs& pa, const PassArgs& pb, const PassArgs& pc,
                                 const float2* tab_a, const float2* tab_b, const L2PipeArgs& q,
                                 int num_sms, cudaStream_t st) {
  constexpr size_t smem = CA::TILE_BYTES + (size_t)(CA::TW_PAD + CB::TW_PAD) * sizeof(float2) +
                          (size_t)CA::RL * sizeof(float4) + 2 * 64 + 64;
  static_assert(sizeof(TileInfo) <= 64, "tile record slot");
  const bool twoch = pb.P == 1;
  auto kern = twoch ? l2pipe_kernel<CA, CB, true> : l2pipe_kernel<CA, CB, false>;
  static bool attr_done[2][16] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 16 && !attr_done[twoch][dev]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    attr_done[twoch][dev] = true;
  }
  // persistent: exactly the CTAs that are resident together (2 per SM), never more -- the
  // dependency waits rely on every earlier ticket being held by a running CTA
  int per_sm = 0;
  cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, CA::NT, smem);
  if (e != cudaSuccess) return e;
  if (per_sm < 1) return cudaErrorLaunchOutOfResources;
  kern<<<num_sms * (per_sm > 2 ? 2 : per_sm), CA::NT, smem, st>>>(pa, pb, pc, tab_a, tab_b, q);
  return cudaGetLastError();
}

}  // namespace pbk
