// pbk_l2pipe.cuh -- the three middle passes of a 3-level plan as ONE persistent kernel that keeps
// their intermediate results in the 126 MB L2.
//
// After the first forward level a 2^n-point column is 2^l1 independent contiguous blocks (one per
// first-level bin k1, N / 2^l1 rows each: 16 MB for BASELINE config 2).  The next three passes --
// FWD level 2, MID level 3 (fft * chirp * ifft), INV level 2 -- work inside a block.  Run as three
// launches over the whole array they cost three HBM round trips; here their tiles are handed out
// by a ticket counter in a software-pipelined order
//     A(b0) | A(b1) B(b0) | A(b2) B(b1) C(b0) | A(b3) B(b2) C(b1) | ...
// (A, B, C = the three passes; inside a slot each pass is one run of tickets) so that what pass A writes
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
// `hook()` runs on every thread right after the tile's first barrier (the place where the plain
// kernels prepare their next tile record): thread 0 prepares the next work item there
template <class C, class Sync, class Hook>
__device__ __forceinline__ void l2p_tile_fwd(const PassArgs& p, const FastTile& T, float4* tile,
                                             const float2* tws, float4* G4, int tid, Sync sync,
                                             Hook hook) {
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
  hook();
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

template <class C, bool TWOCH, class Sync, class Hook>
__device__ __forceinline__ void l2p_tile_mid(const PassArgs& p, const FastTile& T, float4* tile,
                                             const float2* tws, int tid, Sync sync, Hook hook) {
  constexpr int RL = C::RL;
  constexpr int LTASKS = (C::L / RL) * C::PW;
  constexpr int LITERS = (LTASKS + C::NT - 1) / C::NT;
  const int pr = tid & (C::PW - 1);
  const unsigned rb_in = (unsigned)(p.min.a_row * 8), rb_out = (unsigned)(p.mout.a_row * 8);
  fwd_first<C, LK_PLANAR>(T, tile, tws, tid, rb_in);
  sync();
  hook();
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

template <class C, class Sync, class Hook>
__device__ __forceinline__ void l2p_tile_inv(const PassArgs& p, const FastTile& T, float4* tile,
                                             const float2* tws, float4* G4, int tid, Sync sync,
                                             Hook hook) {
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
  hook();
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

// a decoded ticket, prepared by thread 0 one work item ahead
struct L2Item {
  long long t;        // ticket (>= total: no more work)
  long long idx;      // work item within (block, pass)
  int phase, blk;
  int valid;          // 0: a ticket that maps to nothing (pipeline fill / drain, uneven passes)
  int dep_ok;         // the dependency counter had already reached its target when it was polled
};

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
  TileInfo* sinfo = reinterpret_cast<TileInfo*>(G4 + CA::RL);     // [2]: the NEXT item's tiles
  __shared__ L2Item snext;
  const int tid = threadIdx.x;
  for (int i = tid; i < CA::TW_TOTAL; i += CA::NT) tws_a[i] = tab_a[i];
  for (int i = tid; i < CB::TW_TOTAL; i += CA::NT) tws_b[i] = tab_b[i];

  const long long items_b = q.tiles_b / 2;
  const long long per_slot = 2 * q.tiles_a + items_b;
  const long long total = (long long)(q.nblocks + 2) * per_slot;
  const int g = tid / CB::NT, gtid = tid - g * CB::NT;       // half-CTA groups of pass B

  // ---- thread 0 only: ticket prefetch, deferred publication, preparation of the next item ----
  unsigned tnext = 0;                    // the ticket after the prepared one (atomic in flight)
  int pub_blk = -1, pub_phase = 0;       // finished item whose completion is not published yet
  unsigned pub_cnt = 0;
  auto publish = [&]() {                 // the item's stores precede a CTA barrier this thread passed
    if (pub_blk >= 0) {
      __threadfence();
      atomicAdd(q.done + 2 * pub_blk + pub_phase, pub_cnt);
      pub_blk = -1;
    }
  };
  auto dep_ptr = [&](int phase, int blk) { return q.done + 2 * blk + (phase - 1); };
  auto dep_need = [&](int phase) { return (unsigned)(phase == 1 ? q.tiles_a : q.tiles_b); };
  auto prepare = [&](long long t) {      // decode ticket t into snext / sinfo (non-blocking)
    L2Item it;
    it.t = t;
    it.valid = 0;
    it.dep_ok = 1;
    it.phase = 0; it.blk = 0; it.idx = 0;
    if (t < total) {
      // a slot is A(slot), then B(slot - 1), then C(slot - 2), each pass as one run of tickets: by
      // the time the first B ticket of a block is taken, the last A ticket of that block was taken
      // two pass-runs (several waves of CTAs) earlier, so the dependency polls almost never wait
      const int slot = (int)(t / per_slot);
      const long long r = t - (long long)slot * per_slot;
      it.phase = r < q.tiles_a ? 0 : r < q.tiles_a + items_b ? 1 : 2;
      it.idx = r - (it.phase == 0 ? 0 : it.phase == 1 ? q.tiles_a : q.tiles_a + items_b);
      it.blk = slot - it.phase;
      it.valid = it.blk >= 0 && it.blk < q.nblocks;
      if (it.valid) {
        if (it.phase > 0) {
          it.dep_ok = *(const volatile unsigned*)dep_ptr(it.phase, it.blk) >= dep_need(it.phase);
          __threadfence();   // acquire; the CTA barrier before the tile extends it to every thread
        }
        if (it.phase == 0)
          fast_tile_info<CA, EPI_SCRATCH>(pa, (long long)it.blk * q.tiles_a + it.idx, sinfo[0], 64, 8);
        else if (it.phase == 2)
          fast_tile_info<CA, EPI_SCRATCH>(pc, (long long)it.blk * q.tiles_a + it.idx, sinfo[0], 64, 8);
        else
          for (int h = 0; h < 2; ++h)
            fast_tile_info<CB, EPI_SCRATCH>(pb, (long long)it.blk * q.tiles_b + 2 * it.idx + h,
                                            sinfo[h], 64, 8);
      }
    }
    snext = it;
  };
  if (tid == 0) {
    const unsigned t0 = atomicAdd(q.ticket, 1u);
    tnext = atomicAdd(q.ticket, 1u);
    prepare(t0);
  }

  for (;;) {
    __syncthreads();     // snext / sinfo are ready; tile buffer and G are free
    const L2Item cur = snext;
    if (cur.t >= total) break;
    if (!cur.valid) {    // nothing to do for this ticket: prepare the next one (all have read snext)
      __syncthreads();
      if (tid == 0) {
        publish();
        const unsigned t2 = tnext;
        tnext = atomicAdd(q.ticket, 1u);
        prepare(t2);
      }
      continue;
    }
    TileInfo ti = sinfo[cur.phase == 1 ? g : 0];
    // pass-B tiles only meet at half-CTA barriers: make sure the other half has read the records
    // before thread 0 rewrites them in its hook
    if (cur.phase == 1) __syncthreads();
    if (!cur.dep_ok) {   // rare: the producer pass of this block was not finished at prefetch time
      if (tid == 0) {
        publish();       // nothing this CTA still owes may be what the producers wait for
        const volatile unsigned* d = dep_ptr(cur.phase, cur.blk);
        const unsigned need = dep_need(cur.phase);
        unsigned spins = 0;
        while (*d < need) {
          __nanosleep(32);
          if (++spins > (1u << 24)) { atomicExch(q.err, 1u); break; }
        }
        __threadfence();
      }
      __syncthreads();
    }
    auto hook = [&]() {  // after the tile's first barrier: every thread holds cur / ti in registers
      if (tid == 0) {
        publish();                           // the PREVIOUS item: its stores are long since out
        const unsigned t2 = tnext;
        tnext = atomicAdd(q.ticket, 1u);     // (returns while this tile is being computed)
        prepare(t2);
      }
    };
    if (cur.phase == 0) {
      const FastTile T = l2p_context<CA>(pa, ti, tid & (CA::PW - 1));
      l2p_tile_fwd<CA>(pa, T, tile, tws_a, G4, tid, CtaSync(), hook);
    } else if (cur.phase == 2) {
      const FastTile T = l2p_context<CA>(pc, ti, tid & (CA::PW - 1));
      l2p_tile_inv<CA>(pc, T, tile, tws_a, G4, tid, CtaSync(), hook);
    } else {
      const FastTile T = l2p_context<CB>(pb, ti, gtid & (CB::PW - 1));
      l2p_tile_mid<CB, TWOCH>(pb, T, tile + (size_t)g * CB::L * CB::PW, tws_b, gtid,
                              NamedSync<CB::NT>{1 + g}, hook);
    }
    if (tid == 0 && cur.phase < 2) {   // published one item later (or before this CTA blocks / exits)
      pub_blk = cur.blk;
      pub_phase = cur.phase;
      pub_cnt = cur.phase == 1 ? 2u : 1u;
    }
  }
  if (tid == 0) publish();
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
