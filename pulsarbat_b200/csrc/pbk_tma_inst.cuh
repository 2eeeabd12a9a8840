// pbk_tma_inst.cuh -- shared body of the instantiation units of the TMA-pipelined pass kernels
// (pbk_tma.cuh); one translation unit per tile length, like pbk_fast_inst.cuh.
#pragma once
#include <algorithm>

#include "pbk_tma.cuh"
#include "pbk_tma_launch.h"

namespace pbk {

template <class C>
static void tma_cfg_info(TmaInfo* info) {
  using T_ = TmaCfg<C>;
  info->log2pw = C::LOG2PW;
  info->log2L = C::LOG2L;
  info->box_rows = T_::BOX_ROWS;
  info->groups = T_::NG;
  info->buffers = T_::NBUF;
  info->threads = T_::CTA_THREADS;
  info->smem = T_::SMEM_BYTES;
  info->tsum_ok = (C::stride(0) * C::PW) % C::NT == 0;
  info->mid_only = false;
}

template <int MODE, class C, int LOADK, int EPI, bool TWOCH = false, bool TSUM = false>
static cudaError_t tma_launch_variant(const PassArgs& a, const CUtensorMap& tm,
                                      const float2* d_tables, long long ntiles, int num_sms,
                                      cudaStream_t st) {
  using T_ = TmaCfg<C>;
  auto kern = tma_pass_kernel<MODE, C, LOADK, EPI, TWOCH, TSUM>;
  static bool attr_done[16] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 16 && !attr_done[dev]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)T_::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    attr_done[dev] = true;
  }
  constexpr int NG = T_::NG;
  unsigned grid;
  if (TSUM) {
    // virtual CTAs (= groups) as in cfg_launch_variant: q adjacent ones share a run, and a run
    // spans at least two whole groups of summed rows; rounded to whole CTAs of NG groups
    const long long q = a.tsum_q;
    long long per = std::max<long long>(1, std::min<long long>((long long)num_sms * NG / q,
                                                               (ntiles / q) >> (a.tsum_log2 + 1)));
    long long vgrid = q * per;
    while (vgrid % NG) vgrid += q;      // q and NG are powers of two
    grid = (unsigned)(vgrid / NG);
  } else {
    grid = (unsigned)std::min<long long>((ntiles + NG - 1) / NG, num_sms);
  }
  return pdl_launch(kern, dim3(grid), dim3(T_::CTA_THREADS), T_::SMEM_BYTES, st, a, tm, d_tables,
                    ntiles);
}

template <class C>
static cudaError_t tma_cfg_launch(int mode, const PassArgs& a, const CUtensorMap& tm,
                                  const float2* d_tables, long long ntiles, int num_sms,
                                  cudaStream_t st) {
  switch (mode) {
    case MODE_FWD:
      if (a.load_kind == LOAD_PLANAR)
        return tma_launch_variant<MODE_FWD, C, LK_PLANAR, EPI_SCRATCH>(a, tm, d_tables, ntiles,
                                                                       num_sms, st);
      return tma_launch_variant<MODE_FWD, C, LK_C64, EPI_SCRATCH>(a, tm, d_tables, ntiles, num_sms,
                                                                  st);
    case MODE_MID:
      if (a.P == 1)
        return tma_launch_variant<MODE_MID, C, LK_PLANAR, EPI_SCRATCH, true>(a, tm, d_tables,
                                                                             ntiles, num_sms, st);
      return tma_launch_variant<MODE_MID, C, LK_PLANAR, EPI_SCRATCH>(a, tm, d_tables, ntiles,
                                                                     num_sms, st);
    default:
      if (!a.final_epi)
        return tma_launch_variant<MODE_INV, C, LK_PLANAR, EPI_SCRATCH>(a, tm, d_tables, ntiles,
                                                                       num_sms, st);
      if (a.tsum_log2 > 0) {
        if (a.epi_kind == EPI_INTENSITY)
          return tma_launch_variant<MODE_INV, C, LK_PLANAR, EPI_INTENSITY, false, true>(
              a, tm, d_tables, ntiles, num_sms, st);
        return tma_launch_variant<MODE_INV, C, LK_PLANAR, EPI_STOKES_I, false, true>(
            a, tm, d_tables, ntiles, num_sms, st);
      }
      switch (a.epi_kind) {
        case EPI_C64:
          return tma_launch_variant<MODE_INV, C, LK_PLANAR, EPI_C64>(a, tm, d_tables, ntiles,
                                                                     num_sms, st);
        case EPI_INTENSITY:
          return tma_launch_variant<MODE_INV, C, LK_PLANAR, EPI_INTENSITY>(a, tm, d_tables, ntiles,
                                                                           num_sms, st);
        default:
          return tma_launch_variant<MODE_INV, C, LK_PLANAR, EPI_STOKES_I>(a, tm, d_tables, ntiles,
                                                                          num_sms, st);
      }
  }
}

// tile lengths whose only TMA kernel is the scratch-to-scratch MID pass
template <class C>
static cudaError_t tma_cfg_launch_mid(int mode, const PassArgs& a, const CUtensorMap& tm,
                                      const float2* d_tables, long long ntiles, int num_sms,
                                      cudaStream_t st) {
  if (mode != MODE_MID || a.final_epi) return cudaErrorInvalidValue;
  if (a.P == 1)
    return tma_launch_variant<MODE_MID, C, LK_PLANAR, EPI_SCRATCH, true>(a, tm, d_tables, ntiles,
                                                                         num_sms, st);
  return tma_launch_variant<MODE_MID, C, LK_PLANAR, EPI_SCRATCH>(a, tm, d_tables, ntiles, num_sms,
                                                                 st);
}

}  // namespace pbk
