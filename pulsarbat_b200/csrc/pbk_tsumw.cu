// Instantiations of the warp-private time-summing last pass (pbk_tsumw.cuh) for 2^8-point tiles.
#include <algorithm>

#include "pbk_tsumw.cuh"
#include "pbk_tsumw_launch.h"

namespace pbk {

using Cfg = FastCfg<16, 16, 1, 1, 4, 256, 2>;

template <int EPI>
static cudaError_t tsumw_launch_epi(const PassArgs& a, const CUtensorMap& tm,
                                    const float2* d_tables, long long ntiles, int num_sms,
                                    cudaStream_t st) {
  using T_ = TsumwCfg<Cfg>;
  auto kern = tsum_warp_kernel<Cfg, EPI>;
  static bool attr_done[16] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 16 && !attr_done[dev]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)T_::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    attr_done[dev] = true;
  }
  // virtual CTAs (= groups) as in tma_launch_variant<.., TSUM>: q adjacent ones share a run, a run
  // spans at least two whole groups of summed rows; rounded to whole CTAs of NG groups
  constexpr int NG = T_::NG;
  const long long q = a.tsum_q;
  const long long per = std::max<long long>(1, std::min<long long>((long long)num_sms * NG / q,
                                                                   (ntiles / q) >> (a.tsum_log2 + 1)));
  long long vgrid = q * per;
  while (vgrid % NG) vgrid += q;
  return pdl_launch(kern, dim3((unsigned)(vgrid / NG)), dim3(T_::CTA_THREADS), T_::SMEM_BYTES, st,
                    a, tm, d_tables, ntiles);
}

bool tsumw_supported(int log2L, int log2pw) { return log2L == Cfg::LOG2L && log2pw == Cfg::LOG2PW; }

cudaError_t tsumw_launch(const PassArgs& a, const CUtensorMap& tm, const float2* d_tables,
                         long long ntiles, int num_sms, cudaStream_t st) {
  if (a.epi_kind == EPI_INTENSITY)
    return tsumw_launch_epi<EPI_INTENSITY>(a, tm, d_tables, ntiles, num_sms, st);
  return tsumw_launch_epi<EPI_STOKES_I>(a, tm, d_tables, ntiles, num_sms, st);
}

}  // namespace pbk
