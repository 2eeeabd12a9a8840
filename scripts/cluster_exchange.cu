// Microbenchmark 5 (B200): cost of a cluster-wide FFT tile.  A cluster of CS CTAs holds one tile of
// (CS*256) rows x 256 B; CTA c loads rows c + CS*m (m < 256) into its 64 KiB of shared memory,
// cluster.sync, then every CTA gathers, for its share of positions, the CS values that sit at the
// same offset in the CS shared memories (DSMEM, CS-1 of them remote), combines them and stores CS
// rows to global memory.  This is the data movement of a 2^10/2^11-point level done in ONE pass
// (local 256-point FFT + cross-CTA radix-CS stage); no FFT arithmetic here.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o cluster_exchange cluster_exchange.cu
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
namespace cg = cooperative_groups;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

constexpr int kThreads = 256;
constexpr int kLc = 256;            // local rows per CTA
constexpr int kChunk16 = 16;        // 256 B = 16 float4 per row

// rows of the tile are `row_stride16` float4 apart; tiles are adjacent 256 B column chunks
template <int CS, int MODE>   // MODE 0: gather remote after sync (FWD-like); 1: also scatter back + 2nd sync (MID-like)
__global__ void __launch_bounds__(kThreads, 2) cluster_pass(const float4* __restrict__ in,
                                                            float4* __restrict__ out,
                                                            long long row_stride16,
                                                            long long ntiles, long long tiles_per_row,
                                                            long long group16) {
  extern __shared__ float4 sm[];   // [kLc][16]
  cg::cluster_group cluster = cg::this_cluster();
  const int c = (int)cluster.block_rank();
  const long long ncl = gridDim.x / CS, cl = blockIdx.x / CS;
  const int j = threadIdx.x & 15, r0 = threadIdx.x >> 4;   // 16 rows per sweep
  const float4* rsm[CS];
#pragma unroll
  for (int q = 0; q < CS; ++q) rsm[q] = cluster.map_shared_rank(sm, q);
  for (long long t = cl; t < ntiles; t += ncl) {
    const long long g = t / tiles_per_row, col = t % tiles_per_row;
    const float4* src = in + g * group16 + col * kChunk16;
    float4* dst = out + g * group16 + col * kChunk16;
    // load local rows c + CS*m
    float4 v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int m = r0 + 16 * i;
      v[i] = __ldcg(src + (long long)(c + CS * m) * row_stride16 + j);
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) sm[(r0 + 16 * i) * 16 + j] = v[i];
    cluster.sync();
    // cross stage: this CTA owns positions m in [c*kLc/CS, (c+1)*kLc/CS)
    constexpr int OWN = kLc / CS;                 // 32 (CS=8) or 64 (CS=4)
    for (int mm = r0; mm < OWN; mm += 16) {
      const int m = c * OWN + mm;
      float4 w[CS];
#pragma unroll
      for (int q = 0; q < CS; ++q) w[q] = rsm[q][m * 16 + j];
      float4 s = w[0];
#pragma unroll
      for (int q = 1; q < CS; ++q) { s.x += w[q].x; s.y += w[q].y; s.z += w[q].z; s.w += w[q].w; }
      if (MODE == 1) {
        // scatter a combined value back to every CTA at the same position (in-place exchange)
#pragma unroll
        for (int q = 0; q < CS; ++q) {
          float4 o = w[q];
          o.x += s.x;
          const_cast<float4*>(rsm[q])[m * 16 + j] = o;
        }
      } else {
#pragma unroll
        for (int q = 0; q < CS; ++q) {
          float4 o = w[q];
          o.x += s.x;
          dst[(long long)(m + kLc * q) * row_stride16 + j] = o;   // output rows k0 + Lc*j
        }
      }
    }
    cluster.sync();
    if (MODE == 1) {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int m = r0 + 16 * i;
        dst[(long long)(c + CS * m) * row_stride16 + j] = sm[m * 16 + j];
      }
      cluster.sync();
    }
  }
}

template <int CS, int MODE>
static float run(const float4* a, float4* b, long long row_stride, long long ntiles,
                 long long tiles_per_row, long long group_bytes, int ctas_per_sm) {
  auto kern = cluster_pass<CS, MODE>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  cudaLaunchConfig_t cfg = {};
  int grid = (148 * ctas_per_sm / CS) * CS;
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = 64 * 1024;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int maxcl = 0;
  CK(cudaOccupancyMaxActiveClusters(&maxcl, kern, &cfg));
  if (maxcl * CS < grid) { grid = maxcl * CS; cfg.gridDim = dim3(grid); }
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  auto go = [&]() {
    CK(cudaLaunchKernelEx(&cfg, kern, a, b, row_stride / 16, ntiles, tiles_per_row,
                          group_bytes / 16));
  };
  go();
  CK(cudaDeviceSynchronize());
  cudaEventRecord(e0);
  go();
  cudaEventRecord(e1);
  CK(cudaEventSynchronize(e1));
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  printf("    (grid %d CTAs = %d clusters of %d, max active clusters %d)\n", grid, grid / CS, CS,
         maxcl);
  return ms;
}

int main() {
  const size_t bytes = 4ull << 30;
  float4 *a, *b;
  CK(cudaMalloc(&a, bytes));
  CK(cudaMalloc(&b, bytes));
  CK(cudaMemset(a, 1, bytes));
  CK(cudaMemset(b, 0, bytes));
  // array = 2^22 rows x 1 KiB.  strided level: rows of a tile are 4 GiB/(CS*256) apart, tiles =
  // (inner row n2) x (4 column chunks).  consecutive level: tile = CS*256 consecutive rows.
  printf("cluster tile passes over 4 GiB (read + write = 8.59 GB)\n");
  for (int pass = 0; pass < 2; ++pass) {
    {
      const int CS = 8;
      const long long L = CS * 256;
      long long row_stride, ntiles, tpr = 4, group;
      if (pass == 0) { row_stride = (long long)(bytes / L); group = 1024; ntiles = (row_stride / 1024) * tpr; }
      else { row_stride = 1024; group = L * 1024; ntiles = (long long)(bytes / group) * tpr; }
      for (int cps : {2, 3}) {
        float ms0 = run<8, 0>(a, b, row_stride, ntiles, tpr, group, cps);
        printf("  %s CS=8 gather      ctas/SM %d: %.3f ms -> %.0f GB/s\n", pass ? "consecutive" : "strided    ", cps, ms0, 2.0 * bytes / ms0 / 1e6);
        float ms1 = run<8, 1>(a, b, row_stride, ntiles, tpr, group, cps);
        printf("  %s CS=8 gather+scat ctas/SM %d: %.3f ms -> %.0f GB/s\n", pass ? "consecutive" : "strided    ", cps, ms1, 2.0 * bytes / ms1 / 1e6);
      }
    }
    {
      const int CS = 4;
      const long long L = CS * 256;
      long long row_stride, ntiles, tpr = 4, group;
      if (pass == 0) { row_stride = (long long)(bytes / L); group = 1024; ntiles = (row_stride / 1024) * tpr; }
      else { row_stride = 1024; group = L * 1024; ntiles = (long long)(bytes / group) * tpr; }
      for (int cps : {2, 3}) {
        float ms0 = run<4, 0>(a, b, row_stride, ntiles, tpr, group, cps);
        printf("  %s CS=4 gather      ctas/SM %d: %.3f ms -> %.0f GB/s\n", pass ? "consecutive" : "strided    ", cps, ms0, 2.0 * bytes / ms0 / 1e6);
        float ms1 = run<4, 1>(a, b, row_stride, ntiles, tpr, group, cps);
        printf("  %s CS=4 gather+scat ctas/SM %d: %.3f ms -> %.0f GB/s\n", pass ? "consecutive" : "strided    ", cps, ms1, 2.0 * bytes / ms1 / 1e6);
      }
    }
  }
  return 0;
}
