// Microbenchmark: strided-chunk copy bandwidth on B200.
// A "tile" = ROWS rows x CHUNK bytes; rows are `stride` bytes apart; adjacent tiles (adjacent
// blockIdx) take adjacent chunks of the same rows -- the access pattern of the FFT passes.
// Prints GB/s (read+write) for each (chunk, stride, rows, threads) combination.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

template <int CH16>  // chunk in units of 16 bytes
__global__ void copy_tiles(const float4* __restrict__ in, float4* __restrict__ out, int rows,
                           long long stride16, long long tiles_per_row, long long ntiles,
                           long long group16) {
  for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const long long g = t / tiles_per_row, c = t % tiles_per_row;
    const float4* src = in + g * group16 + c * CH16;
    float4* dst = out + g * group16 + c * CH16;
    for (int i = threadIdx.x; i < rows * CH16; i += blockDim.x) {
      const int r = i / CH16, j = i % CH16;
      dst[r * stride16 + j] = __ldcs(src + r * stride16 + j);
    }
  }
}

int main() {
  const size_t bytes = 4ull << 30;
  float4 *a, *b;
  cudaMalloc(&a, bytes); cudaMalloc(&b, bytes);
  cudaMemset(a, 1, bytes); cudaMemset(b, 0, bytes);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int rows = 2048;
  printf("chunkB strideB threads ctas/SM   GB/s(read+write)\n");
  for (int pass = 0; pass < 2; ++pass) {
    // pass 0: stride 2 MiB (pass A/C pattern: row r of a tile is r*2MiB away), row width 1 KiB..
    // pass 1: stride 1 KiB (MID pattern: consecutive 1 KiB rows)
    for (int chunk : {32, 64, 128, 256, 512, 1024}) {
      for (int threads : {256, 512}) {
        for (int cps : {2, 4}) {
          long long stride = pass == 0 ? (2ll << 20) : 1024;
          long long rowwidth = 1024;            // bytes of one full row
          long long tiles_per_row = rowwidth / chunk;
          long long group_bytes, ngroups;
          if (pass == 0) { group_bytes = 1024; ngroups = (2ll << 20) / 1024; }  // n2 index
          else { group_bytes = rows * 1024ll; ngroups = bytes / group_bytes; }
          long long ntiles = ngroups * tiles_per_row;
          int grid = 148 * cps;
          auto run = [&]() {
            switch (chunk) {
              case 32: copy_tiles<2><<<grid, threads>>>(a, b, rows, stride / 16, tiles_per_row, ntiles, group_bytes / 16); break;
              case 64: copy_tiles<4><<<grid, threads>>>(a, b, rows, stride / 16, tiles_per_row, ntiles, group_bytes / 16); break;
              case 128: copy_tiles<8><<<grid, threads>>>(a, b, rows, stride / 16, tiles_per_row, ntiles, group_bytes / 16); break;
              case 256: copy_tiles<16><<<grid, threads>>>(a, b, rows, stride / 16, tiles_per_row, ntiles, group_bytes / 16); break;
              case 512: copy_tiles<32><<<grid, threads>>>(a, b, rows, stride / 16, tiles_per_row, ntiles, group_bytes / 16); break;
              default: copy_tiles<64><<<grid, threads>>>(a, b, rows, stride / 16, tiles_per_row, ntiles, group_bytes / 16); break;
            }
          };
          run(); cudaDeviceSynchronize();
          cudaEventRecord(e0); run(); cudaEventRecord(e1); cudaEventSynchronize(e1);
          float ms; cudaEventElapsedTime(&ms, e0, e1);
          cudaError_t err = cudaGetLastError();
          printf("%6d %8lld %5d %3d   %8.1f %s\n", chunk, stride, threads, cps,
                 2.0 * bytes / ms / 1e6, err == cudaSuccess ? "" : cudaGetErrorString(err));
        }
      }
    }
  }
  return 0;
}
