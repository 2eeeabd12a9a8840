// Host -> device bandwidth probe (why end-to-end throughput stops scaling at 4 GPUs).
// For every GPU subset of interest, every listed GPU copies its own 1 GiB host buffer to the device
// repeatedly for ~0.4 s, all of them at the same time (one host thread + stream per GPU); the
// aggregate and the per-GPU rates are printed.  Host buffer kinds:
//   pinned   cudaHostAlloc(cudaHostAllocDefault)
//   wc       cudaHostAlloc(cudaHostAllocWriteCombined)
//   reg      aligned_alloc + madvise(MADV_HUGEPAGE) + cudaHostRegister
// and placements: "any" (allocated by the main thread) or "local" (allocated by a thread bound to
// the CPUs of the GPU's NUMA node, read from sysfs, first-touch).
// Build: nvcc -O2 -o scripts/h2d_probe.bin scripts/h2d_probe.cu -lpthread
#include <cuda_runtime.h>
#include <sched.h>
#include <sys/mman.h>
#include <unistd.h>

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>
#include <thread>
#include <vector>

static const size_t kBytes = 1ull << 30;

static std::vector<int> node_cpus(int dev, int* node_out) {
  char bus[32] = {0};
  cudaDeviceGetPCIBusId(bus, sizeof(bus), dev);
  for (char* c = bus; *c; ++c) *c = (char)tolower(*c);
  std::ifstream f(std::string("/sys/bus/pci/devices/") + bus + "/numa_node");
  int node = -1;
  f >> node;
  *node_out = node;
  std::vector<int> cpus;
  if (node < 0) return cpus;
  std::ifstream g("/sys/devices/system/node/node" + std::to_string(node) + "/cpulist");
  std::string s;
  g >> s;
  size_t pos = 0;
  while (pos < s.size()) {
    size_t e = s.find(',', pos);
    if (e == std::string::npos) e = s.size();
    std::string part = s.substr(pos, e - pos);
    size_t d = part.find('-');
    int lo = atoi(part.c_str()), hi = d == std::string::npos ? lo : atoi(part.c_str() + d + 1);
    for (int c = lo; c <= hi; ++c) cpus.push_back(c);
    pos = e + 1;
  }
  return cpus;
}

static void* host_alloc(int kind) {
  void* p = nullptr;
  if (kind == 0) { if (cudaHostAlloc(&p, kBytes, cudaHostAllocDefault) != cudaSuccess) return nullptr; }
  else if (kind == 1) { if (cudaHostAlloc(&p, kBytes, cudaHostAllocWriteCombined) != cudaSuccess) return nullptr; }
  else {
    p = aligned_alloc(2 << 20, kBytes);
    if (!p) return nullptr;
    madvise(p, kBytes, MADV_HUGEPAGE);
    memset(p, 1, kBytes);
    if (cudaHostRegister(p, kBytes, cudaHostRegisterDefault) != cudaSuccess) { free(p); return nullptr; }
    return p;
  }
  memset(p, 1, kBytes);
  return p;
}
static void host_free(void* p, int kind) {
  if (!p) return;
  if (kind == 2) { cudaHostUnregister(p); free(p); } else cudaFreeHost(p);
}

int main() {
  int ndev = 0;
  cudaGetDeviceCount(&ndev);
  printf("h2d_probe: %d GPU(s), %ld host CPUs online\n", ndev, sysconf(_SC_NPROCESSORS_ONLN));
  std::vector<void*> dbuf(ndev);
  for (int d = 0; d < ndev; ++d) {
    int node;
    auto cpus = node_cpus(d, &node);
    char bus[32] = {0};
    cudaDeviceGetPCIBusId(bus, sizeof(bus), d);
    printf("  gpu %d pci %s numa_node %d (%zu cpus)\n", d, bus, node, cpus.size());
    cudaSetDevice(d);
    cudaMalloc(&dbuf[d], kBytes);
  }
  std::vector<std::vector<int>> subsets;
  for (int d = 0; d < ndev; ++d) subsets.push_back({d});
  if (ndev >= 2) subsets.push_back({0, 1});
  if (ndev >= 8) subsets.push_back({0, 4});
  if (ndev >= 4) subsets.push_back({0, 1, 2, 3});
  if (ndev >= 8) { subsets.push_back({4, 5, 6, 7}); subsets.push_back({0, 2, 4, 6}); }
  if (ndev >= 3) { std::vector<int> all; for (int d = 0; d < ndev; ++d) all.push_back(d); subsets.push_back(all); }
  const char* kinds[3] = {"pinned", "wc", "reg"};
  for (int placement = 0; placement < 2; ++placement)
    for (int kind = 0; kind < 3; ++kind) {
      std::vector<void*> hbuf(ndev, nullptr);
      for (int d = 0; d < ndev; ++d) {
        std::thread t([&, d] {
          cudaSetDevice(d);
          if (placement == 1) {
            int node;
            auto cpus = node_cpus(d, &node);
            if (!cpus.empty()) {
              cpu_set_t set;
              CPU_ZERO(&set);
              for (int c : cpus) CPU_SET(c, &set);
              sched_setaffinity(0, sizeof(set), &set);
            }
          }
          hbuf[d] = host_alloc(kind);
        });
        t.join();
      }
      for (auto& sub : subsets) {
        std::vector<double> rate(ndev, 0.0);
        std::atomic<int> ready{0};
        std::atomic<bool> go{false};
        std::vector<std::thread> th;
        for (int d : sub)
          th.emplace_back([&, d] {
            cudaSetDevice(d);
            cudaStream_t st;
            cudaStreamCreate(&st);
            if (!hbuf[d]) { ready++; return; }
            cudaMemcpyAsync(dbuf[d], hbuf[d], kBytes, cudaMemcpyHostToDevice, st);
            cudaStreamSynchronize(st);
            ready++;
            while (!go.load()) {}
            auto t0 = std::chrono::steady_clock::now();
            int n = 0;
            double dt = 0;
            do {
              cudaMemcpyAsync(dbuf[d], hbuf[d], kBytes, cudaMemcpyHostToDevice, st);
              cudaStreamSynchronize(st);
              ++n;
              dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            } while (dt < 0.4);
            rate[d] = n * (double)kBytes / dt / 1e9;
            cudaStreamDestroy(st);
          });
        while (ready.load() < (int)sub.size()) {}
        go = true;
        for (auto& t : th) t.join();
        double tot = 0;
        std::string per, names;
        for (int d : sub) {
          tot += rate[d];
          char b[32];
          snprintf(b, sizeof(b), " %.1f", rate[d]);
          per += b;
          names += (names.empty() ? "" : ",") + std::to_string(d);
        }
        printf("%-6s %-5s gpus {%s}: aggregate %7.1f GB/s  per GPU%s\n", kinds[kind],
               placement ? "local" : "any", names.c_str(), tot, per.c_str());
        fflush(stdout);
      }
      for (int d = 0; d < ndev; ++d) {
        cudaSetDevice(d);
        host_free(hbuf[d], kind);
      }
    }
  return 0;
}
