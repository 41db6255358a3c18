// Microbenchmark behind DESIGN.md §6 "K2+K3 fusion through a thread-block cluster" (VERDICT r1 item 3-ii/iii).
//
// A fused channel-IDFT + resample kernel would transpose the intermediate v[tau][t1] inside a cluster: every CTA
// transforms 1/C of the rows and scatters 16-byte units (both polarisations of one (tau, t1)) into the shared
// memory of the CTA that owns column t1 - (C-1)/C of the 357 MB intermediate of one SKA-Low step crosses DSMEM.
// This program measures exactly that access pattern (st.shared::cluster.v4 of 16-byte units, each warp store
// covering 512 contiguous bytes of ONE peer, peers visited round-robin) against the same stores into the CTA's own
// shared memory, for cluster sizes 2/4/8/16, and counts how many SMs the cluster launch can occupy at all.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dsmem_transpose dsmem_transpose.cu && ./dsmem_transpose
#include <cooperative_groups.h>
#include <cstdio>
#include <cuda_runtime.h>
#include <set>
namespace cg = cooperative_groups;

constexpr int kThreads = 256;
constexpr int kTileBytes = 64 * 1024;          // per-CTA landing buffer (a ring slot of the fused design)

__global__ void __launch_bounds__(kThreads, 1) scatter(int csize, int remote, int iters, unsigned* smid, float* sink) {
  extern __shared__ __align__(16) unsigned char smem[];
  cg::cluster_group cl = cg::this_cluster();
  const int rank = (int)cl.block_rank();
  if (threadIdx.x == 0) {
    unsigned id;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(id));
    smid[blockIdx.x] = id;
  }
  float4* base = reinterpret_cast<float4*>(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float4 v = make_float4(threadIdx.x, rank, 1.f, 2.f);
  cl.sync();
  for (int it = 0; it < iters; ++it) {
    // one "row" of this CTA = csize pieces of 512 B, piece p goes to peer (rank + p) mod csize
#pragma unroll 4
    for (int p = 0; p < 8; ++p) {
      const int peer = remote ? (rank + p + warp) % csize : rank;
      float4* dst = cl.map_shared_rank(base, peer) + ((it * 8 + p) * 8 + warp) % (kTileBytes / 512) * 32 + lane;
      *dst = v;
      v.x += 1.f;
    }
  }
  cl.sync();
  if (threadIdx.x == 0) sink[blockIdx.x] = base[0].x;
}

int main() {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  unsigned* smid; float* sink;
  cudaMalloc(&smid, 4096 * 4); cudaMalloc(&sink, 4096 * 4);
  cudaFuncSetAttribute(scatter, cudaFuncAttributeMaxDynamicSharedMemorySize, kTileBytes);
  cudaFuncSetAttribute(scatter, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  printf("{\"sms\": %d, \"results\": [", sms);
  bool first = true;
  const int sizes[5] = {1, 2, 4, 8, 16};
  for (int csize : sizes) {
    for (int remote = 0; remote < 2; ++remote) {
      if (csize == 1 && remote) continue;
      cudaLaunchConfig_t cfg = {};
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = csize; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = kTileBytes;
      // how many clusters are co-resident with 1 CTA per SM?
      int nclusters = 0;
      cfg.gridDim = dim3(csize * 64);
      cudaOccupancyMaxActiveClusters(&nclusters, scatter, &cfg);
      if (nclusters < 1) { cudaGetLastError(); continue; }
      const int grid = nclusters * csize;
      cfg.gridDim = dim3(grid);
      const int iters = 4096;
      cudaMemset(smid, 0xff, 4096 * 4);
      cudaLaunchKernelEx(&cfg, scatter, csize, remote, 64, smid, sink);     // warm-up
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0); cudaEventCreate(&e1);
      cudaEventRecord(e0);
      cudaLaunchKernelEx(&cfg, scatter, csize, remote, iters, smid, sink);
      cudaEventRecord(e1);
      cudaError_t err = cudaEventSynchronize(e1);
      float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
      unsigned h[4096];
      cudaMemcpy(h, smid, grid * 4, cudaMemcpyDeviceToHost);
      std::set<unsigned> used(h, h + grid);
      const double bytes = (double)grid * kThreads * 16.0 * 8 * iters;
      printf("%s{\"cluster\": %d, \"remote\": %d, \"ctas\": %d, \"sms_used\": %d, \"ms\": %.4f, \"GBps_total\": %.1f, "
             "\"GBps_per_sm\": %.2f, \"err\": \"%s\"}", first ? "" : ", ", csize, remote, grid, (int)used.size(), ms,
             bytes / ms / 1e6, bytes / ms / 1e6 / grid, cudaGetErrorString(err));
      first = false;
    }
  }
  printf("]}\n");
  return 0;
}
