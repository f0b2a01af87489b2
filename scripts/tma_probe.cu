// Micro-benchmark: how fast does one SM (and all 148 at once) stage constants from L2 with 1-D TMA bulk copies?
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_probe scripts/tma_probe.cu && ./tma_probe
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <stdint.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// total bytes in `pieces` copies issued by `issuers` threads (round-robin), one mbarrier per piece or one for all;
// src_stride: byte offset between the sources of consecutive CTAs (0 = everybody reads the same lines)
__global__ void probe(const char* src, int total, int piece, int issuers, int per_piece_bar, long long src_stride,
                      long long* out, int reps) {
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  char* dst = reinterpret_cast<char*>(smem + 1024);
  const int n = total / piece;
  const char* my = src + blockIdx.x * src_stride;
  long long acc = 0;
  for (int r = 0; r < reps; ++r) {
    if (threadIdx.x == 0) {
      for (int i = 0; i < 64; ++i) mbar_init(&bars[i], 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const long long t0 = clock64();
    if (threadIdx.x < issuers) {
      if (!per_piece_bar) {
        if (threadIdx.x == 0) mbar_expect_tx(&bars[0], total);
      }
    }
    __syncthreads();
    if (threadIdx.x < issuers) {
      for (int i = threadIdx.x; i < n; i += issuers) {
        uint64_t* b = per_piece_bar ? &bars[i % 64] : &bars[0];
        if (per_piece_bar) mbar_expect_tx(b, piece);
        bulk_g2s(dst + (size_t)i * piece, my + (size_t)i * piece, piece, b);
      }
    }
    if (threadIdx.x == 0) {
      if (per_piece_bar) { for (int i = 0; i < n && i < 64; ++i) mbar_wait(&bars[i], 0); }
      else mbar_wait(&bars[0], 0);
      acc += clock64() - t0;
    }
    __syncthreads();
    asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_u32(&bars[0])));
    __syncthreads();
  }
  if (threadIdx.x == 0) out[blockIdx.x] = acc / reps;
}
int main() {
  const int total = 128 * 1024;
  char* src; cudaMalloc(&src, 64ll << 20); cudaMemset(src, 1, 64ll << 20);
  long long* out; cudaMalloc(&out, 148 * 8);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, total + 1024);
  long long h[148];
  printf("total = %d KB per CTA; cycles (mean over CTAs) and B/cycle/SM\n", total / 1024);
  for (int grid : {1, 148})
    for (long long stride : {0ll, (long long)total})
      for (int per_bar : {0, 1})
        for (int issuers : {1, 4})
          for (int piece : {2048, 8192, 32768}) {
            if (per_bar && total / piece > 64) continue;
            probe<<<grid, 128, total + 1024>>>(src, total, piece, issuers, per_bar, stride, out, 3);  // warm L2
            probe<<<grid, 128, total + 1024>>>(src, total, piece, issuers, per_bar, stride, out, 10);
            cudaDeviceSynchronize();
            cudaMemcpy(h, out, grid * 8, cudaMemcpyDeviceToHost);
            double m = 0; for (int i = 0; i < grid; ++i) m += h[i]; m /= grid;
            printf("grid %3d src_stride %7lld bar/piece %d issuers %d piece %5d B: %8.0f cycles  %6.1f B/cyc/SM  (%s)\n", grid, stride, per_bar,
                   issuers, piece, m, total / m, cudaGetErrorString(cudaGetLastError()));
          }
  // small transfers: latency of one copy of X bytes
  for (int grid : {1, 148})
    for (int sz : {2048, 8192, 32768}) {
      probe<<<grid, 128, total + 1024>>>(src, sz, sz, 1, 0, 0, out, 10);
      cudaDeviceSynchronize();
      cudaMemcpy(h, out, grid * 8, cudaMemcpyDeviceToHost);
      double m = 0; for (int i = 0; i < grid; ++i) m += h[i]; m /= grid;
      printf("single copy grid %3d %5d B: %8.0f cycles\n", grid, sz, m);
    }
  return 0;
}
