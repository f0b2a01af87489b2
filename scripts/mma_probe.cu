// Micro-benchmark: issue -> commit -> mbarrier latency and back-to-back rate of tcgen05.mma kind::tf32 (K = 8)
// with K-major no-swizzle operands, as used by lqs_tc.cuh / lmi_tc.cuh.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I rayen_b200/csrc -o build/mma_probe scripts/mma_probe.cu
#include <cstdio>
#include "lqs_tc.cuh"
using namespace rayen;
__global__ void probe(int n_mma, int N, int poll_threads, long long* out) {
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
  uint32_t* slot = reinterpret_cast<uint32_t*>(smem + 64);
  float* a = reinterpret_cast<float*>(smem + 1024);        // 128 x 32 tf32
  float* b = a + 128 * 32;                                  // 256 x 32
  for (int i = threadIdx.x; i < (128 + 256) * 32; i += blockDim.x) a[i] = 1.0f;
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(slot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = *slot;
  const uint32_t LBO_A = 16 * 128, LBO_B = (N / 8) * 128, SBO = 128;
  const uint32_t idesc = umma_idesc_tf32(128, N);
  long long t0 = 0;
  for (int rep = 0; rep < 4; ++rep) {
    __syncthreads();
    t0 = clock64();
    if (threadIdx.x == 0) {
      for (int i = 0; i < n_mma; ++i) {
        const int ks = i & 3;
        umma_tf32(tm + (i & 1) * 256, umma_smem_desc(smem_u32(a) + 2 * ks * LBO_A, LBO_A, SBO),
                  umma_smem_desc(smem_u32(b) + 2 * ks * LBO_B, LBO_B, SBO), idesc, 1u);
      }
      umma_commit(bar);
    }
    if (threadIdx.x < poll_threads) mbar_wait(bar, rep & 1);
    tc_fence_after();
    if (threadIdx.x == 0) out[rep] = clock64() - t0;
  }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm));
}
int main() {
  long long* out; cudaMalloc(&out, 64);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 60000);
  long long h[4];
  for (int N : {64, 96, 128, 256})
    for (int poll : {1, 256})
      for (int n : {1, 3, 12, 48, 192}) {
        probe<<<1, 256, 60000>>>(n, N, poll, out);
        cudaDeviceSynchronize();
        cudaMemcpy(h, out, 32, cudaMemcpyDeviceToHost);
        printf("N %3d poll_threads %3d n_mma %3d: %6lld cycles (last rep)  -> %.1f cycles/MMA  [%s]\n", N, poll, n, h[3], (double)h[3] / n,
               cudaGetErrorString(cudaGetLastError()));
      }
  return 0;
}
