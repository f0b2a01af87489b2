// Probe: one tcgen05 TF32 GEMM  D[128 x N] = A[128 x K] * B[N x K]^T  (K-major operands, no swizzle),
// 3xTF32 split, result read back with tcgen05.ld and compared with a double-precision CPU product.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <stdint.h>
#include <vector>

constexpr int M = 128, N = 96, K = 32;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // version = 1 (Blackwell)
  return d;                // layout_type = 0 (no swizzle), base_offset = 0
}
__device__ __forceinline__ float tf32_rna(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// operand tile in smem: [kc = k/4][row group g = r/8][r%8][4 floats]
__device__ __forceinline__ int op_off(int rows, int r, int k) { return ((k >> 2) * (rows >> 3) + (r >> 3)) * 32 + (r & 7) * 4 + (k & 3); }

__global__ void probe(const float* A, const float* B, float* D) {
  extern __shared__ __align__(1024) unsigned char smem[];
  float* a_hi = (float*)smem;                 // M*K
  float* a_lo = a_hi + M * K;
  float* b_hi = a_lo + M * K;                 // N*K
  float* b_lo = b_hi + N * K;
  uint64_t* bar = (uint64_t*)(b_lo + N * K);
  uint32_t* tmem_slot = (uint32_t*)(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;

  for (int i = tid; i < M * K; i += blockDim.x) {
    int r = i / K, k = i % K;
    float x = A[i], h = tf32_rna(x);
    a_hi[op_off(M, r, k)] = h;
    a_lo[op_off(M, r, k)] = tf32_rna(x - h);
  }
  for (int i = tid; i < N * K; i += blockDim.x) {
    int r = i / K, k = i % K;
    float x = B[i], h = tf32_rna(x);
    b_hi[op_off(N, r, k)] = h;
    b_lo[op_off(N, r, k)] = tf32_rna(x - h);
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(smem_u32(tmem_slot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;

  if (tid == 0) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    const uint32_t lbo_a = (M / 8) * 128, lbo_b = (N / 8) * 128, sbo = 128;
    int first = 1;
    for (int ks = 0; ks < K / 8; ++ks) {
      for (int term = 0; term < 3; ++term) {
        const float* pa = (term == 1) ? a_lo : a_hi;
        const float* pb = (term == 2) ? b_lo : b_hi;
        const uint64_t da = make_desc(smem_u32(pa) + 2 * ks * lbo_a, lbo_a, sbo);
        const uint64_t db = make_desc(smem_u32(pb) + 2 * ks * lbo_b, lbo_b, sbo);
        const uint32_t acc = first ? 0u : 1u;
        asm volatile(
            "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem),
            "l"(da), "l"(db), "r"(idesc), "r"(acc)
            : "memory");
        first = 0;
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
  }
  // everyone waits for the MMAs
  {
    uint32_t ok = 0;
    while (!ok) {
      asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}\n"
                   : "=r"(ok) : "r"(smem_u32(bar)) : "memory");
    }
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (warp < 4) {
    const int row = warp * 32 + (tid & 31);
    for (int c0 = 0; c0 < N; c0 += 16) {
      uint32_t r[16];
      const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
            "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
          : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int j = 0; j < 16; ++j) D[row * N + c0 + j] = __uint_as_float(r[j]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem));
}

int main() {
  std::vector<float> A(M * K), B(N * K), D(M * N);
  srand(1);
  for (auto& x : A) x = (rand() / (float)RAND_MAX) * 2 - 1;
  for (auto& x : B) x = ((rand() / (float)RAND_MAX) * 2 - 1) * 3;
  float *dA, *dB, *dD;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  cudaMemset(dD, 0, D.size() * 4);
  size_t smem = (2 * M * K + 2 * N * K) * 4 + 64;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  probe<<<1, 128, smem>>>(dA, dB, dD);
  cudaError_t e = cudaDeviceSynchronize();
  printf("kernel: %s\n", cudaGetErrorString(e));
  cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
  double worst = 0, scale = 0;
  for (int i = 0; i < M; ++i)
    for (int j = 0; j < N; ++j) {
      double ref = 0;
      for (int k = 0; k < K; ++k) ref += (double)A[i * K + k] * (double)B[j * K + k];
      worst = fmax(worst, fabs(ref - D[i * N + j]));
      scale = fmax(scale, fabs(ref));
    }
  printf("max abs err %.3e  (max |ref| %.3f)  rel %.3e\n", worst, scale, worst / scale);
  printf("D[0][0..3] = %f %f %f %f\n", D[0], D[1], D[2], D[3]);
  return 0;
}
