// Phase-latency probe of lmi_warp.cuh (development aid): one warp per CTA, one CTA per SM -- the situation of a short
// work list -- on random symmetric matrices; prints the median cycles of every phase.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I rayen_b200/csrc -o scripts/bin/lw_probe scripts/lw_probe.cu
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "lmi_warp.cuh"

using namespace rayen;

__global__ void __launch_bounds__(256, 1) probe_kernel(const float* __restrict__ FWg, int n, const float* __restrict__ usg,
                                                       long long* __restrict__ out, float* __restrict__ sink, int warps) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* fw = reinterpret_cast<float*>(smem_raw);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < n * kLwMatWords; i += blockDim.x) fw[i] = FWg[i];
  float* scr = fw + n * kLwMatWords + warp * kLwScratch;
  for (int i = lane; i < 128; i += 32) scr[kLwScrUs + i] = usg[i];
  __syncthreads();
  if (warp >= warps) return;
  long long t[10];
  float kprior[kLwMT] = {1e9f, 1e9f, 1e9f, 1e9f};
  t[0] = clock64();
  LwFilter F;
  F.contract(fw, n, scr + kLwScrUs, lane);
  t[1] = clock64();
  const unsigned pass = F.passes(kprior, scr + kLwScrCb, lane);
  t[2] = clock64();
  LwSolver<true> S;
  S.scr = scr;
  S.lane = lane;
  S.contract_one(fw, n, scr + kLwScrUs, 1);
  t[3] = clock64();
  S.tridiagonalize();
  t[4] = clock64();
  const float lam = S.lambda_max_relu();
  t[5] = clock64();
  const float qj = S.eigenvector(lam);
  t[6] = clock64();
  const float g = S.eig_gradient(fw, n, qj);
  t[7] = clock64();
  if (lane == 0) {
    long long* o = out + (blockIdx.x * 8 + warp) * 8;
    for (int i = 0; i < 7; ++i) o[i] = t[i + 1] - t[i];
    o[7] = pass;
  }
  sink[(blockIdx.x * 8 + warp) * 32 + lane] = g + lam;
}

int main(int argc, char** argv) {
  const int n = 32;
  const int warps = argc > 1 ? atoi(argv[1]) : 1;
  std::vector<float> FW(n * kLwMatWords, 0.f), us(128);
  srand(1);
  for (int a = 0; a < n; ++a)
    for (int i = 0; i < 32; ++i)
      for (int j = 0; j <= i; ++j) {
        const float x = (rand() / (float)RAND_MAX * 2.f - 1.f);
        FW[a * kLwMatWords + i * kLwRowStride + j] = x;
        FW[a * kLwMatWords + j * kLwRowStride + i] = x;
      }
  for (auto& x : us) x = (rand() / (float)RAND_MAX * 2.f - 1.f) * 0.2f;
  float *dF, *dU, *dS;
  long long* dO;
  cudaMalloc(&dF, FW.size() * 4);
  cudaMalloc(&dU, 512);
  cudaMalloc(&dS, 148 * 8 * 32 * 4);
  cudaMalloc(&dO, 148 * 8 * 8 * 8);
  cudaMemcpy(dF, FW.data(), FW.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dU, us.data(), 512, cudaMemcpyHostToDevice);
  const size_t smem = n * kLwMatWords * 4 + 8 * kLwScratch * 4;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int rep = 0; rep < 3; ++rep) probe_kernel<<<148, 256, smem>>>(dF, n, dU, dO, dS, warps);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
  std::vector<long long> o(148 * 8 * 8);
  cudaMemcpy(o.data(), dO, o.size() * 8, cudaMemcpyDeviceToHost);
  const char* names[7] = {"contract4", "ldlt4", "contract1", "tridiag", "sturm", "eigvec", "grad"};
  printf("%d warp(s) per SM, n = %d: median cycles per phase (warp 0 of every CTA)\n", warps, n);
  for (int ph = 0; ph < 7; ++ph) {
    std::vector<long long> v;
    for (int b = 0; b < 148; ++b) v.push_back(o[(b * 8) * 8 + ph]);
    std::sort(v.begin(), v.end());
    printf("  %-10s %8lld\n", names[ph], v[74]);
  }
  return 0;
}
