// Launch-floor probe (development aid): what a chain of small kernels costs on this GPU, as a function of the dynamic
// shared memory each asks for (a change of the L1/shared carve-out between consecutive kernels drains the SMs) and of
// programmatic dependent launch.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/bin/launch_probe scripts/launch_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>

__global__ void empty_kernel(int* sink, int pdl) {
  extern __shared__ unsigned char smem[];
  if (pdl) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
  }
  if (sink == reinterpret_cast<int*>(1)) sink[threadIdx.x] = smem[threadIdx.x];
}

static void launch(int grid, int block, size_t smem, cudaStream_t st, bool pdl) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  int* sink = nullptr;
  int flag = pdl ? 1 : 0;
  cudaLaunchKernelEx(&cfg, empty_kernel, sink, flag);
}

static float time_chain(const std::vector<size_t>& smems, const std::vector<int>& blocks, int reps, bool pdl, bool graph) {
  cudaStream_t st;
  cudaStreamCreate(&st);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaGraphExec_t exec = nullptr;
  if (graph) {
    cudaGraph_t g;
    cudaStreamBeginCapture(st, cudaStreamCaptureModeGlobal);
    for (size_t i = 0; i < smems.size(); ++i) launch(148, blocks[i], smems[i], st, pdl);
    cudaStreamEndCapture(st, &g);
    cudaGraphInstantiate(&exec, g, 0);
  }
  auto run = [&](int n) {
    for (int r = 0; r < n; ++r) {
      if (graph) cudaGraphLaunch(exec, st);
      else for (size_t i = 0; i < smems.size(); ++i) launch(148, blocks[i], smems[i], st, pdl);
    }
  };
  run(50);
  cudaStreamSynchronize(st);
  cudaEventRecord(e0, st);
  run(reps);
  cudaEventRecord(e1, st);
  cudaStreamSynchronize(st);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaStreamDestroy(st);
  return ms * 1e3f / reps;
}

int main() {
  cudaFuncSetAttribute(empty_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  struct Case { const char* name; std::vector<size_t> smems; std::vector<int> blocks; };
  std::vector<Case> cases = {
      {"1 kernel, no smem", {0}, {256}},
      {"1 kernel, 193 KB", {193 * 1024}, {352}},
      {"3 kernels, no smem", {0, 0, 0}, {352, 256, 128}},
      {"3 kernels, 193 KB each", {193 * 1024, 193 * 1024, 193 * 1024}, {352, 256, 128}},
      {"3 kernels, 193 / 150 / 18 KB (the cfg5 step)", {193 * 1024, 150 * 1024, 18 * 1024}, {352, 256, 128}},
      {"3 kernels, 193 / 193 / 0 KB", {193 * 1024, 193 * 1024, 0}, {352, 256, 128}},
      {"2 kernels, 193 / 18 KB", {193 * 1024, 18 * 1024}, {352, 128}},
  };
  printf("%-48s %10s %10s %10s %10s\n", "chain (us per chain)", "plain", "pdl", "graph", "graph+pdl");
  for (auto& c : cases) {
    const float a = time_chain(c.smems, c.blocks, 2000, false, false);
    const float b = time_chain(c.smems, c.blocks, 2000, true, false);
    const float g = time_chain(c.smems, c.blocks, 2000, false, true);
    const float h = time_chain(c.smems, c.blocks, 2000, true, true);
    printf("%-48s %10.2f %10.2f %10.2f %10.2f\n", c.name, a, b, g, h);
  }
  return 0;
}
