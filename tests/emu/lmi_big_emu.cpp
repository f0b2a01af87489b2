// CPU build of rayen_b200/csrc/lmi_big.cuh (contraction GEMM + one-CTA-per-sample eigen-solver of the big-LMI path)
// under the host SIMT emulator (tests/emu/simt_emu.h).  Test infrastructure only; the product never loads it.
#define RAYEN_EMU 1
#include <algorithm>
#include "simt_emu.h"

#include "../../include/rayen_b200.h"

using std::max;
using std::min;

namespace rayen {
__attribute__((aligned(16))) float lmib_smem[1 << 18];
constexpr int kFamShift = 24;
inline int make_tag(int fam, int idx) { return (fam << kFamShift) | idx; }
inline int tag_family(int tag) { return tag >> kFamShift; }
constexpr float kNormEps = 1e-12f;
}  // namespace rayen

#include "../../rayen_b200/csrc/lmi_big.cuh"

using namespace rayen;

extern "C" int emu_lmib_contract(const float* V, long long ldv, const float* F, int n, int p4, float* C, long long Bc, const float* C0) {
  const int tiles_n = (p4 + kLbTileN - 1) / kLbTileN;
  const int tiles_m = static_cast<int>((Bc + kLbTileM - 1) / kLbTileM);
  emu_launch(tiles_n * tiles_m, kLbGemmThreads, [&] { lmib_contract_kernel(V, ldv, F, n, p4, C, Bc, C0); });
  return 0;
}

extern "C" int emu_lmib_solve(const float* blob, int n, int k, int r, int p4, int off_lmib, int off_y0, const float* S,
                              const float* v, long long ldv, float* y, float* kappa_io, int* active_io, float* dkappa,
                              long long Bc, int mode, int flags, int threads, int grid, float* Sw, int* grad_list,
                              int* grad_count) {
  LmiBigDev P{};
  P.blob = blob; P.n = n; P.k = k; P.r = r; P.p4 = p4; P.off_lmib = off_lmib; P.off_y0 = off_y0;
  if (lmib_smem_bytes(r) > sizeof(lmib_smem)) return -2;
  if (threads < r) return -4;
#define LB_RUN(T) emu_launch(grid, T, [&] { lmib_solve_kernel<T>(P, S, v, ldv, y, kappa_io, active_io, dkappa, Bc, mode, flags, Sw, grad_list, grad_count); })
  if (threads == 64) LB_RUN(64);
  else if (threads == 128) LB_RUN(128);
  else if (threads == 256) LB_RUN(256);
  else if (threads == 320) LB_RUN(320);
  else return -3;
  return 0;
}

extern "C" int emu_lmib_grad_gemm(const float* Sw, const int* list, const int* count, const float* F, int n, int p4,
                                  float* dkappa, long long Bc) {
  const int tiles_n = (n + kLbGradTile - 1) / kLbGradTile;
  const int tiles_m = static_cast<int>((Bc + kLbGradTile - 1) / kLbGradTile);
  emu_launch(tiles_n * tiles_m, kLbGradThreads, [&] { lmib_grad_gemm_kernel(Sw, list, count, F, n, p4, dkappa); });
  return 0;
}
