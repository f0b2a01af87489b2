// CPU build of rayen_b200/csrc/wide.cuh under the host SIMT emulator (tests/emu/simt_emu.h): the kernels' source is
// included as it is (tests/test_wide_emulated.py strips only its two #include lines); the handful of helpers it takes
// from common.cuh / lqs.cuh are restated below because those headers carry inline PTX.
#include "simt_emu.h"

#include "../../include/rayen_b200.h"

namespace rayen {
__attribute__((aligned(16))) float wide_smem[1 << 20];
constexpr int kFamShift = 24;
inline int make_tag(int fam, int idx) { return (fam << kFamShift) | idx; }
inline int tag_family(int tag) { return tag >> kFamShift; }
inline int tag_index(int tag) { return tag & ((1 << kFamShift) - 1); }
constexpr float kNormEps = 1e-12f;
inline void group_argmax(float& val, int& tag, int width) {   // common.cuh
  for (int off = width >> 1; off > 0; off >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, val, off);
    const int ot = __shfl_xor_sync(0xffffffffu, tag, off);
    if (ov > val || (ov == val && ot < tag)) {
      val = ov;
      tag = ot;
    }
  }
}
inline float soc_root(float A, float hb, float cq, float* root_out) {   // lqs.cuh
  const float disc = fmaxf(fmaf(hb, hb, A * cq), 0.f);
  const float root = sqrtf(disc);
  if (root_out) *root_out = root;
  if (hb >= 0.f) return (hb + root) / A;
  const float den = root - hb;
  return cq / den;
}
}  // namespace rayen

#include "wide_stripped.cuh"

using namespace rayen;

static WideDev make_dev(const float* blob, const int* h, int n, int k, int off_y0, int n_is_identity) {
  WideDev w{};
  w.blob = blob; w.n = n; w.k = k;
  w.r_pad = h[1]; w.n_tasks = h[2]; w.off_tasks = h[3]; w.off_wt = h[4]; w.off_nt = h[5]; w.off_nrow = h[6];
  w.k32 = h[7]; w.np = h[8]; w.off_items = h[9]; w.n_quad = h[10]; w.n_soc = h[11]; w.n_rounds = h[12]; w.off_rounds = h[13];
  w.off_y0 = off_y0; w.n_is_identity = n_is_identity;
  return w;
}

extern "C" int emu_wide_forward(const float* blob, long long off_wide, int n, int k, int off_y0, int n_is_identity,
                                const float* v, long long ldv, float* y, float* kappa, int* active, long long B, int mode,
                                int grid, int ts) {
  const int* h = reinterpret_cast<const int*>(blob + off_wide);
  if (h[0] != kWideMagic || h[14] != kWideVersion) return -1;
  if (wide_fwd_smem_bytes(n, ts) > sizeof(wide_smem)) return -2;
  const WideDev w = make_dev(blob, h, n, k, off_y0, n_is_identity);
  if (ts == 16)
    emu_launch(grid, kWideThreads, [&] { if (n >= kWideBlockedN) wide_forward_kernel<16, true>(w, v, ldv, y, kappa, active, B, mode); else wide_forward_kernel<16, false>(w, v, ldv, y, kappa, active, B, mode); });
  else if (ts == 8)
    emu_launch(grid, kWideThreads, [&] { if (n >= kWideBlockedN) wide_forward_kernel<8, true>(w, v, ldv, y, kappa, active, B, mode); else wide_forward_kernel<8, false>(w, v, ldv, y, kappa, active, B, mode); });
  else if (ts == 4)
    emu_launch(grid, kWideThreads, [&] { wide_forward_kernel<4, true>(w, v, ldv, y, kappa, active, B, mode); });
  else
    return -3;
  return 0;
}

extern "C" int emu_wide_backward(const float* blob, long long off_wide, int n, int k, int off_y0, int n_is_identity,
                                 const float* v, long long ldv, const float* gy, const float* kappa, const int* active,
                                 float* gv, long long ldgv, long long B, int mode, int grid, int threads) {
  const int* h = reinterpret_cast<const int*>(blob + off_wide);
  if (h[0] != kWideMagic) return -1;
  const WideDev w = make_dev(blob, h, n, k, off_y0, n_is_identity);
  if (threads == 128)
    emu_launch(grid, 128, [&] { wide_backward_kernel<128>(w, v, ldv, gy, kappa, active, gv, ldgv, B, mode, nullptr); });
  else if (threads == 256)
    emu_launch(grid, 256, [&] { wide_backward_kernel<256>(w, v, ldv, gy, kappa, active, gv, ldgv, B, mode, nullptr); });
  else
    return -3;
  return 0;
}
