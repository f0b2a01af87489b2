// CPU build of the PTX-free part of rayen_b200/csrc/lmi_warp.cuh (the definiteness filter and the one-warp-per-matrix
// eigen-solver) under the host SIMT emulator (tests/emu/simt_emu.h).  Test infrastructure only.
#define RAYEN_EMU 1
#include <algorithm>
#include <type_traits>
#include "simt_emu.h"

#include "../../include/rayen_b200.h"

using std::max;
using std::min;

namespace rayen {
constexpr int kFamShift = 24;
inline int make_tag(int fam, int idx) { return (fam << kFamShift) | idx; }
inline int tag_family(int tag) { return tag >> kFamShift; }
}  // namespace rayen

#include "../../rayen_b200/csrc/lmi_warp.cuh"

using namespace rayen;

// One block of `warps` warps; warp w takes chunks w, w + warps, ... of 4 consecutive samples (dense mode) or of the
// work list.  FW: n matrices of 32 rows x 36 words.
extern "C" int emu_lmi_warp(const float* FW, int n, int k, const float* y0, const float* nmat, int nstride, int n_is_identity,
                            int mode, const float* v, long long ldv, float* y, float* kappa_io, int* active_io,
                            float* dkappa, long long B, const int* work_list, long long n_list, int use_filter,
                            int with_grad, int warps, int solves_per_warp, int* fail_list, int* fail_count, int mt) {
  static float scratch[32 * kLwScratch];
  if (warps < 1 || warps > 32) return -1;
  const long long total = work_list ? n_list : B;
  if (mt != 2 && mt != kLwMT) return -2;
  const long long n_chunks = (total + mt - 1) / mt;
  auto body = [&](auto tag) {
    constexpr int MT = decltype(tag)::value;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    LwCtx C;
    C.FW = FW; C.scr = scratch + warp * kLwScratch; C.y0 = y0; C.nmat = nmat;
    C.n = n; C.k = k; C.nstride = nstride; C.n_is_identity = n_is_identity; C.mode = mode;
    int solve_budget = solves_per_warp;
    for (long long c = warp; c < n_chunks; c += warps) {
      long long b[MT];
      bool valid[MT];
      for (int m = 0; m < MT; ++m) {
        const long long idx = c * MT + m;
        valid[m] = idx < total;
        b[m] = valid[m] ? (work_list ? work_list[idx] : idx) : 0;
      }
      if (with_grad)
        lw_process_chunk<true, MT>(C, b, valid, v, ldv, y, kappa_io, active_io, dkappa, lane, use_filter != 0, solve_budget,
                                   fail_list, fail_count);
      else
        lw_process_chunk<false, MT>(C, b, valid, v, ldv, y, kappa_io, active_io, dkappa, lane, use_filter != 0, solve_budget,
                                    fail_list, fail_count);
    }
  };
  emu_launch(1, warps * 32, [&] {
    if (mt == 2) body(std::integral_constant<int, 2>{});
    else body(std::integral_constant<int, kLwMT>{});
  });
  return 0;
}
