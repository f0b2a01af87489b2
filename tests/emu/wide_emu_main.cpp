// Stand-alone driver of the emulated wide kernels for ThreadSanitizer (tests/test_wide_emulated.py builds it with
// -fsanitize=thread): a missing __syncthreads / __syncwarp in wide.cuh is a data race between the OS threads that play
// the CUDA threads, and TSan reports it -- the CPU stand-in for compute-sanitizer's racecheck.
//   usage: wide_emu_main <input file> <output file>
//   input : 16 int64 {n, k, off_y0, n_is_identity, off_wide, blob_words, B, cols, mode, ts, bwd_threads, grid_f, grid_b, 0, 0, 0},
//           blob [blob_words] f32, v [B x cols] f32, gy [B x k] f32
//   output: y [B x k] f32, kappa [B] f32, active [B] i32, gv [B x cols] f32
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "wide_emu.cpp"

int main(int argc, char** argv) {
  if (argc != 3) return 2;
  FILE* f = fopen(argv[1], "rb");
  if (!f) return 3;
  long long h[16];
  if (fread(h, sizeof(long long), 16, f) != 16) return 4;
  const int n = (int)h[0], k = (int)h[1], off_y0 = (int)h[2], ident = (int)h[3];
  const long long off_wide = h[4], words = h[5], B = h[6], cols = h[7];
  const int mode = (int)h[8], ts = (int)h[9], bthreads = (int)h[10], grid_f = (int)h[11], grid_b = (int)h[12];
  std::vector<float> blob(words), v(B * cols), gy(B * k), y(B * k), kap(B), gv(B * cols);
  std::vector<int> act(B);
  if (fread(blob.data(), 4, words, f) != (size_t)words || fread(v.data(), 4, B * cols, f) != (size_t)(B * cols) ||
      fread(gy.data(), 4, B * k, f) != (size_t)(B * k))
    return 5;
  fclose(f);
  int rc = emu_wide_forward(blob.data(), off_wide, n, k, off_y0, ident, v.data(), cols, y.data(), kap.data(), act.data(), B,
                            mode, grid_f, ts);
  if (rc) return 10 - rc;
  rc = emu_wide_backward(blob.data(), off_wide, n, k, off_y0, ident, v.data(), cols, gy.data(), kap.data(), act.data(),
                         gv.data(), cols, B, mode, grid_b, bthreads);
  if (rc) return 20 - rc;
  FILE* o = fopen(argv[2], "wb");
  if (!o) return 6;
  fwrite(y.data(), 4, B * k, o);
  fwrite(kap.data(), 4, B, o);
  fwrite(act.data(), 4, B, o);
  fwrite(gv.data(), 4, B * cols, o);
  fclose(o);
  return 0;
}
