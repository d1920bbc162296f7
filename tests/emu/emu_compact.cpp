// TEST INFRASTRUCTURE -- runs the source of the operator compaction kernels (csrc/compact.cu) on the
// host. Nothing here is linked into the product libraries; the product path never sees PTB_HOST_EMU.
#define PTB_HOST_EMU 1
#include <algorithm>
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

#include <cuda_runtime.h>
#ifndef __launch_bounds__
#define __launch_bounds__(...)
#endif

struct EmuIdx
{
  unsigned x = 0, y = 0, z = 0;
};
static thread_local EmuIdx threadIdx, blockIdx;
static EmuIdx blockDim, gridDim;
static std::barrier<>* emu_barrier = nullptr;
static std::barrier<>* emu_warp[32];
static int emu_xi[32][32];
static inline void __syncthreads() { emu_barrier->arrive_and_wait(); }
static inline unsigned __ballot_sync(unsigned, bool pred)
{
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  emu_xi[w][l] = pred ? 1 : 0;
  emu_warp[w]->arrive_and_wait();
  unsigned m = 0;
  for (int i = 0; i < 32; ++i)
    m |= emu_xi[w][i] ? 1u << i : 0u;
  emu_warp[w]->arrive_and_wait();
  return m;
}
using std::max;
using std::min;

// the fixture passes a copy of csrc/compact.cu in which `__shared__` reads `static` (one CTA at a time)
#include PTB_EMU_COMPACT_SOURCE

namespace
{
template <typename K, typename... Args>
void emu_launch(K kernel, unsigned grid, unsigned block, Args... args)
{
  gridDim.x = grid, blockDim.x = block;
  for (unsigned b = 0; b < grid; ++b)
  {
    std::barrier<> bar(block);
    emu_barrier = &bar;
    std::vector<std::unique_ptr<std::barrier<>>> wb;
    for (unsigned w = 0; w < (block + 31) / 32; ++w)
    {
      wb.push_back(std::make_unique<std::barrier<>>(32));
      emu_warp[w] = wb.back().get();
    }
    std::vector<std::thread> th;
    for (unsigned t = 0; t < block; ++t)
      th.emplace_back([=] {
        threadIdx.x = t, blockIdx.x = b;
        kernel(args...);
        emu_barrier->arrive_and_drop();
      });
    for (auto& x : th)
      x.join();
  }
}
} // namespace

extern "C" {

// stage 1: counts + both scans (the caller sizes the compacted arrays from mat_off_z[S], xoff_z[S])
int emu_compact_offsets(int32_t n_rows, int32_t n_slices, const int64_t* mat_off, const double* vals,
                        const int32_t* cdelta, const double* dinv, double tol, int64_t* cnt_w,
                        int64_t* cnt_x, int64_t* mat_off_z, int64_t* xoff_z)
{
  using namespace ptb;
  emu_launch(compact_count, (n_slices + 7) / 8, CP_THREADS, n_rows, n_slices, mat_off, vals, cdelta, dinv,
             tol, cnt_w, cnt_x);
  emu_launch(scan_exclusive, 1, 1024, static_cast<std::int64_t>(n_slices), (const std::int64_t*)cnt_w, mat_off_z);
  emu_launch(scan_exclusive, 1, 1024, static_cast<std::int64_t>(n_slices), (const std::int64_t*)cnt_x, xoff_z);
  return 0;
}

int emu_compact_copy(int32_t n_rows, int32_t n_slices, const int64_t* mat_off, const double* vals,
                     const int32_t* cdelta, const double* dinv, double tol, const int32_t* colsx,
                     const int64_t* xoff,
                     const int64_t* mat_off_z, const int64_t* xoff_z, double* vals_z,
                     int32_t* cdelta_z, int32_t* colsx_z)
{
  using namespace ptb;
  emu_launch(compact_copy, (n_slices + 7) / 8, CP_THREADS, n_rows, n_slices, mat_off, vals, cdelta, dinv, tol,
             colsx, xoff, mat_off_z, xoff_z, vals_z, cdelta_z, colsx_z);
  return 0;
}
}
