// Zero-column compaction of the assembled scalar operator for the SpMV (opt-in, PTB_SPMV_COMPACT=1).
//
// The reference's sparsity pattern (fem::create_matrix, poisson_problem.cpp:122-123) holds every
// (row, column) coupled through a cell, whatever the value. On the reference's own mesh -- the
// Kuhn split of a lattice -- 8 of the 15 stored entries of an interior P1 Poisson row are exact
// zeros (the operator is the 7-point stencil, KAT K3), and PETSc multiplies them like any other.
// After assembly this pass builds a second SELL-32 copy of the matrix in which a position k of a
// slice is dropped when all 32 rows hold 0.0 there. Dropping whole positions keeps the
// translation-invariant column deltas of layout.h intact, so the SpMV kernels run unchanged on
// the compacted arrays; y is bit-identical (the dropped terms were +-0 * p). An optional tolerance
// (PTB_SPMV_COMPACT_TOL, relative to the row's diagonal) also drops rounding residue of analytic
// zeros; then y changes at that level. The assembled values
// the C ABI hands out (ptb_get_matrix_values) stay the full pattern.
//   count   one warp per slice: kept positions / kept explicit positions
//   scan    exclusive prefix sums -> mat_off, xoff of the compacted copy
//   copy    one warp per slice: values, deltas and explicit column indices of the kept positions
// Measured on the B200 in round 2 (Poisson 20 M DOFs: 0.505 -> 0.388 ms per SpMV, profiles/r02/
// bench_poisson20M_compact*.json) and kept opt-in; tests/emu also runs it on the host
// against the uncompacted operator.
#include "kernels.h"
#include <climits>
#include <cstdlib>

namespace ptb
{
namespace
{

constexpr int CP_THREADS = 256;

// A position survives when some row of the slice holds more than `floor` there; floor = tol * |a_rr|
// of the lane's row (tol = 0: exact zeros only).
__device__ __forceinline__ bool position_kept(const double* __restrict__ vals, std::int64_t mo, int k,
                                              int lane, double floor)
{
  return __ballot_sync(0xffffffffu, fabs(vals[mo + static_cast<std::int64_t>(k) * 32 + lane]) > floor) != 0u;
}
__device__ __forceinline__ double row_floor(const double* __restrict__ dinv, std::int32_t n_rows,
                                            std::int32_t s, int lane, double tol)
{
  const std::int32_t row = s * 32 + lane;
  return tol > 0.0 && row < n_rows ? tol / fabs(dinv[row]) : 0.0;
}

__global__ void __launch_bounds__(CP_THREADS)
compact_count(std::int32_t n_rows, std::int32_t n_slices, const std::int64_t* __restrict__ mat_off,
              const double* __restrict__ vals, const std::int32_t* __restrict__ cdelta,
              const double* __restrict__ dinv, double tol, std::int64_t* __restrict__ cnt_w,
              std::int64_t* __restrict__ cnt_x)
{
  const int lane = threadIdx.x & 31;
  const std::int32_t s = blockIdx.x * (CP_THREADS / 32) + (threadIdx.x >> 5);
  if (s >= n_slices)
    return;
  const std::int64_t mo = mat_off[s];
  const int w = static_cast<int>((mat_off[s + 1] - mo) >> 5);
  const double floor = row_floor(dinv, n_rows, s, lane, tol);
  int kept = 0, kept_x = 0;
  for (int k = 0; k < w; ++k)
    if (position_kept(vals, mo, k, lane, floor))
    {
      ++kept;
      kept_x += cdelta[(mo >> 5) + k] == INT32_MIN ? 1 : 0;
    }
  if (lane == 0)
  {
    cnt_w[s] = 32 * static_cast<std::int64_t>(kept);
    cnt_x[s] = 32 * static_cast<std::int64_t>(kept_x);
  }
}

// Exclusive prefix sum of n values into out[0..n] (out[n] = total), one CTA: every thread sums a
// contiguous chunk, the chunk sums are scanned in shared memory, every thread writes its chunk.
__global__ void __launch_bounds__(1024)
scan_exclusive(std::int64_t n, const std::int64_t* __restrict__ in, std::int64_t* __restrict__ out)
{
  __shared__ std::int64_t part[1024];
  const std::int64_t chunk = (n + blockDim.x - 1) / blockDim.x;
  const std::int64_t lo = min(n, chunk * threadIdx.x), hi = min(n, lo + chunk);
  std::int64_t sum = 0;
  for (std::int64_t i = lo; i < hi; ++i)
    sum += in[i];
  part[threadIdx.x] = sum;
  __syncthreads();
  if (threadIdx.x == 0)
  {
    std::int64_t run = 0;
    for (unsigned t = 0; t < blockDim.x; ++t)
    {
      const std::int64_t v = part[t];
      part[t] = run;
      run += v;
    }
    out[n] = run;
  }
  __syncthreads();
  std::int64_t run = part[threadIdx.x];
  for (std::int64_t i = lo; i < hi; ++i)
  {
    out[i] = run;
    run += in[i];
  }
}

__global__ void __launch_bounds__(CP_THREADS)
compact_copy(std::int32_t n_rows, std::int32_t n_slices, const std::int64_t* __restrict__ mat_off,
             const double* __restrict__ vals, const std::int32_t* __restrict__ cdelta,
             const double* __restrict__ dinv, double tol,
             const std::int32_t* __restrict__ colsx, const std::int64_t* __restrict__ xoff,
             const std::int64_t* __restrict__ mat_off_z, const std::int64_t* __restrict__ xoff_z,
             double* __restrict__ vals_z, std::int32_t* __restrict__ cdelta_z,
             std::int32_t* __restrict__ colsx_z)
{
  const int lane = threadIdx.x & 31;
  const std::int32_t s = blockIdx.x * (CP_THREADS / 32) + (threadIdx.x >> 5);
  if (s >= n_slices)
    return;
  const std::int64_t mo = mat_off[s], moz = mat_off_z[s];
  const std::int64_t xo = xoff[s], xoz = xoff_z[s];
  const int w = static_cast<int>((mat_off[s + 1] - mo) >> 5);
  const double floor = row_floor(dinv, n_rows, s, lane, tol);
  int j = 0, ix = 0, jx = 0; // kept so far; explicit positions seen / kept so far
  for (int k = 0; k < w; ++k)
  {
    const std::int32_t d = cdelta[(mo >> 5) + k];
    const bool explicit_k = d == INT32_MIN;
    if (position_kept(vals, mo, k, lane, floor))
    {
      vals_z[moz + static_cast<std::int64_t>(j) * 32 + lane] = vals[mo + static_cast<std::int64_t>(k) * 32 + lane];
      if (lane == 0)
        cdelta_z[(moz >> 5) + j] = d;
      if (explicit_k)
      {
        colsx_z[xoz + static_cast<std::int64_t>(jx) * 32 + lane] = colsx[xo + static_cast<std::int64_t>(ix) * 32 + lane];
        ++jx;
      }
      ++j;
    }
    ix += explicit_k ? 1 : 0;
  }
}

} // namespace

#ifndef PTB_HOST_EMU // host side: device build only
void compact_operator(ptb_ctx* c)
{
  if (c->bs != 1 || c->cdelta.p == nullptr)
    return;
  const std::int32_t S = c->n_slices;
  c->zcnt_w.alloc(static_cast<std::size_t>(S));
  c->zcnt_x.alloc(static_cast<std::size_t>(S));
  c->mat_off_z.alloc(static_cast<std::size_t>(S) + 1);
  c->xoff_z.alloc(static_cast<std::size_t>(S) + 1);
  const int grid = (S + CP_THREADS / 32 - 1) / (CP_THREADS / 32);
  // PTB_SPMV_COMPACT_TOL (default 0 = exact zeros only): entries below tol * |a_rr| count as zero.
  // With fused multiply-adds the analytic zeros of the lattice operator may come out as rounding
  // residue (1e-17 |a_rr|) instead of 0.0; a tolerance of 1e-14 treats them as what they are.
  const char* te = std::getenv("PTB_SPMV_COMPACT_TOL");
  const double tol = te && *te ? std::atof(te) : 0.0;
  compact_count<<<grid, CP_THREADS, 0, c->stream>>>(c->n_owned, S, c->mat_off.p, c->vals.p, c->cdelta.p,
                                                     c->dinv.p, tol, c->zcnt_w.p, c->zcnt_x.p);
  scan_exclusive<<<1, 1024, 0, c->stream>>>(S, c->zcnt_w.p, c->mat_off_z.p);
  scan_exclusive<<<1, 1024, 0, c->stream>>>(S, c->zcnt_x.p, c->xoff_z.p);
  PTB_CUDA(cudaGetLastError());
  // the totals size the compacted arrays (first assembly only: later ones reuse the allocation
  // unless the zero structure grew)
  std::int64_t tot[2];
  PTB_CUDA(cudaMemcpyAsync(&tot[0], c->mat_off_z.p + S, sizeof(std::int64_t), cudaMemcpyDeviceToHost, c->stream));
  PTB_CUDA(cudaMemcpyAsync(&tot[1], c->xoff_z.p + S, sizeof(std::int64_t), cudaMemcpyDeviceToHost, c->stream));
  PTB_CUDA(cudaStreamSynchronize(c->stream));
  if (c->vals_z.n < static_cast<std::size_t>(tot[0]))
    c->vals_z.alloc(static_cast<std::size_t>(tot[0]));
  if (c->cdelta_z.n < static_cast<std::size_t>(tot[0] / 32))
    c->cdelta_z.alloc(static_cast<std::size_t>(tot[0] / 32));
  if (c->colsx_z.n < static_cast<std::size_t>(std::max<std::int64_t>(tot[1], 1)))
    c->colsx_z.alloc(static_cast<std::size_t>(std::max<std::int64_t>(tot[1], 1)));
  compact_copy<<<grid, CP_THREADS, 0, c->stream>>>(c->n_owned, S, c->mat_off.p, c->vals.p, c->cdelta.p,
                                                    c->dinv.p, tol, c->colsx.p, c->xoff.p, c->mat_off_z.p, c->xoff_z.p, c->vals_z.p,
                                                    c->cdelta_z.p, c->colsx_z.p);
  PTB_CUDA(cudaGetLastError());
  c->launches += 4;
  c->compact_nnz = tot[0];
  c->have_compact = true;
}
#endif // PTB_HOST_EMU

} // namespace ptb
