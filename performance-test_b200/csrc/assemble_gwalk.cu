// P1 cell-vector assembly along the one-vertex-per-step star walk (layout.h SellLayout::walk1) with
// the new vertex gathered straight from the coordinate array a few steps ahead. Reference region:
//   fem::assemble_vector + bc->set   poisson_problem.cpp:150-155, elasticity_problem.cpp:221-229
//
// Nothing is staged in shared memory except the slice's column list (1.9 KB): step j issues the
// two 16-byte loads of the vertex that step j + GW_AHEAD brings in, so the gather latency is
// covered by four steps of arithmetic. Default for P1 since round 2 (PTB_VEC_GWALK=0 selects the
// staged-star kernel of assemble.cu): measured 0.586 -> 0.426 ms (Poisson, 4.0 M DOFs) and
// 0.456 -> 0.420 ms (elasticity, 1.33 M nodes), profiles/r02/assembly_ab_4M.json.
// The MATRIX and matrix-free-operator kernels of the same family were measured in the same run and
// lost against the staged star walk (Poisson matrix 0.865 vs 0.612 ms; elasticity 1.05 vs 0.91 ms
// for assemble_matrix_p1_walk3; matrix-free CG 2.5 vs 4.1 G DOF-it/s) and were removed.
#include "geom.cuh"
#include "envopt.h"
#include "kernels.h"
#include <climits>
#include <cstdlib>

namespace ptb
{
namespace
{

// step words in flight per thread: GW_CHUNK = 2 * AHEAD (the next chunk is requested AHEAD steps before its first use)
// gathers in flight per thread (steps of lookahead) = template parameter AHEAD of the kernel. Measured
// (profiles/r02/vector_gwalk_lookahead_ab.txt, Poisson 20 M DOFs): AHEAD 4 2.38 ms, AHEAD 8 3.26 ms
// (142 registers instead of 102) -> 4; PTB_GWALK_AHEAD=8 keeps the other instantiation reachable.

// A vertex in flight: the two 16-byte halves of its padded coordinates (+ the source term).
template <int NF>
struct InFlight
{
  double2 xy, z_;
  double f[NF > 0 ? NF : 1];
  std::int32_t cw; // the column word it was gathered through (top bit: flag set by the stager)
};

// Issue the gather of the vertex a later step brings in. Padding / no-load steps fetch offset 0
// (always a valid column) so the code stays branch-free.
template <int NF>
__device__ __forceinline__ InFlight<NF> gw_issue(std::uint32_t word, const std::int32_t* C,
                                                 const double* __restrict__ xdof,
                                                 const double* __restrict__ f)
{
  const bool loads = word != ADJ_INVALID_DEV && ((word >> 16) & 3u) != 3u;
  const int slot = loads ? static_cast<int>(word & 0xFFu) : 0;
  const std::int32_t cw = C[slot * 32];
  const std::int64_t col = cw & INT32_MAX; // the top bit may carry the Dirichlet flag
  const double2* p = reinterpret_cast<const double2*>(xdof + 4 * col);
  InFlight<NF> v;
  v.cw = cw;
  v.xy = __ldg(p);
  v.z_ = __ldg(p + 1);
#pragma unroll
  for (int a = 0; a < NF; ++a)
    v.f[a] = __ldg(f + col * NF + a);
  return v;
}

struct StepBits
{
  bool p0, p1, p2, compute;
  int nw, old;
};
__device__ __forceinline__ StepBits gw_decode(std::uint32_t word)
{
  const bool valid = word != ADJ_INVALID_DEV;
  const unsigned pos = valid ? (word >> 16) & 3u : 3u;
  return {pos == 0u, pos == 1u, pos == 2u, valid && ((word >> 18) & 1u) != 0u,
          static_cast<int>(word & 0xFFu), static_cast<int>((word >> 8) & 0xFFu)};
}

// ------------------------------------------------------------------------------------------
// Cell vector, P1, BS = 1 or 3: b[row*BS + a] = sum_cells |det|/120 (sum_j f_j + f_own).
// One warp = one slice, one thread = one block row: the BS components share the walk, the gathered
// coordinates and |det| (until round 2 every component had its own warp and repeated all three;
// 1.06 ms at 10 M DOFs). Warps are independent (each keeps its own copy of the column list:
// 1.9 KB), no barrier, no accumulators in shared memory.
// ------------------------------------------------------------------------------------------
template <int BS, int WARPS, int GW_AHEAD = 4>
__global__ void __launch_bounds__(WARPS * 32)
assemble_vector_p1_gwalk(VectorArgs A, const std::uint32_t* __restrict__ walk1,
                         const std::int64_t* __restrict__ walk1_off)
{
  extern __shared__ double smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int GW_CHUNK = 2 * GW_AHEAD;
  const std::int32_t slice = blockIdx.x * WARPS + warp;
  if (slice >= A.n_slices)
    return;
  const std::int64_t mo = A.mat_off[slice], so = walk1_off[slice];
  const int w = static_cast<int>((A.mat_off[slice + 1] - mo) >> 5);
  const int w1 = static_cast<int>((walk1_off[slice + 1] - so) >> 5);
  const std::int32_t row = slice * 32 + lane;
  const bool live = row < A.n_rows;
  std::int32_t* C = reinterpret_cast<std::int32_t*>(smem) + static_cast<std::size_t>(warp) * A.max_w * 32 + lane;

  const std::uint32_t word0 = w1 > 0 ? __ldg(walk1 + so + lane) : ADJ_INVALID_DEV;
  const std::uint32_t* wp = walk1 + so + 32 + lane;
  const int nsteps = w1 - 1;
  std::uint32_t wd[GW_CHUNK];
#pragma unroll
  for (int j = 0; j < GW_CHUNK; ++j)
    wd[j] = j < nsteps ? __ldg(wp + j * 32) : ADJ_INVALID_DEV;
  const bool bc_row = live && A.bc[row];
  double f_own[BS];
#pragma unroll
  for (int a = 0; a < BS; ++a)
    f_own[a] = live ? __ldg(A.f + static_cast<std::int64_t>(row) * BS + a) : 0.0;
  const Vec3 X0 = live ? load_point(A.xdof, row) : Vec3{0.0, 0.0, 0.0};
  for (int k0 = 0; k0 < w; k0 += 16)
  {
    std::int32_t c[16];
#pragma unroll
    for (int j = 0; j < 16; ++j)
      c[j] = k0 + j < w ? __ldg(A.cols + mo + (k0 + j) * 32 + lane) : -1;
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if (c[j] >= 0)
        C[(k0 + j) * 32] = c[j];
  }

  const bool valid0 = word0 != ADJ_INVALID_DEV;
  const int o0 = valid0 ? word0 & 0xFFu : 0, o1 = valid0 ? (word0 >> 8) & 0xFFu : 0,
            o2 = valid0 ? (word0 >> 16) & 0xFFu : 0;
  const std::int64_t c0 = C[o0 * 32], c1 = C[o1 * 32], c2 = C[o2 * 32];
  Vec3 e0 = load_point(A.xdof, c0) - X0;
  Vec3 e1 = load_point(A.xdof, c1) - X0;
  Vec3 e2 = load_point(A.xdof, c2) - X0;
  double f0[BS], f1[BS], f2[BS], sum[BS];
#pragma unroll
  for (int a = 0; a < BS; ++a)
  {
    f0[a] = __ldg(A.f + c0 * BS + a), f1[a] = __ldg(A.f + c1 * BS + a), f2[a] = __ldg(A.f + c2 * BS + a);
    sum[a] = 0.0;
  }
  InFlight<BS> q[GW_AHEAD];
#pragma unroll
  for (int j = 0; j < GW_AHEAD; ++j)
    q[j] = gw_issue<BS>(wd[j], C, A.xdof, A.f);
  auto cell = [&](bool compute) {
    const double det = dot(e0, cross(e1, e2));
    const double wgt = compute ? fabs(det) * (1.0 / 120.0) : 0.0;
#pragma unroll
    for (int a = 0; a < BS; ++a)
      sum[a] = fma(wgt, ((f_own[a] + f0[a]) + (f1[a] + f2[a])) + f_own[a], sum[a]);
  };
  cell(valid0);
  auto step = [&](std::uint32_t word, const InFlight<BS>& v) {
    const StepBits S = gw_decode(word);
    const Vec3 xn = {v.xy.x, v.xy.y, v.z_.x};
    if (S.p0)
      e0 = xn - X0;
    if (S.p1)
      e1 = xn - X0;
    if (S.p2)
      e2 = xn - X0;
#pragma unroll
    for (int a = 0; a < BS; ++a)
    {
      f0[a] = S.p0 ? v.f[a] : f0[a];
      f1[a] = S.p1 ? v.f[a] : f1[a];
      f2[a] = S.p2 ? v.f[a] : f2[a];
    }
    cell(S.compute);
  };
  for (int k0 = 0; k0 < nsteps; k0 += GW_CHUNK)
  {
    std::uint32_t nx[GW_CHUNK];
#pragma unroll
    for (int j = 0; j < GW_CHUNK; ++j)
      nx[j] = k0 + GW_CHUNK + j < nsteps ? __ldg(wp + (k0 + GW_CHUNK + j) * 32) : ADJ_INVALID_DEV;
#pragma unroll
    for (int j = 0; j < GW_CHUNK; ++j)
    {
      const InFlight<BS> cur = q[j % GW_AHEAD];
      const std::uint32_t ahead = j + GW_AHEAD < GW_CHUNK ? wd[(j + GW_AHEAD) % GW_CHUNK]
                                                          : nx[(j + GW_AHEAD) % GW_CHUNK];
      q[j % GW_AHEAD] = gw_issue<BS>(ahead, C, A.xdof, A.f);
      step(wd[j], cur);
    }
#pragma unroll
    for (int j = 0; j < GW_CHUNK; ++j)
      wd[j] = nx[j];
  }
  if (live)
  {
#pragma unroll
    for (int a = 0; a < BS; ++a)
      A.b[static_cast<std::int64_t>(row) * BS + a] = bc_row ? 0.0 : sum[a];
  }
}

} // namespace

#ifndef PTB_HOST_EMU // launchers: device build only
namespace
{
template <int BS, int WARPS>
void launch_vector_gwalk(ptb_ctx* c, const VectorArgs& A)
{
  const std::size_t smem = static_cast<std::size_t>(c->max_w) * 32 * sizeof(std::int32_t) * WARPS;
  auto kernel = env_int("PTB_GWALK_AHEAD", 4) == 8 ? assemble_vector_p1_gwalk<BS, WARPS, 8>
                                                   : assemble_vector_p1_gwalk<BS, WARPS, 4>;
  PTB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  const std::int64_t warps = A.n_slices;
  kernel<<<static_cast<unsigned>((warps + WARPS - 1) / WARPS), WARPS * 32, smem, c->stream>>>(A, c->walk1.p, c->walk1_off.p);
}

} // namespace

bool launch_assemble_vector_gwalk(ptb_ctx* c, const VectorArgs& A)
{
  if (c->order != 1 || c->walk1.p == nullptr)
    return false;
  const int warps = env_int("PTB_GWALK_WARPS", 4);
  if (c->bs == 1)
  {
    if (warps == 1)
      launch_vector_gwalk<1, 1>(c, A);
    else if (warps == 8)
      launch_vector_gwalk<1, 8>(c, A);
    else
      launch_vector_gwalk<1, 4>(c, A);
  }
  else
  {
    if (warps == 1)
      launch_vector_gwalk<3, 1>(c, A);
    else if (warps == 8)
      launch_vector_gwalk<3, 8>(c, A);
    else
      launch_vector_gwalk<3, 4>(c, A);
  }
  PTB_CUDA(cudaGetLastError());
  c->launches += 1;
  return true;
}

#endif // PTB_HOST_EMU

} // namespace ptb
