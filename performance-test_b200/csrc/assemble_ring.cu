// Elasticity P1 matrix assembly, column-major along the *edge rings* (layout.h build_rings) -- third
// generation of the row-owner gather for the 3x3-block operator, same results to rounding, same
// reference region (fem::assemble_matrix + set_diagonal, elasticity_problem.cpp:203-211 with the
// tabulate_tensor of Elasticity.py:30-40).
//
// Why: assemble_matrix_p1_walk3 (assemble_walk.cu) visits the 24 cells of a vertex star once, but it
// needs three warps per slice (one per component row of the block, each repeating the geometry) and
// read-modify-writes its accumulators in shared memory: ncu (profiles/r02/ncu_walk3_elasticity_10M.csv)
// shows 4 700 issued instructions per warp for ~1 000 FP64 ones, 12 warps per SM, issue slots 56 %
// busy. Here ONE thread owns the whole block row of a vertex and the loop runs over the row's
// stored columns: for column k (neighbour j) the cells around the edge (i, j) are walked as a chain
// v_0, v_1, ... (cell t = (i, j, v_{t-1}, v_t)), so that
//   * the nine entries of T_j = sum_cells c_i (x) c_j / (6|det|) stay in registers until the block
//     is written -- no accumulator in shared memory, nothing indexed dynamically;
//   * consecutive cells share the face (i, j, v): one new edge vector (3 LDS.64) and two new
//     cofactor vectors per cell, n_b of a cell is -n_a of its predecessor;
//   * the material law is applied and the block stored (9 coalesced 256-byte lines) as soon as
//     its ring is closed; the diagonal block follows from sum_j c_j = 0:  T_ii = -sum_{j != i} T_ij.
// A cell is visited three times per row (once per non-owner vertex: 72 cell steps + 14 chain heads of
// 36 FP64 instructions on the Kuhn box against 3 x 24 steps of ~40), but a step costs ~53 issue slots
// instead of 3 x 150, and 14 one-warp CTAs per SM fit (15.1 KB of shared memory per slice: the star's
// edge vectors 11.5 KB + the slice's ring words 3.6 KB, the latter staged by one TMA bulk copy).
// Measured at 10 M DOFs: 1.02 ms against 2.26 ms (profiles/r02/assembly_ring_tma_ab.txt).
// The chains follow the mesh topology and the ascending cell order only, so every off-diagonal block
// is independent of the partition bit for bit; the diagonal block is summed in the row's local
// column order (ghost columns last) and agrees across partitions to rounding (~4 ulp of |A_ii|).
// PTB_ASM_RING=0 selects assemble_matrix_p1_walk3.
#include "geom.cuh"
#include "envopt.h"
#include "kernels.h"
#include "tma.cuh"
#include <climits>

namespace ptb
{
namespace
{

constexpr int RING_CHUNK = 8;   // star columns gathered (coordinates, Dirichlet flags) per trip of the prologue
constexpr int RING_CCHUNK = 16; // column indices loaded per trip
constexpr int RING_WCHUNK = 32; // ring words loaded per trip
constexpr std::uint32_t RING_PAD = 0x80808080u;
// L2 prefetch distance in slices: one generation of resident warps (148 SMs x 14). Measured: with
// 4096 and plain stores the 3 TB/s write stream evicted the prefetched lines before their use (ncu:
// DRAM reads 0.71 -> 1.38 GB, L2 hit rate 16 %); the values are therefore stored evict-first (st.cs).
#ifdef PTB_HOST_EMU
constexpr int RING_PF_DIST = 3; // tests/emu: tiny meshes must reach the prefetch address arithmetic
#else
constexpr int RING_PF_DIST = 2048;
#endif

// Shared memory per slice (one warp), in doubles: E [3 mw][32] edge vectors owner -> column k,
// RS [rw][32] ring words (uint32), NS [mw] chain bytes per column (uint8, padded to 8 bytes), one
// mbarrier; rounded to 16 bytes. E and RS are private per lane (column `lane`); NS is the warp's
// (one __syncwarp after the prologue).
__host__ __device__ inline int ring_smem_doubles(int mw, int rw)
{
  return (mw * 96 + rw * 16 + (mw + 7) / 8 + 1 + 1) & ~1;
}

// TMA = true: the slice's ring words -- one contiguous block of rw x 128 bytes whose global layout
// [word][lane] is the shared-memory layout -- are staged by ONE bulk copy through the TMA unit
// (cp.async.bulk, SASS UBLKCP) issued by lane 0 and awaited on an mbarrier just before the column loop,
// instead of rw LDG + STS per lane through registers.
template <int WARPS, bool TMA>
__global__ void __launch_bounds__(WARPS * 32, 14 / WARPS) // 14 slices per SM fit in shared memory on the Kuhn box
assemble_matrix_p1_ring3(MatrixArgs A, const std::uint32_t* __restrict__ ring,
                         const std::int64_t* __restrict__ ring_off, const std::uint8_t* __restrict__ ring_ns,
                         int max_rw)
{
  extern __shared__ double smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const std::int32_t slice = blockIdx.x * WARPS + warp;
  if (slice >= A.n_slices)
    return;
  const std::int64_t mo = A.mat_off[slice];
  const int w = static_cast<int>((A.mat_off[slice + 1] - mo) >> 5);
  const std::int32_t row = slice * 32 + lane;
  const bool live = row < A.n_rows;
  const int mw = A.max_w;

  double* base = smem + warp * ring_smem_doubles(mw, max_rw);
  double* E = base + lane;                                                      // E[(k*3+d)*32]
  std::uint32_t* RS = reinterpret_cast<std::uint32_t*>(base + mw * 96) + lane;  // RS[q*32]
  std::uint8_t* NS = reinterpret_cast<std::uint8_t*>(base + mw * 96 + max_rw * 16); // NS[k]
#ifndef PTB_HOST_EMU
  std::uint64_t* bar = reinterpret_cast<std::uint64_t*>(base + mw * 96 + max_rw * 16 + (mw + 7) / 8);
  if constexpr (TMA)
  {
    if (lane == 0)
    {
      mbar_init(bar, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
  }
#endif

  // ---- prologue: every global read of the slice is issued here, independent loads together -------
  // (ncu of the first version: a fifth of the stall samples sat in eight serialised load -> store trips)
  const std::int64_t ro = ring_off[slice];
  const int rw = static_cast<int>((ring_off[slice + 1] - ro) >> 5);
  const std::uint32_t* rp = ring + ro + lane;
  std::uint32_t wd[RING_WCHUNK];
  auto load_ring = [&](int q0) {
#pragma unroll
    for (int j = 0; j < RING_WCHUNK; ++j)
      wd[j] = q0 + j < rw ? __ldg(rp + (q0 + j) * 32) : RING_PAD;
  };
  auto store_ring = [&](int q0) {
#pragma unroll
    for (int j = 0; j < RING_WCHUNK; ++j)
      if (q0 + j < rw)
        RS[(q0 + j) * 32] = wd[j];
  };
  std::int32_t c[RING_CCHUNK];
  auto load_cols = [&](int k0) {
#pragma unroll
    for (int j = 0; j < RING_CCHUNK; ++j)
      c[j] = k0 + j < w ? __ldg(A.cols + mo + (k0 + j) * 32 + lane) : -1;
  };
#ifndef PTB_HOST_EMU
  if constexpr (TMA)
  {
    if (lane == 0)
    {
      mbar_expect_tx(bar, static_cast<std::uint32_t>(rw) * 128u);
      if (rw > 0)
        tma_load_1d(base + mw * 96, ring + ro, static_cast<std::uint32_t>(rw) * 128u, bar);
    }
  }
  else
#endif
    load_ring(0);
  load_cols(0);
  for (int k = lane; k < w; k += 32)
    NS[k] = __ldg(ring_ns + (mo >> 5) + k);
  const int len = live ? static_cast<int>(A.rowptr[row + 1] - A.rowptr[row]) : 0;
  const bool bc_row = live && A.bc[row];
  const Vec3 X0 = live ? load_point(A.xdof, row) : Vec3{0.0, 0.0, 0.0};
  {
    // Pull the streams of a slice one warp generation ahead into L2: ring words, column indices
    // (one 128-byte line per lane), row pointers, coordinates. Never past the end of an array.
    const std::int32_t s2 = slice + RING_PF_DIST;
    if (s2 < A.n_slices)
    {
      const std::int64_t mo2 = A.mat_off[s2], ro2 = ring_off[s2];
      const int w2 = static_cast<int>((A.mat_off[s2 + 1] - mo2) >> 5);
      const int rw2 = static_cast<int>((ring_off[s2 + 1] - ro2) >> 5);
      const std::int64_t r2 = static_cast<std::int64_t>(s2) * 32;
      for (int j = lane; j < rw2; j += 32)
        prefetch_l2(ring + ro2 + j * 32);
      for (int j = lane; j < w2; j += 32)
        prefetch_l2(A.cols + mo2 + j * 32);
      if (lane < 8 && r2 + lane * 4 < A.n_rows)
        prefetch_l2(A.xdof + (r2 + lane * 4) * 4);
      if (lane < 2 && r2 + lane * 16 < A.n_rows)
        prefetch_l2(A.rowptr + r2 + lane * 16);
    }
  }
#ifndef PTB_HOST_EMU
  if constexpr (!TMA)
#endif
  {
    store_ring(0);
    for (int q0 = RING_WCHUNK; q0 < rw; q0 += RING_WCHUNK)
    {
      load_ring(q0);
      store_ring(q0);
    }
  }
  int own = -1;
  std::uint64_t bcm0 = 0, bcm1 = 0; // bit k: column k of the row is constrained
  for (int k0 = 0; k0 < w; k0 += RING_CCHUNK)
  {
    if (k0 > 0)
      load_cols(k0);
#pragma unroll
    for (int h = 0; h < RING_CCHUNK; h += RING_CHUNK)
    {
      Vec3 x[RING_CHUNK];
      std::uint8_t b[RING_CHUNK];
#pragma unroll
      for (int j = 0; j < RING_CHUNK; ++j)
        if (c[h + j] >= 0)
        {
          x[j] = load_point(A.xdof, c[h + j]);
          b[j] = __ldg(A.bc + c[h + j]);
        }
#pragma unroll
      for (int j = 0; j < RING_CHUNK; ++j)
        if (c[h + j] >= 0)
        {
          const int k = k0 + h + j;
          const Vec3 d = x[j] - X0;
          E[(k * 3 + 0) * 32] = d.x;
          E[(k * 3 + 1) * 32] = d.y;
          E[(k * 3 + 2) * 32] = d.z;
          const std::uint64_t bit = b[j] ? std::uint64_t(1) << (k & 63) : 0;
          bcm0 |= k < 64 ? bit : 0;
          bcm1 |= k < 64 ? 0 : bit;
          own = c[h + j] == row && k < len ? k : own;
        }
    }
  }
#ifndef PTB_HOST_EMU
  if constexpr (TMA)
    mbar_wait(bar, 0);
#endif
  __syncwarp();
  auto edge = [&](int o) { return Vec3{E[(o * 3 + 0) * 32], E[(o * 3 + 1) * 32], E[(o * 3 + 2) * 32]}; };

  // The walk accumulates 6 T (the 1/6 of the element volume is folded into the material constants):
  // block[a][b] = mu (delta_ab tr T + T[b][a]) + lambda T[a][b]   (Elasticity.py:12-15, 33-39), T[a] = row a.
  // A Dirichlet row / column or a padding entry is written as zeros by zeroing the two constants
  // (every T of a valid mesh is finite).
  constexpr double mu6 = 1.0e6 / (2.0 * (1.0 + 0.3)) / 6.0;
  constexpr double lmbda6 = 1.0e6 * 0.3 / ((1.0 + 0.3) * (1.0 - 2.0 * 0.3)) / 6.0;
  auto store_block = [&](int k, Vec3 T0, Vec3 T1, Vec3 T2, bool zero) {
    const double mu = zero ? 0.0 : mu6, lmbda = zero ? 0.0 : lmbda6;
    const double tr = T0.x + T1.y + T2.z;
    const double v[9] = {mu * (tr + T0.x) + lmbda * T0.x, mu * T1.x + lmbda * T0.y, mu * T2.x + lmbda * T0.z,
                         mu * T0.y + lmbda * T1.x, mu * (tr + T1.y) + lmbda * T1.y, mu * T2.y + lmbda * T1.z,
                         mu * T0.z + lmbda * T2.x, mu * T1.z + lmbda * T2.y, mu * (tr + T2.z) + lmbda * T2.z};
    double* out = A.vals + (mo + k * 32) * 9 + lane;
#pragma unroll
    for (int e = 0; e < 9; ++e)
      store_stream(out + e * 32, v[e]);
    return Vec3{v[0], v[4], v[8]};
  };

  // ---- columns ----------------------------------------------------------------------------------
  Vec3 S0{0.0, 0.0, 0.0}, S1 = S0, S2 = S0; // sum of the off-diagonal raw tensors of the row
  int q = 0;                                // first ring word of the column
  for (int k = 0; k < w; ++k)
  {
    const int ns = NS[k];
    Vec3 T0{0.0, 0.0, 0.0}, T1 = T0, T2 = T0;
    if (ns > 1)
    {
      const Vec3 ej = edge(k);
      std::uint32_t word = RS[q * 32];
      Vec3 ea = edge(word & 0x7Fu);
      Vec3 nap = cross(ea, ej); // n_a of the step before: e_b x e_j with b = this step's a
      for (int t = 1; t < ns; ++t)
      {
        if ((t & 3) == 0)
          word = RS[(q + (t >> 2)) * 32];
        const std::uint32_t beta = (word >> (8 * (t & 3))) & 0xFFu;
        const Vec3 eb = edge(beta & 0x7Fu);
        // cell (i; j, a, b): n_j = e_a x e_b, n_a = e_b x e_j, n_b = e_j x e_a = -nap
        const Vec3 nj = cross(ea, eb), na = cross(eb, ej);
        const double det = dot(ej, nj);
        const double r = (beta & 0x80u) ? 0.0 : rcp_nr1(fabs(det));
        // c_i = -(n_j + n_a + n_b)
        const double qx = r * (nap.x - nj.x - na.x), qy = r * (nap.y - nj.y - na.y), qz = r * (nap.z - nj.z - na.z);
        T0 = Vec3{fma(qx, nj.x, T0.x), fma(qx, nj.y, T0.y), fma(qx, nj.z, T0.z)};
        T1 = Vec3{fma(qy, nj.x, T1.x), fma(qy, nj.y, T1.y), fma(qy, nj.z, T1.z)};
        T2 = Vec3{fma(qz, nj.x, T2.x), fma(qz, nj.y, T2.y), fma(qz, nj.z, T2.z)};
        ea = eb;
        nap = na;
      }
    }
    q += (ns + 3) >> 2;
    S0 = Vec3{S0.x + T0.x, S0.y + T0.y, S0.z + T0.z};
    S1 = Vec3{S1.x + T1.x, S1.y + T1.y, S1.z + T1.z};
    S2 = Vec3{S2.x + T2.x, S2.y + T2.y, S2.z + T2.z};
    const bool bc_col = ((k < 64 ? bcm0 : bcm1) >> (k & 63)) & 1u;
    store_block(k, T0, T1, T2, k >= len || bc_row || bc_col); // (the own column: zeros for now)
  }

  // ---- diagonal block: T_ii = -sum_j T_ij; Dirichlet rows -> identity ---------------------------
  Vec3 diag{1.0, 1.0, 1.0};
  if (own >= 0 && !bc_row)
    diag = store_block(own, Vec3{-S0.x, -S0.y, -S0.z}, Vec3{-S1.x, -S1.y, -S1.z}, Vec3{-S2.x, -S2.y, -S2.z}, false);
  else if (own >= 0)
  {
    double* out = A.vals + (mo + own * 32) * 9 + lane;
    out[0] = out[4 * 32] = out[8 * 32] = 1.0; // the other six entries were written as zeros by the column loop
  }
  if (live)
  {
    double* di = A.dinv + static_cast<std::int64_t>(row) * 3;
    di[0] = 1.0 / diag.x;
    di[1] = 1.0 / diag.y;
    di[2] = 1.0 / diag.z;
  }
}

} // namespace

#ifndef PTB_HOST_EMU // launcher: device build only
namespace
{
template <int WARPS, bool TMA>
bool launch_ring(ptb_ctx* c, const MatrixArgs& A)
{
  const std::size_t smem = static_cast<std::size_t>(ring_smem_doubles(c->max_w, c->ring_max_words)) * WARPS * sizeof(double);
  if (smem > 227 * 1024)
    return false;
  auto kernel = assemble_matrix_p1_ring3<WARPS, TMA>;
  PTB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  kernel<<<(A.n_slices + WARPS - 1) / WARPS, WARPS * 32, smem, c->stream>>>(A, c->ring.p, c->ring_off.p,
                                                                          c->ring_ns.p, c->ring_max_words);
  PTB_CUDA(cudaGetLastError());
  c->launches += 1;
  return true;
}
} // namespace

bool launch_assemble_matrix_ring(ptb_ctx* c, const MatrixArgs& A)
{
  if (c->order != 1 || c->bs != 3 || c->ring.p == nullptr)
    return false;
  // one-warp CTAs measured fastest (profiles/r02/assembly_ring_ab_call57.json: 1.376 / 1.412 / 1.431 ms)
  const bool tma = env_flag("PTB_RING_TMA", true);
  switch (env_int("PTB_RING_WARPS", 1))
  {
  case 4: return tma ? launch_ring<4, true>(c, A) : launch_ring<4, false>(c, A);
  case 2: return tma ? launch_ring<2, true>(c, A) : launch_ring<2, false>(c, A);
  default: return tma ? launch_ring<1, true>(c, A) : launch_ring<1, false>(c, A);
  }
}
#endif // PTB_HOST_EMU

} // namespace ptb
