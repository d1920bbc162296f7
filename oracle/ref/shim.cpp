// TEST INFRASTRUCTURE -- oracle/_ref/libref.so: the reference's own code for the parts of the
// hot path that are self-contained, compiled here from /root/reference where it lies.
//
//   src/cg.h                 included UNCHANGED (linalg::axpy, linalg::cg) against the stub
//                            dolfinx/la/Vector.h of oracle/ref/stubs (array(), inner_product,
//                            squared_norm; MPI ranks are host threads meeting in a barrier)
//   src/mesh.cpp:44-74,82-151        num_entities / num_pdofs / the sizing search   (gen/sizing.inc)
//   src/poisson_problem.cpp:60-71,86-97,100-106   BC marker, f, g lambdas           (gen/lambdas.inc)
//   src/elasticity_problem.cpp:127-138,155-176    BC marker, f lambdas              (gen/lambdas.inc)
//   src/cgpoisson_problem.cpp:32-44  pack_fn / unpack_fn                            (gen/pack.inc)
//
// The gen/*.inc fragments are cut out of the reference by oracle/ref/extract.py at build time and
// are never committed. Everything else in this file (the CSR `action`, the thread world, the
// extern "C" wrappers) is ours: the reference gets its operator from DOLFINx/PETSc, which cannot
// be built here, so what this library pins is exactly: the CG loop, the axpy, the reductions'
// owned-entry convention, the halo pack/unpack, the sizing integers, the BC predicates and the
// source-term formulas. Only tests/, smoke() and bench.py's reference leg may load it.
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <functional>
#include <memory>
#include <span>
#include <stdexcept>
#include <thread>
#include <tuple>
#include <utility>
#include <vector>

// --- names the fragments expect from their translation units -----------------------------------
using MPI_Comm = int;                    // the stub "communicator" is just the number of ranks
namespace dolfinx::MPI
{
inline int size(MPI_Comm comm) { return comm; }
} // namespace dolfinx::MPI
using PetscScalar = double;              // CI builds real, fp64 PETSc (.github/workflows/ccpp.yml:24)
using T = PetscScalar;

// elasticity_problem.cpp:159-166 spells its view type through the mdspan macros
#define MDSPAN_IMPL_STANDARD_NAMESPACE refstub_md
#define MDSPAN_IMPL_PROPOSED_NAMESPACE ex
namespace refstub_md
{
inline constexpr std::size_t dynamic_extent = static_cast<std::size_t>(-1);
template <typename I, std::size_t... E>
struct extents
{
};
template <typename V, typename Ext>
struct mdspan
{
  mdspan(V* p, std::size_t e0, std::size_t e1) : _p(p), _e{e0, e1} {}
  V& operator()(std::size_t i, std::size_t j) const { return _p[i * _e[1] + j]; }
  std::size_t extent(int i) const { return _e[i]; }
  V* _p;
  std::size_t _e[2];
};
namespace ex
{
}
} // namespace refstub_md

#include "cg.h" // /root/reference/src/cg.h, unchanged (-I on the command line)
#include "gen/lambdas.inc"
#include "gen/pack.inc"
#include "gen/sizing.inc"

namespace
{
// what the reference's lambdas receive: x(i, p) over a [3][n] array, extent(1) = n
struct XView
{
  const double* d;
  std::size_t n;
  std::size_t extent(int i) const { return i == 0 ? 3 : n; }
  double operator()(std::size_t i, std::size_t p) const { return d[i * n + p]; }
};
} // namespace

extern "C"
{
// ---- mesh.cpp ---------------------------------------------------------------------------------
void ref_num_entities(std::int64_t i, std::int64_t j, std::int64_t k, int nrefine, std::int64_t out[4])
{
  auto [v, e, f, c] = num_entities(i, j, k, nrefine);
  out[0] = v, out[1] = e, out[2] = f, out[3] = c;
}

int ref_num_pdofs(std::int64_t i, std::int64_t j, std::int64_t k, int nrefine, int order,
                  std::int64_t* out)
{
  try
  {
    *out = num_pdofs(i, j, k, nrefine, order);
  }
  catch (const std::exception&)
  {
    return 1;
  }
  return 0;
}

int ref_cube_sizing(std::uint64_t target_dofs, int target_dofs_total, std::uint64_t dofs_per_node,
                    int order, int num_processes, std::int64_t out[4])
{
  try
  {
    ref_sizing_body(num_processes, target_dofs, target_dofs_total != 0, dofs_per_node, order, out);
  }
  catch (const std::exception&)
  {
    return 1;
  }
  return 0;
}

// ---- problem data: x is [3][n] (the layout DOLFINx hands to these lambdas) ---------------------
void ref_bc_marker(int elasticity, std::int64_t n, const double* x, std::int8_t* marker)
{
  XView xv{x, static_cast<std::size_t>(n)};
  std::vector<std::int8_t> m = elasticity ? ref_elasticity_bc_marker(xv) : ref_poisson_bc_marker(xv);
  std::copy(m.begin(), m.end(), marker);
}

void ref_poisson_source(std::int64_t n, const double* x, double* f, double* g)
{
  XView xv{x, static_cast<std::size_t>(n)};
  auto vf = ref_poisson_f(xv).first;
  auto vg = ref_poisson_g(xv).first;
  std::copy(vf.begin(), vf.end(), f);
  std::copy(vg.begin(), vg.end(), g);
}

/* f comes back as the lambda fills it: [3][n] (component-major). */
void ref_elasticity_source(std::int64_t n, const double* x, double* f)
{
  XView xv{x, static_cast<std::size_t>(n)};
  auto vf = ref_elasticity_f(xv).first;
  std::copy(vf.begin(), vf.end(), f);
}

// ---- cgpoisson_problem.cpp:32-44 ---------------------------------------------------------------
void ref_pack(std::int64_t n_idx, const std::int32_t* idx, const double* in, std::int64_t n_in,
              double* out)
{
  pack_fn(std::span<const T>(in, n_in), std::span<const std::int32_t>(idx, n_idx),
          std::span<T>(out, n_idx));
}

/* op: 0 = std::plus (reverse scatter, :221), 1 = overwrite (forward scatter, :228-229) */
void ref_unpack(std::int64_t n_idx, const std::int32_t* idx, const double* in, std::int64_t n_out,
                double* out, int op)
{
  std::function<T(T, T)> fn = std::plus<T>();
  if (op == 1)
    fn = [](auto, auto y) { return y; };
  unpack_fn(std::span<const T>(in, n_idx), std::span<const std::int32_t>(idx, n_idx),
            std::span<T>(out, n_out), fn);
}

// ---- cg.h --------------------------------------------------------------------------------------
void ref_axpy(std::int32_t n, double alpha, const double* x, const double* y, double* r)
{
  la::Vector<double> vx(n, 0, 1), vy(n, 0, 1), vr(n, 0, 1);
  std::copy(x, x + n, vx.array().begin());
  std::copy(y, y + n, vy.array().begin());
  linalg::axpy(vr, alpha, vx, vy);
  std::copy(vr.array().begin(), vr.array().end(), r);
}

/* One partition of the operator: block CSR over the owned rows (3x3 row-major blocks for bs = 3),
 * local columns (owned then ghost), vectors [owned*bs | ghost*bs], and the Scatterer-style halo
 * lists of include/ptb200.h:ptb_set_halo (block indices). */
struct ref_part
{
  std::int32_t bs, n_owned, n_ghost, n_nbr;
  const std::int64_t* rowptr;
  const std::int32_t* cols;
  const double* vals;
  const double* b; /* [(n_owned+n_ghost)*bs], ghosts current (cg.h:36-37) */
  double* x;       /* in: initial guess, out: solution, [(n_owned+n_ghost)*bs] */
  const std::int32_t *nbr_ranks, *send_displ, *local_indices, *recv_displ, *remote_indices;
};

/* linalg::cg(x, b, action, kmax, rtol) on `nranks` partitions, one host thread per rank.
 * action(p, y) = forward halo update of p (pack_fn -> neighbour copy -> unpack_fn, exactly the
 * sequence of cgpoisson_problem.cpp:223-229) followed by y_owned = A p. iters[q] receives rank
 * q's return value (they are equal: the reductions are identical on every rank). */
int ref_cg(int nranks, const ref_part* parts, int kmax, double rtol, int* iters)
{
  refstub::World world(nranks);
  std::vector<std::vector<double>> sendbuf(nranks), recvbuf(nranks);
  std::vector<const double*> sendptr(nranks);
  // expand block index lists to scalar ones the way common::Scatterer does for bs > 1
  std::vector<std::vector<std::int32_t>> lidx(nranks), ridx(nranks);
  for (int q = 0; q < nranks; ++q)
  {
    const ref_part& P = parts[q];
    const int bs = P.bs;
    for (std::int32_t i = 0; i < P.send_displ[P.n_nbr]; ++i)
      for (int a = 0; a < bs; ++a)
        lidx[q].push_back(P.local_indices[i] * bs + a);
    for (std::int32_t i = 0; i < P.recv_displ[P.n_nbr]; ++i)
      for (int a = 0; a < bs; ++a)
        ridx[q].push_back((P.remote_indices[i] - P.n_owned) * bs + a);
    sendbuf[q].assign(lidx[q].size(), 0.0);
    recvbuf[q].assign(ridx[q].size(), 0.0);
  }
  std::vector<int> failed(nranks, 0);

  auto rank_main = [&](int q)
  {
    refstub::this_rank = {&world, q};
    const ref_part& P = parts[q];
    const int bs = P.bs;
    const std::int32_t local_size = bs * P.n_owned, num_ghosts = bs * P.n_ghost;
    la::Vector<double> x(P.n_owned, P.n_ghost, bs), b(P.n_owned, P.n_ghost, bs);
    std::copy(P.x, P.x + local_size + num_ghosts, x.array().begin());
    std::copy(P.b, P.b + local_size + num_ghosts, b.array().begin());

    auto action = [&](la::Vector<double>& p, la::Vector<double>& y)
    {
      std::span<double> local_data(p.array().data(), local_size);
      std::span<double> remote_data(p.array().data() + local_size, num_ghosts);
      if (nranks > 1)
      {
        pack_fn(local_data, lidx[q], sendbuf[q]);
        world.bar.arrive_and_wait();
        for (int i = 0; i < P.n_nbr; ++i) // "scatter_fwd_begin/end": pull what each neighbour packed for us
        {
          const int nb = P.nbr_ranks[i];
          const ref_part& O = parts[nb];
          int me = 0;
          while (O.nbr_ranks[me] != q)
            ++me;
          const std::size_t cnt = static_cast<std::size_t>(P.recv_displ[i + 1] - P.recv_displ[i]) * bs;
          if (cnt != static_cast<std::size_t>(O.send_displ[me + 1] - O.send_displ[me]) * bs)
            failed[q] = 1;
          else
            std::copy_n(sendbuf[nb].data() + static_cast<std::size_t>(O.send_displ[me]) * bs, cnt,
                        recvbuf[q].data() + static_cast<std::size_t>(P.recv_displ[i]) * bs);
        }
        world.bar.arrive_and_wait();
        unpack_fn(recvbuf[q], ridx[q], remote_data, [](auto, auto v) { return v; });
      }
      std::span<const double> pv = p.array();
      std::span<double> yv = y.array();
      for (std::int32_t row = 0; row < P.n_owned; ++row)
      {
        double acc[3] = {0, 0, 0};
        for (std::int64_t s = P.rowptr[row]; s < P.rowptr[row + 1]; ++s)
        {
          const double* blk = P.vals + s * bs * bs;
          const double* pc = pv.data() + static_cast<std::size_t>(P.cols[s]) * bs;
          for (int a = 0; a < bs; ++a) // same association as oracle.c orc_spmv: acc += (row of block . pc)
          {
            double t = blk[a * bs] * pc[0];
            for (int c = 1; c < bs; ++c)
              t += blk[a * bs + c] * pc[c];
            acc[a] += t;
          }
        }
        for (int a = 0; a < bs; ++a)
          yv[static_cast<std::size_t>(row) * bs + a] = acc[a];
      }
    };

    iters[q] = linalg::cg(x, b, action, kmax, rtol);
    std::copy(x.array().begin(), x.array().end(), P.x);
  };

  if (nranks == 1)
    rank_main(0);
  else
  {
    std::vector<std::thread> th;
    for (int q = 0; q < nranks; ++q)
      th.emplace_back(rank_main, q);
    for (auto& t : th)
      t.join();
  }
  for (int q = 0; q < nranks; ++q)
    if (failed[q])
      return 1;
  return 0;
}
} // extern "C"
