#include "../../include/ptb200_host.h"
#include "../common/intmaps.h"
#include "box_mesh.h"
#include "fem.h"
#include <cstring>
#include <algorithm>
#include <map>
#include <random>
#include <memory>
#include <stdexcept>
#include <string>

using namespace ptb::host;

struct pth_problem
{
  std::string type;
  BoxMesh mesh;
  FunctionSpace V;
  std::vector<std::int32_t> bc_dofs, facet_cells, facet_local;
  std::vector<double> f, g;
  std::vector<std::int64_t> rowptr;
  std::vector<std::int32_t> cols;
};

namespace
{
thread_local std::string g_err;

template <typename F>
int guarded(F&& fn)
{
  try
  {
    fn();
    return 0;
  }
  catch (const std::exception& e)
  {
    g_err = e.what();
    return 1;
  }
}
} // namespace

extern "C" {

const char* pth_last_error(void) { return g_err.c_str(); }

int pth_num_entities(int64_t i, int64_t j, int64_t k, int nrefine, int64_t out4[4])
{
  return guarded([&] {
    const auto e = num_entities(i, j, k, nrefine);
    for (int a = 0; a < 4; ++a)
      out4[a] = e[a];
  });
}

int pth_num_pdofs(int64_t i, int64_t j, int64_t k, int nrefine, int order, int64_t* out)
{
  return guarded([&] { *out = num_pdofs(i, j, k, nrefine, order); });
}

int pth_cube_sizing(uint64_t target_dofs, int target_dofs_total, uint64_t dofs_per_node, int order,
                    uint64_t num_processes, int64_t out4[4])
{
  return guarded([&] {
    const CubeSizing s
        = cube_mesh_sizing(target_dofs, target_dofs_total != 0, dofs_per_node, order, num_processes);
    out4[0] = s.Nx, out4[1] = s.Ny, out4[2] = s.Nz, out4[3] = s.r;
  });
}

int pth_problem_create(const char* problem_type, int order, int64_t nx, int64_t ny, int64_t nz,
                       int rank, int nranks, pth_problem** out)
{
  return guarded([&] {
    const std::string type(problem_type);
    if (type != "poisson" && type != "cgpoisson" && type != "elasticity")
      throw std::runtime_error("Unknown problem type: " + type);
    auto p = std::make_unique<pth_problem>();
    p->type = type;
    p->mesh = create_box_mesh(nx, ny, nz, rank, nranks);
    p->V = create_functionspace(p->mesh, order, type == "elasticity" ? 3 : 1);
    p->bc_dofs = locate_bc_dofs(p->mesh, p->V, type);
    interpolate_rhs(p->V, type, p->f, p->g);
    exterior_facets(p->mesh, p->facet_cells, p->facet_local);
    ptb::RowAdjacency adj;
    ptb::build_row_adjacency(p->V.dofmap.data(), p->mesh.n_cells_local(), p->V.nd, p->V.n_owned,
                             adj);
    ptb::build_pattern(p->V.dofmap.data(), p->V.nd, p->V.n_owned, adj, p->rowptr, p->cols);
    *out = p.release();
  });
}

int pth_problem_create_sizes_only(const char* problem_type, int order, int64_t nx, int64_t ny,
                                  int64_t nz, int rank, int nranks, pth_problem** out)
{
  return guarded([&] {
    const std::string type(problem_type);
    if (type != "poisson" && type != "cgpoisson" && type != "elasticity")
      throw std::runtime_error("Unknown problem type: " + type);
    auto p = std::make_unique<pth_problem>();
    p->type = type;
    p->mesh = create_box_mesh(nx, ny, nz, rank, nranks, false);
    p->V = create_functionspace(p->mesh, order, type == "elasticity" ? 3 : 1, false);
    exterior_facets(p->mesh, p->facet_cells, p->facet_local);
    p->rowptr.assign(1, 0);
    *out = p.release();
  });
}

// Renumber the owned dofs of a single-rank problem. DOLFINx does not number dofs lattice-
// lexicographically: fem::DofMap reorders them (reverse Cuthill-McKee / Gibbs-Poole-Stockmeyer on
// the dof graph), so a drop-in sees banded but NOT translation-invariant numberings. This
// gives the stand-in the same property: "rcm" = reverse Cuthill-McKee over the sparsity graph
// (breadth-first from a minimum-degree dof, neighbours by ascending degree, order reversed),
// "random" = a seeded shuffle (the worst case for the gather of p). Everything that names a dof
// follows: dofmap, dof coordinates, sources, Dirichlet dofs, pattern.
int pth_problem_renumber(pth_problem* p, const char* kind, uint64_t seed)
{
  return guarded([&] {
    if (p->mesh.nranks != 1)
      throw std::runtime_error("pth_problem_renumber: single-rank problems only");
    if (p->V.dofmap.empty() || p->rowptr.size() < 2)
      throw std::runtime_error("pth_problem_renumber: needs the dofmap and the pattern");
    const std::string k(kind);
    const std::int32_t n = p->V.n_owned;
    std::vector<std::int32_t> perm(n); // perm[old] = new
    if (k == "random")
    {
      std::vector<std::int32_t> order(n);
      for (std::int32_t i = 0; i < n; ++i)
        order[i] = i;
      std::mt19937_64 gen(seed);
      for (std::int32_t i = n - 1; i > 0; --i)
        std::swap(order[i], order[static_cast<std::int32_t>(gen() % static_cast<std::uint64_t>(i + 1))]);
      for (std::int32_t i = 0; i < n; ++i)
        perm[order[i]] = i;
    }
    else if (k == "rcm")
    {
      const std::vector<std::int64_t>& rp = p->rowptr;
      const std::vector<std::int32_t>& cl = p->cols;
      auto degree = [&](std::int32_t v) { return static_cast<std::int32_t>(rp[v + 1] - rp[v]); };
      std::vector<std::int32_t> cm;
      cm.reserve(n);
      std::vector<char> seen(n, 0);
      std::vector<std::int32_t> by_degree(n), nb;
      for (std::int32_t i = 0; i < n; ++i)
        by_degree[i] = i;
      std::stable_sort(by_degree.begin(), by_degree.end(),
                       [&](std::int32_t a, std::int32_t b) { return degree(a) < degree(b); });
      for (std::int32_t start : by_degree) // one breadth-first sweep per connected component
      {
        if (seen[start])
          continue;
        seen[start] = 1;
        std::size_t head = cm.size();
        cm.push_back(start);
        while (head < cm.size())
        {
          const std::int32_t v = cm[head++];
          nb.clear();
          for (std::int64_t q = rp[v]; q < rp[v + 1]; ++q)
          {
            const std::int32_t c = cl[q];
            if (c < n && !seen[c])
              seen[c] = 1, nb.push_back(c);
          }
          std::stable_sort(nb.begin(), nb.end(),
                           [&](std::int32_t a, std::int32_t b) { return degree(a) < degree(b); });
          cm.insert(cm.end(), nb.begin(), nb.end());
        }
      }
      for (std::int32_t i = 0; i < n; ++i)
        perm[cm[i]] = n - 1 - i; // reversed
    }
    else
      throw std::runtime_error("pth_problem_renumber: kind must be rcm or random");

    FunctionSpace& V = p->V;
    const int bs = V.bs;
#pragma omp parallel for schedule(static)
    for (std::int64_t i = 0; i < static_cast<std::int64_t>(V.dofmap.size()); ++i)
      if (V.dofmap[i] < n)
        V.dofmap[i] = perm[V.dofmap[i]];
    auto permute = [&](std::vector<double>& v, int width) {
      if (v.empty())
        return;
      std::vector<double> out(v);
#pragma omp parallel for schedule(static)
      for (std::int32_t i = 0; i < n; ++i)
        for (int a = 0; a < width; ++a)
          out[static_cast<std::size_t>(perm[i]) * width + a] = v[static_cast<std::size_t>(i) * width + a];
      v.swap(out);
    };
    permute(V.dof_x, 3);
    permute(p->f, bs);
    permute(p->g, 1);
    for (std::int32_t& d : p->bc_dofs)
      if (d < n)
        d = perm[d];
    std::sort(p->bc_dofs.begin(), p->bc_dofs.end());
    ptb::RowAdjacency adj;
    ptb::build_row_adjacency(V.dofmap.data(), p->mesh.n_cells_local(), V.nd, n, adj);
    ptb::build_pattern(V.dofmap.data(), V.nd, n, adj, p->rowptr, p->cols);
  });
}

void pth_problem_destroy(pth_problem* p) { delete p; }

int pth_problem_scalar(const pth_problem* p, const char* name, int64_t* out)
{
  return guarded([&] {
    const std::map<std::string, std::int64_t> m = {
        {"n_cells", p->mesh.n_cells_local()},
        {"n_cells_owned", p->mesh.n_cells_owned()},
        {"n_ghost_cells_front", p->mesh.n_ghost_cells_front()},
        {"cell_global_offset", p->mesh.cell_global_offset()},
        {"n_cells_global", p->mesh.n_cells_global()},
        {"n_vertices", p->mesh.n_vertices_local()},
        {"nd", p->V.nd},
        {"bs", p->V.bs},
        {"order", p->V.order},
        {"n_owned", p->V.n_owned},
        {"n_ghost", p->V.n_ghost},
        {"n_global", p->V.n_global},
        {"global_offset", p->V.global_offset},
        {"nnz", static_cast<std::int64_t>(p->cols.size())},
        {"n_bc", static_cast<std::int64_t>(p->bc_dofs.size())},
        {"n_facets", static_cast<std::int64_t>(p->facet_cells.size())},
        {"n_nbr", static_cast<std::int64_t>(p->V.nbr_ranks.size())},
        {"rank", p->mesh.rank},
        {"nranks", p->mesh.nranks},
        {"nx", p->mesh.nx},
        {"ny", p->mesh.ny},
        {"nz", p->mesh.nz}};
    const auto it = m.find(name);
    if (it == m.end())
      throw std::runtime_error(std::string("pth_problem_scalar: unknown name ") + name);
    *out = it->second;
  });
}

int pth_problem_array(const pth_problem* p, const char* name, const void** data, int64_t* count,
                      int* dtype)
{
  return guarded([&] {
    const std::string n(name);
    auto f64 = [&](const std::vector<double>& v) { *data = v.data(), *count = v.size(), *dtype = 0; };
    auto i32 = [&](const std::vector<std::int32_t>& v)
    { *data = v.data(), *count = v.size(), *dtype = 1; };
    auto i64 = [&](const std::vector<std::int64_t>& v)
    { *data = v.data(), *count = v.size(), *dtype = 2; };
    if (n == "x") f64(p->mesh.x);
    else if (n == "x_dofmap") i32(p->mesh.x_dofmap);
    else if (n == "dofmap") i32(p->V.dofmap);
    else if (n == "dof_x") f64(p->V.dof_x);
    else if (n == "ghost_global") i64(p->V.ghost_global);
    else if (n == "ghost_owner") i32(p->V.ghost_owner);
    else if (n == "rowptr") i64(p->rowptr);
    else if (n == "cols") i32(p->cols);
    else if (n == "bc_dofs") i32(p->bc_dofs);
    else if (n == "f") f64(p->f);
    else if (n == "g") f64(p->g);
    else if (n == "facet_cells") i32(p->facet_cells);
    else if (n == "facet_local") i32(p->facet_local);
    else if (n == "nbr_ranks") i32(p->V.nbr_ranks);
    else if (n == "send_displ") i32(p->V.send_displ);
    else if (n == "recv_displ") i32(p->V.recv_displ);
    else if (n == "local_indices") i32(p->V.local_indices);
    else if (n == "remote_indices") i32(p->V.remote_indices);
    else throw std::runtime_error("pth_problem_array: unknown name " + n);
  });
}

} // extern "C"
