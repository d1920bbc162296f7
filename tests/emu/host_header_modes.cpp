// TEST INFRASTRUCTURE -- the "sizes and lists only" modes of the host stand-in (what the CLI's
// --device_setup keeps on the host) against the full build: same sizes, ranges, ghost and halo lists.
#include "../../performance-test_b200/host/box_mesh.cpp"
#include "../../performance-test_b200/host/fem.cpp"

extern "C" int header_modes_agree(int order, int bs, long nx, long ny, long nz, int rank, int nranks)
{
  using namespace ptb::host;
  const BoxMesh a = create_box_mesh(nx, ny, nz, rank, nranks), b = create_box_mesh(nx, ny, nz, rank, nranks, false);
  if (!b.x.empty() || !b.x_dofmap.empty() || a.x.empty())
    return 1;
  if (a.l0 != b.l0 || a.l1 != b.l1 || a.L0 != b.L0 || a.L1 != b.L1 || a.P0 != b.P0 || a.P1 != b.P1
      || a.n_vertices_local() != b.n_vertices_local() || a.n_cells_local() != b.n_cells_local())
    return 2;
  const FunctionSpace V = create_functionspace(a, order, bs), W = create_functionspace(b, order, bs, false);
  if (!W.dofmap.empty() || !W.dof_x.empty() || V.dofmap.empty())
    return 3;
  if (V.n_owned != W.n_owned || V.n_ghost != W.n_ghost || V.n_global != W.n_global || V.nd != W.nd
      || V.global_offset != W.global_offset || V.ghost_global != W.ghost_global || V.ghost_owner != W.ghost_owner)
    return 4;
  if (V.nbr_ranks != W.nbr_ranks || V.send_displ != W.send_displ || V.recv_displ != W.recv_displ
      || V.local_indices != W.local_indices || V.remote_indices != W.remote_indices)
    return 5;
  std::vector<std::int32_t> c1, f1, c2, f2;
  exterior_facets(a, c1, f1);
  exterior_facets(b, c2, f2);
  return c1 == c2 && f1 == f2 ? 0 : 6;
}
