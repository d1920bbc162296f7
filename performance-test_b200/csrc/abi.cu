// extern "C" entry points of libptb200.so (include/ptb200.h). The host-only inspection entry points
// (ptb_debug_*, ptb_build_cell_slot_map) live in abi_debug.cpp.
#include "abi_util.h"
#include "comm.h"
#include "envopt.h"
#include "kernels.h"
#include "layout.h"
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <memory>

using namespace ptb;

namespace
{
// Star-walk assembly (assemble_walk.cu): PTB_ASM_WALK=1/0 overrides the built-in default.
bool walk_enabled() { return env_flag("PTB_ASM_WALK", true); }
// Elasticity matrix along the star walk (assemble_matrix_p1_walk3): default since round 2
// (1.40 -> 0.91 ms at 1.33 M nodes, profiles/r02/assembly_ab_4M.json); 0 = first-generation kernel.
bool walk3_enabled() { return env_flag("PTB_ASM_WALK3", true); }
// Elasticity matrix column-major along the edge rings (assemble_matrix_p1_ring3); 0 = walk3.
bool ring_enabled() { return env_flag("PTB_ASM_RING", true); }
// P1 cell vector by direct gather along the single-reload walk (assemble_gwalk.cu): default.
bool gwalk_enabled() { return env_flag("PTB_VEC_GWALK", true); }
// Assembly maps and column layout built by the setup kernels (setup.cu) instead of the host loops.
bool gpu_setup_enabled() { return env_flag("PTB_GPU_SETUP", false); }
using ptb::abi::g_err;
using ptb::abi::guarded;
using ptb::abi::need;

void use_device(ptb_ctx* c) { PTB_CUDA(cudaSetDevice(c->device)); }

struct StageTimer
{
  ptb_ctx* c;
  int stage;
  StageTimer(ptb_ctx* c_, int s) : c(c_), stage(s) { PTB_CUDA(cudaEventRecord(c->ev0, c->stream)); }
  void stop()
  {
    PTB_CUDA(cudaEventRecord(c->ev1, c->stream));
    PTB_CUDA(cudaEventSynchronize(c->ev1));
    float ms = 0.f;
    PTB_CUDA(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    c->stage_ms[stage] = ms;
  }
};

std::int64_t n_local_entries(const ptb_ctx* c)
{
  return (static_cast<std::int64_t>(c->n_owned) + c->n_ghost) * c->bs;
}

// L2 plan of the operator kernels: B200 has 126 MB of L2. The solver vectors
// (x, p, r, y, D^-1) are re-read every CG iteration, the matrix is a pure stream: its loads carry
// the evict_first priority and bypass L1, so the stream cannot push the vectors (strong scaling:
// 10 M DOFs over 8 GPUs = 50 MB per GPU) or the gathered part of p out of L2.
// Measured (profiles/r02/ab_call4/summary.txt, elasticity, 1.25 M DOFs on one GPU = the 8-GPU share
// of config 3): 134.6 -> 122.6 us per CG iteration; 2.5 M DOFs 221.8 -> 214.8; 10 M 789.6 -> 776.2;
// Poisson 20 M unchanged within noise. PTB_L2_POLICY=0 switches the hints off.
// PTB_L2_PIN_MB=m additionally loads the first m MB of the matrix evict_last (resident from one
// iteration to the next): the kernel alone gains (99.6 -> 85.2 us with 36 MB) but inside the loop
// the pinned lines displace the vectors (129.2 us per iteration with 36 MB against 122.6 with 0),
// so the default is 0.
void plan_l2(ptb_ctx* c)
{
  c->l2_mode = env_flag("PTB_L2_POLICY", true) ? 1 : 0;
  c->l2_pin_entries = 0;
  if (!c->l2_mode)
    return;
  const double pin = 1048576.0 * std::max(0, env_int("PTB_L2_PIN_MB", 0));
  const double bytes_per_entry = c->bs == 1 ? 12.0 : 76.0; // value(s) + column index
  c->l2_pin_entries = static_cast<std::int64_t>(pin / bytes_per_entry);
}
// The host copy of the dofmap (integer maps built on the host); a space generated on the device
// (ptb_create_box) has none until someone asks.
void host_dofmap(ptb_ctx* c)
{
  if (!c->h_dofmap.empty())
    return;
  c->h_dofmap.resize(c->dofmap.n);
  PTB_CUDA(cudaStreamSynchronize(c->stream));
  PTB_CUDA(cudaMemcpy(c->h_dofmap.data(), c->dofmap.p, c->dofmap.bytes(), cudaMemcpyDeviceToHost));
}
std::int64_t n_owned_entries(const ptb_ctx* c) { return static_cast<std::int64_t>(c->n_owned) * c->bs; }

void alloc_vectors(ptb_ctx* c)
{
  const std::int64_t nl = n_local_entries(c), no = n_owned_entries(c);
  c->b.alloc(no), c->dinv.alloc(no), c->ones.alloc(no), c->r.alloc(no), c->y.alloc(no);
  c->x.alloc(nl), c->p.alloc(nl);
  c->x.zero(c->stream), c->p.zero(c->stream), c->b.zero(c->stream);
  launch_fill(c, c->ones.p, no, 1.0);
  launch_fill(c, c->dinv.p, no, 1.0);
}
} // namespace

std::int64_t ptb_ctx::device_bytes() const
{
  return xyz.bytes() + xyz3.bytes() + x_dofmap.bytes() + dofmap.bytes() + bc.bytes() + rowptr.bytes()
         + mat_off.bytes() + adj_off.bytes() + cols.bytes() + vals.bytes() + adj.bytes()
         + cell_g.bytes() + adjso.bytes() + adjrot.bytes() + walk.bytes() + ring.bytes() + ring_off.bytes() + ring_ns.bytes() + walk1.bytes() + walk1_off.bytes() + xdof.bytes() + dof_vertex.bytes() + frow_ids.bytes() + frow_ptr.bytes() + fent.bytes() + f.bytes()
         + g.bytes() + b.bytes() + dinv.bytes() + ones.bytes() + x.bytes() + p.bytes() + r.bytes()
         + y.bytes() + cg.bytes() + partials.bytes() + tickets.bytes() + send_idx.bytes()
         + recv_idx.bytes() + send_buf.bytes() + recv_buf.bytes() + peer.window.bytes() + slice_order.bytes() + cdelta.bytes() + colsx.bytes() + xoff.bytes() + zcnt_w.bytes() + zcnt_x.bytes() + mat_off_z.bytes() + xoff_z.bytes() + vals_z.bytes() + cdelta_z.bytes() + colsx_z.bytes()
         + peer.src_index.bytes();
}

namespace
{
VectorArgs vector_args(ptb_ctx* c)
{
  return VectorArgs{c->n_owned, c->n_slices, c->xyz.p, c->x_dofmap.p, c->dofmap.p, c->bc.p,
                    c->adj_off.p, c->adj.p, c->adjrot.p, c->xdof.p, c->mat_off.p, c->cols.p,
                    c->max_w, c->f.p, c->b.p};
}

// y = A p with the configured operator. `fused` only applies to the assembled operator in peer
// mode; the caller has updated the ghosts of p otherwise.
void apply_operator(ptb_ctx* c, const double* p, double* y, CgState* st, unsigned int epoch,
                    bool fused)
{
  if (c->operator_mode == PTB_OP_MATRIX_FREE)
  {
    launch_action_matrix_free(c, vector_args(c), p, y, st ? &st->py : nullptr);
    if (st != nullptr)
      launch_publish_py(c, st, epoch); // peer mode: hand the local p.y to the window all-reduce
  }
  else
    launch_spmv(c, p, y, st, epoch, fused);
}
} // namespace

extern "C" {

int ptb_create(int device, ptb_ctx** out)
{
  return guarded(nullptr, [&] {
    need(out != nullptr, "ptb_create: out is NULL");
    int ndev = 0;
    const cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
      throw std::runtime_error(
          std::string("ptb_create: no CUDA device (there is no CPU fallback): ")
          + cudaGetErrorString(e));
    need(device >= 0 && device < ndev, "ptb_create: device index out of range");
    auto c = std::make_unique<ptb_ctx>();
    c->device = device;
    PTB_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    PTB_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
      throw std::runtime_error("ptb_create: kernels are built for sm_100a only; found "
                               + std::string(prop.name));
    c->num_sms = prop.multiProcessorCount;
    PTB_CUDA(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
    c->stream = c->own_stream;
    PTB_CUDA(cudaEventCreate(&c->ev0));
    PTB_CUDA(cudaEventCreate(&c->ev1));
    c->cg.alloc(2);
    c->partials.alloc(static_cast<std::size_t>(3) * c->num_sms * 8);
    c->tickets.alloc(4);
    c->tickets.zero(c->stream);
    PTB_CUDA(cudaMallocHost(&c->h_cg, 2 * sizeof(CgState)));
    PTB_CUDA(cudaMallocHost(&c->h_scalar, 4 * sizeof(double)));
    PTB_CUDA(cudaStreamSynchronize(c->stream));
    *out = c.release();
  });
}

void ptb_destroy(ptb_ctx* c)
{
  if (!c)
    return;
  cudaSetDevice(c->device);
  try
  {
    peer_disconnect(c);
    comm_destroy(c);
  }
  catch (...)
  {
  }
  if (c->h_cg)
    cudaFreeHost(c->h_cg);
  if (c->h_scalar)
    cudaFreeHost(c->h_scalar);
  if (c->ev0)
    cudaEventDestroy(c->ev0);
  if (c->ev1)
    cudaEventDestroy(c->ev1);
  for (int i = 0; i < ptb_ctx::N_BIN_STREAMS; ++i)
  {
    if (c->bin_streams[i])
      cudaStreamDestroy(c->bin_streams[i]);
    if (c->bin_join[i])
      cudaEventDestroy(c->bin_join[i]);
  }
  if (c->bin_fork)
    cudaEventDestroy(c->bin_fork);
  if (c->own_stream)
    cudaStreamDestroy(c->own_stream);
  delete c;
}

const char* ptb_last_error(const ptb_ctx* c) { return c ? c->err.c_str() : g_err.c_str(); }

int ptb_set_stream(ptb_ctx* c, void* s)
{
  return guarded(c, [&] { c->stream = s ? static_cast<cudaStream_t>(s) : c->own_stream; });
}

int ptb_set_mesh(ptb_ctx* c, int64_t n_vertices, const double* x, int64_t n_cells,
                 const int32_t* x_dofmap)
{
  return guarded(c, [&] {
    use_device(c);
    need(n_vertices > 0 && n_cells > 0 && x && x_dofmap, "ptb_set_mesh: empty mesh");
    need(n_vertices <= INT32_MAX, "ptb_set_mesh: too many vertices for int32 indices");
    c->n_vertices = n_vertices, c->n_cells = n_cells;
    c->x_dofmap.upload(x_dofmap, static_cast<std::size_t>(n_cells) * 4, c->stream);
    c->xyz.alloc(static_cast<std::size_t>(n_vertices) * 4);
    c->xyz3.upload(x, static_cast<std::size_t>(n_vertices) * 3, c->stream); // one contiguous DMA
    launch_pad_xyz(c);
    PTB_CUDA(cudaStreamSynchronize(c->stream));
    c->have_mesh = true;
    c->matrix_assembled = c->vector_assembled = false;
  });
}

int ptb_update_geometry(ptb_ctx* c, const double* x)
{
  return guarded(c, [&] {
    use_device(c);
    need(c->have_mesh && x, "ptb_update_geometry: call ptb_set_mesh first");
    c->xyz3.upload(x, static_cast<std::size_t>(c->n_vertices) * 3, c->stream);
    launch_pad_xyz(c);
    if (c->have_space)
      launch_gather_xdof(c);
    PTB_CUDA(cudaStreamSynchronize(c->stream));
    // the operator, its Jacobi diagonal and b belong to the old coordinates
    c->matrix_assembled = c->vector_assembled = false;
    c->have_compact = false;
  });
}

int ptb_set_space(ptb_ctx* c, int problem, int order, int bs, int32_t n_owned, int32_t n_ghost,
                  const int32_t* dofmap)
{
  return guarded(c, [&] {
    use_device(c);
    need(c->have_mesh, "ptb_set_space: call ptb_set_mesh first");
    need(problem == PTB_POISSON || problem == PTB_ELASTICITY, "ptb_set_space: unknown problem");
    need(order >= 1 && order <= 3, "Order not supported");
    need((problem == PTB_POISSON && bs == 1) || (problem == PTB_ELASTICITY && bs == 3),
         "ptb_set_space: bs must be 1 for Poisson and 3 for elasticity");
    need(n_owned > 0 && n_ghost >= 0 && dofmap, "ptb_set_space: empty space");
    peer_disconnect(c); // peers hold IPC mappings of the old x / p: export and connect again
    c->problem = problem, c->order = order, c->bs = bs;
    c->operator_mode = PTB_OP_ASSEMBLED;
    c->nd = (order + 1) * (order + 2) * (order + 3) / 6;
    c->n_owned = n_owned, c->n_ghost = n_ghost;
    need(static_cast<std::uint64_t>(c->n_cells) * c->nd <= 0xFFFFFFFEull,
         "ptb_set_space: n_cells * nd exceeds the 32-bit pair index");
    c->dof_x.release();
    c->h_dofmap.assign(dofmap, dofmap + static_cast<std::size_t>(c->n_cells) * c->nd);
    c->dofmap.upload(c->h_dofmap, c->stream);
    c->bc.alloc(static_cast<std::size_t>(n_owned) + n_ghost);
    c->bc.zero(c->stream);
    c->h_bc_dofs.clear();
    // Geometry by dof index: local dofs 0..3 of a Lagrange cell sit on its vertices.
    {
      const std::int64_t nl = static_cast<std::int64_t>(n_owned) + n_ghost;
      std::vector<std::int32_t> dv(nl, -1);
      std::vector<std::int32_t> xd(static_cast<std::size_t>(c->n_cells) * 4);
      PTB_CUDA(cudaMemcpy(xd.data(), c->x_dofmap.p, xd.size() * sizeof(std::int32_t),
                          cudaMemcpyDeviceToHost));
      const int nd = c->nd;
      bool bad = false;
#pragma omp parallel for schedule(static) reduction(|| : bad)
      for (std::int64_t cell = 0; cell < c->n_cells; ++cell)
        for (int i = 0; i < nd; ++i)
        {
          const std::int32_t d = dofmap[cell * nd + i];
          if (d < 0 || d >= nl)
            bad = true;
          else if (i < 4)
            dv[d] = xd[cell * 4 + i]; // every writer of dv[d] stores the same vertex
        }
      need(!bad, "ptb_set_space: dofmap entry out of range");
      c->dof_vertex.upload(dv, c->stream);
      c->xdof.alloc(static_cast<std::size_t>(nl) * 4);
      launch_gather_xdof(c);
    }
    alloc_vectors(c);
    PTB_CUDA(cudaStreamSynchronize(c->stream));
    c->have_space = true, c->have_pattern = false, c->have_source = false;
    c->matrix_assembled = c->vector_assembled = false;
  });
}

int ptb_create_box(ptb_ctx* c, int problem, int bs, int order, int64_t nx, int64_t ny, int64_t nz, int rank,
                   int nranks, int64_t sizes[4])
{
  return guarded(c, [&] {
    use_device(c);
    need(problem == PTB_POISSON || problem == PTB_ELASTICITY, "ptb_create_box: unknown problem");
    need((problem == PTB_POISSON && bs == 1) || (problem == PTB_ELASTICITY && bs == 3),
         "ptb_create_box: bs must be 1 for Poisson and 3 for elasticity");
    need(nx >= 1 && ny >= 1 && nz >= 1, "ptb_create_box: box dimensions must be positive");
    need(nranks >= 1 && rank >= 0 && rank < nranks && nz >= nranks,
         "ptb_create_box: need 0 <= rank < nranks <= nz");
    need(order >= 1 && order <= 3, "Order not supported");
    peer_disconnect(c);
    gpu_create_box(c, order, nx, ny, nz, rank, nranks);
    c->have_mesh = true;
    c->problem = problem, c->order = order, c->bs = bs;
    c->nd = (order + 1) * (order + 2) * (order + 3) / 6;
    c->operator_mode = PTB_OP_ASSEMBLED;
    c->h_dofmap.clear(); // downloaded on demand (ptb_set_pattern's host build)
    c->bc.alloc(static_cast<std::size_t>(c->n_owned) + c->n_ghost);
    c->bc.zero(c->stream);
    c->h_bc_dofs.clear();
    c->xdof.alloc((static_cast<std::size_t>(c->n_owned) + c->n_ghost) * 4);
    launch_gather_xdof(c);
    alloc_vectors(c);
    PTB_CUDA(cudaStreamSynchronize(c->stream));
    c->have_space = true, c->have_pattern = false, c->have_source = false;
    c->matrix_assembled = c->vector_assembled = false;
    if (sizes)
      sizes[0] = c->n_vertices, sizes[1] = c->n_cells, sizes[2] = c->n_owned, sizes[3] = c->n_ghost;
  });
}

int ptb_get_mesh(ptb_ctx* c, double* x, int32_t* x_dofmap)
{
  return guarded(c, [&] {
    use_device(c);
    need(c->have_mesh, "ptb_get_mesh: no mesh set");
    PTB_CUDA(cudaStreamSynchronize(c->stream));
    if (x)
      PTB_CUDA(cudaMemcpy(x, c->xyz3.p, c->xyz3.bytes(), cudaMemcpyDeviceToHost));
    if (x_dofmap)
      PTB_CUDA(cudaMemcpy(x_dofmap, c->x_dofmap.p, c->x_dofmap.bytes(), cudaMemcpyDeviceToHost));
  });
}

int ptb_get_dofmap(ptb_ctx* c, int32_t* dofmap)
{
  return guarded(c, [&] {
    use_device(c);
    need(c->have_space && dofmap, "ptb_get_dofmap: no space set / NULL output");
    PTB_CUDA(cudaStreamSynchronize(c->stream));
    PTB_CUDA(cudaMemcpy(dofmap, c->dofmap.p, c->dofmap.bytes(), cudaMemcpyDeviceToHost));
  });
}

int ptb_get_dof_coordinates(ptb_ctx* c, double* dof_x)
{
  return guarded(c, [&] {
    use_device(c);
    need(c->have_space && dof_x, "ptb_get_dof_coordinates: no space set / NULL output");
    PTB_CUDA(cudaStreamSynchronize(c->stream));
    if (c->dof_x.p)
      PTB_CUDA(cudaMemcpy(dof_x, c->dof_x.p, c->dof_x.bytes(), cudaMemcpyDeviceToHost));
    else
    {
      // order 1 (or a caller-supplied space): the vertex-dof coordinates, unpadded
      need(c->order == 1, "ptb_get_dof_coordinates: only for order 1 or a space made by ptb_create_box");
      const std::size_t nl = static_cast<std::size_t>(c->n_owned) + c->n_ghost;
      PTB_CUDA(cudaMemcpy2D(dof_x, 3 * sizeof(double), c->xdof.p, 4 * sizeof(double), 3 * sizeof(double), nl,
                            cudaMemcpyDeviceToHost));
    }
  });
}

int ptb_set_pattern(ptb_ctx* c, const int64_t* rowptr, const int32_t* cols)
{
  return guarded(c, [&] {
    use_device(c);
    need(c->have_space, "ptb_set_pattern: call ptb_set_space first");
    need(rowptr && cols, "ptb_set_pattern: NULL pattern");
    const std::int32_t N = c->n_owned;
    c->nnz = rowptr[N];
    host_dofmap(c);
    const std::vector<std::int32_t>& dm = c->h_dofmap;
    SellLayout L;
    // The adjacency side (cell lists, slot words, star walk) comes from the host build below or,
    // opt-in for P1, from the device (setup.cu); the column side is always laid out here.
    bool dev_maps = gpu_setup_enabled();
    auto host_layout = [&](bool with_adjacency) {
      std::int64_t max_so = 0;
      if (with_adjacency)
      {
        build_row_adjacency(dm.data(), c->n_cells, c->nd, N, c->h_adj);
        max_so = build_slot_offsets(dm.data(), c->nd, N, c->h_adj, rowptr, cols, c->h_so);
        need(max_so >= 0, "ptb_set_pattern: a cell's (row, col) pair is missing from the pattern");
      }
      else
      {
        c->h_adj.ptr.assign(static_cast<std::size_t>(N) + 1, 0);
        c->h_adj.pairs.clear(), c->h_so.clear();
      }
      L = SellLayout();
      build_sell_layout(N, c->nd, rowptr, cols, c->h_adj, c->h_so, max_so, L);
      c->n_slices = L.n_slices, c->max_w = L.max_w, c->max_wa = L.max_wa;
      c->so_bits = L.so_bits, c->so_words = L.so_words;
    };
    host_layout(!dev_maps);
    c->h_rowptr.assign(rowptr, rowptr + N + 1);
    c->rowptr.upload(c->h_rowptr, c->stream);
    c->mat_off.upload(L.mat_off, c->stream);
    c->cols.upload(L.cols, c->stream);
    if (c->bs == 1)
    {
      compress_columns(N, static_cast<std::int64_t>(N) + c->n_ghost, rowptr, L);
      c->cdelta.upload(L.cdelta, c->stream);
      c->colsx.upload(L.colsx, c->stream);
      c->xoff.upload(L.xoff, c->stream);
      c->cols_explicit_frac
          = L.cols.empty() ? 0.0 : static_cast<double>(L.colsx.size()) / L.cols.size();
    }
    const auto want_walk = [&] {
      return (c->bs == 1 && walk_enabled() && L.max_w <= 32) || (c->bs == 3 && walk3_enabled());
    };
    c->walk.release();
    c->ring.release(), c->ring_off.release(), c->ring_ns.release();
    c->walk_loads_per_step = 0.0;
    c->maps_on_device = false;
    if (dev_maps)
    {
      // opt-in (PTB_GPU_SETUP=1): adj_off, the rotated slot words and the walk are built by the
      // setup kernels from the uploaded dofmap and columns. A pattern the device build cannot
      // express (missing pair, offsets beyond a byte, a star of more than 64 cells) falls back to
      // the host build, which also owns the error messages.
      int max_wa = 0;
      if (c->nd == 4 && L.max_w < 255 && gpu_setup_p1(c, want_walk(), c->bs == 3 && ring_enabled(), &max_wa))
      {
        c->max_wa = max_wa;
        c->adj.release(), c->adjso.release();
        c->maps_on_device = true;
      }
      else if (c->nd != 4 && L.max_w <= 256 && gpu_setup_pk(c, &max_wa))
      {
        // P2/P3: 8-bit offsets, as the host chooses for rows of at most 256 columns
        c->max_wa = max_wa;
        c->adjrot.release();
        c->maps_on_device = true;
      }
      else
      {
        c->walk.release();
        dev_maps = false;
        host_layout(true);
      }
    }
    if (!dev_maps)
    {
      c->adj_off.upload(L.adj_off, c->stream);
      if (!L.adjrot.empty())
      {
        // P1: the rotated one-word slot map is all the kernels need
        c->adjrot.upload(L.adjrot, c->stream);
        c->adj.release(), c->adjso.release();
      }
      else
      {
        c->adjrot.release();
        c->adj.upload(L.adj, c->stream);
        c->adjso.upload(L.adjso, c->stream);
      }
      if (!L.adjrot.empty() && want_walk())
      {
        // star-walk assembly kernels (assemble_walk.cu)
        const WalkStats ws = build_walk(N, c->h_adj, c->h_so, L);
        c->walk_loads_per_step = ws.steps ? static_cast<double>(ws.loads) / ws.steps : 0.0;
        c->walk.upload(L.walk, c->stream);
      }
      if (!L.adjrot.empty() && c->bs == 3 && ring_enabled())
      {
        // edge rings of the column-major elasticity kernel (assemble_ring.cu)
        const std::int64_t nb = build_rings(N, rowptr, c->h_adj, c->h_so, L);
        if (!L.ring.empty())
        {
          c->ring.upload(L.ring, c->stream);
          c->ring_off.upload(L.ring_off, c->stream);
          c->ring_ns.upload(L.ring_ns, c->stream);
          c->ring_bytes_per_row = N ? static_cast<double>(nb) / N : 0.0;
          std::int64_t mrw = 0;
          for (std::int32_t s = 0; s < L.n_slices; ++s)
            mrw = std::max(mrw, (L.ring_off[s + 1] - L.ring_off[s]) / 32);
          c->ring_max_words = static_cast<int>(mrw);
        }
      }
    }
    c->pk_bin_slices.release(), c->pk_bin_off.clear(), c->pk_bin_w.clear();
    if (c->order > 1)
    {
      // width classes for the binned P2/P3 matrix kernel (assemble_pk.cu)
      std::vector<std::int32_t> list;
      build_width_bins(L, list, c->pk_bin_off, c->pk_bin_w);
      c->pk_bin_slices.upload(list, c->stream);
    }
    c->walk1.release(), c->walk1_off.release();
    if (gwalk_enabled() && !L.adjrot.empty() && (c->bs == 3 || L.max_w <= 32))
    {
      // one-vertex-per-step walk for the direct-gather vector kernel (assemble_gwalk.cu)
      if (L.walk.empty())
        build_walk(N, c->h_adj, c->h_so, L);
      build_walk_single(N, c->h_adj, L);
      c->walk1.upload(L.walk1, c->stream);
      c->walk1_off.upload(L.walk1_off, c->stream);
    }
    {
      // visiting order of the operator kernels (ghost-reading slices last, clustered groups)
      std::vector<std::int32_t> inner;
      const char* env = std::getenv("PTB_SLICE_CLUSTER");
      build_slice_order(L, N, 8, env && env[0] == '1', inner, c->n_interior_slices);
      c->slice_order.upload(inner, c->stream);
    }
    c->vals.alloc(L.cols.size() * c->bs * c->bs);
    c->vals.zero(c->stream);
    PTB_CUDA(cudaStreamSynchronize(c->stream));
    // The host copy of the compressed slot map is only kept for inspection (parity tests); at
    // HBM-capacity sizes (config C5: 0.8 G pairs) it would cost ~10 GB of host memory per rank.
    if (c->h_adj.pairs.size() > (std::size_t(1) << 28))
    {
      std::vector<std::uint32_t>().swap(c->h_adj.pairs);
      std::vector<std::uint16_t>().swap(c->h_so);
    }
    plan_l2(c);
    c->balance[0].grid = c->balance[1].grid = -1; // plans belong to the old pattern
    c->have_pattern = true;
    c->matrix_assembled = false;
    c->have_compact = false;
    std::vector<std::int32_t>().swap(c->h_cols); // ptb_build_pattern keeps its own copy after this call
  });
}

int ptb_build_pattern(ptb_ctx* c, int64_t* nnz)
{
  std::vector<std::int64_t> rowptr;
  std::vector<std::int32_t> cols;
  bool done = false;
  const int rc = guarded(c, [&] {
    use_device(c);
    need(c->have_space, "ptb_build_pattern: call ptb_set_space first");
    DevBuf<std::int64_t> rp;
    DevBuf<std::int32_t> cl;
    if (!gpu_build_pattern(c, rowptr, cols, rp, cl))
    {
      // a row beyond the device kernels' capacity: same pattern from the host builder
      host_dofmap(c);
      RowAdjacency adj;
      build_row_adjacency(c->h_dofmap.data(), c->n_cells, c->nd, c->n_owned, adj);
      build_pattern(c->h_dofmap.data(), c->nd, c->n_owned, adj, rowptr, cols);
      return;
    }
    if (!gpu_setup_enabled())
      return;
    // PTB_GPU_SETUP=1: the pattern never leaves the device on its way into the layouts -- column
    // side (setup.cu gpu_setup_columns) and adjacency side (gpu_setup_p1 / gpu_setup_pk) are built
    // there; this is ptb_set_pattern without its host loops.
    c->nnz = rowptr[c->n_owned];
    SellLayout L; // host side keeps the slice offsets only
    gpu_setup_columns(c, rp, cl, L.mat_off);
    L.n_slices = c->n_slices, L.max_w = c->max_w;
    const bool want_walk
        = (c->bs == 1 && walk_enabled() && c->max_w <= 32) || (c->bs == 3 && walk3_enabled());
    int max_wa = 0;
    c->walk.release();
    c->pk_bin_slices.release(), c->pk_bin_off.clear(), c->pk_bin_w.clear();
    if (c->nd == 4)
    {
      if (c->max_w >= 255 || !gpu_setup_p1(c, want_walk, c->bs == 3 && ring_enabled(), &max_wa))
        return; // not expressible in one-byte offsets: ptb_set_pattern below rebuilds everything
      c->adj.release(), c->adjso.release();
    }
    else
    {
      if (c->max_w > 256 || !gpu_setup_pk(c, &max_wa))
        return;
      c->adjrot.release();
      // width classes for the binned P2/P3 matrix kernel (assemble_pk.cu)
      std::vector<std::int32_t> list;
      build_width_bins(L, list, c->pk_bin_off, c->pk_bin_w);
      c->pk_bin_slices.upload(list, c->stream);
    }
    c->max_wa = max_wa;
    c->so_bits = 8, c->so_words = (c->nd + 3) / 4;
    c->h_rowptr = rowptr;
    c->h_adj.ptr.assign(static_cast<std::size_t>(c->n_owned) + 1, 0);
    c->h_adj.pairs.clear(), c->h_so.clear();
    c->walk1.release(), c->walk1_off.release();
    c->walk_loads_per_step = 0.0;
    c->vals.alloc(c->cols.n * c->bs * c->bs);
    c->vals.zero(c->stream);
    PTB_CUDA(cudaStreamSynchronize(c->stream));
    c->maps_on_device = true;
    plan_l2(c);
    c->balance[0].grid = c->balance[1].grid = -1; // plans belong to the old pattern
    c->have_pattern = true;
    c->matrix_assembled = false;
    c->have_compact = false;
    done = true;
  });
  if (rc != 0)
    return rc;
  if (!done)
  {
    const int rc2 = ptb_set_pattern(c, rowptr.data(), cols.data());
    if (rc2 != 0)
      return rc2;
  }
  c->h_cols.swap(cols);
  if (nnz)
    *nnz = c->nnz;
  return 0;
}

int ptb_get_pattern(ptb_ctx* c, int64_t* rowptr, int32_t* cols)
{
  return guarded(c, [&] {
    need(c->have_pattern && c->h_cols.size() == static_cast<std::size_t>(c->nnz),
         "ptb_get_pattern: the pattern was not built by ptb_build_pattern");
    if (rowptr)
      std::copy(c->h_rowptr.begin(), c->h_rowptr.end(), rowptr);
    if (cols)
      std::copy(c->h_cols.begin(), c->h_cols.end(), cols);
  });
}

int ptb_set_bc(ptb_ctx* c, int32_t n_bc, const int32_t* bc_dofs)
{
  return guarded(c, [&] {
    use_device(c);
    need(c->have_space, "ptb_set_bc: call ptb_set_space first");
    const std::int64_t nl = static_cast<std::int64_t>(c->n_owned) + c->n_ghost;
    std::vector<std::uint8_t> m(nl, 0);
    for (std::int32_t k = 0; k < n_bc; ++k)
    {
      need(bc_dofs[k] >= 0 && bc_dofs[k] < nl, "ptb_set_bc: dof index out of range");
      m[bc_dofs[k]] = 1;
    }
    c->h_bc_dofs.assign(bc_dofs, bc_dofs + n_bc);
    c->bc.upload(m, c->stream);
    PTB_CUDA(cudaStreamSynchronize(c->stream));
    c->matrix_assembled = c->vector_assembled = false;
  });
}

int ptb_locate_bc(ptb_ctx* c, int32_t* n_bc)
{
  return guarded(c, [&] {
    use_device(c);
    need(c->have_space, "ptb_locate_bc: call ptb_set_space first");
    const std::int64_t nl = static_cast<std::int64_t>(c->n_owned) + c->n_ghost;
    launch_locate_bc(c);
    std::vector<std::uint8_t> m(static_cast<std::size_t>(nl));
    PTB_CUDA(cudaMemcpyAsync(m.data(), c->bc.p, m.size(), cudaMemcpyDeviceToHost, c->stream));
    PTB_CUDA(cudaStreamSynchronize(c->stream));
    c->h_bc_dofs.clear();
    for (std::int64_t d = 0; d < nl; ++d)
      if (m[d])
        c->h_bc_dofs.push_back(static_cast<std::int32_t>(d));
    if (n_bc)
      *n_bc = static_cast<std::int32_t>(c->h_bc_dofs.size());
    c->matrix_assembled = c->vector_assembled = false;
  });
}

int ptb_get_bc(ptb_ctx* c, int32_t* bc_dofs)
{
  return guarded(c, [&] {
    need(c->have_space && bc_dofs, "ptb_get_bc: no space set / NULL output");
    std::copy(c->h_bc_dofs.begin(), c->h_bc_dofs.end(), bc_dofs);
  });
}

int ptb_set_exterior_facets(ptb_ctx* c, int64_t n_facets, const int32_t* cells,
                            const int32_t* local_facets)
{
  return guarded(c, [&] {
    use_device(c);
    need(c->have_space, "ptb_set_exterior_facets: call ptb_set_space first");
    for (std::int64_t k = 0; k < n_facets; ++k)
      need(cells[k] >= 0 && cells[k] < c->n_cells, "exterior facet: cell index out of range");
    std::vector<std::int32_t> ids, ptr, ent;
    if (c->h_dofmap.empty() && n_facets > 0)
    {
      // the dofmap was generated on the device (ptb_create_box): fetch the rows of the facets'
      // cells only -- surface-sized, not the whole map
      DevBuf<std::int32_t> d_cells, d_rows;
      d_cells.upload(cells, static_cast<std::size_t>(n_facets), c->stream);
      d_rows.alloc(static_cast<std::size_t>(n_facets) * c->nd);
      launch_gather_dofmap_rows(c, n_facets, d_cells.p, d_rows.p);
      std::vector<std::int32_t> rows(d_rows.n);
      PTB_CUDA(cudaMemcpyAsync(rows.data(), d_rows.p, d_rows.bytes(), cudaMemcpyDeviceToHost, c->stream));
      PTB_CUDA(cudaStreamSynchronize(c->stream));
      build_facet_rows_gathered(n_facets, cells, local_facets, rows.data(), c->nd, c->order, c->n_owned, ids,
                                ptr, ent);
    }
    else
    {
      host_dofmap(c);
      build_facet_rows(n_facets, cells, local_facets, c->h_dofmap.data(), c->nd, c->order, c->n_owned, ids,
                       ptr, ent);
    }
    c->n_frows = static_cast<std::int32_t>(ids.size());
    c->frow_ids.upload(ids, c->stream);
    c->frow_ptr.upload(ptr, c->stream);
    c->fent.upload(ent, c->stream);
    PTB_CUDA(cudaStreamSynchronize(c->stream));
    c->vector_assembled = false;
  });
}

int ptb_set_source(ptb_ctx* c, const double* f, const double* g)
{
  return guarded(c, [&] {
    use_device(c);
    need(c->have_space && f, "ptb_set_source: call ptb_set_space first / f is NULL");
    c->f.upload(f, n_local_entries(c), c->stream);
    if (g)
      c->g.upload(g, static_cast<std::size_t>(c->n_owned) + c->n_ghost, c->stream);
    else
      c->g.release();
    PTB_CUDA(cudaStreamSynchronize(c->stream));
    c->have_source = true;
    c->vector_assembled = false;
  });
}

int ptb_interpolate_source(ptb_ctx* c, const double* dof_x)
{
  return guarded(c, [&] {
    use_device(c);
    need(c->have_space, "ptb_interpolate_source: call ptb_set_space first");
    need(dof_x || c->order == 1 || c->dof_x.p,
         "ptb_interpolate_source: dof coordinates are required for order > 1");
    const std::size_t nl = static_cast<std::size_t>(c->n_owned) + c->n_ghost;
    c->f.alloc(nl * c->bs);
    if (c->problem == PTB_POISSON)
      c->g.alloc(nl);
    else
      c->g.release();
    if (dof_x)
    {
      DevBuf<double> X;
      X.upload(dof_x, nl * 3, c->stream);
      launch_interpolate_source(c, X.p, 3);
      PTB_CUDA(cudaStreamSynchronize(c->stream)); // X dies here
    }
    else
    {
      if (c->order == 1)
        launch_interpolate_source(c, c->xdof.p, 4);
      else
        launch_interpolate_source(c, c->dof_x.p, 3);
      PTB_CUDA(cudaStreamSynchronize(c->stream));
    }
    c->have_source = true;
    c->vector_assembled = false;
  });
}

int ptb_get_source(ptb_ctx* c, double* f, double* g)
{
  return guarded(c, [&] {
    use_device(c);
    need(c->have_source && f, "ptb_get_source: no source set / NULL output");
    PTB_CUDA(cudaStreamSynchronize(c->stream));
    PTB_CUDA(cudaMemcpy(f, c->f.p, c->f.bytes(), cudaMemcpyDeviceToHost));
    if (g && c->g.p)
      PTB_CUDA(cudaMemcpy(g, c->g.p, c->g.bytes(), cudaMemcpyDeviceToHost));
  });
}

int ptb_set_halo(ptb_ctx* c, int n_nbr, const int32_t* nbr_ranks, const int32_t* send_displ,
                 const int32_t* local_indices, const int32_t* recv_displ,
                 const int32_t* remote_indices)
{
  return guarded(c, [&] {
    use_device(c);
    need(c->have_space, "ptb_set_halo: call ptb_set_space first");
    peer_disconnect(c); // src_index of a connected peer set belongs to the old lists
    c->nbr_ranks.assign(nbr_ranks, nbr_ranks + n_nbr);
    c->send_displ.assign(send_displ, send_displ + n_nbr + 1);
    c->recv_displ.assign(recv_displ, recv_displ + n_nbr + 1);
    const std::int64_t ns = c->send_displ.back(), nr = c->recv_displ.back();
    for (std::int64_t i = 0; i < ns; ++i)
      need(local_indices[i] >= 0 && local_indices[i] < c->n_owned,
           "ptb_set_halo: send index is not an owned dof");
    for (std::int64_t i = 0; i < nr; ++i)
      need(remote_indices[i] >= c->n_owned && remote_indices[i] < c->n_owned + c->n_ghost,
           "ptb_set_halo: receive index is not a ghost dof");
    c->send_idx.upload(local_indices, ns, c->stream);
    c->recv_idx.upload(remote_indices, nr, c->stream);
    c->send_buf.alloc(ns * c->bs);
    c->recv_buf.alloc(nr * c->bs);
    PTB_CUDA(cudaStreamSynchronize(c->stream));
  });
}

int ptb_nccl_unique_id(void* out128)
{
  return guarded(nullptr, [&] { nccl_unique_id(out128); });
}

int ptb_comm_init(ptb_ctx* c, int rank, int nranks, const void* id)
{
  return guarded(c, [&] {
    use_device(c);
    comm_init(c, rank, nranks, id);
  });
}

int ptb_peer_export(ptb_ctx* c, void* handles192)
{
  return guarded(c, [&] {
    use_device(c);
    need(handles192 != nullptr, "ptb_peer_export: NULL buffer");
    peer_export(c, handles192);
  });
}

int ptb_peer_connect(ptb_ctx* c, int rank, int nranks, const void* all_handles,
                     const int32_t* src_index)
{
  return guarded(c, [&] {
    use_device(c);
    need(all_handles != nullptr, "ptb_peer_connect: NULL handles");
    peer_connect(c, rank, nranks, all_handles, src_index);
  });
}

// ---------------------------------------------------------------------------------------------
// hot calls
// ---------------------------------------------------------------------------------------------

int ptb_assemble_matrix(ptb_ctx* c)
{
  return guarded(c, [&] {
    use_device(c);
    need(c->have_pattern, "ptb_assemble_matrix: pattern not set");
    StageTimer t(c, PTB_STAGE_ASSEMBLE_MATRIX);
    MatrixArgs A{c->n_owned, c->n_slices, c->so_bits, c->so_words, c->xyz.p, c->x_dofmap.p,
                 c->dofmap.p, c->bc.p, c->rowptr.p, c->mat_off.p, c->adj_off.p, c->cols.p,
                 c->adj.p, c->adjso.p, c->adjrot.p, c->xdof.p, c->max_w, c->vals.p, c->dinv.p};
    launch_assemble_matrix(c, A);
    c->have_compact = false;
    if (env_flag("PTB_SPMV_COMPACT", false))
      compact_operator(c); // part of the assembly stage: the SpMV then skips all-zero positions
    t.stop();
    c->matrix_assembled = true;
  });
}

int ptb_assemble_vector(ptb_ctx* c)
{
  return guarded(c, [&] {
    use_device(c);
    need(c->have_pattern && c->have_source, "ptb_assemble_vector: pattern / source not set");
    StageTimer t(c, PTB_STAGE_ASSEMBLE_VECTOR);
    VectorArgs A{c->n_owned, c->n_slices, c->xyz.p, c->x_dofmap.p, c->dofmap.p, c->bc.p,
                 c->adj_off.p, c->adj.p, c->adjrot.p, c->xdof.p, c->mat_off.p, c->cols.p, c->max_w, c->f.p,
                 c->b.p};
    FacetArgs F{c->n_frows, c->xyz.p, c->x_dofmap.p, c->dofmap.p, c->bc.p, c->frow_ids.p,
                c->frow_ptr.p, c->fent.p, c->g.p, c->b.p};
    launch_assemble_vector(c, A, F);
    t.stop();
    c->vector_assembled = true;
  });
}

int ptb_cg_solve(ptb_ctx* c, int kmax, double rtol, int precond, int* iterations,
                 double* rel_residual)
{
  return guarded(c, [&] {
    use_device(c);
    const bool mf = c->operator_mode == PTB_OP_MATRIX_FREE;
    need(c->matrix_assembled || mf, "ptb_cg_solve: matrix not assembled");
    need(c->have_pattern, "ptb_cg_solve: pattern not set");
    need(precond == PTB_PC_NONE || precond == PTB_PC_JACOBI, "ptb_cg_solve: unknown preconditioner");
    need(!(mf && precond == PTB_PC_JACOBI && !c->matrix_assembled),
         "ptb_cg_solve: Jacobi needs the assembled diagonal (call ptb_assemble_matrix once)");
    need(kmax >= 0, "ptb_cg_solve: kmax < 0");
    need(!mf || c->bs == 1, "matrix-free operator: built for the scalar Poisson space only");
    StageTimer t(c, PTB_STAGE_SOLVE);
    const double* dinv = precond == PTB_PC_JACOBI ? c->dinv.p : c->ones.p;
    CgState* st = c->cg.p;
    if (!c->have_x0)
      c->x.zero(c->stream);
    // r0 = b - A x0 (cg.h:46-47): the action is evaluated even for x0 = 0, like the reference.
    halo_forward(c, c->x.p);
    apply_operator(c, c->x.p, c->y.p, nullptr, 0, false);
    const unsigned int e0 = next_red_epoch(c);
    launch_cg_init(c, dinv, &st[1], e0);
    allreduce_sum(c, &st[1].rr, 2);
    launch_cg_finish_init(c, &st[1], rtol, e0);

    static const bool allow_fused = [] {
      const char* e = std::getenv("PTB_FUSED_HALO");
      return !(e && e[0] == '0');
    }();
    const bool fused = allow_fused && !mf && c->peer.enabled && !c->nbr_ranks.empty();
    // The whole loop in one cooperative kernel (cg.cu cg_loop) when the per-GPU problem is small
    // enough for launch gaps and kernel ramps to matter: measured on one GPU (elasticity,
    // profiles/r02/ab_call4/summary.txt) 134.6 -> 125.8 us per iteration at 1.25 M DOFs, a tie at
    // 2.5 M, 4 % slower at 10 M. PTB_CG_PERSISTENT=1 / 0 forces it on / off. Needs the assembled
    // operator and, across GPUs, the peer-memory path with the fused halo (no NCCL inside a kernel).
    static const int persistent_env = env_int("PTB_CG_PERSISTENT", -1);
    const std::int64_t persistent_max = env_int("PTB_CG_PERSISTENT_MAX_DOFS", 3000000);
    // Across GPUs every rank must take the same path (the two paths consume the peer epochs
    // differently), and the ranks' row counts differ: there the caller decides from the global size
    // (ptb_set_cg_persistent), here only a single GPU decides by itself.
    const bool persistent = persistent_env >= 0    ? persistent_env == 1
                            : c->cg_persistent >= 0 ? c->cg_persistent == 1
                            : c->nranks == 1 && !c->peer.enabled
                                ? c->bs == 3 && static_cast<std::int64_t>(c->n_owned) * c->bs <= persistent_max
                                : false;
    bool looped = false;
    if (persistent && !mf && kmax > 0 && (c->nranks == 1 || fused) && !c->nccl_comm)
    {
      const unsigned int ebase = c->peer.red_epoch + 1u, lbase = c->loop_epoch + 1u;
      need(ebase + 2u * static_cast<unsigned int>(kmax) > ebase, "ptb_cg_solve: reduction epochs would wrap");
      need(lbase + 3u * static_cast<unsigned int>(kmax) > lbase, "ptb_cg_solve: barrier epochs would wrap");
      looped = launch_cg_loop(c, dinv, 0, kmax, ebase, lbase, fused);
      if (looped)
      {
        PTB_CUDA(cudaMemcpyAsync(c->h_cg, st, 2 * sizeof(CgState), cudaMemcpyDeviceToHost, c->stream));
        PTB_CUDA(cudaStreamSynchronize(c->stream));
        const CgState fin = c->h_cg[0].k >= c->h_cg[1].k ? c->h_cg[0] : c->h_cg[1];
        // the loop consumed one halo epoch and two reduction epochs per iteration
        c->peer.halo_epoch += static_cast<unsigned long long>(fin.k);
        c->peer.red_epoch += 2u * static_cast<unsigned int>(fin.k);
        c->loop_epoch += 3u * static_cast<unsigned int>(fin.k); // three grid barriers per iteration
        halo_forward(c, c->x.p);
        peer_neighbour_barrier(c);
        t.stop();
        if (iterations)
          *iterations = fin.k;
        if (rel_residual)
          *rel_residual = std::sqrt(fin.rnorm / fin.rnorm0);
        return;
      }
    }

    // Iterations are queued in batches; the stopping flag of batch j is read back while batch
    // j + 1 is already running, so the device never waits for the host. Kernels of iterations
    // after convergence return immediately.
    const int batch = 8;
    int it = 0;
    int pending = -1;
    bool done = false;
    struct EventPair // destroyed on every exit path, a throwing PTB_CUDA inside the loop included
    {
      cudaEvent_t e[2] = {nullptr, nullptr};
      ~EventPair()
      {
        for (cudaEvent_t x : e)
          if (x)
            cudaEventDestroy(x);
      }
      cudaEvent_t& operator[](int i) { return e[i]; }
    } evs;
    PTB_CUDA(cudaEventCreateWithFlags(&evs[0], cudaEventDisableTiming));
    PTB_CUDA(cudaEventCreateWithFlags(&evs[1], cudaEventDisableTiming));
    int slot = 0;
    while (it < kmax && !done)
    {
      const int n = std::min(batch, kmax - it);
      for (int j = 0; j < n; ++j)
      {
        ++it;
        CgState* cur = &st[it & 1];
        CgState* nxt = &st[(it + 1) & 1];
        if (!fused)
          halo_forward(c, c->p.p);
        const unsigned int ea = next_red_epoch(c), eb = next_red_epoch(c);
        apply_operator(c, c->p.p, c->y.p, cur, ea, fused);
        allreduce_sum(c, &cur->py, 1);
        launch_cg_update(c, dinv, cur, ea, eb);
        allreduce_sum(c, &cur->rr, 2);
        launch_cg_direction(c, dinv, cur, nxt, eb);
      }
      PTB_CUDA(cudaMemcpyAsync(&c->h_cg[slot], &st[(it + 1) & 1], sizeof(CgState),
                               cudaMemcpyDeviceToHost, c->stream));
      PTB_CUDA(cudaEventRecord(evs[slot], c->stream));
      if (pending >= 0)
      {
        PTB_CUDA(cudaEventSynchronize(evs[pending]));
        done = c->h_cg[pending].conv != 0;
      }
      pending = slot;
      slot ^= 1;
    }
    PTB_CUDA(cudaMemcpyAsync(&c->h_cg[0], &st[(it + 1) & 1], sizeof(CgState),
                             cudaMemcpyDeviceToHost, c->stream));
    halo_forward(c, c->x.p); // leave the ghosts of the solution current (cg.h:36-37)
    peer_neighbour_barrier(c); // peers may still be reading x
    t.stop();
    const CgState fin = c->h_cg[0];
    if (iterations)
      *iterations = kmax == 0 ? 0 : fin.k;
    if (rel_residual)
      *rel_residual = std::sqrt(fin.rnorm / fin.rnorm0);
  });
}

int ptb_set_cg_persistent(ptb_ctx* c, int mode)
{
  return guarded(c, [&] {
    need(mode >= -1 && mode <= 1, "ptb_set_cg_persistent: mode must be -1 (auto), 0 or 1");
    c->cg_persistent = mode;
  });
}

int ptb_set_operator_mode(ptb_ctx* c, int mode)
{
  return guarded(c, [&] {
    need(mode == PTB_OP_ASSEMBLED || mode == PTB_OP_MATRIX_FREE, "ptb_set_operator_mode: unknown mode");
    need(mode == PTB_OP_ASSEMBLED || !c->have_space || c->bs == 1,
         "matrix-free operator: built for the scalar Poisson space only");
    c->operator_mode = mode;
  });
}

int ptb_apply_operator(ptb_ctx* c, const double* p_host, double* y_host)
{
  return guarded(c, [&] {
    use_device(c);
    need(c->matrix_assembled || (c->operator_mode == PTB_OP_MATRIX_FREE && c->have_pattern),
         "ptb_apply_operator: matrix not assembled");
    need(c->operator_mode == PTB_OP_ASSEMBLED || c->bs == 1,
         "matrix-free operator: built for the scalar Poisson space only");
    PTB_CUDA(cudaMemcpyAsync(c->p.p, p_host, n_local_entries(c) * sizeof(double),
                             cudaMemcpyHostToDevice, c->stream));
    StageTimer t(c, PTB_STAGE_SPMV);
    halo_forward(c, c->p.p);
    apply_operator(c, c->p.p, c->y.p, nullptr, 0, false);
    peer_neighbour_barrier(c); // peers may still be reading p
    t.stop();
    PTB_CUDA(cudaMemcpyAsync(y_host, c->y.p, n_owned_entries(c) * sizeof(double),
                             cudaMemcpyDeviceToHost, c->stream));
    PTB_CUDA(cudaStreamSynchronize(c->stream));
  });
}

// ---------------------------------------------------------------------------------------------
// data in / out
// ---------------------------------------------------------------------------------------------

int ptb_set_rhs(ptb_ctx* c, const double* b)
{
  return guarded(c, [&] {
    use_device(c);
    need(c->have_space && b, "ptb_set_rhs: space not set / NULL");
    PTB_CUDA(cudaMemcpyAsync(c->b.p, b, n_owned_entries(c) * sizeof(double),
                             cudaMemcpyHostToDevice, c->stream));
    PTB_CUDA(cudaStreamSynchronize(c->stream));
  });
}

int ptb_set_initial_guess(ptb_ctx* c, const double* x)
{
  return guarded(c, [&] {
    use_device(c);
    need(c->have_space, "ptb_set_initial_guess: space not set");
    c->have_x0 = x != nullptr;
    if (x)
      PTB_CUDA(cudaMemcpyAsync(c->x.p, x, n_local_entries(c) * sizeof(double),
                               cudaMemcpyHostToDevice, c->stream));
    else
      c->x.zero(c->stream);
    PTB_CUDA(cudaStreamSynchronize(c->stream));
  });
}

int ptb_get_matrix_values(ptb_ctx* c, double* vals)
{
  return guarded(c, [&] {
    use_device(c);
    need(c->matrix_assembled && vals, "ptb_get_matrix_values: matrix not assembled");
    DevBuf<double> csr;
    csr.alloc(static_cast<std::size_t>(c->nnz) * c->bs * c->bs);
    launch_sell_to_csr(c, csr.p);
    PTB_CUDA(cudaMemcpyAsync(vals, csr.p, csr.bytes(), cudaMemcpyDeviceToHost, c->stream));
    PTB_CUDA(cudaStreamSynchronize(c->stream));
  });
}

int ptb_get_diagonal_inverse(ptb_ctx* c, double* dinv)
{
  return guarded(c, [&] {
    use_device(c);
    need(c->matrix_assembled && dinv, "ptb_get_diagonal_inverse: matrix not assembled");
    PTB_CUDA(cudaMemcpyAsync(dinv, c->dinv.p, c->dinv.bytes(), cudaMemcpyDeviceToHost, c->stream));
    PTB_CUDA(cudaStreamSynchronize(c->stream));
  });
}

int ptb_get_rhs(ptb_ctx* c, double* b)
{
  return guarded(c, [&] {
    use_device(c);
    need(c->have_space && b, "ptb_get_rhs: space not set");
    PTB_CUDA(cudaMemcpyAsync(b, c->b.p, c->b.bytes(), cudaMemcpyDeviceToHost, c->stream));
    PTB_CUDA(cudaStreamSynchronize(c->stream));
  });
}

int ptb_get_solution(ptb_ctx* c, double* x)
{
  return guarded(c, [&] {
    use_device(c);
    need(c->have_space && x, "ptb_get_solution: space not set");
    PTB_CUDA(cudaMemcpyAsync(x, c->x.p, c->x.bytes(), cudaMemcpyDeviceToHost, c->stream));
    PTB_CUDA(cudaStreamSynchronize(c->stream));
  });
}

int ptb_solution_norm(ptb_ctx* c, double* norm)
{
  return guarded(c, [&] {
    use_device(c);
    need(c->have_space && norm, "ptb_solution_norm: space not set");
    double* d = &c->cg.p[0].py; // scratch: not live outside a solve
    launch_sqnorm(c, c->x.p, n_owned_entries(c), d);
    allreduce_sum(c, d, 1);
    PTB_CUDA(cudaMemcpyAsync(c->h_scalar, d, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    PTB_CUDA(cudaStreamSynchronize(c->stream));
    *norm = std::sqrt(c->h_scalar[0]);
  });
}

int ptb_get_slot_offsets(ptb_ctx* c, int64_t* n_pairs, int64_t* pair_ptr, uint32_t* pairs,
                         uint16_t* offsets)
{
  return guarded(c, [&] {
    need(c->have_pattern, "ptb_get_slot_offsets: pattern not set");
    need(!c->maps_on_device, "ptb_get_slot_offsets: the maps were built on the device (PTB_GPU_SETUP=1); "
                             "use ptb_get_p1_maps");
    need(!c->h_adj.pairs.empty() || c->h_adj.ptr.back() == 0,
         "ptb_get_slot_offsets: the host copy is not retained above 2^28 pairs");
    if (n_pairs)
      *n_pairs = static_cast<std::int64_t>(c->h_adj.pairs.size());
    if (pair_ptr)
      std::copy(c->h_adj.ptr.begin(), c->h_adj.ptr.end(), pair_ptr);
    if (pairs)
      std::copy(c->h_adj.pairs.begin(), c->h_adj.pairs.end(), pairs);
    if (offsets)
      std::copy(c->h_so.begin(), c->h_so.end(), offsets);
  });
}

int ptb_get_p1_maps(ptb_ctx* c, int64_t* adj_off, uint32_t* adjrot, uint32_t* walk, int* have_walk,
                    int* built_on_device)
{
  return guarded(c, [&] {
    use_device(c);
    need(c->have_pattern && c->adjrot.p != nullptr, "ptb_get_p1_maps: no P1 pattern set");
    need(adj_off != nullptr, "ptb_get_p1_maps: NULL adj_off");
    PTB_CUDA(cudaStreamSynchronize(c->stream));
    PTB_CUDA(cudaMemcpy(adj_off, c->adj_off.p, c->adj_off.bytes(), cudaMemcpyDeviceToHost));
    if (adjrot)
      PTB_CUDA(cudaMemcpy(adjrot, c->adjrot.p, c->adjrot.bytes(), cudaMemcpyDeviceToHost));
    if (walk && c->walk.p)
      PTB_CUDA(cudaMemcpy(walk, c->walk.p, c->walk.bytes(), cudaMemcpyDeviceToHost));
    if (have_walk)
      *have_walk = c->walk.p != nullptr;
    if (built_on_device)
      *built_on_device = c->maps_on_device;
  });
}

int ptb_get_p1_rings(ptb_ctx* c, int64_t* ring_off, uint8_t* ring_ns, uint32_t* ring, int* have_rings)
{
  return guarded(c, [&] {
    use_device(c);
    need(c->have_pattern && have_rings != nullptr, "ptb_get_p1_rings: no pattern set / NULL have_rings");
    *have_rings = c->ring.p != nullptr || (c->ring_off.p != nullptr && c->ring.n == 0);
    if (!*have_rings)
      return;
    PTB_CUDA(cudaStreamSynchronize(c->stream));
    if (ring_off)
      PTB_CUDA(cudaMemcpy(ring_off, c->ring_off.p, c->ring_off.bytes(), cudaMemcpyDeviceToHost));
    if (ring_ns)
      PTB_CUDA(cudaMemcpy(ring_ns, c->ring_ns.p, c->ring_ns.bytes(), cudaMemcpyDeviceToHost));
    if (ring && c->ring.p)
      PTB_CUDA(cudaMemcpy(ring, c->ring.p, c->ring.bytes(), cudaMemcpyDeviceToHost));
  });
}

int ptb_time_kernel(ptb_ctx* c, int which, int reps, double* ms_avg)
{
  return guarded(c, [&] {
    use_device(c);
    need(c->matrix_assembled && c->have_source, "ptb_time_kernel: assemble first");
    need(reps > 0 && ms_avg, "ptb_time_kernel: reps must be positive");
    // benign scalars: alpha = 0, beta = 1, never converged
    CgState s{};
    s.py = 1.0, s.rr = 1.0, s.rz = 1.0, s.rz_old = 0.0, s.rnorm0 = 1.0, s.rtol2 = 0.0, s.rnorm = 1.0;
    s.alpha = 0.0;
    CgState sd = s;
    sd.rz_old = 1.0;
    PTB_CUDA(cudaMemcpyAsync(&c->cg.p[0], &s, sizeof(s), cudaMemcpyHostToDevice, c->stream));
    PTB_CUDA(cudaMemcpyAsync(&c->cg.p[1], &sd, sizeof(s), cudaMemcpyHostToDevice, c->stream));
    MatrixArgs MA{c->n_owned, c->n_slices, c->so_bits, c->so_words, c->xyz.p, c->x_dofmap.p,
                  c->dofmap.p, c->bc.p, c->rowptr.p, c->mat_off.p, c->adj_off.p, c->cols.p,
                  c->adj.p, c->adjso.p, c->adjrot.p, c->xdof.p, c->max_w, c->vals.p, c->dinv.p};
    VectorArgs VA{c->n_owned, c->n_slices, c->xyz.p, c->x_dofmap.p, c->dofmap.p, c->bc.p,
                  c->adj_off.p, c->adj.p, c->adjrot.p, c->xdof.p, c->mat_off.p, c->cols.p, c->max_w, c->f.p,
                 c->b.p};
    FacetArgs FA{0, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    auto one = [&] {
      switch (which)
      {
      case PTB_KERNEL_SPMV: launch_spmv(c, c->p.p, c->y.p, &c->cg.p[0], 0); break;
      case PTB_KERNEL_CG_UPDATE: launch_cg_update(c, c->dinv.p, &c->cg.p[0], 0, 0); break;
      case PTB_KERNEL_CG_DIRECTION:
        launch_cg_direction(c, c->dinv.p, &c->cg.p[1], reinterpret_cast<CgState*>(c->partials.p), 0);
        break;
      case PTB_KERNEL_ASSEMBLE_MATRIX: launch_assemble_matrix(c, MA); break;
      case PTB_KERNEL_ASSEMBLE_VECTOR: launch_assemble_vector(c, VA, FA); break;
      default: throw std::runtime_error("ptb_time_kernel: unknown kernel");
      }
    };
    one();
    PTB_CUDA(cudaEventRecord(c->ev0, c->stream));
    for (int i = 0; i < reps; ++i)
      one();
    PTB_CUDA(cudaEventRecord(c->ev1, c->stream));
    PTB_CUDA(cudaEventSynchronize(c->ev1));
    float ms = 0.f;
    PTB_CUDA(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    *ms_avg = ms / reps;
    c->vector_assembled = c->vector_assembled && which != PTB_KERNEL_ASSEMBLE_VECTOR;
  });
}

double ptb_stage_ms(const ptb_ctx* c, int stage)
{
  return (c && stage >= 0 && stage < PTB_STAGE_COUNT) ? c->stage_ms[stage] : -1.0;
}
int64_t ptb_launch_count(const ptb_ctx* c) { return c ? c->launches : 0; }
int64_t ptb_device_bytes(const ptb_ctx* c) { return c ? c->device_bytes() : 0; }
double ptb_cols_explicit_fraction(const ptb_ctx* c) { return c ? c->cols_explicit_frac : 1.0; }
int64_t ptb_spmv_stored_entries(const ptb_ctx* c)
{
  if (!c)
    return 0;
  return c->have_compact ? c->compact_nnz
                         : static_cast<std::int64_t>(c->vals.n) / std::max(1, c->bs * c->bs);
}

} // extern "C"
