// Internal context of libptb200.so. Device data layout (DESIGN.md "Data layout in HBM"):
//
//   rows are the owned block dofs; 32 consecutive rows form a *slice* (one warp);
//   every per-row list is stored slice-major, column-major inside the slice (SELL-32), so a warp
//   reading "entry k of my row" touches 32 consecutive words:
//     matrix   cols[mat_off[s] + k*32 + lane]           vals[(mat_off[s] + k*32)*bs2 + e*32 + lane]
//     cells    adj [adj_off[s] + k*32 + lane]           (pair = cell*nd + local index)
//     slots    adjso[(adj_off[s] + k*32)*nw + w*32 + lane]   (4 x uint8 or 2 x uint16 per word)
#pragma once
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/ptb200.h"
#include "../../include/ptb200_debug.h"
#include "../common/intmaps.h"

namespace ptb
{

struct CudaError : std::runtime_error
{
  using std::runtime_error::runtime_error;
};

#define PTB_CUDA(call)                                                                            \
  do                                                                                              \
  {                                                                                               \
    cudaError_t e__ = (call);                                                                     \
    if (e__ != cudaSuccess)                                                                       \
      throw ::ptb::CudaError(std::string(#call) + ": " + cudaGetErrorString(e__));                \
  } while (0)

/// Owning device buffer.
template <typename T>
struct DevBuf
{
  T* p = nullptr;
  std::size_t n = 0;
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { release(); }
  void release()
  {
    if (p)
      cudaFree(p);
    p = nullptr, n = 0;
  }
  void alloc(std::size_t count)
  {
    if (count == n && p)
      return;
    release();
    if (count)
      PTB_CUDA(cudaMalloc(&p, count * sizeof(T)));
    n = count;
  }
  void upload(const T* h, std::size_t count, cudaStream_t s)
  {
    alloc(count);
    if (count)
      PTB_CUDA(cudaMemcpyAsync(p, h, count * sizeof(T), cudaMemcpyHostToDevice, s));
  }
  void upload(const std::vector<T>& h, cudaStream_t s) { upload(h.data(), h.size(), s); }
  void zero(cudaStream_t s)
  {
    if (n)
      PTB_CUDA(cudaMemsetAsync(p, 0, n * sizeof(T), s));
  }
  std::size_t bytes() const { return n * sizeof(T); }
};

/// Scalars of the CG loop, double-buffered by iteration parity (DESIGN.md "CG").
struct CgState
{
  double py;     // p.y                 (written by spmv's last block, then allreduced)
  double rr;     // r.r   after update  (written by update's last block, allreduced with rz)
  double rz;     // r.z   after update
  double rz_old; // r.z entering this iteration
  double rnorm0; // |r0|^2
  double rtol2;
  double rnorm;  // |r|^2 entering this iteration
  double alpha;  // step length of this iteration (written by cg_update for cg_direction)
  int k;         // iterations completed
  int conv;      // 1 once the stopping rule fired
};

} // namespace ptb

struct ptb_ctx
{
  int device = 0;
  cudaStream_t own_stream = nullptr, stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  std::string err;
  double stage_ms[PTB_STAGE_COUNT] = {0, 0, 0, 0};
  std::int64_t launches = 0;
  int grid_cache[32] = {}; // one-wave grid sizes per kernel (0 = not queried yet)
  int num_sms = 148;

  // problem description
  int problem = PTB_POISSON, order = 1, bs = 1, nd = 4;
  std::int64_t n_vertices = 0, n_cells = 0;
  std::int32_t n_owned = 0, n_ghost = 0;
  std::int64_t nnz = 0; // block nonzeros
  bool have_mesh = false, have_space = false, have_pattern = false, have_source = false,
       matrix_assembled = false, vector_assembled = false;

  // geometry + dofmaps
  ptb::DevBuf<double> xyz;        // [n_vertices][4] padded
  ptb::DevBuf<double> xyz3;       // [n_vertices][3] staging of the caller's layout
  ptb::DevBuf<double> dof_x;      // [n_local dofs][3] dof coordinates of a device-generated P2/P3 space (box.cu)
  ptb::DevBuf<double> xdof;       // [n_local dofs][4]: coordinates of vertex dofs, by dof index
  ptb::DevBuf<std::int32_t> dof_vertex; // [n_local dofs] geometry vertex of a vertex dof, else -1
  ptb::DevBuf<std::int32_t> x_dofmap, dofmap;
  ptb::DevBuf<std::uint8_t> bc;   // marker per local block dof
  std::vector<std::int32_t> h_bc_dofs, h_dofmap; // host copy of the dofmap for the integer maps

  // pattern (CSR, kept for inspection) + SELL layouts
  std::vector<std::int64_t> h_rowptr;
  std::int32_t n_slices = 0;
  int max_w = 0, max_wa = 0, so_bits = 8, so_words = 1;
  ptb::DevBuf<std::int64_t> rowptr, mat_off, adj_off;
  ptb::DevBuf<std::int32_t> cols;      // SELL
  ptb::DevBuf<std::int32_t> cdelta, colsx; // compressed column indices for the scalar SpMV
  ptb::DevBuf<std::int64_t> xoff;
  double cols_explicit_frac = 1.0;
  // zero-column compaction of the scalar operator for the SpMV (compact.cu, PTB_SPMV_COMPACT=1)
  ptb::DevBuf<std::int64_t> zcnt_w, zcnt_x, mat_off_z, xoff_z;
  ptb::DevBuf<double> vals_z;
  ptb::DevBuf<std::int32_t> cdelta_z, colsx_z;
  bool have_compact = false;
  std::int64_t compact_nnz = 0; // stored entries (padding included) of the compacted copy
  // L2 plan of the operator kernels (abi.cu plan_l2): hints on/off, pinned prefix in mat_off units
  int l2_mode = 0;
  std::int64_t l2_pin_entries = 0;
  // balanced work split of the operator kernels for small problems (cg.cu ensure_balance): one
  // plan per kernel family (0 = spmv_sell, 1 = cg_loop), rebuilt when grid / roles / operator change
  struct Balance
  {
    int grid = -1, npull = -2;
    bool compact = false, ok = false;
    int longest_run = 0; // slices of the longest CTA run
    ptb::DevBuf<std::int32_t> ounit, begin;
  } balance[2];
  ptb::DevBuf<std::int32_t> slice_order; // slices without ghost columns first (fused halo)
  std::int32_t n_interior_slices = 0;
  ptb::DevBuf<double> vals;            // SELL, bs2 planes per entry
  ptb::DevBuf<std::uint32_t> adj, adjso, adjrot;
  ptb::DevBuf<std::uint32_t> walk;     // P1 star walk (layout.h), uploaded when PTB_ASM_WALK=1
  double walk_loads_per_step = 0.0;
  std::vector<std::int32_t> h_cols;    // CSR columns, kept only when ptb_build_pattern built them
  bool maps_on_device = false;         // adj_off / adjrot / walk built by setup.cu (PTB_GPU_SETUP=1)
  ptb::DevBuf<std::uint32_t> walk1;    // one-vertex-per-step walk (layout.h), PTB_ASM_GWALK=1
  ptb::DevBuf<std::int64_t> walk1_off;
  // P1 edge rings (layout.h build_rings) of the column-major elasticity kernel (assemble_ring.cu)
  ptb::DevBuf<std::uint32_t> ring;
  ptb::DevBuf<std::int64_t> ring_off;
  ptb::DevBuf<std::uint8_t> ring_ns;
  double ring_bytes_per_row = 0.0;
  int ring_max_words = 0; // ring words per lane of the longest slice
  // host copies of the compressed slot map (parity inspection)
  ptb::RowAdjacency h_adj;
  std::vector<std::uint16_t> h_so;

  // exterior facets: CSR over boundary rows
  std::int32_t n_frows = 0;
  ptb::DevBuf<std::int32_t> frow_ids, frow_ptr, fent; // fent = {cell, local_facet*nd + li} pairs

  // reference tensors of the P2/P3 element (element_tables.h), uploaded on first use
  ptb::DevBuf<double> tab_S, tab_M, tab_MF, tab_KF;
  // P2/P3 matrix assembly by row-length bins (PTB_PK_BINS=1): slices grouped by width class
  ptb::DevBuf<std::int32_t> pk_bin_slices;
  std::vector<std::int32_t> pk_bin_off; // [n_bins + 1] into pk_bin_slices
  std::vector<int> pk_bin_w;            // accumulator width of each bin
  int tab_order = 0;
  ptb::DevBuf<double> cell_g;           // P2/P3 geometry factors per cell (assemble_pk.cu cell_geometry_pk)
  // side streams of the binned launches (assemble_pk.cu run_bins): the row-length classes are
  // independent kernels, the small ones run beside the large ones; created on first use
  static constexpr int N_BIN_STREAMS = 6;
  cudaStream_t bin_streams[N_BIN_STREAMS] = {};
  cudaEvent_t bin_fork = nullptr, bin_join[N_BIN_STREAMS] = {};

  // matrix-free operator mode (ptb_set_operator_mode) and its per-CTA dot partials
  int operator_mode = 0;
  ptb::DevBuf<double> mf_partials;

  // vectors
  ptb::DevBuf<double> f, g, b, dinv, ones, x, p, r, y;
  bool have_x0 = false;

  // CG scalars + reductions
  ptb::DevBuf<ptb::CgState> cg;          // [2]
  ptb::DevBuf<double> partials;          // [3 * max_grid]
  ptb::DevBuf<unsigned int> tickets;     // [4]
  ptb::DevBuf<unsigned long long> loop_slots; // [grid + 1][4] LL records of the persistent loop's barrier
  unsigned int loop_epoch = 0;                // last barrier epoch used (monotone across solves)
  ptb::DevBuf<unsigned long long> loop_trace; // PTB_LOOP_TRACE diagnostic (cg.cu launch_cg_loop)
  int cg_persistent = -1;                     // ptb_set_cg_persistent: -1 auto, 0 off, 1 on
  ptb::CgState* h_cg = nullptr;          // pinned [2]
  double* h_scalar = nullptr;            // pinned [4]

  // halo / comm
  int rank = 0, nranks = 1;
  void* nccl_comm = nullptr;
  std::vector<std::int32_t> nbr_ranks, send_displ, recv_displ;
  ptb::DevBuf<std::int32_t> send_idx, recv_idx;
  ptb::DevBuf<double> send_buf, recv_buf;

  // peer-memory (NVLink P2P) communication state, see peer.cuh
  struct Peer
  {
    bool enabled = false;
    ptb::DevBuf<unsigned char> window;    // this rank's PeerWindow
    void* win[16] = {};                   // every rank's window (device pointers)
    void* nbr_x[8] = {};                  // neighbours' x and p vectors
    void* nbr_p[8] = {};
    ptb::DevBuf<std::int32_t> src_index;  // owner-local index per receive entry
    ptb::DevBuf<unsigned long long> ready; // [256] completion epochs of the fused halo pullers
    std::vector<void*> opened;            // IPC mappings to close
    unsigned long long halo_epoch = 0;
    unsigned int red_epoch = 0;
  } peer;

  std::int64_t device_bytes() const;
};
