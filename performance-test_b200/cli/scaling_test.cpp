// dolfinx-scaling-test (B200): command-line drop-in for the reference's driver.
//
// Keeps the reference's surface:
//   * the ten options of src/main.cpp:54-74 (unregistered options are ignored, :76-82), plus the
//     PETSc-style solver options the in-scope comparison needs: -ksp_rtol, -ksp_max_it,
//     -pc_type {none,jacobi}, -ksp_type cg (gamg / hypre are refused: AMG is out of scope);
//   * the mesh line of src/mesh.cpp:190-194, the "Test problem summary" block of
//     src/main.cpp:186-205, the ZZZ timer names of each problem type, the "Summary of timings"
//     table (src/main.cpp:226) and the two "***" lines (src/main.cpp:232-233).
// What it runs instead of DOLFINx/PETSc: the host stand-in for setup and libptb200.so (CUDA) for
// ZZZ Assemble matrix / ZZZ Assemble vector / ZZZ Solve. One process per GPU; ranks rendezvous
// over TCP (bootstrap.h) and talk NVLink peer memory (default) or NCCL (--comm nccl).
#include "../../include/ptb200.h"
#include "../common/intmaps.h"
#include "../host/box_mesh.h"
#include "../host/fem.h"
#include "bootstrap.h"
#include "timing.h"
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <iomanip>
#include <iostream>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <sys/wait.h>
#include <tuple>
#include <unistd.h>

using namespace ptb::host;
using ptb::cli::Bootstrap;
using ptb::cli::Timer;
using T = double;

namespace
{

// src/main.cpp:31-50
std::string int64_to_human(std::int64_t n)
{
  double r = static_cast<double>(n);
  const std::string name[] = {"", "thousand", "million", "billion", "trillion"};
  int i = 0;
  for (; r > 1000.0; ++i)
    r /= 1000.0;
  if (i > 4)
    throw std::runtime_error("number too big");
  std::stringstream s;
  if (i > 0)
    s << " (" << std::setprecision(3) << r << " " << name[i] << ")";
  return s.str();
}

struct Options
{
  std::string problem_type = "poisson", mesh_type = "cube", scaling_type = "weak", output = "",
              scatterer = "neighbor", comm = "peer", pc_type = "jacobi";
  bool help = false, memory_profiling = false, subcomm_partition = false, device_setup = false;
  std::size_t ndofs = 50000, order = 1;
  int nprocs = 0; // > 0: fork that many ranks on this node
  double ksp_rtol = 1e-8;
  int ksp_max_it = 10000;
  bool rtol_set = false, maxit_set = false;
};

const char* USAGE = R"(Allowed options:
  -h [ --help ]                     print usage message
  --problem_type arg (=poisson)     problem (poisson, cgpoisson, or elasticity)
  --mesh_type arg (=cube)           mesh (cube or unstructured)
  --memory_profiling                turn on memory logging
  --subcomm_partition               Use sub-communicator for partitioning
  --scaling_type arg (=weak)        scaling (weak or strong)
  --output arg                      output directory (no output unless this is set)
  --ndofs arg (=50000)              number of degrees of freedom
  --order arg (=1)                  polynomial order
  --scatterer arg (=neighbor)       scatterer for CG (neighbor or p2p)
B200 additions:
  --nprocs arg                      start this many ranks (one per GPU) on this node
  --comm arg (=peer)                multi-GPU transport: peer (NVLink peer memory) or nccl
  --device_setup                    mesh, dofmap, sparsity, boundary conditions and RHS on the GPU
  -ksp_rtol arg (=1e-8)  -ksp_max_it arg (=10000)  -pc_type arg (=jacobi) none|jacobi
)";

Options parse(int argc, char* argv[])
{
  Options o;
  for (int i = 1; i < argc; ++i)
  {
    std::string a = argv[i], val;
    bool has_val = false;
    const auto eq = a.find('=');
    if (a.rfind("--", 0) == 0 && eq != std::string::npos)
      val = a.substr(eq + 1), a = a.substr(0, eq), has_val = true;
    auto value = [&]() -> std::string {
      if (has_val)
        return val;
      if (i + 1 >= argc)
        throw std::runtime_error("the required argument for option '" + a + "' is missing");
      return argv[++i];
    };
    if (a == "--help" || a == "-h") o.help = true;
    else if (a == "--problem_type") o.problem_type = value();
    else if (a == "--mesh_type") o.mesh_type = value();
    else if (a == "--memory_profiling") o.memory_profiling = true;
    else if (a == "--subcomm_partition") o.subcomm_partition = true;
    else if (a == "--scaling_type") o.scaling_type = value();
    else if (a == "--output") o.output = value();
    else if (a == "--ndofs") o.ndofs = std::stoull(value());
    else if (a == "--order") o.order = std::stoull(value());
    else if (a == "--scatterer") o.scatterer = value();
    else if (a == "--nprocs") o.nprocs = std::stoi(value());
    else if (a == "--comm") o.comm = value();
    else if (a == "--device_setup") o.device_setup = true;
    else if (a == "-ksp_rtol") o.ksp_rtol = std::stod(value()), o.rtol_set = true;
    else if (a == "-ksp_max_it") o.ksp_max_it = std::stoi(value()), o.maxit_set = true;
    else if (a == "-pc_type")
    {
      o.pc_type = value();
      if (o.pc_type != "none" && o.pc_type != "jacobi")
        throw std::runtime_error("-pc_type " + o.pc_type
                                 + " is not available: only none and jacobi are (AMG is out of scope)");
    }
    else if (a == "-ksp_type")
    {
      if (value() != "cg")
        throw std::runtime_error("only -ksp_type cg is available");
    }
    // anything else: ignored like boost's allow_unregistered / PETSc's options database
  }
  return o;
}

void ok(ptb_ctx* c, int rc)
{
  if (rc != 0)
    throw std::runtime_error(ptb_last_error(c));
}

struct Gpu
{
  ptb_ctx* c = nullptr;
  explicit Gpu(int device)
  {
    if (ptb_create(device, &c) != 0)
      throw std::runtime_error(ptb_last_error(nullptr));
  }
  ~Gpu() { ptb_destroy(c); }
};

struct Vector // la::Vector stand-in: [owned*bs | ghost*bs]
{
  std::vector<T> array;
  std::int64_t n_owned_entries = 0;
};

using SolverFunction = std::function<int(Vector& u, const Vector& b)>;

// {poisson,elastic,cgpoisson}::problem of the reference (src/poisson_problem.cpp:31,
// src/elasticity_problem.cpp:99, src/cgpoisson_problem.cpp:49): returns (b, u, solver_function).
std::tuple<std::shared_ptr<Vector>, std::shared_ptr<Vector>, SolverFunction>
problem(const std::string& type, const BoxMesh& mesh, int order, const Options& opt,
        Bootstrap& boot, std::shared_ptr<Gpu> gpu, std::int64_t& ndofs_global, double& solve_seconds)
{
  const bool elasticity = type == "elasticity";
  const bool cgp = type == "cgpoisson";
  ptb_ctx* c = gpu->c;

  // --device_setup (opt-in): the arrays of every setup region are generated
  // on the device (ptb_create_box, ptb_locate_bc, ptb_interpolate_source, ptb_build_pattern); the
  // host keeps the sizes and the surface-sized lists (ghost/halo lists, exterior facets).
  const bool dev = opt.device_setup;
  Timer t0("ZZZ FunctionSpace");
  auto V = std::make_shared<FunctionSpace>(create_functionspace(mesh, order, elasticity ? 3 : 1, !dev));
  if (dev)
  {
    std::int64_t sizes[4];
    ok(c, ptb_create_box(c, elasticity ? PTB_ELASTICITY : PTB_POISSON, V->bs, order, mesh.nx, mesh.ny, mesh.nz,
                         mesh.rank, mesh.nranks, sizes));
    if (sizes[0] != mesh.n_vertices_local() || sizes[1] != mesh.n_cells_local() || sizes[2] != V->n_owned
        || sizes[3] != V->n_ghost)
      throw std::runtime_error("device-generated slab does not match the host sizing");
  }
  t0.stop();
  t0.flush();
  ndofs_global = V->n_global * V->bs;

  std::unique_ptr<Timer> t1;
  if (!elasticity)
    t1 = std::make_unique<Timer>("ZZZ Assemble");

  Timer t2("ZZZ Create boundary conditions");
  std::vector<std::int32_t> bdofs;
  if (dev)
  {
    std::int32_t n_bc = 0;
    ok(c, ptb_locate_bc(c, &n_bc));
  }
  else
    bdofs = locate_bc_dofs(mesh, *V, type);
  t2.stop();
  t2.flush();

  Timer t3("ZZZ Create RHS function");
  std::vector<double> f, g;
  if (dev)
    ok(c, ptb_interpolate_source(c, nullptr));
  else
    interpolate_rhs(*V, type, f, g);
  t3.stop();
  t3.flush();

  {
    std::unique_ptr<Timer> tf;
    if (elasticity)
      tf = std::make_unique<Timer>("ZZZ Create forms");
    // create_matrix: sparsity pattern (poisson_problem.cpp:122-123) + device setup
    std::vector<std::int32_t> fc, fl;
    exterior_facets(mesh, fc, fl);
    if (dev)
    {
      std::int64_t nnz = 0;
      ok(c, ptb_build_pattern(c, &nnz)); // with PTB_GPU_SETUP=1 the layouts and maps as well
      ok(c, ptb_set_exterior_facets(c, static_cast<std::int64_t>(fc.size()), fc.data(), fl.data()));
    }
    else
    {
      ptb::RowAdjacency adj;
      std::vector<std::int64_t> rowptr;
      std::vector<std::int32_t> cols;
      ptb::build_row_adjacency(V->dofmap.data(), mesh.n_cells_local(), V->nd, V->n_owned, adj);
      ptb::build_pattern(V->dofmap.data(), V->nd, V->n_owned, adj, rowptr, cols);
      ok(c, ptb_set_mesh(c, mesh.n_vertices_local(), mesh.x.data(), mesh.n_cells_local(),
                         mesh.x_dofmap.data()));
      ok(c, ptb_set_space(c, elasticity ? PTB_ELASTICITY : PTB_POISSON, order, V->bs, V->n_owned,
                          V->n_ghost, V->dofmap.data()));
      ok(c, ptb_set_pattern(c, rowptr.data(), cols.data()));
      ok(c, ptb_set_bc(c, static_cast<std::int32_t>(bdofs.size()), bdofs.data()));
      ok(c, ptb_set_exterior_facets(c, static_cast<std::int64_t>(fc.size()), fc.data(), fl.data()));
      ok(c, ptb_set_source(c, f.data(), g.empty() ? nullptr : g.data()));
    }
    if (boot.world() > 1)
    {
      ok(c, ptb_set_halo(c, static_cast<int>(V->nbr_ranks.size()), V->nbr_ranks.data(),
                         V->send_displ.data(), V->local_indices.data(), V->recv_displ.data(),
                         V->remote_indices.data()));
      if (opt.comm == "nccl")
      {
        char id[128] = {};
        if (boot.rank() == 0)
          ok(nullptr, ptb_nccl_unique_id(id));
        const auto all = boot.allgather(id, 128);
        ok(c, ptb_comm_init(c, boot.rank(), boot.world(), all[0].data()));
      }
      else
      {
        // IPC handles, then the owners' send lists (what Scatterer construction exchanges)
        char h[192];
        ok(c, ptb_peer_export(c, h));
        const auto handles = boot.allgather(h, 192);
        std::vector<char> flat;
        for (auto& b : handles)
          flat.insert(flat.end(), b.begin(), b.end());
        // blob: n_nbr, nbr ranks, send_displ, local_indices
        std::vector<std::int32_t> blob;
        blob.push_back(static_cast<std::int32_t>(V->nbr_ranks.size()));
        blob.insert(blob.end(), V->nbr_ranks.begin(), V->nbr_ranks.end());
        blob.insert(blob.end(), V->send_displ.begin(), V->send_displ.end());
        blob.insert(blob.end(), V->local_indices.begin(), V->local_indices.end());
        const auto lists = boot.allgather(blob.data(), blob.size() * sizeof(std::int32_t));
        std::vector<std::int32_t> src;
        for (std::int32_t r : V->nbr_ranks)
        {
          const auto* q = reinterpret_cast<const std::int32_t*>(lists[r].data());
          const int nn = q[0];
          const std::int32_t *nbr = q + 1, *sd = q + 1 + nn, *li = q + 1 + nn + nn + 1;
          int j = 0;
          while (j < nn && nbr[j] != boot.rank())
            ++j;
          if (j == nn)
            throw std::runtime_error("halo lists of neighbouring ranks do not match");
          src.insert(src.end(), li + sd[j], li + sd[j + 1]);
        }
        ok(c, ptb_peer_connect(c, boot.rank(), boot.world(), flat.data(), src.data()));
      }
    }
  }

  auto b = std::make_shared<Vector>();
  b->n_owned_entries = static_cast<std::int64_t>(V->n_owned) * V->bs;
  b->array.assign(static_cast<std::size_t>(V->n_owned + V->n_ghost) * V->bs, 0.0);

  if (!cgp)
  {
    Timer t4("ZZZ Assemble matrix");
    ok(c, ptb_assemble_matrix(c));
    t4.stop();
    t4.flush();
  }
  {
    Timer t5("ZZZ Assemble vector");
    ok(c, ptb_assemble_vector(c));
    ok(c, ptb_get_rhs(c, b->array.data()));
    t5.stop();
    t5.flush();
  }
  if (cgp)
  {
    // cgpoisson never assembles A (cgpoisson_problem.cpp:139-141): the operator is the matrix-free
    // action of form M (Poisson.py:33)
    ok(c, ptb_set_operator_mode(c, PTB_OP_MATRIX_FREE));
  }
  if (t1)
  {
    t1->stop();
    t1->flush();
  }
  if (elasticity)
  {
    Timer t6("ZZZ Create near-nullspace"); // only feeds GAMG in the reference: nothing to build
    t6.stop();
    t6.flush();
  }

  auto u = std::make_shared<Vector>();
  u->n_owned_entries = b->n_owned_entries;
  u->array.assign(b->array.size(), 0.0);

  // cgpoisson: linalg::cg(u, b, action, 100, 1e-6) (cgpoisson_problem.cpp:233), no preconditioner
  const int kmax = cgp && !opt.maxit_set ? 100 : opt.ksp_max_it;
  const double rtol = cgp && !opt.rtol_set ? 1e-6 : opt.ksp_rtol;
  const int pc = (cgp || opt.pc_type == "none") ? PTB_PC_NONE : PTB_PC_JACOBI;
  const std::int64_t nglob = ndofs_global;
  const int rank = boot.rank();
  const int world = boot.world();
  const bool block_rows = opt.problem_type == "elasticity"; // the loop pays for block rows
  SolverFunction solver_function = [gpu, kmax, rtol, pc, cgp, nglob, rank, world, block_rows,
                                    &solve_seconds](Vector& u, const Vector& b) {
    ptb_ctx* c = gpu->c;
    int its = 0;
    double rel = 0.0;
    // across GPUs every rank must run the same form of the loop: decide from the global size
    // (include/ptb200.h ptb_set_cg_persistent); a single GPU decides by itself
    if (world > 1)
      ok(c, ptb_set_cg_persistent(c, block_rows && nglob / world <= 3000000 ? 1 : 0));
    ok(c, ptb_set_rhs(c, b.array.data()));
    Timer tcg;
    ok(c, ptb_cg_solve(c, kmax, rtol, pc, &its, &rel));
    tcg.stop();
    solve_seconds = tcg.elapsed().count();
    ok(c, ptb_get_solution(c, u.array.data()));
    if (cgp && rank == 0)
    {
      const double gdofs = (its * static_cast<double>(nglob)) / solve_seconds / 1e9;
      std::cout << "CG matrix-free action processed: " << gdofs << " Gdof/s\n";
    }
    return its;
  };
  return {b, u, solver_function};
}

void solve(const Options& opt, Bootstrap& boot, int local_rank)
{
  if (opt.help)
  {
    if (boot.rank() == 0)
      std::cout << USAGE << std::endl;
    return;
  }
  bool strong_scaling;
  if (opt.scaling_type == "strong")
    strong_scaling = true;
  else if (opt.scaling_type == "weak")
    strong_scaling = false;
  else
    throw std::runtime_error("Scaling type '" + opt.scaling_type + "` unknown");
  if (opt.problem_type != "poisson" && opt.problem_type != "cgpoisson"
      && opt.problem_type != "elasticity")
    throw std::runtime_error("Unknown problem type: " + opt.problem_type);
  if (opt.mesh_type != "cube")
    throw std::runtime_error("mesh_type '" + opt.mesh_type
                             + "': only the cube mesh is built here (spoke mesh is out of scope)");
  if (opt.comm != "peer" && opt.comm != "nccl")
    throw std::runtime_error("--comm must be peer or nccl");

  const std::size_t num_processes = boot.world();
  const int ndofs_per_node = (opt.problem_type == "elasticity") ? 3 : 1;
  const int order = static_cast<int>(opt.order);

  Timer t0("ZZZ Create Mesh");
  const CubeSizing sz
      = cube_mesh_sizing(opt.ndofs, strong_scaling, ndofs_per_node, order, num_processes);
  if (boot.rank() == 0)
  {
    std::cout << "UnitCube (" << sz.Nx << "x" << sz.Ny << "x" << sz.Nz << ") to be refined " << sz.r
              << " times" << std::endl;
    if (sz.r > 0)
      std::cout << "  [b200] generated directly as the (" << (sz.Nx << sz.r) << "x"
                << (sz.Ny << sz.r) << "x" << (sz.Nz << sz.r)
                << ") box: same entity counts as the refined mesh" << std::endl;
  }
  // --device_setup: sizes and ranges only; the arrays are generated on the device in problem()
  const BoxMesh mesh = create_box_mesh(sz.Nx << sz.r, sz.Ny << sz.r, sz.Nz << sz.r, boot.rank(),
                                       boot.world(), !opt.device_setup);
  t0.stop();
  t0.flush();

  Timer t_ent("ZZZ Create facets and facet->cell connectivity");
  // facets are implicit in the structured stand-in; the exterior-facet list is built in problem()
  t_ent.stop();
  t_ent.flush();

  auto gpu = std::make_shared<Gpu>(local_rank);
  std::int64_t num_dofs = 0;
  double solve_seconds = 0.0;
  auto [b, u, solver_function]
      = problem(opt.problem_type, mesh, order, opt, boot, gpu, num_dofs, solve_seconds);

  if (boot.rank() == 0)
  {
    const std::int64_t num_cells = mesh.n_cells_global();
    std::cout << "----------------------------------------------------------------" << std::endl;
    std::cout << "Test problem summary" << std::endl;
    std::cout << "  dolfinx version: n/a (performance-test_b200 host stand-in)" << std::endl;
    std::cout << "  dolfinx hash:    n/a" << std::endl;
    std::cout << "  ufl hash:        n/a (hand-written sm_100a element kernels)" << std::endl;
    std::cout << "  petsc version:   n/a (cg.h-style CG + " << opt.pc_type << " on the GPU)"
              << std::endl;
    std::cout << "  Problem type:    " << opt.problem_type << std::endl;
    std::cout << "  Scaling type:    " << opt.scaling_type << std::endl;
    std::cout << "  Num processes:   " << num_processes << std::endl;
    std::cout << "  Num cells:       " << num_cells << int64_to_human(num_cells) << std::endl;
    std::cout << "  Total degrees of freedom:               " << num_dofs
              << int64_to_human(num_dofs) << std::endl;
    std::cout << "  Average degrees of freedom per process: " << num_dofs / boot.world()
              << std::endl;
    std::cout << "----------------------------------------------------------------" << std::endl;
  }

  Timer t5("ZZZ Solve");
  const int num_iter = solver_function(*u, *b);
  t5.stop();
  t5.flush();

  if (!opt.output.empty() && boot.rank() == 0)
    std::cout << "[b200] --output: XDMF output is out of scope, nothing written" << std::endl;

  ptb::cli::list_timings(boot.rank(), [&](std::vector<double>& v) {
    const auto all = boot.allgather(v.data(), v.size() * sizeof(double));
    for (auto& blob : all)
    {
      const auto* q = reinterpret_cast<const double*>(blob.data());
      for (std::size_t i = 0; i < v.size() && i * sizeof(double) < blob.size(); ++i)
        v[i] = std::max(v[i], q[i]);
    }
  });

  double norm = 0.0;
  ok(gpu->c, ptb_solution_norm(gpu->c, &norm));
  if (boot.rank() == 0)
  {
    std::cout << "*** Number of Krylov iterations: " << num_iter << std::endl;
    std::cout << "*** Solution norm:  " << norm << std::endl;
    std::cout << "[b200] ZZZ Solve throughput: "
              << num_iter * static_cast<double>(num_dofs) / solve_seconds / 1e9
              << " G DOF-iterations/s" << std::endl;
  }
  boot.barrier(); // nobody frees device memory that a peer may still map
}

} // namespace

int main(int argc, char* argv[])
{
  Options opt;
  try
  {
    opt = parse(argc, argv);
  }
  catch (const std::exception& e)
  {
    std::cerr << "error: " << e.what() << std::endl;
    return 1;
  }
  auto env_int = [](const char* k, int d) {
    const char* v = std::getenv(k);
    return v ? std::atoi(v) : d;
  };
  int rank = env_int("RANK", 0), world = env_int("WORLD_SIZE", 1), local = env_int("LOCAL_RANK", rank);
  std::string addr = std::getenv("MASTER_ADDR") ? std::getenv("MASTER_ADDR") : "127.0.0.1";
  int port = env_int("MASTER_PORT", 29500) + 1; // +1: do not collide with a launcher's own store

  if (opt.nprocs > 1 && world == 1)
  {
    // own launcher: fork one rank per GPU before any CUDA call
    world = opt.nprocs;
    port = 20000 + (getpid() % 20000);
    std::vector<pid_t> kids;
    for (int r = 1; r < world; ++r)
    {
      const pid_t p = fork();
      if (p == 0)
      {
        rank = local = r;
        kids.clear();
        break;
      }
      kids.push_back(p);
    }
    if (rank == 0 && !kids.empty())
    {
      int rc = 0;
      try
      {
        Timer ti("Init MPI"); // name kept from src/main.cpp:245
        Bootstrap boot(0, world, addr, port);
        ti.stop();
        ti.flush();
        solve(opt, boot, 0);
      }
      catch (const std::exception& e)
      {
        std::cerr << "terminate called after throwing: " << e.what() << std::endl;
        rc = 1;
      }
      for (pid_t p : kids)
      {
        int st = 0;
        waitpid(p, &st, 0);
        rc |= (WIFEXITED(st) ? WEXITSTATUS(st) : 1);
      }
      return rc;
    }
  }
  try
  {
    Timer ti("Init MPI");
    Bootstrap boot(rank, world, addr, port);
    ti.stop();
    ti.flush();
    solve(opt, boot, local);
  }
  catch (const std::exception& e)
  {
    // the reference lets exceptions escape main (src/main.cpp:243-275): non-zero exit
    std::cerr << "terminate called after throwing: " << e.what() << std::endl;
    return 1;
  }
  return 0;
}
