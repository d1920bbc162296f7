/* TEST INFRASTRUCTURE (oracle) -- never linked, imported or called by the product path.
 *
 * CPU restatement of the reference's hot path (FEniCS/performance-test), used only by tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg.
 *
 * Pinning: the reference holds no golden vectors or known-answer tests for this path
 * (.github/workflows/ccpp.yml:56-197 only checks exit codes) and as a whole cannot be built here
 * (DOLFINx, Basix, FFCx, PETSc, MPI are not vendored and not installed). Its own self-contained
 * code CAN: src/cg.h is compiled unchanged into oracle/_ref/libref.so (oracle/ref/), and orc_cg
 * with precond = 0 reproduces it bit for bit (tests/test_ref_pin.py: same iteration counts, same
 * x). Everything else in this file is "parity unpinned": the Jacobi extension of the loop, and
 * what is restated from the published behaviour of the un-vendored
 * dependencies (DOLFINx main / FFCx main / PETSc, unpinned by the reference: ccpp.yml:31-50):
 * fem::assemble_matrix / assemble_vector / set_diagonal / DirichletBC::set and
 * MatSetValuesLocal(ADD_VALUES), as called at src/poisson_problem.cpp:125-157 and
 * src/elasticity_problem.cpp:199-231. The oracle is pinned by the analytic known-answer tests
 * of SURVEY 4.3 (tests/test_oracle_kats.py).
 *
 * Build: gcc -O3 -fopenmp -shared -fPIC (oracle/Makefile). With nthreads == 1 every loop runs in
 * the reference's sequential order; nthreads > 1 is only for the timed CPU baseline.
 */
#include <math.h>
#include <omp.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_POISSON 0
#define ORC_ELASTICITY 1

/* ------------------------------------------------------------------------------------------ */
/* tabulate_tensor equivalents (FFCx-generated code in the reference)           */
/* ------------------------------------------------------------------------------------------ */

/* Affine geometry: J[a][b] = sum_v x_v[a] * dgeo[v][b]; returns detJ, fills K = J^-1. */
static double geometry(const double* cd /*[4][3]*/, const double* dgeo /*[4][3]*/, double J[3][3],
                       double K[3][3])
{
  for (int a = 0; a < 3; ++a)
    for (int b = 0; b < 3; ++b)
    {
      double s = 0.0;
      for (int v = 0; v < 4; ++v)
        s += cd[3 * v + a] * dgeo[3 * v + b];
      J[a][b] = s;
    }
  const double c00 = J[1][1] * J[2][2] - J[1][2] * J[2][1];
  const double c01 = J[1][2] * J[2][0] - J[1][0] * J[2][2];
  const double c02 = J[1][0] * J[2][1] - J[1][1] * J[2][0];
  const double det = J[0][0] * c00 + J[0][1] * c01 + J[0][2] * c02;
  K[0][0] = c00 / det;
  K[1][0] = c01 / det;
  K[2][0] = c02 / det;
  K[0][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) / det;
  K[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) / det;
  K[2][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) / det;
  K[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) / det;
  K[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) / det;
  K[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) / det;
  return det;
}

/* a = inner(grad u, grad v) dx   (src/Poisson.py:31). A is [nd][nd], zeroed by the caller. */
static void tabulate_a_poisson(double* A, const double* cd, int nd, int nq, const double* w,
                               const double* dphi /*[nq][nd][3]*/, const double* dgeo, double* g)
{
  double J[3][3], K[3][3];
  const double det = geometry(cd, dgeo, J, K);
  const double scale = fabs(det);
  for (int q = 0; q < nq; ++q)
  {
    /* physical gradients g[i][a] = sum_b K[b][a] dphi[q][i][b] */
    for (int i = 0; i < nd; ++i)
      for (int a = 0; a < 3; ++a)
      {
        double s = 0.0;
        for (int b = 0; b < 3; ++b)
          s += K[b][a] * dphi[(q * nd + i) * 3 + b];
        g[3 * i + a] = s;
      }
    const double wq = w[q] * scale;
    for (int i = 0; i < nd; ++i)
      for (int j = 0; j < nd; ++j)
        A[i * nd + j]
            += wq * (g[3 * i] * g[3 * j] + g[3 * i + 1] * g[3 * j + 1] + g[3 * i + 2] * g[3 * j + 2]);
  }
}

/* a = inner(sigma(u), eps(v)) dx, sigma = 2 mu eps + lambda tr(eps) I   (src/Elasticity.py:30-39).
 * Blocked element: local index 3*i + comp, A is [3nd][3nd]. Done "the long way" through explicit
 * strain/stress tensors so that it shares nothing with the closed form in the CUDA kernel. */
static void tabulate_a_elasticity(double* A, const double* cd, int nd, int nq, const double* w,
                                  const double* dphi, const double* dgeo, double* g)
{
  const double E = 1.0e6, nu = 0.3; /* src/Elasticity.py:12-15 */
  const double mu = E / (2.0 * (1.0 + nu));
  const double lmbda = E * nu / ((1.0 + nu) * (1.0 - 2.0 * nu));
  double J[3][3], K[3][3];
  const double det = geometry(cd, dgeo, J, K);
  const double scale = fabs(det);
  const int n = 3 * nd;
  for (int q = 0; q < nq; ++q)
  {
    for (int i = 0; i < nd; ++i)
      for (int a = 0; a < 3; ++a)
      {
        double s = 0.0;
        for (int b = 0; b < 3; ++b)
          s += K[b][a] * dphi[(q * nd + i) * 3 + b];
        g[3 * i + a] = s;
      }
    const double wq = w[q] * scale;
    for (int j = 0; j < nd; ++j)
      for (int b = 0; b < 3; ++b)
      {
        /* trial function u = phi_j e_b: grad u[c][d] = delta_cb g_j[d] */
        double eu[3][3], sig[3][3], tr = 0.0;
        for (int c = 0; c < 3; ++c)
          for (int d = 0; d < 3; ++d)
            eu[c][d] = 0.5 * ((c == b ? g[3 * j + d] : 0.0) + (d == b ? g[3 * j + c] : 0.0));
        for (int c = 0; c < 3; ++c)
          tr += eu[c][c];
        for (int c = 0; c < 3; ++c)
          for (int d = 0; d < 3; ++d)
            sig[c][d] = 2.0 * mu * eu[c][d] + (c == d ? lmbda * tr : 0.0);
        for (int i = 0; i < nd; ++i)
          for (int a = 0; a < 3; ++a)
          {
            double s = 0.0;
            for (int c = 0; c < 3; ++c)
              for (int d = 0; d < 3; ++d)
              {
                const double ev
                    = 0.5 * ((c == a ? g[3 * i + d] : 0.0) + (d == a ? g[3 * i + c] : 0.0));
                s += sig[c][d] * ev;
              }
            A[(3 * i + a) * n + (3 * j + b)] += wq * s;
          }
      }
  }
}

/* L = f v dx (Poisson.py:32 first term; Elasticity.py:40): be[bs*i+k] += w |detJ| f_h[k](x_q) phi_i.
 * fe is the cell's packed coefficient [nd][bs] (what pack_coefficients produces). */
static void tabulate_L_cell(double* be, const double* cd, const double* fe, int nd, int bs, int nq,
                            const double* w, const double* phi /*[nq][nd]*/, const double* dgeo)
{
  double J[3][3], K[3][3];
  const double scale = fabs(geometry(cd, dgeo, J, K));
  for (int q = 0; q < nq; ++q)
  {
    double fq[3] = {0.0, 0.0, 0.0};
    for (int j = 0; j < nd; ++j)
      for (int k = 0; k < bs; ++k)
        fq[k] += fe[bs * j + k] * phi[q * nd + j];
    for (int i = 0; i < nd; ++i)
      for (int k = 0; k < bs; ++k)
        be[bs * i + k] += w[q] * scale * fq[k] * phi[q * nd + i];
  }
}

/* L = g v ds on local facet lf (Poisson.py:32 second term). Scale = |J t1 x J t2|. */
static void tabulate_L_facet(double* be, const double* cd, const double* ge, int nd, int lf, int nq,
                             const double* w, const double* phi_f /*[4][nq][nd]*/,
                             const double* dgeo, const double* facet_t /*[4][2][3]*/)
{
  double J[3][3], K[3][3];
  geometry(cd, dgeo, J, K);
  double a[3], b[3];
  for (int r = 0; r < 3; ++r)
  {
    a[r] = b[r] = 0.0;
    for (int c = 0; c < 3; ++c)
    {
      a[r] += J[r][c] * facet_t[(lf * 2 + 0) * 3 + c];
      b[r] += J[r][c] * facet_t[(lf * 2 + 1) * 3 + c];
    }
  }
  const double cx = a[1] * b[2] - a[2] * b[1], cy = a[2] * b[0] - a[0] * b[2],
               cz = a[0] * b[1] - a[1] * b[0];
  const double scale = sqrt(cx * cx + cy * cy + cz * cz);
  const double* phi = phi_f + (size_t)lf * nq * nd;
  for (int q = 0; q < nq; ++q)
  {
    double gq = 0.0;
    for (int j = 0; j < nd; ++j)
      gq += ge[j] * phi[q * nd + j];
    for (int i = 0; i < nd; ++i)
      be[i] += w[q] * scale * gq * phi[q * nd + i];
  }
}

/* ------------------------------------------------------------------------------------------ */
/* fem::assemble_matrix + MatSetValuesBlockedLocal(ADD_VALUES)                                 */
/* (src/poisson_problem.cpp:129-131, src/elasticity_problem.cpp:203-205)    */
/* ------------------------------------------------------------------------------------------ */

static int64_t find_col(const int32_t* cols, int64_t lo, int64_t hi, int32_t c)
{
  while (lo < hi)
  {
    const int64_t mid = (lo + hi) >> 1;
    if (cols[mid] < c)
      lo = mid + 1;
    else
      hi = mid;
  }
  return lo;
}

/* vals: block CSR, block (row r, slot s) entry (a, b) at vals[9*s + 3*a + b] (bs = 3) or vals[s].
 * Rows >= n_rows (ghost rows) are dropped: the caller passes all cells adjacent to its owned rows
 * (ghost-cell layer instead of the MatAssembly stash exchange).
 * slot: optional precomputed cell->CSR-slot map [ncells][nd][nd] (int64, -1 = not owned); when
 * NULL every insertion searches the row like MatSetValuesLocal does.
 * Returns 0, or 1 if a (row, col) pair is missing from the pattern. */
int orc_assemble_matrix(int problem, int nd, int bs, int nq, const double* w, const double* dphi,
                        const double* dgeo, int64_t ncells, const double* x,
                        const int32_t* x_dofmap, const int32_t* dofmap, int32_t n_rows,
                        const int64_t* rowptr, const int32_t* cols, const int8_t* bc_marker,
                        const int64_t* slot, double* vals, int nthreads)
{
  const int n = nd * bs;
  const int bs2 = bs * bs;
  int err = 0;
  if (nthreads < 1)
    nthreads = 1;
#pragma omp parallel num_threads(nthreads) reduction(| : err)
  {
    double* Ae = (double*)malloc(sizeof(double) * n * n);
    double* g = (double*)malloc(sizeof(double) * 3 * nd);
    /* static contiguous chunks: with one thread this is the reference's cell order */
#pragma omp for schedule(static)
    for (int64_t c = 0; c < ncells; ++c)
    {
      double cd[12];
      for (int v = 0; v < 4; ++v)
        for (int a = 0; a < 3; ++a)
          cd[3 * v + a] = x[3 * (int64_t)x_dofmap[4 * c + v] + a];
      memset(Ae, 0, sizeof(double) * n * n);
      if (problem == ORC_POISSON)
        tabulate_a_poisson(Ae, cd, nd, nq, w, dphi, dgeo, g);
      else
        tabulate_a_elasticity(Ae, cd, nd, nq, w, dphi, dgeo, g);
      const int32_t* dofs = dofmap + c * nd;
      /* zero rows/cols of constrained dofs before insertion (all bs components constrained) */
      for (int i = 0; i < nd; ++i)
        if (bc_marker[dofs[i]])
          for (int k = 0; k < bs; ++k)
            for (int col = 0; col < n; ++col)
            {
              Ae[(bs * i + k) * n + col] = 0.0;
              Ae[col * n + (bs * i + k)] = 0.0;
            }
      for (int i = 0; i < nd; ++i)
      {
        const int32_t r = dofs[i];
        if (r >= n_rows)
          continue;
        for (int j = 0; j < nd; ++j)
        {
          int64_t s;
          if (slot)
            s = slot[(c * nd + i) * nd + j];
          else
          {
            s = find_col(cols, rowptr[r], rowptr[r + 1], dofs[j]);
            if (s >= rowptr[r + 1] || cols[s] != dofs[j])
              s = -1;
          }
          if (s < 0)
          {
            err |= 1;
            continue;
          }
          for (int a = 0; a < bs; ++a)
            for (int b = 0; b < bs; ++b)
            {
              const double v = Ae[(bs * i + a) * n + (bs * j + b)];
              if (nthreads > 1)
              {
#pragma omp atomic
                vals[bs2 * s + bs * a + b] += v;
              }
              else
                vals[bs2 * s + bs * a + b] += v;
            }
        }
      }
    }
    free(Ae);
    free(g);
  }
  return err;
}

/* fem::set_diagonal (src/poisson_problem.cpp:134-135): A[d,d] = 1.0 (INSERT) for owned BC dofs. */
int orc_set_diagonal(int bs, int32_t n_rows, const int64_t* rowptr, const int32_t* cols,
                     int32_t n_bc, const int32_t* bc_dofs, double* vals)
{
  const int bs2 = bs * bs;
  for (int32_t k = 0; k < n_bc; ++k)
  {
    const int32_t d = bc_dofs[k];
    if (d >= n_rows)
      continue;
    const int64_t s = find_col(cols, rowptr[d], rowptr[d + 1], d);
    if (s >= rowptr[d + 1] || cols[s] != d)
      return 1;
    for (int a = 0; a < bs; ++a)
      vals[bs2 * s + bs * a + a] = 1.0;
  }
  return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* fem::assemble_vector (cells, then exterior facets) + DirichletBC::set                        */
/* (src/poisson_problem.cpp:147-155, src/elasticity_problem.cpp:221-229)  */
/* ------------------------------------------------------------------------------------------ */
int orc_assemble_vector(int nd, int bs, int nq_l, const double* w_l, const double* phi_l,
                        int nq_f, const double* w_f, const double* phi_f, const double* dgeo,
                        const double* facet_t, int64_t ncells, const double* x,
                        const int32_t* x_dofmap, const int32_t* dofmap, int32_t n_rows,
                        const double* f, const double* g, int64_t n_facets,
                        const int32_t* facet_cells, const int32_t* facet_local, int32_t n_bc,
                        const int32_t* bc_dofs, double* b /*[n_rows*bs]*/)
{
  const int n = nd * bs;
  double* be = (double*)malloc(sizeof(double) * n);
  double* fe = (double*)malloc(sizeof(double) * n);
  for (int64_t c = 0; c < ncells; ++c)
  {
    double cd[12];
    for (int v = 0; v < 4; ++v)
      for (int a = 0; a < 3; ++a)
        cd[3 * v + a] = x[3 * (int64_t)x_dofmap[4 * c + v] + a];
    const int32_t* dofs = dofmap + c * nd;
    for (int j = 0; j < nd; ++j) /* pack_coefficients */
      for (int k = 0; k < bs; ++k)
        fe[bs * j + k] = f[(int64_t)bs * dofs[j] + k];
    memset(be, 0, sizeof(double) * n);
    tabulate_L_cell(be, cd, fe, nd, bs, nq_l, w_l, phi_l, dgeo);
    for (int i = 0; i < nd; ++i)
      if (dofs[i] < n_rows)
        for (int k = 0; k < bs; ++k)
          b[(int64_t)bs * dofs[i] + k] += be[bs * i + k];
  }
  if (g)
    for (int64_t k = 0; k < n_facets; ++k)
    {
      const int64_t c = facet_cells[k];
      double cd[12];
      for (int v = 0; v < 4; ++v)
        for (int a = 0; a < 3; ++a)
          cd[3 * v + a] = x[3 * (int64_t)x_dofmap[4 * c + v] + a];
      const int32_t* dofs = dofmap + c * nd;
      for (int j = 0; j < nd; ++j)
        fe[j] = g[dofs[j]];
      memset(be, 0, sizeof(double) * n);
      tabulate_L_facet(be, cd, fe, nd, facet_local[k], nq_f, w_f, phi_f, dgeo, facet_t);
      for (int i = 0; i < nd; ++i)
        if (dofs[i] < n_rows)
          b[dofs[i]] += be[i];
    }
  /* bc->set(b, nullopt): b[bc] = g_bc = 0 (u0 = 0, src/poisson_problem.cpp:53-54,155) */
  for (int32_t k = 0; k < n_bc; ++k)
    if (bc_dofs[k] < n_rows)
      for (int a = 0; a < bs; ++a)
        b[(int64_t)bs * bc_dofs[k] + a] = 0.0;
  free(be);
  free(fe);
  return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* Operator application y = A p on owned rows (PETSc MatMult in the reference's KSP,            */
/* src/poisson_problem.cpp:177; the `action` concept of src/cg.h:38-39)                         */
/* ------------------------------------------------------------------------------------------ */
void orc_spmv(int bs, int32_t n_rows, const int64_t* rowptr, const int32_t* cols,
              const double* vals, const double* p, double* y, int nthreads)
{
  if (nthreads < 1)
    nthreads = 1;
  if (bs == 1)
  {
#pragma omp parallel for schedule(static) num_threads(nthreads)
    for (int32_t r = 0; r < n_rows; ++r)
    {
      double s = 0.0;
      for (int64_t k = rowptr[r]; k < rowptr[r + 1]; ++k)
        s += vals[k] * p[cols[k]];
      y[r] = s;
    }
    return;
  }
#pragma omp parallel for schedule(static) num_threads(nthreads)
  for (int32_t r = 0; r < n_rows; ++r)
  {
    double s[3] = {0.0, 0.0, 0.0};
    for (int64_t k = rowptr[r]; k < rowptr[r + 1]; ++k)
    {
      const double* v = vals + 9 * k;
      const double* pc = p + 3 * (int64_t)cols[k];
      for (int a = 0; a < 3; ++a)
        s[a] += v[3 * a] * pc[0] + v[3 * a + 1] * pc[1] + v[3 * a + 2] * pc[2];
    }
    y[3 * (int64_t)r] = s[0], y[3 * (int64_t)r + 1] = s[1], y[3 * (int64_t)r + 2] = s[2];
  }
}

static double dot(const double* a, const double* b, int64_t n, int nthreads)
{
  double s = 0.0;
  if (nthreads <= 1)
  {
    /* la::inner_product / squared_norm (cg.h:53,65,74) reduce the owned entries with
     * std::transform_reduce; libstdc++ evaluates it four entries at a time, ((p0+p1)+(p2+p3))
     * added to the running sum, then the remainder one by one. Restated here so that this
     * loop and the reference's cg.h compiled into oracle/_ref produce the same bits. */
    int64_t i = 0;
    for (; i + 4 <= n; i += 4)
    {
      const double v1 = a[i] * b[i] + a[i + 1] * b[i + 1];
      const double v2 = a[i + 2] * b[i + 2] + a[i + 3] * b[i + 3];
      s = s + (v1 + v2);
    }
    for (; i < n; ++i)
      s = s + a[i] * b[i];
    return s;
  }
#pragma omp parallel for schedule(static) reduction(+ : s) num_threads(nthreads)
  for (int64_t i = 0; i < n; ++i)
    s += a[i] * b[i];
  return s;
}

/* r = alpha*x + y over the whole array (src/cg.h:18-25) */
static void axpy(double* r, double alpha, const double* x, const double* y, int64_t n,
                 int nthreads)
{
#pragma omp parallel for schedule(static) num_threads(nthreads > 1 ? nthreads : 1)
  for (int64_t i = 0; i < n; ++i)
    r[i] = alpha * x[i] + y[i];
}

/* linalg::cg (src/cg.h:38-86) on a single partition, with the optional Jacobi extension the
 * north star asks for: precond = 0 reproduces cg.h exactly (z == r, rz == rnorm);
 * precond = 1 uses z = D^-1 r, alpha = (r.z)/(p.y), beta = (r.z)_new/(r.z). Stopping rule is
 * cg.h's in both cases: ||r||^2 / ||r0||^2 < rtol^2 after the x and r updates (cg.h:74-79).
 * Returns the iteration count; *rel_res = ||r|| / ||r0||. */
int orc_cg(int bs, int32_t n_rows, const int64_t* rowptr, const int32_t* cols, const double* vals,
           const double* b, double* x, int kmax, double rtol, int precond, double* rel_res,
           int nthreads)
{
  const int64_t n = (int64_t)n_rows * bs;
  double* r = (double*)malloc(sizeof(double) * n);
  double* y = (double*)malloc(sizeof(double) * n);
  double* p = (double*)malloc(sizeof(double) * n);
  double* z = (double*)malloc(sizeof(double) * n);
  double* dinv = NULL;
  if (precond)
  {
    dinv = (double*)malloc(sizeof(double) * n);
    for (int32_t row = 0; row < n_rows; ++row)
    {
      const int64_t s = find_col(cols, rowptr[row], rowptr[row + 1], row);
      for (int a = 0; a < bs; ++a)
        dinv[(int64_t)bs * row + a] = 1.0 / vals[bs * bs * s + bs * a + a];
    }
  }
  /* r0 = b - A x0 (cg.h:46-47) */
  orc_spmv(bs, n_rows, rowptr, cols, vals, x, y, nthreads);
  axpy(r, -1.0, y, b, n, nthreads);
  for (int64_t i = 0; i < n; ++i)
    z[i] = precond ? dinv[i] * r[i] : r[i];
  memcpy(p, z, sizeof(double) * n); /* cg.h:50 */
  const double rnorm0 = dot(r, r, n, nthreads);
  const double rtol2 = rtol * rtol;
  double rnorm = rnorm0;
  double rz = precond ? dot(r, z, n, nthreads) : rnorm0;
  int k = 0;
  while (k < kmax)
  {
    ++k;
    orc_spmv(bs, n_rows, rowptr, cols, vals, p, y, nthreads);           /* cg.h:62 */
    const double alpha = rz / dot(p, y, n, nthreads);                   /* cg.h:65 */
    axpy(x, alpha, p, x, n, nthreads);                                  /* cg.h:68 */
    axpy(r, -alpha, y, r, n, nthreads);                                 /* cg.h:71 */
    const double rnorm_new = dot(r, r, n, nthreads);                    /* cg.h:74 */
    double rz_new = rnorm_new;
    if (precond)
    {
#pragma omp parallel for schedule(static) num_threads(nthreads > 1 ? nthreads : 1)
      for (int64_t i = 0; i < n; ++i)
        z[i] = dinv[i] * r[i];
      rz_new = dot(r, z, n, nthreads);
    }
    const double beta = rz_new / rz;                                    /* cg.h:75 */
    rz = rz_new;
    rnorm = rnorm_new;
    if (rnorm / rnorm0 < rtol2)                                         /* cg.h:78 */
      break;
    axpy(p, beta, p, precond ? z : r, n, nthreads);                     /* cg.h:82 */
  }
  if (rel_res)
    *rel_res = sqrt(rnorm / rnorm0);
  free(r), free(y), free(p), free(z), free(dinv);
  return k;
}

int orc_max_threads(void) { return omp_get_max_threads(); }

/* ------------------------------------------------------------------------------------------ */
/* The solve on P partitions at once: the analogue of `mpirun -np P` for the timed CPU arm       */
/* (BASELINE.md section 3 (ii)). One OpenMP thread plays one rank: it owns a block of rows       */
/* [owned | ghost] exactly as a DOLFINx rank does, updates the ghosts of p before every operator */
/* application with the Scatterer lists (pack -> neighbour copy -> unpack,                       */
/* cgpoisson_problem.cpp:32-44,223-229) and reduces the dot products over ranks in rank order    */
/* (MPI_Allreduce in la::inner_product). Same loop as orc_cg (cg.h:38-86 + optional Jacobi).     */
/* ------------------------------------------------------------------------------------------ */
typedef struct
{
  int32_t bs, n_owned, n_ghost, n_nbr;
  const int64_t* rowptr;
  const int32_t* cols;
  const double* vals;
  const double* b; /* owned entries */
  double* x;       /* [(n_owned + n_ghost) * bs] in: initial guess, out: solution */
  const int32_t *nbr_ranks, *send_displ, *local_indices, *recv_displ, *remote_indices;
} orc_part;

static void part_spmv(const orc_part* P, const double* p, double* y)
{
  orc_spmv(P->bs, P->n_owned, P->rowptr, P->cols, P->vals, p, y, 1);
}

int orc_cg_partitioned(int nparts, const orc_part* parts, int kmax, double rtol, int precond,
                       int* iterations, double* rel_res)
{
  double* red = (double*)calloc((size_t)4 * 2 * nparts, sizeof(double)); /* ring of 4 reductions x 2 values */
  double** sendbuf = (double**)calloc(nparts, sizeof(double*));
  int k_out = 0, bad = 0;
  double rel_out = 0.0;
#pragma omp parallel num_threads(nparts)
  {
    const int q = omp_get_thread_num();
    const orc_part* P = &parts[q];
    const int bs = P->bs;
    const int64_t n = (int64_t)P->n_owned * bs, nl = (int64_t)(P->n_owned + P->n_ghost) * bs;
    double* r = (double*)malloc(sizeof(double) * n);
    double* y = (double*)malloc(sizeof(double) * n);
    double* z = (double*)malloc(sizeof(double) * n);
    double* p = (double*)calloc(nl, sizeof(double));
    double* dinv = (double*)malloc(sizeof(double) * n);
    sendbuf[q] = (double*)malloc(sizeof(double) * (size_t)(P->send_displ[P->n_nbr] * bs + 1));
    for (int32_t row = 0; row < P->n_owned; ++row)
    {
      const int64_t s = find_col(P->cols, P->rowptr[row], P->rowptr[row + 1], row);
      for (int a = 0; a < bs; ++a)
        dinv[(int64_t)bs * row + a] = precond ? 1.0 / P->vals[bs * bs * s + bs * a + a] : 1.0;
    }
    int ring = 0;
#define ORC_ALLREDUCE2(v0, v1, s0, s1)                                                            \
  do                                                                                              \
  {                                                                                               \
    double* slot = red + (size_t)(ring & 3) * 2 * nparts;                                         \
    slot[2 * q] = (v0), slot[2 * q + 1] = (v1);                                                   \
    _Pragma("omp barrier") (s0) = 0.0, (s1) = 0.0;                                                \
    for (int t = 0; t < nparts; ++t)                                                              \
      (s0) += slot[2 * t], (s1) += slot[2 * t + 1];                                               \
    ++ring;                                                                                       \
  } while (0)
#define ORC_HALO(v)                                                                               \
  do                                                                                              \
  {                                                                                               \
    for (int32_t i = 0; i < P->send_displ[P->n_nbr]; ++i)                                         \
      for (int a = 0; a < bs; ++a)                                                                \
        sendbuf[q][(int64_t)i * bs + a] = (v)[(int64_t)P->local_indices[i] * bs + a];             \
    _Pragma("omp barrier") for (int i = 0; i < P->n_nbr; ++i)                                     \
    {                                                                                             \
      const orc_part* O = &parts[P->nbr_ranks[i]];                                                \
      int me = 0;                                                                                 \
      while (me < O->n_nbr && O->nbr_ranks[me] != q)                                              \
        ++me;                                                                                     \
      if (me == O->n_nbr                                                                          \
          || O->send_displ[me + 1] - O->send_displ[me] != P->recv_displ[i + 1] - P->recv_displ[i]) \
      {                                                                                           \
        bad = 1;                                                                                  \
        continue;                                                                                 \
      }                                                                                           \
      const double* src = sendbuf[P->nbr_ranks[i]] + (int64_t)O->send_displ[me] * bs;             \
      for (int32_t j = P->recv_displ[i]; j < P->recv_displ[i + 1]; ++j)                           \
        for (int a = 0; a < bs; ++a)                                                              \
          (v)[(int64_t)P->remote_indices[j] * bs + a] = src[(int64_t)(j - P->recv_displ[i]) * bs + a]; \
    }                                                                                             \
    _Pragma("omp barrier")                                                                        \
  } while (0)

    /* r0 = b - A x0 (cg.h:46-47) */
    ORC_HALO(P->x);
    part_spmv(P, P->x, y);
    for (int64_t i = 0; i < n; ++i)
    {
      r[i] = -1.0 * y[i] + P->b[i];
      z[i] = dinv[i] * r[i];
      p[i] = z[i];
    }
    double rnorm0, rz, rnorm;
    ORC_ALLREDUCE2(dot(r, r, n, 1), dot(r, z, n, 1), rnorm0, rz);
    rnorm = rnorm0;
    const double rtol2 = rtol * rtol;
    int k = 0;
    while (k < kmax)
    {
      ++k;
      ORC_HALO(p);
      part_spmv(P, p, y);                                        /* cg.h:62 */
      double py, unused;
      ORC_ALLREDUCE2(dot(p, y, n, 1), 0.0, py, unused);          /* cg.h:65 */
      (void)unused;
      const double alpha = rz / py;
      for (int64_t i = 0; i < n; ++i)
      {
        P->x[i] = alpha * p[i] + P->x[i];                        /* cg.h:68 */
        r[i] = -alpha * y[i] + r[i];                             /* cg.h:71 */
        z[i] = dinv[i] * r[i];
      }
      double rnorm_new, rz_new;
      ORC_ALLREDUCE2(dot(r, r, n, 1), dot(r, z, n, 1), rnorm_new, rz_new); /* cg.h:74 */
      const double beta = rz_new / rz;                           /* cg.h:75 */
      rz = rz_new, rnorm = rnorm_new;
      if (rnorm / rnorm0 < rtol2)                                /* cg.h:78 */
        break;
      for (int64_t i = 0; i < n; ++i)
        p[i] = beta * p[i] + z[i];                               /* cg.h:82 */
    }
    ORC_HALO(P->x); /* ghosts of the solution current on return (KrylovSolver::solve) */
    if (q == 0)
      k_out = k, rel_out = sqrt(rnorm / rnorm0);
    free(r), free(y), free(z), free(p), free(dinv);
#undef ORC_ALLREDUCE2
#undef ORC_HALO
  }
  for (int q = 0; q < nparts; ++q)
    free(sendbuf[q]);
  free(sendbuf), free(red);
  if (iterations)
    *iterations = k_out;
  if (rel_res)
    *rel_res = rel_out;
  return bad;
}
