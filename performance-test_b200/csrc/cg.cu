// ZZZ Solve on the device: linalg::cg of the reference (src/cg.h:38-86) with the optional Jacobi
// extension, fused into three kernels per iteration:
//
//   spmv_sell        y = A p  (the `action` of cg.h:62) + local p.y                 cg.h:62,65
//   cg_update        alpha = rz/py; x += alpha p; r -= alpha y; local r.r, r.z      cg.h:65-74
//   cg_direction     beta = rz'/rz; stopping rule; p = beta p + D^-1 r              cg.h:75-82
//
// Scalars never visit the host: they live in two CgState records indexed by iteration parity, so
// the kernel that writes the next iteration's record never races with readers of the current one.
// Dot products use the deterministic last-block reduction of reduce.cuh (la::inner_product /
// squared_norm sum owned entries only: cg.h:53,65,74).
#include "kernels.h"
#include "peer.cuh"
#include "reduce.cuh"

namespace ptb
{
namespace
{

constexpr int SPMV_THREADS = 256;
constexpr int VEC_THREADS = 256;

template <int BS>
__global__ void __launch_bounds__(SPMV_THREADS)
spmv_sell(SpmvArgs A, const double* __restrict__ p, double* __restrict__ y, CgState* st,
          double* partials, unsigned int* ticket, PeerView P, unsigned int epoch)
{
  __shared__ double red[32];
  if (st != nullptr && st->conv)
    return;
  const int lane = threadIdx.x & 31;
  const int warps_per_cta = SPMV_THREADS / 32;
  const std::int32_t warp0 = blockIdx.x * warps_per_cta + (threadIdx.x >> 5);
  const std::int32_t stride = gridDim.x * warps_per_cta;
  double dotv = 0.0;
  for (std::int32_t slice = warp0; slice < A.n_slices; slice += stride)
  {
    const std::int64_t mo = A.mat_off[slice];
    const int w = static_cast<int>((A.mat_off[slice + 1] - mo) >> 5);
    const std::int32_t row = slice * 32 + lane;
    if constexpr (BS == 1)
    {
      const std::int32_t* __restrict__ cp = A.cols + mo + lane;
      const double* __restrict__ vp = A.vals + mo + lane;
      double sum = 0.0;
      int k = 0;
      for (; k + 4 <= w; k += 4)
      {
        const std::int32_t c0 = cp[(k + 0) * 32], c1 = cp[(k + 1) * 32], c2 = cp[(k + 2) * 32],
                           c3 = cp[(k + 3) * 32];
        const double v0 = vp[(k + 0) * 32], v1 = vp[(k + 1) * 32], v2 = vp[(k + 2) * 32],
                     v3 = vp[(k + 3) * 32];
        sum += v0 * __ldg(p + c0);
        sum += v1 * __ldg(p + c1);
        sum += v2 * __ldg(p + c2);
        sum += v3 * __ldg(p + c3);
      }
      for (; k < w; ++k)
        sum += vp[k * 32] * __ldg(p + cp[k * 32]);
      if (row < A.n_rows)
      {
        y[row] = sum;
        dotv += sum * __ldg(p + row);
      }
    }
    else
    {
      const std::int32_t* __restrict__ cp = A.cols + mo + lane;
      const double* __restrict__ vp = A.vals + mo * 9 + lane;
      double s0 = 0.0, s1 = 0.0, s2 = 0.0;
      for (int k = 0; k < w; ++k)
      {
        const std::int64_t c = cp[k * 32];
        const double* __restrict__ v = vp + static_cast<std::int64_t>(k) * 9 * 32;
        const double p0 = __ldg(p + 3 * c), p1 = __ldg(p + 3 * c + 1), p2 = __ldg(p + 3 * c + 2);
        s0 += v[0 * 32] * p0 + v[1 * 32] * p1 + v[2 * 32] * p2;
        s1 += v[3 * 32] * p0 + v[4 * 32] * p1 + v[5 * 32] * p2;
        s2 += v[6 * 32] * p0 + v[7 * 32] * p1 + v[8 * 32] * p2;
      }
      if (row < A.n_rows)
      {
        const std::int64_t r3 = 3 * static_cast<std::int64_t>(row);
        y[r3] = s0, y[r3 + 1] = s1, y[r3 + 2] = s2;
        dotv += s0 * __ldg(p + r3) + s1 * __ldg(p + r3 + 1) + s2 * __ldg(p + r3 + 2);
      }
    }
  }
  if (st != nullptr)
  {
    double v[1] = {dotv}, out[1];
    if (grid_sum_last_block<1>(v, partials, ticket, red, out) && threadIdx.x == 0)
    {
      if (P.nranks > 1)
        peer_publish(P, epoch, out[0], 0.0); // all-reduce of p.y: every rank gets every partial
      else
        st->py = out[0];
    }
  }
}

// Global sums for the consumer kernels: one thread per CTA collects the nranks partials from the
// local window (peer mode) or reads the locally reduced / NCCL-reduced values.
__device__ __forceinline__ void global_sums(const PeerView& P, unsigned int epoch, double l0,
                                            double l1, double* sh, double& s0, double& s1)
{
  if (P.nranks > 1)
  {
    if (threadIdx.x == 0)
      peer_collect(P, epoch, sh[0], sh[1]);
    __syncthreads();
    s0 = sh[0], s1 = sh[1];
  }
  else
    s0 = l0, s1 = l1;
}

// r = b - y (cg.h:47), p = z = D^-1 r (cg.h:50), local r.r and r.z.
__global__ void __launch_bounds__(VEC_THREADS)
cg_init(std::int64_t n, const double* __restrict__ b, const double* __restrict__ y,
        const double* __restrict__ dinv, double* __restrict__ r, double* __restrict__ p,
        CgState* st, double* partials, unsigned int* ticket, PeerView P, unsigned int epoch)
{
  __shared__ double red[64];
  double v[2] = {0.0, 0.0};
  for (std::int64_t i = blockIdx.x * static_cast<std::int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<std::int64_t>(gridDim.x) * blockDim.x)
  {
    const double ri = -1.0 * y[i] + b[i];
    const double zi = dinv[i] * ri;
    r[i] = ri;
    p[i] = zi;
    v[0] += ri * ri;
    v[1] += ri * zi;
  }
  double out[2];
  if (grid_sum_last_block<2>(v, partials, ticket, red, out) && threadIdx.x == 0)
  {
    if (P.nranks > 1)
      peer_publish(P, epoch, out[0], out[1]);
    else
      st->rr = out[0], st->rz = out[1];
  }
}

__global__ void cg_finish_init(CgState* st, double rtol, PeerView P, unsigned int epoch)
{
  // st->rr, st->rz hold the (all-reduced) initial sums (cg.h:53-55)
  if (P.nranks > 1)
    peer_collect(P, epoch, st->rr, st->rz);
  st->rnorm0 = st->rr;
  st->rnorm = st->rr;
  st->rz_old = st->rz;
  st->rtol2 = rtol * rtol;
  st->py = 0.0;
  st->k = 0;
  st->conv = 0;
}

__global__ void __launch_bounds__(VEC_THREADS)
cg_update(std::int64_t n, const double* __restrict__ p, const double* __restrict__ y,
          const double* __restrict__ dinv, double* __restrict__ x, double* __restrict__ r,
          CgState* cur, double* partials, unsigned int* ticket, PeerView P, unsigned int epoch_in,
          unsigned int epoch_out)
{
  __shared__ double red[64];
  __shared__ double sh[2];
  if (cur->conv)
    return;
  double py, unused;
  global_sums(P, epoch_in, cur->py, 0.0, sh, py, unused);
  const double alpha = cur->rz_old / py; // cg.h:65
  double v[2] = {0.0, 0.0};
  for (std::int64_t i = blockIdx.x * static_cast<std::int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<std::int64_t>(gridDim.x) * blockDim.x)
  {
    x[i] = alpha * p[i] + x[i];             // cg.h:68
    const double ri = -alpha * y[i] + r[i]; // cg.h:71
    r[i] = ri;
    v[0] += ri * ri;                        // cg.h:74
    v[1] += ri * (dinv[i] * ri);
  }
  double out[2];
  if (grid_sum_last_block<2>(v, partials, ticket, red, out) && threadIdx.x == 0)
  {
    if (P.nranks > 1)
      peer_publish(P, epoch_out, out[0], out[1]);
    else
      cur->rr = out[0], cur->rz = out[1];
  }
}

__global__ void __launch_bounds__(VEC_THREADS)
cg_direction(std::int64_t n, const double* __restrict__ r, const double* __restrict__ dinv,
             double* __restrict__ p, const CgState* cur, CgState* nxt, PeerView P,
             unsigned int epoch)
{
  __shared__ double sh[2];
  const bool first = blockIdx.x == 0 && threadIdx.x == 0;
  if (cur->conv)
  {
    if (first)
      *nxt = *cur;
    return;
  }
  double rr, rz;
  global_sums(P, epoch, cur->rr, cur->rz, sh, rr, rz);
  const double beta = rz / cur->rz_old;                 // cg.h:75
  const bool converged = rr / cur->rnorm0 < cur->rtol2; // cg.h:78
  if (first)
  {
    CgState s = *cur;
    s.rz_old = rz;
    s.rnorm = rr;
    s.k = cur->k + 1;
    s.conv = converged ? 1 : 0;
    *nxt = s;
  }
  if (converged)
    return;
  for (std::int64_t i = blockIdx.x * static_cast<std::int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<std::int64_t>(gridDim.x) * blockDim.x)
    p[i] = beta * p[i] + dinv[i] * r[i]; // cg.h:82
}

__global__ void fill_kernel(double* v, std::int64_t n, double value)
{
  for (std::int64_t i = blockIdx.x * static_cast<std::int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<std::int64_t>(gridDim.x) * blockDim.x)
    v[i] = value;
}

// pack_fn / unpack_fn of cgpoisson_problem.cpp:32-44 (block indices, bs values each)
__global__ void pack_kernel(const double* __restrict__ v, const std::int32_t* __restrict__ idx,
                            std::int64_t n, int bs, double* __restrict__ out)
{
  const std::int64_t i = blockIdx.x * static_cast<std::int64_t>(blockDim.x) + threadIdx.x;
  if (i < n * bs)
    out[i] = v[static_cast<std::int64_t>(idx[i / bs]) * bs + i % bs];
}
__global__ void unpack_kernel(const double* __restrict__ in, const std::int32_t* __restrict__ idx,
                              std::int64_t n, int bs, double* __restrict__ v)
{
  const std::int64_t i = blockIdx.x * static_cast<std::int64_t>(blockDim.x) + threadIdx.x;
  if (i < n * bs)
    v[static_cast<std::int64_t>(idx[i / bs]) * bs + i % bs] = in[i];
}

__global__ void __launch_bounds__(VEC_THREADS)
sqnorm_kernel(std::int64_t n, const double* __restrict__ v, double* out, double* partials,
              unsigned int* ticket, PeerView P, unsigned int epoch)
{
  __shared__ double red[32];
  double s[1] = {0.0};
  for (std::int64_t i = blockIdx.x * static_cast<std::int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<std::int64_t>(gridDim.x) * blockDim.x)
    s[0] += v[i] * v[i];
  double o[1];
  if (grid_sum_last_block<1>(s, partials, ticket, red, o) && threadIdx.x == 0)
  {
    if (P.nranks > 1)
    {
      double t0, t1;
      peer_publish(P, epoch, o[0], 0.0);
      peer_collect(P, epoch, t0, t1);
      o[0] = t0;
    }
    *out = o[0];
  }
}

int vec_grid(const ptb_ctx* c, std::int64_t n)
{
  const std::int64_t need = (n + VEC_THREADS - 1) / VEC_THREADS;
  const std::int64_t cap = static_cast<std::int64_t>(c->num_sms) * 8;
  return static_cast<int>(std::max<std::int64_t>(1, std::min(need, cap)));
}

} // namespace

int cg_grid(const ptb_ctx* c)
{
  const std::int64_t need = (c->n_slices + SPMV_THREADS / 32 - 1) / (SPMV_THREADS / 32);
  const std::int64_t cap = static_cast<std::int64_t>(c->num_sms) * 8;
  return static_cast<int>(std::max<std::int64_t>(1, std::min(need, cap)));
}

void launch_spmv(ptb_ctx* c, const double* p, double* y, CgState* st, unsigned int epoch)
{
  SpmvArgs A{c->n_owned, c->n_slices, c->mat_off.p, c->cols.p, c->vals.p};
  const int grid = cg_grid(c);
  const PeerView P = peer_view(c);
  if (c->bs == 1)
    spmv_sell<1><<<grid, SPMV_THREADS, 0, c->stream>>>(A, p, y, st, c->partials.p, c->tickets.p,
                                                       P, epoch);
  else
    spmv_sell<3><<<grid, SPMV_THREADS, 0, c->stream>>>(A, p, y, st, c->partials.p, c->tickets.p,
                                                       P, epoch);
  PTB_CUDA(cudaGetLastError());
  c->launches += 1;
}

void launch_cg_init(ptb_ctx* c, const double* dinv, CgState* st, unsigned int epoch)
{
  const std::int64_t n = static_cast<std::int64_t>(c->n_owned) * c->bs;
  cg_init<<<vec_grid(c, n), VEC_THREADS, 0, c->stream>>>(n, c->b.p, c->y.p, dinv, c->r.p, c->p.p,
                                                         st, c->partials.p, c->tickets.p + 1,
                                                         peer_view(c), epoch);
  PTB_CUDA(cudaGetLastError());
  c->launches += 1;
}

void launch_cg_finish_init(ptb_ctx* c, CgState* st, double rtol, unsigned int epoch)
{
  cg_finish_init<<<1, 1, 0, c->stream>>>(st, rtol, peer_view(c), epoch);
  PTB_CUDA(cudaGetLastError());
  c->launches += 1;
}

void launch_cg_update(ptb_ctx* c, const double* dinv, CgState* cur, unsigned int epoch_in,
                      unsigned int epoch_out)
{
  const std::int64_t n = static_cast<std::int64_t>(c->n_owned) * c->bs;
  cg_update<<<vec_grid(c, n), VEC_THREADS, 0, c->stream>>>(n, c->p.p, c->y.p, dinv, c->x.p, c->r.p,
                                                           cur, c->partials.p, c->tickets.p + 1,
                                                           peer_view(c), epoch_in, epoch_out);
  PTB_CUDA(cudaGetLastError());
  c->launches += 1;
}

void launch_cg_direction(ptb_ctx* c, const double* dinv, const CgState* cur, CgState* nxt,
                         unsigned int epoch)
{
  const std::int64_t n = static_cast<std::int64_t>(c->n_owned) * c->bs;
  cg_direction<<<vec_grid(c, n), VEC_THREADS, 0, c->stream>>>(n, c->r.p, dinv, c->p.p, cur, nxt,
                                                              peer_view(c), epoch);
  PTB_CUDA(cudaGetLastError());
  c->launches += 1;
}

void launch_fill(ptb_ctx* c, double* v, std::int64_t n, double value)
{
  if (n == 0)
    return;
  fill_kernel<<<vec_grid(c, n), VEC_THREADS, 0, c->stream>>>(v, n, value);
  PTB_CUDA(cudaGetLastError());
  c->launches += 1;
}

void launch_pack(ptb_ctx* c, const double* v, const std::int32_t* idx, std::int64_t n, int bs,
                 double* out)
{
  if (n == 0)
    return;
  pack_kernel<<<static_cast<int>((n * bs + 255) / 256), 256, 0, c->stream>>>(v, idx, n, bs, out);
  PTB_CUDA(cudaGetLastError());
  c->launches += 1;
}

void launch_unpack(ptb_ctx* c, const double* in, const std::int32_t* idx, std::int64_t n, int bs,
                   double* v)
{
  if (n == 0)
    return;
  unpack_kernel<<<static_cast<int>((n * bs + 255) / 256), 256, 0, c->stream>>>(in, idx, n, bs, v);
  PTB_CUDA(cudaGetLastError());
  c->launches += 1;
}

void launch_sqnorm(ptb_ctx* c, const double* v, std::int64_t n, double* out_dev)
{
  sqnorm_kernel<<<vec_grid(c, n), VEC_THREADS, 0, c->stream>>>(n, v, out_dev, c->partials.p,
                                                               c->tickets.p + 2, peer_view(c),
                                                               next_red_epoch(c));
  PTB_CUDA(cudaGetLastError());
  c->launches += 1;
}

} // namespace ptb
