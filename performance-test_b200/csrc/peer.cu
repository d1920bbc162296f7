// Peer-memory communication: IPC bootstrap and the halo kernels (see peer.cuh).
#include "comm.h"
#include "kernels.h"
#include "peer.cuh"
#include <cstring>

namespace ptb
{
namespace
{
// Forward scatter owner -> ghost as a PULL: signal "my vector is complete" to the neighbours,
// wait for theirs, then copy the ghost values straight out of the owners' vectors over NVLink
// (the Scatterer's pack / send / recv / unpack of cgpoisson_problem.cpp:224-229 in one kernel).
__global__ void __launch_bounds__(256)
halo_pull(PeerView P, PeerHalo H, double* __restrict__ v, int which, unsigned long long epoch)
{
  if (blockIdx.x == 0 && threadIdx.x < H.n_nbr)
  {
    __threadfence_system(); // the producer kernel's stores are complete; order them before the flag
    st_release_sys(&P.win[H.nbr_rank[threadIdx.x]]->halo_flag[P.rank], epoch);
  }
  if (threadIdx.x < H.n_nbr)
  {
    const unsigned long long* flag = &P.win[P.rank]->halo_flag[H.nbr_rank[threadIdx.x]];
    while (ld_acquire_sys(flag) < epoch)
    {
    }
  }
  __syncthreads();
  const std::int64_t n = static_cast<std::int64_t>(H.recv_displ[H.n_nbr]) * H.bs;
  for (std::int64_t i = blockIdx.x * static_cast<std::int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<std::int64_t>(gridDim.x) * blockDim.x)
  {
    const std::int32_t j = static_cast<std::int32_t>(i / H.bs), c = static_cast<std::int32_t>(i % H.bs);
    int nb = 0;
    while (j >= H.recv_displ[nb + 1])
      ++nb;
    const double* src = which == 0 ? H.peer_x[nb] : H.peer_p[nb];
    v[static_cast<std::int64_t>(H.remote_indices[j]) * H.bs + c]
        = __ldcv(src + static_cast<std::int64_t>(H.src_index[j]) * H.bs + c);
  }
}

// Neighbour barrier: nobody returns before all its neighbours have arrived (protects vectors that
// peers may still be reading from being overwritten by the next host call).
__global__ void peer_barrier(PeerView P, PeerHalo H, unsigned long long epoch)
{
  if (threadIdx.x < H.n_nbr)
  {
    __threadfence_system();
    st_release_sys(&P.win[H.nbr_rank[threadIdx.x]]->halo_flag[P.rank], epoch);
    const unsigned long long* flag = &P.win[P.rank]->halo_flag[H.nbr_rank[threadIdx.x]];
    while (ld_acquire_sys(flag) < epoch)
    {
    }
  }
}
} // namespace

PeerView peer_view(const ptb_ctx* c)
{
  PeerView V{};
  V.rank = c->rank;
  V.nranks = c->peer.enabled ? c->nranks : 1;
  for (int r = 0; r < PTB_MAX_RANKS; ++r)
    V.win[r] = static_cast<PeerWindow*>(c->peer.win[r]);
  return V;
}

PeerHalo peer_halo(const ptb_ctx* c)
{
  PeerHalo H{};
  H.n_nbr = static_cast<int>(c->nbr_ranks.size());
  H.bs = c->bs;
  for (int i = 0; i < H.n_nbr; ++i)
  {
    H.nbr_rank[i] = c->nbr_ranks[i];
    H.peer_x[i] = static_cast<const double*>(c->peer.nbr_x[i]);
    H.peer_p[i] = static_cast<const double*>(c->peer.nbr_p[i]);
  }
  for (int i = 0; i <= H.n_nbr; ++i)
    H.recv_displ[i] = c->recv_displ[i];
  H.remote_indices = c->recv_idx.p;
  H.src_index = c->peer.src_index.p;
  return H;
}

void peer_export(ptb_ctx* c, void* handles)
{
  if (!c->have_space)
    throw std::runtime_error("ptb_peer_export: call ptb_set_space first");
  if (!c->peer.window.p)
  {
    c->peer.window.alloc(sizeof(PeerWindow));
    c->peer.window.zero(c->stream);
    PTB_CUDA(cudaStreamSynchronize(c->stream));
  }
  cudaIpcMemHandle_t h[3];
  PTB_CUDA(cudaIpcGetMemHandle(&h[0], c->x.p));
  PTB_CUDA(cudaIpcGetMemHandle(&h[1], c->p.p));
  PTB_CUDA(cudaIpcGetMemHandle(&h[2], c->peer.window.p));
  static_assert(sizeof(h) == 192, "three 64-byte IPC handles");
  std::memcpy(handles, h, sizeof(h));
}

void peer_disconnect(ptb_ctx* c)
{
  for (void* p : c->peer.opened)
    cudaIpcCloseMemHandle(p);
  c->peer.opened.clear();
  c->peer.enabled = false;
}

void peer_connect(ptb_ctx* c, int rank, int nranks, const void* all_handles,
                  const std::int32_t* src_index)
{
  if (nranks < 1 || nranks > PTB_MAX_RANKS || rank < 0 || rank >= nranks)
    throw std::runtime_error("ptb_peer_connect: bad rank / nranks (max 16 ranks)");
  if (!c->peer.window.p)
    throw std::runtime_error("ptb_peer_connect: call ptb_peer_export first");
  if (static_cast<int>(c->nbr_ranks.size()) > PTB_MAX_NBR)
    throw std::runtime_error("ptb_peer_connect: more than 8 neighbours");
  peer_disconnect(c);
  c->rank = rank, c->nranks = nranks;
  const auto* H = static_cast<const cudaIpcMemHandle_t*>(all_handles);
  auto open = [&](const cudaIpcMemHandle_t& h) {
    void* p = nullptr;
    PTB_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    c->peer.opened.push_back(p);
    return p;
  };
  for (int r = 0; r < nranks; ++r)
    c->peer.win[r] = r == rank ? static_cast<void*>(c->peer.window.p) : open(H[3 * r + 2]);
  for (std::size_t i = 0; i < c->nbr_ranks.size(); ++i)
  {
    const int r = c->nbr_ranks[i];
    if (r < 0 || r >= nranks || r == rank)
      throw std::runtime_error("ptb_peer_connect: neighbour rank out of range");
    c->peer.nbr_x[i] = open(H[3 * r + 0]);
    c->peer.nbr_p[i] = open(H[3 * r + 1]);
  }
  const std::int64_t n_recv = c->recv_displ.empty() ? 0 : c->recv_displ.back();
  if (n_recv > 0 && !src_index)
    throw std::runtime_error("ptb_peer_connect: src_index is NULL");
  c->peer.src_index.upload(src_index, n_recv, c->stream);
  c->peer.ready.alloc(256);
  c->peer.ready.zero(c->stream);
  PTB_CUDA(cudaStreamSynchronize(c->stream));
  c->peer.enabled = nranks > 1;
}

void peer_halo_forward(ptb_ctx* c, double* v)
{
  const int which = v == c->x.p ? 0 : 1;
  if (v != c->x.p && v != c->p.p)
    throw std::runtime_error("peer halo: only the solution and search-direction vectors are exported");
  const PeerHalo H = peer_halo(c);
  const std::int64_t n = static_cast<std::int64_t>(H.recv_displ[H.n_nbr]) * H.bs;
  const int grid = static_cast<int>(std::max<std::int64_t>(1, std::min<std::int64_t>((n + 255) / 256, c->num_sms)));
  halo_pull<<<grid, 256, 0, c->stream>>>(peer_view(c), H, v, which, ++c->peer.halo_epoch);
  PTB_CUDA(cudaGetLastError());
  c->launches += 1;
}

void peer_neighbour_barrier(ptb_ctx* c)
{
  if (!c->peer.enabled)
    return;
  peer_barrier<<<1, 32, 0, c->stream>>>(peer_view(c), peer_halo(c), ++c->peer.halo_epoch);
  PTB_CUDA(cudaGetLastError());
  c->launches += 1;
}

} // namespace ptb
