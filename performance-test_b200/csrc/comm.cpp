#include "comm.h"
#include "kernels.h"
#include <cstring>
#include <dlfcn.h>
#include <mutex>

namespace ptb
{
namespace
{
// Minimal NCCL surface (nccl.h 2.27/2.28; signatures stable since 2.7).
using ncclComm_t = void*;
struct ncclUniqueId
{
  char internal[128];
};
constexpr int ncclFloat64 = 8, ncclSum = 0;
struct Nccl
{
  void* h = nullptr;
  int (*GetUniqueId)(ncclUniqueId*) = nullptr;
  int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};

Nccl& nccl()
{
  static Nccl N;
  static std::once_flag once;
  std::call_once(once, [] {
    for (const char* name : {"libnccl.so.2", "libnccl.so"})
      if ((N.h = dlopen(name, RTLD_NOW | RTLD_GLOBAL)))
        break;
    if (!N.h)
      return;
    auto sym = [&](const char* s) { return dlsym(N.h, s); };
    N.GetUniqueId = reinterpret_cast<decltype(N.GetUniqueId)>(sym("ncclGetUniqueId"));
    N.CommInitRank = reinterpret_cast<decltype(N.CommInitRank)>(sym("ncclCommInitRank"));
    N.CommDestroy = reinterpret_cast<decltype(N.CommDestroy)>(sym("ncclCommDestroy"));
    N.AllReduce = reinterpret_cast<decltype(N.AllReduce)>(sym("ncclAllReduce"));
    N.Send = reinterpret_cast<decltype(N.Send)>(sym("ncclSend"));
    N.Recv = reinterpret_cast<decltype(N.Recv)>(sym("ncclRecv"));
    N.GroupStart = reinterpret_cast<decltype(N.GroupStart)>(sym("ncclGroupStart"));
    N.GroupEnd = reinterpret_cast<decltype(N.GroupEnd)>(sym("ncclGroupEnd"));
    N.GetErrorString = reinterpret_cast<decltype(N.GetErrorString)>(sym("ncclGetErrorString"));
  });
  if (!N.h || !N.GetUniqueId || !N.CommInitRank || !N.AllReduce || !N.Send || !N.Recv
      || !N.GroupStart || !N.GroupEnd)
    throw std::runtime_error("NCCL (libnccl.so.2) is not available in this process");
  return N;
}

void check(int rc, const char* what)
{
  if (rc != 0)
  {
    Nccl& N = nccl();
    throw std::runtime_error(std::string(what) + ": "
                             + (N.GetErrorString ? N.GetErrorString(rc) : "NCCL error"));
  }
}
} // namespace

void nccl_unique_id(void* out128)
{
  ncclUniqueId id;
  check(nccl().GetUniqueId(&id), "ncclGetUniqueId");
  std::memcpy(out128, id.internal, 128);
}

void comm_init(ptb_ctx* c, int rank, int nranks, const void* id128)
{
  if (nranks < 1 || rank < 0 || rank >= nranks)
    throw std::runtime_error("comm_init: bad rank / nranks");
  c->rank = rank, c->nranks = nranks;
  // Choosing NCCL is a decision for the whole job: a rank whose peer-memory setup succeeded must
  // not keep reducing through its windows while another rank, whose setup failed, waits in NCCL.
  peer_disconnect(c);
  if (nranks == 1)
    return;
  ncclUniqueId id;
  std::memcpy(id.internal, id128, 128);
  ncclComm_t comm = nullptr;
  check(nccl().CommInitRank(&comm, nranks, id, rank), "ncclCommInitRank");
  c->nccl_comm = comm;
}

void comm_destroy(ptb_ctx* c)
{
  if (c->nccl_comm)
    nccl().CommDestroy(c->nccl_comm);
  c->nccl_comm = nullptr;
}

void allreduce_sum(ptb_ctx* c, double* dev, int n)
{
  if (c->nranks == 1 || c->peer.enabled) // peer mode: the kernels reduce through the windows
    return;
  if (!c->nccl_comm)
    throw std::runtime_error("allreduce: communicator not initialised (ptb_comm_init)");
  check(nccl().AllReduce(dev, dev, n, ncclFloat64, ncclSum, c->nccl_comm, c->stream),
        "ncclAllReduce");
}

void halo_forward(ptb_ctx* c, double* v)
{
  if (c->nranks == 1 || c->nbr_ranks.empty())
    return;
  if (c->peer.enabled)
    return peer_halo_forward(c, v);
  if (!c->nccl_comm)
    throw std::runtime_error("halo: communicator not initialised (ptb_comm_init)");
  Nccl& N = nccl();
  const int bs = c->bs;
  const std::int64_t n_send = c->send_displ.back(), n_recv = c->recv_displ.back();
  launch_pack(c, v, c->send_idx.p, n_send, bs, c->send_buf.p);
  check(N.GroupStart(), "ncclGroupStart");
  for (std::size_t i = 0; i < c->nbr_ranks.size(); ++i)
  {
    const std::int64_t s0 = c->send_displ[i], s1 = c->send_displ[i + 1];
    const std::int64_t r0 = c->recv_displ[i], r1 = c->recv_displ[i + 1];
    if (s1 > s0)
      check(N.Send(c->send_buf.p + s0 * bs, (s1 - s0) * bs, ncclFloat64, c->nbr_ranks[i],
                   c->nccl_comm, c->stream),
            "ncclSend");
    if (r1 > r0)
      check(N.Recv(c->recv_buf.p + r0 * bs, (r1 - r0) * bs, ncclFloat64, c->nbr_ranks[i],
                   c->nccl_comm, c->stream),
            "ncclRecv");
  }
  check(N.GroupEnd(), "ncclGroupEnd");
  launch_unpack(c, c->recv_buf.p, c->recv_idx.p, n_recv, bs, v);
}

} // namespace ptb
