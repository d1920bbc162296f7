// Peer-memory (NVLink P2P) primitives used inside the CG kernels -- the NCCL-free communication
// path. Every rank owns a PeerWindow in its own HBM; peers write into it
// with plain remote stores over NVLink and the owner polls its local copy.
//
//  * all-reduce of 1-2 doubles (la::inner_product's MPI_Allreduce, cg.h:53,65,74): LL-style
//    slots. Every 8-byte word carries 4 bytes of payload and the 4-byte epoch, so a single
//    aligned 8-byte store publishes data and flag atomically -- no fence on the critical path.
//    Each rank writes its partial sums into slot [epoch & 3][its rank] of EVERY rank's window;
//    consumers add the nranks partials in rank order, so all ranks get bit-identical sums.
//  * halo readiness: one monotone 64-bit epoch per source rank, written with st.release.sys
//    after the producer's vector is complete, polled with ld.acquire.sys by the consumer before
//    it pulls the ghost values straight out of the owner's vector.
#pragma once
#include <cstdint>

namespace ptb
{

constexpr int PTB_MAX_RANKS = 16;
constexpr int PTB_MAX_NBR = 8;

struct PeerWindow
{
  unsigned long long halo_flag[PTB_MAX_RANKS]; // [source rank] last published halo epoch
  unsigned long long red[4][PTB_MAX_RANKS][4]; // [slot][source rank][{lo0,hi0,lo1,hi1} | epoch<<32]
};

struct PeerView
{
  int rank, nranks; // nranks == 1: single GPU, nothing below is touched
  PeerWindow* win[PTB_MAX_RANKS]; // win[rank] is the local window
};

struct PeerHalo
{
  int n_nbr, bs;
  int nbr_rank[PTB_MAX_NBR];
  int recv_displ[PTB_MAX_NBR + 1];
  const double* peer_x[PTB_MAX_NBR]; // neighbours' solution vectors (owned part is read)
  const double* peer_p[PTB_MAX_NBR]; // neighbours' search directions
  const std::int32_t* remote_indices; // ghost positions in my vector, per receive entry
  const std::int32_t* src_index;      // owner-local index of each receive entry
};

} // namespace ptb
