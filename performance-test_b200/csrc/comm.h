// NCCL plumbing for the partitioned solve: ghost update of p before every
// operator application (Scatterer forward scatter, cgpoisson_problem.cpp:224-229) and the global
// dot products (MPI_Allreduce inside la::inner_product / squared_norm, cg.h:53,65,74).
// NCCL is bound with dlopen so the library loads (and the integer side is testable) on machines
// without it, and so that it shares the NCCL already loaded by the host process.
#pragma once
#include "ctx.h"

namespace ptb
{
void nccl_unique_id(void* out128);
void comm_init(ptb_ctx* c, int rank, int nranks, const void* id128);
void comm_destroy(ptb_ctx* c);
/// In-place sum over ranks of n doubles at dev (stream-ordered). No-op on one rank.
void allreduce_sum(ptb_ctx* c, double* dev, int n);
/// Forward scatter: owned values of v -> ghost entries of v on the neighbours (stream-ordered).
void halo_forward(ptb_ctx* c, double* v);

// peer.cu: NCCL-free path over NVLink peer memory
void peer_export(ptb_ctx* c, void* handles192);
void peer_connect(ptb_ctx* c, int rank, int nranks, const void* all_handles,
                  const std::int32_t* src_index);
void peer_disconnect(ptb_ctx* c);
void peer_halo_forward(ptb_ctx* c, double* v);
void peer_neighbour_barrier(ptb_ctx* c);
} // namespace ptb
