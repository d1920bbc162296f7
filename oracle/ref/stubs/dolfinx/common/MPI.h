// TEST INFRASTRUCTURE. Stand-in for <dolfinx/common/MPI.h>: the "communicator" of the stub world
// is a set of host threads (one per emulated MPI rank) that meet in a barrier. Only what
// oracle/ref/stubs/dolfinx/la/Vector.h needs for its all-reduce lives here.
#pragma once
#include <barrier>
#include <cstddef>
#include <vector>

namespace refstub
{
struct World
{
  explicit World(int n) : nranks(n), bar(n), slots(n, 0.0) {}
  int nranks;
  std::barrier<> bar;
  std::vector<double> slots;
};

struct Rank
{
  World* world = nullptr;
  int rank = 0;
};

// The rank the calling thread plays (set by the shim before it enters linalg::cg).
inline thread_local Rank this_rank;

// MPI_Allreduce(SUM, 1 double): partial sums added in rank order, so every rank gets the same bits.
inline double allreduce_sum(double local)
{
  World* w = this_rank.world;
  if (!w or w->nranks == 1)
    return local;
  w->slots[this_rank.rank] = local;
  w->bar.arrive_and_wait();
  double s = 0;
  for (int q = 0; q < w->nranks; ++q)
    s += w->slots[q];
  w->bar.arrive_and_wait();
  return s;
}
} // namespace refstub
