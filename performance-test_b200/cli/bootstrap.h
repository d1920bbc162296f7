// Minimal TCP rendezvous for the standalone harness (the reference uses MPI; this image has none).
// Ranks are started by any launcher that exports RANK, WORLD_SIZE, LOCAL_RANK, MASTER_ADDR and
// MASTER_PORT (e.g. `python -m torch.distributed.run --no-python`, or this binary's own --nprocs).
// Only setup data crosses it (NCCL id / IPC handles / halo lists, timing table); nothing on the
// hot path.
#pragma once
#include <cstddef>
#include <string>
#include <vector>

namespace ptb::cli
{
class Bootstrap
{
public:
  Bootstrap(int rank, int world, const std::string& addr, int port);
  ~Bootstrap();
  int rank() const { return _rank; }
  int world() const { return _world; }
  /// Every rank contributes a blob; every rank receives all blobs in rank order.
  std::vector<std::vector<char>> allgather(const void* data, std::size_t n);
  void barrier();

private:
  int _rank, _world;
  int _listen = -1;
  std::vector<int> _peers; // rank 0: socket per rank; others: [0] = socket to rank 0
};
} // namespace ptb::cli
