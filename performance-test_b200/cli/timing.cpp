#include "timing.h"
#include <algorithm>
#include <cstdio>
#include <iostream>

namespace ptb::cli
{
std::map<std::string, TimingRecord>& timing_registry()
{
  static std::map<std::string, TimingRecord> reg;
  return reg;
}
std::vector<std::string>& timing_order()
{
  static std::vector<std::string> order;
  return order;
}

void list_timings(int rank, const std::function<void(std::vector<double>&)>& reduce_max)
{
  auto& reg = timing_registry();
  std::vector<std::string> names = timing_order();
  std::sort(names.begin(), names.end()); // DOLFINx lists timers alphabetically
  std::vector<double> tot;
  for (auto& n : names)
    tot.push_back(reg[n].total);
  if (reduce_max)
    reduce_max(tot);
  if (rank != 0)
    return;
  std::size_t wname = 25;
  for (auto& n : names)
    wname = std::max(wname, n.size());
  std::printf("\n[MPI_MAX] Summary of timings %*s |  reps  wall avg  wall tot\n",
              static_cast<int>(wname - 19), "");
  std::printf("%s\n", std::string(wname + 38, '-').c_str());
  for (std::size_t i = 0; i < names.size(); ++i)
  {
    const auto& r = reg[names[i]];
    std::printf("%-*s |  %4d  %8.6f  %8.6f\n", static_cast<int>(wname + 10), names[i].c_str(),
                r.reps, tot[i] / std::max(1, r.reps), tot[i]);
  }
  std::cout << std::flush;
}
} // namespace ptb::cli
