// Stand-in for dolfinx::common::Timer / dolfinx::list_timings (used at every stage of the
// reference: src/main.cpp:130,142-146,208-211,226; src/poisson_problem.cpp:33,49,51,82,125,146).
// Same usage pattern: construct with a name (starts), stop(), flush(); list_timings prints the
// "Summary of timings" table, reduced (max) over ranks by the caller-supplied reducer.
#pragma once
#include <chrono>
#include <functional>
#include <map>
#include <string>
#include <vector>

namespace ptb::cli
{

struct TimingRecord
{
  int reps = 0;
  double total = 0.0;
};

std::map<std::string, TimingRecord>& timing_registry();
std::vector<std::string>& timing_order();

class Timer
{
public:
  explicit Timer(std::string name = "") : _name(std::move(name)) { start(); }
  ~Timer() { flush(); } // a running timer registers itself on destruction (cgpoisson relies on it)
  void start()
  {
    _t0 = std::chrono::steady_clock::now();
    _running = true;
  }
  void stop()
  {
    if (_running)
      _acc += std::chrono::steady_clock::now() - _t0;
    _running = false;
  }
  std::chrono::duration<double> elapsed() const { return _acc; }
  void flush()
  {
    stop();
    if (_flushed || _name.empty())
      return;
    auto& reg = timing_registry();
    if (!reg.count(_name))
      timing_order().push_back(_name);
    reg[_name].reps += 1;
    reg[_name].total += std::chrono::duration<double>(_acc).count();
    _flushed = true;
  }

private:
  std::string _name;
  std::chrono::steady_clock::time_point _t0;
  std::chrono::duration<double> _acc{0};
  bool _running = false, _flushed = false;
};

/// Print the table on rank 0. `reduce_max` maps a vector of per-rank totals to the max over ranks.
void list_timings(int rank, const std::function<void(std::vector<double>&)>& reduce_max);

} // namespace ptb::cli
