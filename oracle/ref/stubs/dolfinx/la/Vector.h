// TEST INFRASTRUCTURE. The minimum of dolfinx::la::Vector<T> that /root/reference/src/cg.h touches,
// written from DOLFINx's documented behaviour: array() is [owned*bs | ghost*bs];
// la::inner_product / la::squared_norm reduce the OWNED entries, then all-reduce over the ranks.
// This is not DOLFINx code; it exists so that cg.h can be compiled UNCHANGED into oracle/_ref.
#pragma once
#include "../common/MPI.h"
#include <cstdint>
#include <numeric>
#include <span>
#include <vector>

namespace dolfinx::la
{
template <typename T>
class Vector
{
public:
  using value_type = T;
  Vector(std::int32_t size_local, std::int32_t num_ghosts, int bs)
      : _size_local(size_local), _num_ghosts(num_ghosts), _bs(bs),
        _x(static_cast<std::size_t>(bs) * (size_local + num_ghosts), T(0))
  {
  }
  Vector(const Vector&) = default;
  Vector(Vector&&) = default;
  Vector& operator=(const Vector&) = default;
  Vector& operator=(Vector&&) = default;

  std::span<const T> array() const { return std::span<const T>(_x); }
  std::span<T> array() { return std::span<T>(_x); }
  std::int32_t size_local() const { return _size_local; }
  std::int32_t num_ghosts() const { return _num_ghosts; }
  int bs() const { return _bs; }

private:
  std::int32_t _size_local, _num_ghosts;
  int _bs;
  std::vector<T> _x;
};

template <class V>
auto inner_product(const V& a, const V& b)
{
  using T = typename V::value_type;
  const std::size_t n = static_cast<std::size_t>(a.bs()) * a.size_local();
  std::span<const T> xa = a.array(), xb = b.array();
  const T local = std::transform_reduce(xa.begin(), xa.begin() + n, xb.begin(), T(0), std::plus{},
                                        [](T u, T v) -> T { return u * v; });
  return static_cast<T>(refstub::allreduce_sum(local));
}

template <class V>
auto squared_norm(const V& a)
{
  return inner_product(a, a);
}
} // namespace dolfinx::la
