// Device helpers shared by the assembly kernels: small vector algebra, padded point loads and
// the P1 cofactor geometry (see assemble.cu for the scheme).
#pragma once
#include <cstdint>

namespace ptb
{
namespace
{

struct Vec3
{
  double x, y, z;
};
__device__ __forceinline__ Vec3 operator-(Vec3 a, Vec3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ Vec3 cross(Vec3 a, Vec3 b)
{
  return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
// The same without FMA contraction: with `a.y * b.z - a.z * b.y` contracted into fma(a.y, b.z, -(a.z * b.y))
// a component that cancels analytically (both products of the same factors, as for the axis-aligned
// edges of the reference's lattice) comes out as the rounding error of one product, ~1e-17 of the
// entry's scale, instead of 0.0. Two rounded products and a subtraction give the exact zero, so the
// assembled lattice operator holds exact zeros where the 7-point stencil has none (compact.cu drops them).
__device__ __forceinline__ Vec3 cross_rn(Vec3 a, Vec3 b)
{
#ifdef PTB_HOST_EMU // tests/emu: the products go through volatile so that a harness built with
                    // -ffp-contract=fast (the FMA variant of the tests) cannot contract them either
  volatile double p0 = a.y * b.z, q0 = a.z * b.y, p1 = a.z * b.x, q1 = a.x * b.z, p2 = a.x * b.y, q2 = a.y * b.x;
  return {p0 - q0, p1 - q1, p2 - q2};
#else
  return {__dsub_rn(__dmul_rn(a.y, b.z), __dmul_rn(a.z, b.y)), __dsub_rn(__dmul_rn(a.z, b.x), __dmul_rn(a.x, b.z)),
          __dsub_rn(__dmul_rn(a.x, b.y), __dmul_rn(a.y, b.x))};
#endif
}
__device__ __forceinline__ double dot(Vec3 a, Vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ double comp(Vec3 a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }

__device__ __forceinline__ Vec3 load_point(const double* __restrict__ xyz4, std::int64_t v)
{
  // padded [n][4]: two 16-byte loads
  const double2* p = reinterpret_cast<const double2*>(xyz4 + 4 * v);
  const double2 a = __ldg(p), b = __ldg(p + 1);
  return {a.x, a.y, b.x};
}

__device__ __forceinline__ int sel4(int4 v, int i)
{
  return i == 0 ? v.x : (i == 1 ? v.y : (i == 2 ? v.z : v.w));
}

// P1 geometry seen from the owner (local vertex 0 after rotation): scaled gradients
// c_t = det * grad(phi_t) (cofactor vectors) from the three edge vectors. Ae[0][t] = c_0.c_t/(6|det|).
struct P1Geom
{
  Vec3 c0, c1, c2, c3;
  double det;
};
__device__ __forceinline__ P1Geom p1_geometry(Vec3 e1, Vec3 e2, Vec3 e3)
{
  P1Geom G;
  G.c1 = cross(e2, e3);
  G.c2 = cross(e3, e1);
  G.c3 = cross(e1, e2);
  G.det = dot(e1, G.c1);
  G.c0 = {-(G.c1.x + G.c2.x + G.c3.x), -(G.c1.y + G.c2.y + G.c3.y), -(G.c1.z + G.c2.z + G.c3.z)};
  return G;
}

// 1/d for a normal, finite d: MUFU seed + one cubic and one quadratic Newton step. No slow path
// (an element volume of a valid mesh is never denormal), so callers stay branch-free.
__device__ __forceinline__ double rcp_nr(double d)
{
#ifdef PTB_HOST_EMU // tests/emu runs this directory's kernel sources on the host (test harness only)
  return 1.0 / d;
#else
  double x;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(d));
  double e = fma(-d, x, 1.0);
  e = fma(e, e, e);
  x = fma(x, e, x);
  e = fma(-d, x, 1.0);
  return fma(x, e, x);
#endif
}

// 1/d for a normal, finite d: MUFU seed (~2^-21) + one cubic Newton step (-> 2^-63 before rounding).
// The second step of rcp_nr only tightens the last ulp; the row's blocks are sums of 4-8
// such terms against a 1e-12 bound.
__device__ __forceinline__ double rcp_nr1(double d)
{
#ifdef PTB_HOST_EMU
  return 1.0 / d;
#else
  double x;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(d));
  double e = fma(-d, x, 1.0);
  e = fma(e, e, e);
  return fma(x, e, x);
#endif
}

// Evict-first store for streams that are written once and not read by the kernel.
__device__ __forceinline__ void store_stream(double* p, double v)
{
#ifdef PTB_HOST_EMU
  *p = v;
#else
  __stcs(p, v);
#endif
}

__device__ __forceinline__ void prefetch_l2(const void* p)
{
#ifdef PTB_HOST_EMU
  (void)p;
#else
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#endif
}

} // namespace
} // namespace ptb
