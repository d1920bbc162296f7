// 1-D bulk copies global -> shared through the TMA unit (cp.async.bulk, SASS UBLKCP) completing on an
// mbarrier: the PTX wrappers shared by the TMA-staged operator variant (cg.cu spmv_sell_tma) and the
// ring-word staging of the elasticity matrix kernel (assemble_ring.cu). Source and destination 16-byte
// aligned, size a multiple of 16 bytes. Device build only.
#pragma once
#include <cstdint>

namespace ptb
{
namespace
{
#ifndef PTB_HOST_EMU
__device__ __forceinline__ std::uint32_t smem_u32(const void* p)
{
  return static_cast<std::uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(std::uint64_t* bar, std::uint32_t count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(std::uint64_t* bar, std::uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, std::uint32_t bytes,
                                            std::uint64_t* bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(std::uint64_t* bar, std::uint32_t parity)
{
  std::uint32_t ok = 0;
  const std::uint32_t a = smem_u32(bar);
  do
  {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                 "selp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok)
                 : "r"(a), "r"(parity)
                 : "memory");
  } while (!ok);
}
#endif // PTB_HOST_EMU
} // namespace
} // namespace ptb
