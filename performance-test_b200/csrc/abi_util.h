// Error convention of the C ABI, shared by abi.cu and abi_debug.cpp: every entry point returns
// 0 / non-zero and nothing throws across the boundary; the text goes to the context, or to a
// thread-local slot for calls without one (ptb_last_error).
#pragma once
#include "ctx.h"
#include <stdexcept>
#include <string>

namespace ptb::abi
{

inline thread_local std::string g_err;

template <typename F>
int guarded(ptb_ctx* c, F&& fn)
{
  try
  {
    fn();
    return 0;
  }
  catch (const std::exception& e)
  {
    (c ? c->err : g_err) = e.what();
    return 1;
  }
}

inline void need(bool ok, const char* msg)
{
  if (!ok)
    throw std::runtime_error(msg);
}

} // namespace ptb::abi
