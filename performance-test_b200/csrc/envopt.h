// Environment switches of the A/B and opt-in paths.
#pragma once
#include <cstdlib>

namespace ptb
{

inline int env_int(const char* name, int dflt)
{
  const char* e = std::getenv(name);
  return e && *e ? std::atoi(e) : dflt;
}
inline bool env_flag(const char* name, bool dflt)
{
  const char* e = std::getenv(name);
  return e && *e ? e[0] == '1' : dflt;
}

} // namespace ptb
