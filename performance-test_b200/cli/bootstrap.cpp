#include "bootstrap.h"
#include <arpa/inet.h>
#include <chrono>
#include <cstdint>
#include <cstring>
#include <netinet/in.h>
#include <netinet/tcp.h>
#include <stdexcept>
#include <sys/socket.h>
#include <thread>
#include <unistd.h>

namespace ptb::cli
{
namespace
{
void send_all(int fd, const void* p, std::size_t n)
{
  const char* c = static_cast<const char*>(p);
  while (n > 0)
  {
    const ssize_t k = ::send(fd, c, n, 0);
    if (k <= 0)
      throw std::runtime_error("bootstrap: send failed");
    c += k, n -= k;
  }
}
void recv_all(int fd, void* p, std::size_t n)
{
  char* c = static_cast<char*>(p);
  while (n > 0)
  {
    const ssize_t k = ::recv(fd, c, n, 0);
    if (k <= 0)
      throw std::runtime_error("bootstrap: recv failed");
    c += k, n -= k;
  }
}
} // namespace

Bootstrap::Bootstrap(int rank, int world, const std::string& addr, int port)
    : _rank(rank), _world(world)
{
  if (world == 1)
    return;
  sockaddr_in sa{};
  sa.sin_family = AF_INET;
  sa.sin_port = htons(static_cast<std::uint16_t>(port));
  if (inet_pton(AF_INET, addr.c_str(), &sa.sin_addr) != 1)
    throw std::runtime_error("bootstrap: MASTER_ADDR must be an IPv4 address");
  const int one = 1;
  if (rank == 0)
  {
    _listen = ::socket(AF_INET, SOCK_STREAM, 0);
    setsockopt(_listen, SOL_SOCKET, SO_REUSEADDR, &one, sizeof(one));
    if (::bind(_listen, reinterpret_cast<sockaddr*>(&sa), sizeof(sa)) != 0
        || ::listen(_listen, world) != 0)
      throw std::runtime_error("bootstrap: cannot listen on MASTER_PORT");
    _peers.assign(world, -1);
    for (int i = 1; i < world; ++i)
    {
      const int fd = ::accept(_listen, nullptr, nullptr);
      if (fd < 0)
        throw std::runtime_error("bootstrap: accept failed");
      setsockopt(fd, IPPROTO_TCP, TCP_NODELAY, &one, sizeof(one));
      std::int32_t r = -1;
      recv_all(fd, &r, sizeof(r));
      if (r <= 0 || r >= world || _peers[r] != -1)
        throw std::runtime_error("bootstrap: bad peer rank");
      _peers[r] = fd;
    }
  }
  else
  {
    int fd = -1;
    for (int attempt = 0; attempt < 600; ++attempt)
    {
      fd = ::socket(AF_INET, SOCK_STREAM, 0);
      if (::connect(fd, reinterpret_cast<sockaddr*>(&sa), sizeof(sa)) == 0)
        break;
      ::close(fd);
      fd = -1;
      std::this_thread::sleep_for(std::chrono::milliseconds(100));
    }
    if (fd < 0)
      throw std::runtime_error("bootstrap: cannot reach rank 0");
    setsockopt(fd, IPPROTO_TCP, TCP_NODELAY, &one, sizeof(one));
    const std::int32_t r = rank;
    send_all(fd, &r, sizeof(r));
    _peers = {fd};
  }
}

Bootstrap::~Bootstrap()
{
  for (int fd : _peers)
    if (fd >= 0)
      ::close(fd);
  if (_listen >= 0)
    ::close(_listen);
}

std::vector<std::vector<char>> Bootstrap::allgather(const void* data, std::size_t n)
{
  std::vector<std::vector<char>> all(_world);
  all[_rank].assign(static_cast<const char*>(data), static_cast<const char*>(data) + n);
  if (_world == 1)
    return all;
  if (_rank == 0)
  {
    for (int r = 1; r < _world; ++r)
    {
      std::uint64_t len = 0;
      recv_all(_peers[r], &len, sizeof(len));
      all[r].resize(len);
      recv_all(_peers[r], all[r].data(), len);
    }
    for (int r = 1; r < _world; ++r)
      for (int s = 0; s < _world; ++s)
      {
        const std::uint64_t len = all[s].size();
        send_all(_peers[r], &len, sizeof(len));
        send_all(_peers[r], all[s].data(), len);
      }
  }
  else
  {
    const std::uint64_t len = n;
    send_all(_peers[0], &len, sizeof(len));
    send_all(_peers[0], data, n);
    for (int s = 0; s < _world; ++s)
    {
      std::uint64_t l = 0;
      recv_all(_peers[0], &l, sizeof(l));
      all[s].resize(l);
      recv_all(_peers[0], all[s].data(), l);
    }
  }
  return all;
}

void Bootstrap::barrier()
{
  const char c = 0;
  allgather(&c, 1);
}
} // namespace ptb::cli
