#include "layout.h"
#include <algorithm>
#include <stdexcept>

namespace ptb
{

void build_sell_layout(std::int32_t n_rows, int nd, const std::int64_t* rowptr,
                       const std::int32_t* cols, const RowAdjacency& adj,
                       const std::vector<std::uint16_t>& so, std::int64_t max_so, SellLayout& L)
{
  const std::int32_t S = (n_rows + 31) / 32;
  L.n_slices = S;
  L.so_bits = max_so < 256 ? 8 : 16;
  const int per_word = 32 / L.so_bits;
  L.so_words = (nd + per_word - 1) / per_word;
  L.mat_off.assign(S + 1, 0);
  L.adj_off.assign(S + 1, 0);
  int max_w = 0, max_wa = 0;
#pragma omp parallel for schedule(static) reduction(max : max_w, max_wa)
  for (std::int32_t s = 0; s < S; ++s)
  {
    std::int64_t w = 0, wa = 0;
    for (std::int32_t r = 32 * s; r < std::min(n_rows, 32 * s + 32); ++r)
    {
      w = std::max(w, rowptr[r + 1] - rowptr[r]);
      wa = std::max(wa, adj.ptr[r + 1] - adj.ptr[r]);
    }
    L.mat_off[s + 1] = 32 * w;
    L.adj_off[s + 1] = 32 * wa;
    max_w = std::max<int>(max_w, w);
    max_wa = std::max<int>(max_wa, wa);
  }
  L.max_w = max_w, L.max_wa = max_wa;
  for (std::int32_t s = 0; s < S; ++s)
  {
    L.mat_off[s + 1] += L.mat_off[s];
    L.adj_off[s + 1] += L.adj_off[s];
  }
  L.cols.resize(static_cast<std::size_t>(L.mat_off[S]));
  // P1 with short rows: the rotated one-word form replaces adj + adjso (the kernels read nothing else)
  const bool rot = nd == 4 && max_so < 255;
  if (rot)
    L.adjrot.resize(static_cast<std::size_t>(L.adj_off[S]));
  else
  {
    L.adj.resize(static_cast<std::size_t>(L.adj_off[S]));
    L.adjso.resize(static_cast<std::size_t>(L.adj_off[S]) * L.so_words);
  }
  const std::uint32_t* pairs = adj.pairs.data();
#pragma omp parallel for schedule(static)
  for (std::int32_t s = 0; s < S; ++s)
  {
    const std::int64_t mo = L.mat_off[s], w = (L.mat_off[s + 1] - mo) / 32;
    const std::int64_t ao = L.adj_off[s], wa = (L.adj_off[s + 1] - ao) / 32;
    for (int lane = 0; lane < 32; ++lane)
    {
      const std::int32_t r = 32 * s + lane;
      const bool live = r < n_rows;
      const std::int64_t len = live ? rowptr[r + 1] - rowptr[r] : 0;
      const std::int32_t pad = len > 0 ? cols[rowptr[r]] : 0;
      for (std::int64_t k = 0; k < w; ++k)
        L.cols[mo + k * 32 + lane] = k < len ? cols[rowptr[r] + k] : pad;
      const std::int64_t alen = live ? adj.ptr[r + 1] - adj.ptr[r] : 0;
      for (std::int64_t k = 0; k < wa; ++k)
      {
        const bool on = k < alen;
        if (rot)
        {
          std::uint32_t word = ADJ_INVALID;
          if (on)
          {
            const std::int64_t q = adj.ptr[r] + k;
            const int li = pairs[q] & 3;
            word = 0;
            for (int t = 0; t < 4; ++t)
              word |= static_cast<std::uint32_t>(so[q * 4 + ((li + t) & 3)]) << (8 * t);
          }
          L.adjrot[ao + k * 32 + lane] = word;
          continue;
        }
        L.adj[ao + k * 32 + lane] = on ? pairs[adj.ptr[r] + k] : ADJ_INVALID;
        for (int wd = 0; wd < L.so_words; ++wd)
        {
          std::uint32_t word = 0;
          if (on)
            for (int t = 0; t < per_word; ++t)
            {
              const int j = wd * per_word + t;
              if (j < nd)
                word |= static_cast<std::uint32_t>(so[(adj.ptr[r] + k) * nd + j])
                        << (t * L.so_bits);
            }
          L.adjso[(ao + k * 32) * L.so_words + wd * 32 + lane] = word;
        }
      }
    }
  }
}

WalkStats build_walk(std::int32_t n_rows, const RowAdjacency& adj,
                     const std::vector<std::uint16_t>& so, SellLayout& L)
{
  if (L.adjrot.empty())
    throw std::runtime_error("build_walk: needs the rotated P1 slot words");
  L.walk.assign(L.adjrot.size(), ADJ_INVALID);
  const std::uint32_t* pairs = adj.pairs.data();
  std::int64_t steps = 0, loads = 0;
#pragma omp parallel reduction(+ : steps, loads)
  {
    std::vector<std::uint8_t> o;          // [c][3] offsets of the non-owner vertices, rotation order
    std::vector<std::uint64_t> m;         // [c][4] the same as a 256-bit set
    std::vector<char> visited;
    // Rows with the same star (same offsets in the same cell order) have the same walk: on a
    // structured mesh almost every row repeats its predecessor, so the last result is kept.
    std::vector<std::uint8_t> last_o;
    std::vector<std::uint32_t> last_words;
    std::int64_t last_loads = 0;
#pragma omp for schedule(static)
    for (std::int32_t r = 0; r < n_rows; ++r)
    {
      const std::int64_t q0 = adj.ptr[r];
      const int c = static_cast<int>(adj.ptr[r + 1] - q0);
      if (c == 0)
        continue;
      o.resize(static_cast<std::size_t>(c) * 3);
      for (int j = 0; j < c; ++j)
      {
        const int li = pairs[q0 + j] & 3;
        for (int t = 1; t < 4; ++t)
          o[j * 3 + t - 1] = static_cast<std::uint8_t>(so[(q0 + j) * 4 + ((li + t) & 3)]);
      }
      if (o == last_o)
      {
        const std::int64_t b0 = L.adj_off[r >> 5] + (r & 31);
        for (int k = 0; k < c; ++k)
          L.walk[b0 + static_cast<std::int64_t>(k) * 32] = last_words[k];
        steps += c, loads += last_loads;
        continue;
      }
      m.assign(static_cast<std::size_t>(c) * 4, 0);
      visited.assign(c, 0);
      for (int j = 0; j < c * 3; ++j)
        m[(j / 3) * 4 + (o[j] >> 6)] |= std::uint64_t(1) << (o[j] & 63);
      const std::int64_t loads_before = loads;
      auto shared = [&](int a, int b) {
        return __builtin_popcountll(m[a * 4] & m[b * 4]) + __builtin_popcountll(m[a * 4 + 1] & m[b * 4 + 1])
               + __builtin_popcountll(m[a * 4 + 2] & m[b * 4 + 2])
               + __builtin_popcountll(m[a * 4 + 3] & m[b * 4 + 3]);
      };
      const std::int64_t base = L.adj_off[r >> 5] + (r & 31);
      int pos[3] = {o[0], o[1], o[2]};
      int cur = 0;
      visited[0] = 1;
      L.walk[base] = pos[0] | (pos[1] << 8) | (pos[2] << 16) | (7u << 24);
      steps += c, loads += 3;
      for (int step = 1; step < c; ++step)
      {
        int best = -1, best_sh = -1;
        for (int j = 0; j < c; ++j)
        {
          if (visited[j])
            continue;
          const int sh = shared(cur, j);
          if (sh > best_sh)
            best = j, best_sh = sh;
          if (sh >= 2)
            break; // a face neighbour: nothing later in the list can be preferred
        }
        // vertices of the new cell that are already held keep their position
        bool held[3] = {false, false, false}, old[3] = {false, false, false};
        for (int t = 0; t < 3; ++t)
          for (int p = 0; p < 3; ++p)
            if (!held[p] && !old[t] && pos[p] == o[best * 3 + t])
              held[p] = true, old[t] = true;
        unsigned mask = 0;
        int p = 0;
        for (int t = 0; t < 3; ++t)
        {
          if (old[t])
            continue;
          while (held[p])
            ++p;
          pos[p] = o[best * 3 + t];
          held[p] = true;
          mask |= 1u << p;
          ++loads;
        }
        L.walk[base + static_cast<std::int64_t>(step) * 32]
            = pos[0] | (pos[1] << 8) | (pos[2] << 16) | (mask << 24);
        visited[best] = 1;
        cur = best;
      }
      last_o = o;
      last_words.resize(c);
      for (int k = 0; k < c; ++k)
        last_words[k] = L.walk[base + static_cast<std::int64_t>(k) * 32];
      last_loads = loads - loads_before;
    }
  }
  return {steps, loads};
}

namespace
{
// Chains of one row (SellLayout::ring): out[k] = bytes of column k. o = [c][3] in-row offsets of the
// non-owner vertices of the row's cells (ascending cell order, rotation order inside a cell).
void ring_chains(const std::vector<std::uint8_t>& o, int c, int len,
                 std::vector<std::vector<std::uint8_t>>& out)
{
  out.resize(len);
  std::vector<int> va, vb; // the two other vertices of every ring cell of the column
  std::vector<char> used;
  for (int k = 0; k < len; ++k)
  {
    std::vector<std::uint8_t>& bytes = out[k];
    bytes.clear();
    va.clear(), vb.clear();
    for (int j = 0; j < c; ++j)
      for (int t = 0; t < 3; ++t)
        if (o[j * 3 + t] == k)
        {
          va.push_back(o[j * 3 + (t + 1) % 3]);
          vb.push_back(o[j * 3 + (t + 2) % 3]);
        }
    const int n = static_cast<int>(va.size());
    used.assign(n, 0);
    auto degree = [&](int v) {
      int d = 0;
      for (int j = 0; j < n; ++j)
        d += !used[j] && (va[j] == v || vb[j] == v);
      return d;
    };
    auto next_with = [&](int v, int skip) {
      for (int j = 0; j < n; ++j)
        if (!used[j] && j != skip && (va[j] == v || vb[j] == v))
          return j;
      return -1;
    };
    int left = n;
    while (left > 0)
    {
      int start = -1, v0 = 0, v1 = 0;
      for (int j = 0; j < n && start < 0; ++j)
      {
        if (used[j])
          continue;
        if (degree(va[j]) == 1)
          start = j, v0 = va[j], v1 = vb[j];
        else if (degree(vb[j]) == 1)
          start = j, v0 = vb[j], v1 = va[j];
      }
      if (start < 0)
      {
        for (int j = 0; j < n && start < 0; ++j)
          if (!used[j])
            start = j;
        const int ja = next_with(va[start], start), jb = next_with(vb[start], start);
        if (ja <= jb)
          v1 = va[start], v0 = vb[start];
        else
          v1 = vb[start], v0 = va[start];
      }
      bytes.push_back(static_cast<std::uint8_t>(v0 | 0x80));
      bytes.push_back(static_cast<std::uint8_t>(v1));
      used[start] = 1, --left;
      int cur = v1;
      for (int j = next_with(cur, -1); j >= 0; j = next_with(cur, -1))
      {
        cur = va[j] == cur ? vb[j] : va[j];
        bytes.push_back(static_cast<std::uint8_t>(cur));
        used[j] = 1, --left;
      }
    }
  }
}
} // namespace

std::int64_t build_rings(std::int32_t n_rows, const std::int64_t* rowptr, const RowAdjacency& adj,
                         const std::vector<std::uint16_t>& so, SellLayout& L)
{
  L.ring.clear(), L.ring_off.clear(), L.ring_ns.clear();
  if (L.adjrot.empty() || L.max_w > 127)
    return 0;
  const std::int32_t S = L.n_slices;
  const std::uint32_t* pairs = adj.pairs.data();
  L.ring_ns.assign(static_cast<std::size_t>(L.mat_off[S] / 32), 0);
  L.ring_off.assign(static_cast<std::size_t>(S) + 1, 0);
  std::int64_t total = 0;
  // pass 0: chain lengths -> ring_ns, ring_off; pass 1: the bytes. Rows with the same star (same
  // offsets in the same cell order) have the same chains: the last result is kept per thread.
  for (int pass = 0; pass < 2; ++pass)
  {
    if (pass == 1)
    {
      for (std::int32_t s = 0; s < S; ++s)
      {
        std::int64_t words = 0;
        const std::int64_t k0 = L.mat_off[s] / 32, w = (L.mat_off[s + 1] - L.mat_off[s]) / 32;
        for (std::int64_t k = 0; k < w; ++k)
          words += (L.ring_ns[k0 + k] + 3) / 4;
        L.ring_off[s + 1] = L.ring_off[s] + 32 * words;
      }
      L.ring.assign(static_cast<std::size_t>(L.ring_off[S]), 0x80808080u);
    }
#pragma omp parallel reduction(+ : total)
    {
      std::vector<std::uint8_t> o, last_o;
      std::vector<std::vector<std::uint8_t>> chains;
      int last_len = -1;
#pragma omp for schedule(static)
      for (std::int32_t s = 0; s < S; ++s)
      {
        const std::int64_t k0 = L.mat_off[s] / 32, w = (L.mat_off[s + 1] - L.mat_off[s]) / 32;
        for (std::int32_t r = 32 * s; r < std::min(n_rows, 32 * s + 32); ++r)
        {
          const std::int64_t q0 = adj.ptr[r];
          const int c = static_cast<int>(adj.ptr[r + 1] - q0);
          const int len = static_cast<int>(rowptr[r + 1] - rowptr[r]);
          o.resize(static_cast<std::size_t>(c) * 3);
          for (int j = 0; j < c; ++j)
          {
            const int li = pairs[q0 + j] & 3;
            for (int t = 1; t < 4; ++t)
              o[j * 3 + t - 1] = static_cast<std::uint8_t>(so[(q0 + j) * 4 + ((li + t) & 3)]);
          }
          if (o != last_o || len != last_len)
          {
            ring_chains(o, c, len, chains);
            last_o = o, last_len = len;
          }
          if (pass == 0)
          {
            for (int k = 0; k < len; ++k)
            {
              if (chains[k].size() > 255)
                throw std::runtime_error("build_rings: more than 255 chain bytes around one edge");
              L.ring_ns[k0 + k] = std::max<std::uint8_t>(L.ring_ns[k0 + k], static_cast<std::uint8_t>(chains[k].size()));
              total += static_cast<std::int64_t>(chains[k].size());
            }
            continue;
          }
          std::int64_t base = L.ring_off[s] + (r & 31);
          for (std::int64_t k = 0; k < w; ++k)
          {
            const int nw = (L.ring_ns[k0 + k] + 3) / 4;
            if (k < len)
              for (std::size_t t = 0; t < chains[k].size(); ++t)
              {
                std::uint32_t& word = L.ring[base + static_cast<std::int64_t>(t / 4) * 32];
                word = (word & ~(0xFFu << (8 * (t % 4)))) | (static_cast<std::uint32_t>(chains[k][t]) << (8 * (t % 4)));
              }
            base += static_cast<std::int64_t>(nw) * 32;
          }
        }
      }
    }
  }
  return total;
}

void build_walk_single(std::int32_t n_rows, const RowAdjacency& adj, SellLayout& L)
{
  if (L.walk.empty() && L.adj_off[L.n_slices] > 0)
    throw std::runtime_error("build_walk_single: needs the star walk");
  const std::int32_t S = L.n_slices;
  auto steps_of = [&](std::int32_t r) {
    const std::int64_t c = adj.ptr[r + 1] - adj.ptr[r];
    const std::int64_t base = L.adj_off[r >> 5] + (r & 31);
    std::int64_t n = c > 0 ? 1 : 0;
    for (std::int64_t k = 1; k < c; ++k)
      n += std::max(1, __builtin_popcount((L.walk[base + k * 32] >> 24) & 7u));
    return n;
  };
  L.walk1_off.assign(S + 1, 0);
  int max_w1 = 0;
#pragma omp parallel for schedule(static) reduction(max : max_w1)
  for (std::int32_t s = 0; s < S; ++s)
  {
    std::int64_t w = 0;
    for (std::int32_t r = 32 * s; r < std::min(n_rows, 32 * s + 32); ++r)
      w = std::max(w, steps_of(r));
    L.walk1_off[s + 1] = 32 * w;
    max_w1 = std::max<int>(max_w1, w);
  }
  L.max_w1 = max_w1;
  for (std::int32_t s = 0; s < S; ++s)
    L.walk1_off[s + 1] += L.walk1_off[s];
  L.walk1.assign(static_cast<std::size_t>(L.walk1_off[S]), ADJ_INVALID);
#pragma omp parallel for schedule(static)
  for (std::int32_t r = 0; r < n_rows; ++r)
  {
    const std::int64_t c = adj.ptr[r + 1] - adj.ptr[r];
    if (c == 0)
      continue;
    const std::int64_t src = L.adj_off[r >> 5] + (r & 31);
    std::int64_t dst = L.walk1_off[r >> 5] + (r & 31);
    std::uint32_t prev = L.walk[src];
    L.walk1[dst] = prev;
    dst += 32;
    for (std::int64_t k = 1; k < c; ++k)
    {
      const std::uint32_t word = L.walk[src + k * 32];
      const unsigned mask = (word >> 24) & 7u;
      if (mask == 0)
      {
        L.walk1[dst] = (3u << 16) | (1u << 18); // same vertices again: nothing to load
        dst += 32;
      }
      unsigned left = mask;
      for (unsigned p = 0; p < 3; ++p)
        if (mask & (1u << p))
        {
          left &= ~(1u << p);
          const std::uint32_t nw = (word >> (8 * p)) & 0xFFu, old = (prev >> (8 * p)) & 0xFFu;
          L.walk1[dst] = nw | (old << 8) | (p << 16) | (left == 0 ? 1u << 18 : 0u);
          dst += 32;
        }
      prev = word;
    }
  }
}

void compress_columns(std::int32_t n_rows, std::int64_t n_cols, const std::int64_t* rowptr,
                      SellLayout& L)
{
  const std::int32_t S = L.n_slices;
  L.cdelta.assign(static_cast<std::size_t>(L.mat_off[S] / 32), CDELTA_EXPLICIT);
  L.xoff.assign(S + 1, 0);
#pragma omp parallel for schedule(static)
  for (std::int32_t s = 0; s < S; ++s)
  {
    const std::int64_t mo = L.mat_off[s], w = (L.mat_off[s + 1] - mo) / 32;
    const std::int32_t r0 = 32 * s;
    const bool full = r0 + 32 <= n_rows; // a partial last slice stays explicit (r + d may overrun)
    std::int64_t nx = 0;
    for (std::int64_t k = 0; k < w; ++k)
    {
      bool uniform = full;
      std::int64_t d = 0;
      bool have = false;
      for (int lane = 0; lane < 32 && uniform; ++lane)
      {
        const std::int32_t r = r0 + lane;
        if (k >= rowptr[r + 1] - rowptr[r])
          continue; // padding: free to choose
        const std::int64_t dl = static_cast<std::int64_t>(L.cols[mo + k * 32 + lane]) - r;
        if (!have)
          d = dl, have = true;
        else if (dl != d)
          uniform = false;
      }
      if (uniform && !have)
        d = 0; // all padding: point at the row itself
      if (uniform)
        for (int lane = 0; lane < 32; ++lane)
        {
          const std::int64_t cidx = static_cast<std::int64_t>(r0) + lane + d;
          if (cidx < 0 || cidx >= n_cols)
            uniform = false;
        }
      if (uniform)
        L.cdelta[mo / 32 + k] = static_cast<std::int32_t>(d);
      else
        ++nx;
    }
    L.xoff[s + 1] = nx * 32;
  }
  for (std::int32_t s = 0; s < S; ++s)
    L.xoff[s + 1] += L.xoff[s];
  L.colsx.resize(static_cast<std::size_t>(L.xoff[S]));
#pragma omp parallel for schedule(static)
  for (std::int32_t s = 0; s < S; ++s)
  {
    const std::int64_t mo = L.mat_off[s], w = (L.mat_off[s + 1] - mo) / 32;
    std::int64_t j = 0;
    for (std::int64_t k = 0; k < w; ++k)
      if (L.cdelta[mo / 32 + k] == CDELTA_EXPLICIT)
      {
        for (int lane = 0; lane < 32; ++lane)
          L.colsx[L.xoff[s] + j * 32 + lane] = L.cols[mo + k * 32 + lane];
        ++j;
      }
  }
}

void build_width_bins(const SellLayout& L, std::vector<std::int32_t>& list,
                      std::vector<std::int32_t>& off, std::vector<int>& width)
{
  std::vector<int> bounds;
  for (int b : {32, 64, 96, 128, 192})
    if (b < L.max_w)
      bounds.push_back(b);
  for (int b = 256; b < L.max_w; b += 128)
    bounds.push_back(b);
  bounds.push_back(std::max(1, L.max_w));
  width = bounds;
  const std::int32_t S = L.n_slices;
  std::vector<std::vector<std::int32_t>> bins(bounds.size());
  for (std::int32_t s = 0; s < S; ++s)
  {
    const int w = static_cast<int>((L.mat_off[s + 1] - L.mat_off[s]) / 32);
    const std::size_t b = std::lower_bound(bounds.begin(), bounds.end(), w) - bounds.begin();
    bins[std::min(b, bounds.size() - 1)].push_back(s);
  }
  list.clear();
  off.assign(1, 0);
  for (const auto& v : bins)
  {
    list.insert(list.end(), v.begin(), v.end());
    off.push_back(static_cast<std::int32_t>(list.size()));
  }
}

void build_slice_order(const SellLayout& L, std::int32_t n_rows, int group, bool cluster,
                       std::vector<std::int32_t>& order, std::int32_t& n_interior)
{
  const std::int32_t S = L.n_slices;
  std::vector<char> ghost(S, 0), seen(S, 0);
#pragma omp parallel for schedule(static)
  for (std::int32_t s = 0; s < S; ++s)
    for (std::int64_t q = L.mat_off[s]; q < L.mat_off[s + 1]; ++q)
      if (L.cols[q] >= n_rows)
      {
        ghost[s] = 1;
        break;
      }
  order.clear();
  order.reserve(S);
  std::vector<std::int32_t> grp, nb;
  for (int pass = 0; pass < 2; ++pass)
  {
    for (std::int32_t seed = 0; seed < S; ++seed)
    {
      if (seen[seed] || ghost[seed] != pass)
        continue;
      grp.assign(1, seed);
      seen[seed] = 1;
      for (std::size_t head = 0; cluster && head < grp.size() && grp.size() < static_cast<std::size_t>(group); ++head)
      {
        // neighbours of grp[head]: the slices its columns point into (first and last lane suffice
        // to see every stencil direction; all lanes are scanned to stay general)
        const std::int32_t s = grp[head];
        nb.clear();
        for (std::int64_t q = L.mat_off[s]; q < L.mat_off[s + 1]; ++q)
        {
          const std::int32_t c = L.cols[q];
          if (c < n_rows)
            nb.push_back(c >> 5);
        }
        std::sort(nb.begin(), nb.end());
        nb.erase(std::unique(nb.begin(), nb.end()), nb.end());
        for (std::int32_t t : nb)
          if (!seen[t] && ghost[t] == pass && grp.size() < static_cast<std::size_t>(group))
          {
            seen[t] = 1;
            grp.push_back(t);
          }
      }
      order.insert(order.end(), grp.begin(), grp.end());
    }
    if (pass == 0)
      n_interior = static_cast<std::int32_t>(order.size());
  }
}

namespace
{
const int tet_edges[6][2] = {{2, 3}, {1, 3}, {1, 2}, {0, 3}, {0, 2}, {0, 1}};
}

void build_facet_rows(std::int64_t n_facets, const std::int32_t* cells,
                      const std::int32_t* local_facets, const std::int32_t* dofmap, int nd,
                      int order, std::int32_t n_rows, std::vector<std::int32_t>& row_ids,
                      std::vector<std::int32_t>& row_ptr, std::vector<std::int32_t>& ent)
{
  // local dofs on local facet lf (Basix layout: vertices, edges, faces)
  std::vector<int> on[4];
  const int ne = order - 1, nf = (order - 1) * (order - 2) / 2;
  for (int lf = 0; lf < 4; ++lf)
  {
    for (int v = 0; v < 4; ++v)
      if (v != lf)
        on[lf].push_back(v);
    for (int e = 0; e < 6; ++e)
      if (tet_edges[e][0] != lf && tet_edges[e][1] != lf)
        for (int s = 0; s < ne; ++s)
          on[lf].push_back(4 + e * ne + s);
    for (int s = 0; s < nf; ++s)
      on[lf].push_back(4 + 6 * ne + lf * nf + s);
  }
  // (row, cell, local_facet*nd + li) for every facet dof on an owned row, in facet order; a stable
  // sort by row then gives the rows ascending with their entries ascending in k
  struct Ent
  {
    std::int32_t row, cell, code;
  };
  std::vector<Ent> all;
  all.reserve(static_cast<std::size_t>(n_facets) * on[0].size());
  for (std::int64_t k = 0; k < n_facets; ++k)
  {
    const std::int32_t c = cells[k], lf = local_facets[k];
    if (lf < 0 || lf > 3)
      throw std::runtime_error("exterior facet: local facet index out of range");
    for (int li : on[lf])
    {
      const std::int32_t r = dofmap[static_cast<std::int64_t>(c) * nd + li];
      if (r < n_rows)
        all.push_back({r, c, lf * nd + li});
    }
  }
  std::stable_sort(all.begin(), all.end(), [](const Ent& a, const Ent& b) { return a.row < b.row; });
  row_ids.clear(), ent.clear();
  row_ptr.assign(1, 0);
  ent.reserve(all.size() * 2);
  for (std::size_t i = 0; i < all.size(); ++i)
  {
    if (i == 0 || all[i].row != all[i - 1].row)
    {
      if (i > 0)
        row_ptr.push_back(static_cast<std::int32_t>(i));
      row_ids.push_back(all[i].row);
    }
    ent.push_back(all[i].cell), ent.push_back(all[i].code);
  }
  if (!all.empty())
    row_ptr.push_back(static_cast<std::int32_t>(all.size()));
}

void build_facet_rows_gathered(std::int64_t n_facets, const std::int32_t* cells,
                               const std::int32_t* local_facets, const std::int32_t* gathered, int nd,
                               int order, std::int32_t n_rows, std::vector<std::int32_t>& row_ids,
                               std::vector<std::int32_t>& row_ptr, std::vector<std::int32_t>& ent)
{
  // facet k plays the role of "cell k" of the gathered rows; the entries keep the facet order, so
  // putting the real cell index back afterwards gives build_facet_rows' lists exactly
  std::vector<std::int32_t> k(static_cast<std::size_t>(n_facets));
  for (std::int64_t i = 0; i < n_facets; ++i)
    k[i] = static_cast<std::int32_t>(i);
  build_facet_rows(n_facets, k.data(), local_facets, gathered, nd, order, n_rows, row_ids, row_ptr, ent);
  for (std::size_t i = 0; i < ent.size(); i += 2)
    ent[i] = cells[ent[i]];
}

int build_balance_plan(const std::int64_t* mat_off, const std::int32_t* order, std::int32_t n_slices,
                       std::int32_t n_interior, int grid, int npull, std::vector<std::int32_t>& ounit,
                       std::vector<std::int32_t>& begin)
{
  const std::int32_t S = n_slices;
  ounit.assign(static_cast<std::size_t>(S) + 1, 0);
  for (std::int32_t i = 0; i < S; ++i)
    ounit[i + 1] = ounit[i] + static_cast<std::int32_t>((mat_off[order[i] + 1] - mat_off[order[i]]) >> 5);
  begin.clear();
  int longest = 0;
  auto split = [&](std::int32_t a, std::int32_t b, int ctas) {
    // boundaries at the slice edges nearest to the equal-unit cuts
    const std::int64_t lo = ounit[a], len = ounit[b] - ounit[a];
    std::int32_t prev = a;
    for (int t = 0; t <= ctas; ++t)
    {
      const std::int64_t target = lo + len * t / ctas;
      std::int32_t i = static_cast<std::int32_t>(
          std::lower_bound(ounit.begin() + a, ounit.begin() + b + 1, target) - ounit.begin());
      if (i > a && target - ounit[i - 1] < ounit[i] - target)
        --i;
      i = std::max(i, prev);
      if (t == ctas)
        i = b;
      if (t > 0)
        longest = std::max(longest, i - prev);
      begin.push_back(i);
      prev = i;
    }
  };
  if (npull >= 0)
  {
    split(n_interior, S, std::max(npull, 1));
    split(0, n_interior, grid - npull);
  }
  else
    split(0, S, grid);
  return longest;
}

} // namespace ptb
