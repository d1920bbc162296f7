// Integer-side maps shared by the host stand-in and the CUDA C-ABI library.
//
// Everything here is pure integer work on the host (OpenMP): the dof -> (cell, local index)
// adjacency, the CSR sparsity pattern (what DOLFINx's fem::create_sparsity_pattern does for the
// reference at poisson_problem.cpp:122-123 / elasticity_problem.cpp:196-197), and the compressed
// cell -> CSR-slot map (rowptr[dof_i] + in-row offset) that replaces the per-insertion column
// search of MatSetValuesLocal (poisson_problem.cpp:129).
//
// All outputs are deterministic functions of (dofmap, n_rows): tests memcmp them against the
// numpy restatement in oracle/intmaps_ref.py.
#pragma once
#include <cstdint>
#include <vector>

namespace ptb
{

/// dof -> list of "pairs" p = cell*nd + local_index, ascending in p (hence ascending in cell).
/// Only rows < n_rows (the owned block dofs) are listed.
struct RowAdjacency
{
  std::vector<std::int64_t> ptr;    // [n_rows + 1]
  std::vector<std::uint32_t> pairs; // [ptr[n_rows]]
};

/// @param dofmap  [ncells * nd] block dof indices (local numbering: owned first, then ghosts)
void build_row_adjacency(const std::int32_t* dofmap, std::int64_t ncells, int nd,
                         std::int32_t n_rows, RowAdjacency& adj);

/// CSR pattern of the owned rows: row r holds the sorted union of dofmap[c][:] over the cells c
/// adjacent to r. Column indices are local block indices, ascending.
void build_pattern(const std::int32_t* dofmap, int nd, std::int32_t n_rows,
                   const RowAdjacency& adj, std::vector<std::int64_t>& rowptr,
                   std::vector<std::int32_t>& cols);

/// Compressed slot map, aligned with adj.pairs: for pair k (row r, cell c, local index li) and
/// local column lj, off[k*nd + lj] = position of dofmap[c][lj] inside row r of the pattern.
/// The CSR slot of Ae[li][lj] is rowptr[r] + off. Returns -1 if a column is missing from the
/// pattern or an offset exceeds 65535, else the maximum offset found.
std::int64_t build_slot_offsets(const std::int32_t* dofmap, int nd, std::int32_t n_rows,
                                const RowAdjacency& adj, const std::int64_t* rowptr,
                                const std::int32_t* cols, std::vector<std::uint16_t>& off);

/// The uncompressed map the north star names: slot[c*nd*nd + i*nd + j] = CSR position of
/// (dofmap[c][i], dofmap[c][j]) or -1 when row dofmap[c][i] is not owned (>= n_rows).
void build_cell_slot_map(const std::int32_t* dofmap, std::int64_t ncells, int nd,
                         std::int32_t n_rows, const std::int64_t* rowptr,
                         const std::int32_t* cols, std::int64_t* slot);

} // namespace ptb
