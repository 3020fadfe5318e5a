// block_ordering.h -- host-side integer work for the block-sparse (supernodal) Cholesky of a reduced camera system
// that is too large to be dense: fill-reducing ordering, elimination tree, supernodes. No CUDA in here; the functions
// are also exported through the C ABI as pure host helpers (spp_block_ordering / spp_supernodal_stats).
//
// Reference functions replaced: CMatrixOrdering::p_BlockOrdering (src/slam/OrderingMagic.cpp:
// 701-1033, which hands the block graph of A + A^T to SuiteSparse's amd_l2), CUberBlockMatrix::Build_EliminationTree
// (src/slam/BlockMatrix.cpp:9403) and the per-column ereach of CholeskyOf_FBS (include/slam/BlockMatrixFBS.inl:
// 2341-2513). The ordering (amd_exact.cpp) is the approximate-minimum-degree ordering of Amestoy, Davis and Duff with
// the tie-breaking conventions of the SuiteSparse release the reference vendors, so the standalone paths eliminate in
// the reference's order, permutation for permutation (tests/test_ordering_cpu.py, against the reference's own output).
#pragma once

#include <stdint.h>
#include <stddef.h>
#include <vector>

namespace spp {

// Upper block structure in CSC (diagonal present or not; a full symmetric pattern is fine too). order[new position] =
// original block column: the reference's permutation bit for bit (see amd_exact.cpp).
void amd_exact_ordering(size_t n, const uint64_t *col_ptr, const uint64_t *row_idx, std::vector<uint32_t> &order);

struct Supernodes {
	size_t n;                          // block columns
	std::vector<uint32_t> first;       // [n_super + 1] supernode s = permuted block columns first[s] .. first[s + 1]
	std::vector<uint64_t> row_ptr;     // [n_super + 1] structure of s below its own columns: rows[row_ptr[s] .. row_ptr[s + 1]]
	std::vector<uint32_t> rows;        // permuted block rows, ascending
	std::vector<uint32_t> parent;      // [n_super] parent supernode (0xffffffff: root)
	std::vector<uint32_t> level;       // [n_super] height above the leaves (children have smaller levels)
	std::vector<uint32_t> col_super;   // [n] supernode of every permuted block column
	std::vector<uint32_t> col_parent;  // [n] elimination tree (permuted block columns)
	std::vector<uint32_t> col_count;   // [n] blocks in column j of the factor (diagonal included), before amalgamation
	uint64_t nnzb_factor;              // blocks of the factor as stored (amalgamation zeros included)
	uint64_t nnzb_exact;               // blocks of the exact factor
	double flops_blocks;               // sum over columns of count^2 (block operations; x B^3 for flops)
	size_t n_super() const { return first.empty()? 0 : first.size() - 1; }
};

// Symbolic factorisation of P A P^T (A upper CSC in the caller's order, order[new] = old): column structures, maximal
// supernodes, relaxed amalgamation of a supernode into its parent when they are contiguous and the explicit zeros stay
// below relax_zeros of the merged panel (or the merged width is <= relax_small block columns).
void supernodal_symbolic(size_t n, const uint64_t *col_ptr, const uint64_t *row_idx, const std::vector<uint32_t> &order,
	double relax_zeros, size_t relax_small, size_t max_width, Supernodes &out);

// Several ranks: who factors which supernode. The work of a supernode (its panel and the updates it sends: w^3 / 3 +
// w^2 h + w h^2 for w own and h structure columns) is summed over subtrees; supernodes whose subtree weighs more than a
// threshold stay with every rank (owner -1: the top of the tree, an upward-closed set), the subtrees hanging below
// them go to the least loaded rank, heaviest first. The threshold is the one (of a few multiples of total / world) with
// the smallest predicted time = shared work + the heaviest rank's work; a plan is taken when it saves at least
// min_saving of the replicated time. Returns the predicted time as a fraction of the replicated one (1 = replicated,
// owner all -1); work[s] = flops of supernode s. Deterministic: every rank computes the same plan.
double plan_subtree_owners(const Supernodes &sn, int world, double min_saving, std::vector<int> &owner, std::vector<double> &work);

} // namespace spp
