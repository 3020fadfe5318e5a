// block_ordering.cpp -- see block_ordering.h. Pure host code (integer work, one-time per structure).

#include "block_ordering.h"
#include "../../include/spp_b200.h"
#include <algorithm>
#include <stdexcept>
#include <math.h>

namespace spp {

// ---- elimination tree, postorder, supernodes ----------------------------------------------------------------

// for every permuted column j, the permuted rows i < j of the symmetric pattern (unsorted)
static void permuted_upper(size_t n, const uint64_t *col_ptr, const uint64_t *row_idx, const std::vector<uint32_t> &inv,
	std::vector<uint64_t> &ptr, std::vector<uint32_t> &idx)
{
	ptr.assign(n + 1, 0);
	for(size_t c = 0; c < n; ++ c) {
		for(uint64_t k = col_ptr[c]; k < col_ptr[c + 1]; ++ k) {
			const size_t r = row_idx[k];
			if(r >= n) throw std::runtime_error("block ordering: row index out of range");
			if(r == c) continue;
			++ ptr[std::max(inv[r], inv[c]) + 1];
		}
	}
	for(size_t j = 0; j < n; ++ j) ptr[j + 1] += ptr[j];
	idx.resize(ptr[n]);
	std::vector<uint64_t> fill(ptr.begin(), ptr.end() - 1);
	for(size_t c = 0; c < n; ++ c) {
		for(uint64_t k = col_ptr[c]; k < col_ptr[c + 1]; ++ k) {
			const size_t r = row_idx[k];
			if(r == c) continue;
			const uint32_t a = inv[r], b = inv[c];
			idx[fill[std::max(a, b)] ++] = std::min(a, b);
		}
	}
}

static void elimination_tree(size_t n, const std::vector<uint64_t> &ptr, const std::vector<uint32_t> &idx,
	std::vector<uint32_t> &parent)
{
	const uint32_t none = 0xffffffffu;
	parent.assign(n, none);
	std::vector<uint32_t> ancestor(n, none);
	for(size_t j = 0; j < n; ++ j) {
		for(uint64_t k = ptr[j]; k < ptr[j + 1]; ++ k) {
			uint32_t i = idx[k];
			while(i != none && i < j) { // walk to the root of i's subtree, compressing the path to j
				const uint32_t nxt = ancestor[i];
				ancestor[i] = (uint32_t)j;
				if(nxt == none) parent[i] = (uint32_t)j;
				i = nxt;
			}
		}
	}
}

// ---- symbolic factorisation, supernodes -------------------------------------------------------------------

void supernodal_symbolic(size_t n, const uint64_t *col_ptr, const uint64_t *row_idx, const std::vector<uint32_t> &order,
	double relax_zeros, size_t relax_small, size_t max_width, Supernodes &out)
{
	const uint32_t none = 0xffffffffu;
	if(order.size() != n) throw std::runtime_error("supernodal_symbolic: bad ordering");
	std::vector<uint32_t> inv(n, none);
	for(size_t i = 0; i < n; ++ i) {
		if(order[i] >= n || inv[order[i]] != none) throw std::runtime_error("supernodal_symbolic: the ordering is not a permutation");
		inv[order[i]] = (uint32_t)i;
	}
	// permuted lower pattern by column: rows > j of column j
	std::vector<std::vector<uint32_t> > lcol(n);
	{
		std::vector<uint32_t> cnt(n, 0);
		for(size_t c = 0; c < n; ++ c) {
			bool diag = false;
			for(uint64_t k = col_ptr[c]; k < col_ptr[c + 1]; ++ k) {
				const size_t r = row_idx[k];
				if(r > c || r >= n) throw std::runtime_error("supernodal_symbolic: the matrix must be upper block-triangular");
				if(r == c) { diag = true; continue; }
				++ cnt[std::min(inv[r], inv[c])];
			}
			if(!diag) throw std::runtime_error("supernodal_symbolic: missing diagonal block");
		}
		for(size_t j = 0; j < n; ++ j) lcol[j].reserve(cnt[j] + 1);
		for(size_t j = 0; j < n; ++ j) lcol[j].push_back((uint32_t)j);
		for(size_t c = 0; c < n; ++ c) {
			for(uint64_t k = col_ptr[c]; k < col_ptr[c + 1]; ++ k) {
				const size_t r = row_idx[k];
				if(r == c) continue;
				const uint32_t a = inv[r], b = inv[c];
				lcol[std::min(a, b)].push_back(std::max(a, b));
			}
		}
		for(size_t j = 0; j < n; ++ j) {
			std::sort(lcol[j].begin() + 1, lcol[j].end());
			lcol[j].erase(std::unique(lcol[j].begin(), lcol[j].end()), lcol[j].end());
		}
	}
	// column structures of the factor: struct(j) = pattern(A_j) U (struct(children) \ {child}); the structure of a column
	// that is not the first of its maximal supernode is dropped as soon as its parent has consumed it
	out.n = n;
	out.col_parent.assign(n, none);
	out.col_count.assign(n, 0);
	std::vector<uint8_t> starts(n, 1); // column j starts a maximal supernode
	{
		std::vector<std::vector<uint32_t> > children(n);
		std::vector<uint32_t> tmp;
		for(size_t j = 0; j < n; ++ j) {
			std::vector<uint32_t> &s = lcol[j];
			for(uint32_t c : children[j]) {
				const std::vector<uint32_t> &cs = lcol[c];
				tmp.clear();
				std::set_union(s.begin(), s.end(), cs.begin() + 1, cs.end(), std::back_inserter(tmp));
				s.swap(tmp);
			}
			out.col_count[j] = (uint32_t)s.size();
			if(s.size() > 1) {
				out.col_parent[j] = s[1];
				children[s[1]].push_back((uint32_t)j);
			}
			if(j > 0 && out.col_parent[j - 1] == j && out.col_count[j - 1] == out.col_count[j] + 1)
				starts[j] = 0;
			for(uint32_t c : children[j])
				if(!starts[c] || true) { /* kept: amalgamation below may need any supernode's first column */ }
			std::vector<uint32_t>().swap(children[j]);
		}
	}
	out.nnzb_exact = 0;
	out.flops_blocks = 0;
	for(size_t j = 0; j < n; ++ j) {
		out.nnzb_exact += out.col_count[j];
		out.flops_blocks += (double)out.col_count[j] * out.col_count[j];
	}
	// maximal supernodes, then relaxed amalgamation of contiguous child -> parent chains
	std::vector<uint32_t> sfirst; // first column of every (maximal) supernode
	for(size_t j = 0; j < n; ++ j)
		if(starts[j]) sfirst.push_back((uint32_t)j);
	const size_t ns0 = sfirst.size();
	sfirst.push_back((uint32_t)n);
	std::vector<uint8_t> merge_next(ns0, 0); // supernode s is merged with s + 1
	{
		size_t gw = 0; // width and explicit zeros of the group that ends with supernode s
		double gz = 0;
		for(size_t s = 0; s + 1 < ns0; ++ s) {
			const size_t w = sfirst[s + 1] - sfirst[s];
			if(s == 0 || !merge_next[s - 1]) { gw = 0; gz = 0; }
			gw += w;
			const size_t last = sfirst[s + 1] - 1;           // last column of s
			if(out.col_parent[last] != sfirst[s + 1])
				continue;                                       // the next supernode is not the parent
			const size_t h = out.col_count[last] - 1;          // rows below the group
			const size_t wp = sfirst[s + 2] - sfirst[s + 1];
			const size_t hp = out.col_count[sfirst[s + 2] - 1] - 1;
			if(gw + wp > max_width)
				continue;
			const double zeros = gz + (double)gw * (double)(wp + hp - h);
			const double total = (double)(gw + wp) * (double)(gw + wp + hp);
			if(gw + wp <= relax_small || zeros <= relax_zeros * total) {
				merge_next[s] = 1;
				gz = zeros;
			}
		}
	}
	out.first.clear();
	std::vector<uint32_t> top_first; // first column of the top (last) member of every final supernode
	for(size_t s = 0; s < ns0; ++ s) {
		if(s == 0 || !merge_next[s - 1]) out.first.push_back(sfirst[s]);
		if(!merge_next[s] || s + 1 == ns0) top_first.push_back(sfirst[s]);
	}
	out.first.push_back((uint32_t)n);
	const size_t ns = out.first.size() - 1;
	out.col_super.assign(n, 0);
	for(size_t s = 0; s < ns; ++ s)
		for(size_t j = out.first[s]; j < out.first[s + 1]; ++ j) out.col_super[j] = (uint32_t)s;
	out.row_ptr.assign(ns + 1, 0);
	out.rows.clear();
	out.parent.assign(ns, none);
	out.level.assign(ns, 0);
	out.nnzb_factor = 0;
	for(size_t s = 0; s < ns; ++ s) {
		const std::vector<uint32_t> &cs = lcol[top_first[s]];
		const uint32_t end = out.first[s + 1];
		std::vector<uint32_t>::const_iterator it = std::lower_bound(cs.begin(), cs.end(), end);
		out.rows.insert(out.rows.end(), it, cs.end());
		out.row_ptr[s + 1] = out.rows.size();
		const uint64_t wdt = end - out.first[s], h = out.row_ptr[s + 1] - out.row_ptr[s];
		out.nnzb_factor += wdt * (wdt + 1) / 2 + wdt * h;
		if(h) out.parent[s] = out.col_super[out.rows[out.row_ptr[s]]];
	}
	for(size_t s = 0; s < ns; ++ s)
		if(out.parent[s] != none)
			out.level[out.parent[s]] = std::max(out.level[out.parent[s]], out.level[s] + 1);
}

double plan_subtree_owners(const Supernodes &sn, int world, double min_saving, std::vector<int> &owner, std::vector<double> &work)
{
	const size_t ns = sn.n_super();
	owner.assign(ns, -1);
	work.assign(ns, 0.0);
	std::vector<double> sub(ns);
	double f_total = 0;
	for(size_t s = 0; s < ns; ++ s) {
		const double w = 6.0 * (sn.first[s + 1] - sn.first[s]), h = 6.0 * (sn.row_ptr[s + 1] - sn.row_ptr[s]);
		work[s] = sub[s] = w * w * w / 3 + w * w * h + w * h * h;
		f_total += work[s];
	}
	if(world < 2 || ns < 2 || !(f_total > 0))
		return 1.0;
	for(size_t s = 0; s < ns; ++ s) // a postorder: children precede their parent
		if(sn.parent[s] != 0xffffffffu) sub[sn.parent[s]] += sub[s];
	double f_best = f_total * (1 - min_saving);
	bool b_found = false;
	std::vector<int> plan(ns);
	static const double p_theta[] = {0.125, 0.25, 0.5, 1.0, 2.0};
	for(size_t k = 0; k < sizeof(p_theta) / sizeof(p_theta[0]); ++ k) {
		const double f_limit = f_total / world * p_theta[k];
		double f_shared = 0;
		std::vector<size_t> roots;
		for(size_t s = 0; s < ns; ++ s) {
			if(sub[s] > f_limit) {
				plan[s] = -1;
				f_shared += work[s];
			} else {
				plan[s] = -2;
				if(sn.parent[s] == 0xffffffffu || sub[sn.parent[s]] > f_limit) roots.push_back(s);
			}
		}
		std::stable_sort(roots.begin(), roots.end(), [&](size_t a, size_t b) { return sub[a] > sub[b]; });
		std::vector<double> load(world, 0.0);
		for(size_t q = 0; q < roots.size(); ++ q) {
			const int r = int(std::min_element(load.begin(), load.end()) - load.begin());
			load[r] += sub[roots[q]];
			plan[roots[q]] = r;
		}
		for(size_t ss = ns; ss > 0; -- ss) { // parents before children: a subtree inherits the rank of its root
			const size_t s = ss - 1;
			if(plan[s] == -2) plan[s] = plan[sn.parent[s]];
		}
		const double f_time = f_shared + *std::max_element(load.begin(), load.end());
		if(f_time < f_best) {
			f_best = f_time;
			owner = plan;
			b_found = true;
		}
	}
	return b_found? f_best / f_total : 1.0;
}

} // namespace spp

// ---- C ABI: pure host helpers (no context, no GPU) ------------------------------------------------------------

extern "C" int spp_block_ordering(size_t n_block_cols, const uint64_t *p_col_ptr, const uint64_t *p_row_idx, uint64_t *p_order)
{
	if(!p_col_ptr || !p_row_idx || !p_order)
		return SPP_ERR_INVALID;
	try {
		std::vector<uint32_t> order;
		spp::amd_exact_ordering(n_block_cols, p_col_ptr, p_row_idx, order);
		for(size_t i = 0; i < n_block_cols; ++ i) p_order[i] = order[i];
	} catch(const std::bad_alloc&) {
		return SPP_ERR_NOMEM;
	} catch(const std::exception&) {
		return SPP_ERR_INVALID;
	}
	return SPP_OK;
}

extern "C" int spp_block_symbolic_stats(size_t n_block_cols, const uint64_t *p_col_ptr, const uint64_t *p_row_idx,
	const uint64_t *p_order, uint64_t *p_col_count, uint64_t *p_parent, double *p_stats)
{
	if(!p_col_ptr || !p_row_idx)
		return SPP_ERR_INVALID;
	try {
		std::vector<uint32_t> order(n_block_cols);
		for(size_t i = 0; i < n_block_cols; ++ i) {
			if(p_order && p_order[i] >= n_block_cols) return SPP_ERR_INVALID;
			order[i] = p_order? (uint32_t)p_order[i] : (uint32_t)i;
		}
		spp::Supernodes sn;
		spp::supernodal_symbolic(n_block_cols, p_col_ptr, p_row_idx, order, 0.0, 0, (size_t)1 << 30, sn);
		for(size_t j = 0; j < n_block_cols; ++ j) {
			if(p_col_count) p_col_count[j] = sn.col_count[j];
			if(p_parent) p_parent[j] = (sn.col_parent[j] == 0xffffffffu)? UINT64_MAX : sn.col_parent[j];
		}
		if(p_stats) {
			p_stats[0] = (double)sn.nnzb_exact;
			p_stats[1] = sn.flops_blocks;
			p_stats[2] = (double)sn.n_super();
		}
	} catch(const std::bad_alloc&) {
		return SPP_ERR_NOMEM;
	} catch(const std::exception&) {
		return SPP_ERR_INVALID;
	}
	return SPP_OK;
}

extern "C" int spp_block_subtree_owners(size_t n_block_cols, const uint64_t *p_col_ptr, const uint64_t *p_row_idx,
	const uint64_t *p_order, int n_world, double f_min_saving, int32_t *p_owner, double *p_stats)
{
	if(!p_col_ptr || !p_row_idx || n_world < 1)
		return SPP_ERR_INVALID;
	try {
		std::vector<uint32_t> order(n_block_cols);
		for(size_t i = 0; i < n_block_cols; ++ i) {
			if(p_order && p_order[i] >= n_block_cols) return SPP_ERR_INVALID;
			order[i] = p_order? (uint32_t)p_order[i] : (uint32_t)i;
		}
		spp::Supernodes sn;
		spp::supernodal_symbolic(n_block_cols, p_col_ptr, p_row_idx, order, 0.05, 16, (size_t)1 << 30, sn); // the solver's amalgamation defaults
		std::vector<int> owner;
		std::vector<double> work;
		const double f_time = spp::plan_subtree_owners(sn, n_world, f_min_saving, owner, work);
		size_t n_shared = 0;
		for(size_t s = 0; s < owner.size(); ++ s) n_shared += owner[s] < 0;
		if(p_owner)
			for(size_t j = 0; j < n_block_cols; ++ j) p_owner[j] = owner[sn.col_super[j]];
		if(p_stats) {
			p_stats[0] = f_time;
			p_stats[1] = (double)sn.n_super();
			p_stats[2] = (double)n_shared;
		}
	} catch(const std::bad_alloc&) {
		return SPP_ERR_NOMEM;
	} catch(const std::exception&) {
		return SPP_ERR_INVALID;
	}
	return SPP_OK;
}
