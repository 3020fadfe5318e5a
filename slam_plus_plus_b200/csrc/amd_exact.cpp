// amd_exact.cpp -- the fill-reducing ordering of the block graph, permutation for permutation the one the reference
// computes. Pure host code (integer work, once per structure).
//
// The reference hands the block graph of A + A^T (diagonal dropped, every list ascending) to SuiteSparse's amd_l2 with
// default controls (CMatrixOrdering::p_BlockOrdering, src/slam/OrderingMagic.cpp:701-1033; the call is at :1003-1017;
// the third-party routine is vendored under src/cholmod/AMD, version 2.x: amd_2.c, amd_postorder.c, amd_post_tree.c).
// SuiteSparse is a dependency of the reference, not part of the path this library replaces, but the elimination order
// decides the structure of every factor downstream, so north_star asks for the same permutation. The algorithm is the
// published one (Amestoy, Davis, Duff, "An approximate minimum degree ordering algorithm", SIMAX 17(4), 1996, and
// "Algorithm 837: AMD", TOMS 30(3), 2004); what makes the permutation reproducible are its tie-breaking conventions,
// restated here:
//   * degree lists are LIFO: a variable enters its list at the head, the pivot is the head of the lowest list;
//   * the variables of a new element keep the order in which the pivot's lists yield them (its absorbed elements in
//     list order first, then its own variables), and re-enter the degree lists in that order;
//   * after the degree update, a variable's list is [new element, kept elements 2.., kept element 1, variables 2..,
//     variable 1] (the rotation that puts the new element first);
//   * approximate external degree d_i = min(d_i_old + |Lme \ i|, |Lme \ i| + sum of |Le \ Lme| + variables, n_left);
//   * aggressive absorption: an element whose external part |Le \ Lme| is empty is absorbed into the new element;
//   * mass elimination: a variable whose only element is the new one and that has no variables left joins the pivot;
//   * indistinguishable variables are found through hash buckets (sum of the list's members modulo n), bucket chains
//     are LIFO, the first variable of a chain absorbs its later twins;
//   * rows with more than max(16, 10 sqrt(n)) entries are set aside and ordered last, by index;
//   * the assembly tree is postordered with children ascending except that the child with the largest front (the last
//     one of equal size) is visited last; absorbed variables follow one another, by index, before their element.
// List storage differs from SuiteSparse (no in-place garbage collection: the list pool grows instead, which leaves
// the order of every list as it is), as do the hash buckets (arrays of their own rather than the degree lists' spare
// links). tests/test_ordering_cpu.py checks the result against the reference's own output (tests/golden/order_ref.npz,
// written through the unmodified CMatrixOrdering by the reference driver that tests/golden/make_golden_order.py runs).

#include "block_ordering.h"
#include <algorithm>
#include <stdexcept>
#include <limits>
#include <math.h>

namespace spp {

namespace {

const int32_t NONE = -1;
inline int32_t flipped(int32_t x) { return -x - 2; } // an index stored as a link to a parent; flipped(NONE) == NONE

class CQuotientGraph {
public:
	CQuotientGraph(size_t n_nodes, const std::vector<uint64_t> &adj_ptr, const std::vector<uint32_t> &adj)
		:n((int32_t)n_nodes), pool(adj.begin(), adj.end()), start(n), length(n), n_elems(n, 0), weight(n, 1), degree(n),
		mark(n, 1), deg_head(n, NONE), deg_next(n, NONE), deg_prev(n, NONE), bucket(n, NONE), chain(n, NONE), hash_of(n, 0),
		n_eliminated(0), n_dense(0), min_degree(0), pool_end((int32_t)adj.size()), stamp(2), max_front_ext(0)
	{
		if(adj.size() > (size_t)std::numeric_limits<int32_t>::max() / 4)
			throw std::runtime_error("amd: graph too large");
		for(int32_t i = 0; i < n; ++ i) {
			start[i] = (int32_t)adj_ptr[i];
			length[i] = degree[i] = (int32_t)(adj_ptr[i + 1] - adj_ptr[i]);
		}
		stamp_limit = std::numeric_limits<int64_t>::max() - n;
	}

	void Eliminate_All()
	{
		Fill_DegreeLists();
		while(n_eliminated < n) {
			const int32_t me = Pop_Pivot();
			Form_Element(me);
			Measure_ExternalParts();
			Update_Degrees(me);
			Merge_Twins();
			Return_Variables(me);
		}
	}

	void Get_Permutation(std::vector<uint32_t> &order); // order[new] = old

private:
	int32_t n;
	std::vector<int32_t> pool;    // every list: of a variable, its elements then its variables; of an element, its variables
	std::vector<int32_t> start;   // where the list of i begins; flipped(parent) once i is absorbed / eliminated into a parent
	std::vector<int32_t> length;  // entries in the list of i
	std::vector<int32_t> n_elems; // variable: leading entries that are elements; element: flipped(front size); NONE: absorbed variable
	std::vector<int32_t> weight;  // variables in supervariable i (negated while it sits in the element being formed; 0: absorbed / dense)
	std::vector<int32_t> degree;  // approximate external degree of a variable; |Le| of an element
	std::vector<int64_t> mark;    // elements: stamp + |Le \ Lme| during a step, 0 once absorbed; variables: scratch in Merge_Twins
	std::vector<int32_t> deg_head, deg_next, deg_prev;
	std::vector<int32_t> bucket, chain, hash_of;
	int32_t n_eliminated, n_dense, min_degree, pool_end;
	int64_t stamp, stamp_limit;
	int32_t max_front_ext;
	// the step in progress
	int32_t el_first, el_last; // the new element's variables are pool[el_first .. el_last]
	int32_t el_degree, el_pivots, el_had_elems;

	void Renew_Stamp()
	{
		if(stamp < 2 || stamp >= stamp_limit) {
			for(int32_t x = 0; x < n; ++ x)
				if(mark[x] != 0) mark[x] = 1;
			stamp = 2;
		}
	}

	void List_Insert(int32_t i, int32_t d)
	{
		const int32_t h = deg_head[d];
		if(h != NONE) deg_prev[h] = i;
		deg_next[i] = h;
		deg_prev[i] = NONE;
		deg_head[d] = i;
	}

	void List_Remove(int32_t i)
	{
		const int32_t a = deg_prev[i], b = deg_next[i];
		if(b != NONE) deg_prev[b] = a;
		if(a != NONE) deg_next[a] = b;
		else deg_head[degree[i]] = b;
	}

	void Pool_Push(int32_t v)
	{
		if((size_t)pool_end == pool.size())
			pool.resize(std::max((size_t)1024, pool.size() + pool.size() / 2));
		pool[pool_end ++] = v;
	}

	void Fill_DegreeLists()
	{
		int32_t dense = (int32_t)(10.0 * sqrt((double)n));
		dense = std::min(n, std::max(16, dense));
		for(int32_t i = 0; i < n; ++ i) {
			const int32_t d = degree[i];
			if(d == 0) { // no neighbour: an element of its own, a root
				n_elems[i] = flipped(1);
				++ n_eliminated;
				start[i] = NONE;
				mark[i] = 0;
			} else if(d > dense) { // set aside, ordered last
				++ n_dense;
				weight[i] = 0;
				n_elems[i] = NONE;
				++ n_eliminated;
				start[i] = NONE;
			} else
				List_Insert(i, d);
		}
	}

	int32_t Pop_Pivot()
	{
		int32_t d = min_degree;
		while(d < n && deg_head[d] == NONE) ++ d;
		if(d >= n) throw std::runtime_error("amd: degree lists empty");
		min_degree = d;
		const int32_t me = deg_head[d];
		const int32_t nx = deg_next[me];
		if(nx != NONE) deg_prev[nx] = NONE;
		deg_head[d] = nx;
		return me;
	}

	// takes variable i into the element being formed (once)
	bool Claim(int32_t i)
	{
		const int32_t wi = weight[i];
		if(wi <= 0) return false;
		el_degree += wi;
		weight[i] = -wi;
		List_Remove(i);
		return true;
	}

	void Form_Element(int32_t me)
	{
		el_had_elems = n_elems[me];
		el_pivots = weight[me];
		n_eliminated += el_pivots;
		weight[me] = -el_pivots;
		el_degree = 0;
		if(el_had_elems == 0) { // only variables: the element is the pivot's own list, compacted where it lies
			el_first = start[me];
			int32_t out = el_first;
			for(int32_t p = el_first, e = el_first + length[me]; p < e; ++ p) {
				const int32_t i = pool[p];
				if(Claim(i)) pool[out ++] = i;
			}
			el_last = out - 1;
		} else { // the union of the pivot's elements (in list order) and of its variables, appended to the pool
			const int32_t list = start[me], n_own_vars = length[me] - el_had_elems;
			el_first = pool_end;
			for(int32_t k = 0; k <= el_had_elems; ++ k) {
				const bool own = k == el_had_elems;
				const int32_t e = own? me : pool[list + k];
				const int32_t from = own? list + el_had_elems : start[e], count = own? n_own_vars : length[e];
				for(int32_t q = 0; q < count; ++ q) {
					const int32_t i = pool[from + q];
					if(Claim(i)) Pool_Push(i);
				}
				if(!own) { // e is absorbed
					start[e] = flipped(me);
					mark[e] = 0;
				}
			}
			el_last = pool_end - 1;
		}
		degree[me] = el_degree;
		start[me] = el_first;
		length[me] = el_last - el_first + 1;
		n_elems[me] = flipped(el_pivots + el_degree); // the size of the front
		Renew_Stamp();
	}

	// mark[e] - stamp = |Le \ Lme| for every element e next to a variable of the new element
	void Measure_ExternalParts()
	{
		for(int32_t p = el_first; p <= el_last; ++ p) {
			const int32_t i = pool[p], ne = n_elems[i];
			if(ne <= 0) continue;
			const int32_t wi = -weight[i];
			for(int32_t q = start[i], qe = start[i] + ne; q < qe; ++ q) {
				const int32_t e = pool[q];
				int64_t m = mark[e];
				if(m >= stamp) m -= wi;
				else if(m != 0) m = degree[e] + stamp - wi;
				mark[e] = m;
			}
		}
	}

	void Update_Degrees(int32_t me)
	{
		for(int32_t p = el_first; p <= el_last; ++ p) {
			const int32_t i = pool[p];
			const int32_t lo = start[i], mid = lo + n_elems[i], hi = lo + length[i];
			int32_t out = lo, deg = 0;
			uint64_t hash = 0;
			for(int32_t q = lo; q < mid; ++ q) { // elements: drop the absorbed ones, absorb those inside the new element
				const int32_t e = pool[q];
				const int64_t m = mark[e];
				if(m == 0) continue;
				const int64_t ext = m - stamp;
				if(ext > 0) {
					deg += (int32_t)ext;
					pool[out ++] = e;
					hash += (uint64_t)e;
				} else {
					start[e] = flipped(me);
					mark[e] = 0;
				}
			}
			const int32_t n_kept_elems = out - lo, vars_at = out;
			n_elems[i] = n_kept_elems + 1;
			for(int32_t q = mid; q < hi; ++ q) { // variables: the ones still in the graph and outside the new element
				const int32_t j = pool[q], wj = weight[j];
				if(wj > 0) {
					deg += wj;
					pool[out ++] = j;
					hash += (uint64_t)j;
				}
			}
			if(n_kept_elems == 0 && out == vars_at) { // nothing but the new element: eliminated with the pivot
				start[i] = flipped(me);
				const int32_t wi = -weight[i];
				el_degree -= wi;
				el_pivots += wi;
				n_eliminated += wi;
				weight[i] = 0;
				n_elems[i] = NONE;
			} else {
				degree[i] = std::min(degree[i], deg);
				// make room for the new element at the front: first variable to the end, first element behind the elements
				pool[out] = pool[vars_at];
				pool[vars_at] = pool[lo];
				pool[lo] = me;
				length[i] = out - lo + 1;
				const int32_t h = (int32_t)(hash % (uint64_t)n);
				chain[i] = bucket[h];
				bucket[h] = i;
				hash_of[i] = h;
			}
		}
		degree[me] = el_degree;
		max_front_ext = std::max(max_front_ext, el_degree);
		stamp += max_front_ext;
		Renew_Stamp();
	}

	// variables of the new element with identical lists become one supervariable
	void Merge_Twins()
	{
		for(int32_t p = el_first; p <= el_last; ++ p) {
			if(weight[pool[p]] >= 0) continue;
			const int32_t h = hash_of[pool[p]];
			int32_t i = bucket[h];
			bucket[h] = NONE; // a bucket is gone through once
			for(; i != NONE && chain[i] != NONE; i = chain[i], ++ stamp) {
				const int32_t ln = length[i], ne = n_elems[i];
				for(int32_t q = start[i] + 1, qe = start[i] + ln; q < qe; ++ q)
					mark[pool[q]] = stamp;
				int32_t before = i;
				for(int32_t j = chain[i]; j != NONE;) {
					bool same = length[j] == ln && n_elems[j] == ne;
					for(int32_t q = start[j] + 1, qe = start[j] + ln; same && q < qe; ++ q)
						same = mark[pool[q]] == stamp;
					if(same) {
						start[j] = flipped(i);
						weight[i] += weight[j]; // both negative
						weight[j] = 0;
						n_elems[j] = NONE;
						j = chain[j];
						chain[before] = j;
					} else {
						before = j;
						j = chain[j];
					}
				}
			}
		}
	}

	// the surviving variables of the new element go back to the degree lists, in element order
	void Return_Variables(int32_t me)
	{
		const int32_t n_left = n - n_eliminated;
		int32_t out = el_first;
		for(int32_t p = el_first; p <= el_last; ++ p) {
			const int32_t i = pool[p], wi = -weight[i];
			if(wi <= 0) continue;
			weight[i] = wi;
			const int32_t d = std::min(degree[i] + el_degree - wi, n_left - wi);
			List_Insert(i, d);
			min_degree = std::min(min_degree, d);
			degree[i] = d;
			pool[out ++] = i;
		}
		weight[me] = el_pivots;
		length[me] = out - el_first;
		if(length[me] == 0) { // a root
			start[me] = NONE;
			mark[me] = 0;
		}
		if(el_had_elems != 0)
			pool_end = out;
	}
};

void CQuotientGraph::Get_Permutation(std::vector<uint32_t> &order)
{
	// parents: elements point to the element that absorbed them, absorbed variables to their element (path compression)
	std::vector<int32_t> parent(n), front(n);
	for(int32_t i = 0; i < n; ++ i) {
		parent[i] = flipped(start[i]);
		front[i] = flipped(n_elems[i]);
	}
	for(int32_t i = 0; i < n; ++ i) {
		if(weight[i] != 0 || parent[i] == NONE) continue;
		int32_t e = parent[i];
		while(weight[e] == 0) e = parent[e];
		for(int32_t j = i; weight[j] == 0;) {
			const int32_t up = parent[j];
			parent[j] = e;
			j = up;
		}
	}
	// children of every element ascending, the largest front (the last of equals) moved to the end
	std::vector<std::vector<int32_t> > kids(n);
	for(int32_t j = 0; j < n; ++ j)
		if(weight[j] > 0 && parent[j] != NONE) kids[parent[j]].push_back(j);
	for(int32_t i = 0; i < n; ++ i) {
		std::vector<int32_t> &k = kids[i];
		if(k.size() < 2) continue;
		size_t big = 0;
		for(size_t c = 1; c < k.size(); ++ c)
			if(front[k[c]] >= front[k[big]]) big = c;
		std::rotate(k.begin() + big, k.begin() + big + 1, k.end());
	}
	// postorder of the assembly forest, roots ascending
	std::vector<int32_t> rank(n, NONE), by_rank;
	by_rank.reserve(n);
	std::vector<std::pair<int32_t, size_t> > stack;
	for(int32_t r = 0; r < n; ++ r) {
		if(parent[r] != NONE || weight[r] <= 0) continue;
		stack.push_back(std::make_pair(r, (size_t)0));
		while(!stack.empty()) {
			const int32_t f = stack.back().first;
			if(stack.back().second < kids[f].size()) {
				const int32_t c = kids[f][stack.back().second ++];
				stack.push_back(std::make_pair(c, (size_t)0));
			} else {
				rank[f] = (int32_t)by_rank.size();
				by_rank.push_back(f);
				stack.pop_back();
			}
		}
	}
	// positions: each element owns weight[e] consecutive places, its absorbed variables (ascending) first, itself last;
	// the variables that were set aside come after everything else, ascending
	std::vector<int32_t> cursor(n, 0);
	int32_t at = 0;
	for(size_t k = 0; k < by_rank.size(); ++ k) {
		cursor[by_rank[k]] = at;
		at += weight[by_rank[k]];
	}
	order.assign(n, 0);
	std::vector<int32_t> pos(n, NONE);
	for(int32_t i = 0; i < n; ++ i) {
		if(weight[i] != 0) continue;
		const int32_t e = parent[i];
		pos[i] = (e != NONE)? cursor[e] ++ : at ++;
	}
	for(int32_t i = 0; i < n; ++ i) {
		const int32_t k = (weight[i] != 0)? cursor[i] : pos[i];
		if(k < 0 || k >= n) throw std::runtime_error("amd: bad position");
		order[k] = (uint32_t)i;
	}
}

} // namespace

void amd_exact_ordering(size_t n, const uint64_t *col_ptr, const uint64_t *row_idx, std::vector<uint32_t> &order)
{
	order.clear();
	if(!n) return;
	if(n > (size_t)std::numeric_limits<int32_t>::max() / 2) throw std::runtime_error("amd: too many block columns");
	// A + A^T without the diagonal, every list ascending and free of duplicates
	std::vector<uint64_t> ptr(n + 1, 0);
	for(size_t c = 0; c < n; ++ c) {
		for(uint64_t k = col_ptr[c]; k < col_ptr[c + 1]; ++ k) {
			const uint64_t r = row_idx[k];
			if(r >= n) throw std::runtime_error("amd: row index out of range");
			if(r != c) { ++ ptr[c + 1]; ++ ptr[r + 1]; }
		}
	}
	for(size_t i = 0; i < n; ++ i) ptr[i + 1] += ptr[i];
	std::vector<uint32_t> adj(ptr[n]);
	{
		std::vector<uint64_t> fill(ptr.begin(), ptr.end() - 1);
		for(size_t c = 0; c < n; ++ c) {
			for(uint64_t k = col_ptr[c]; k < col_ptr[c + 1]; ++ k) {
				const uint64_t r = row_idx[k];
				if(r != c) { adj[fill[c] ++] = (uint32_t)r; adj[fill[r] ++] = (uint32_t)c; }
			}
		}
	}
	std::vector<uint64_t> uptr(n + 1, 0);
	size_t w = 0;
	for(size_t i = 0; i < n; ++ i) {
		std::sort(adj.begin() + ptr[i], adj.begin() + ptr[i + 1]);
		const size_t b = w;
		for(uint64_t k = ptr[i]; k < ptr[i + 1]; ++ k)
			if(w == b || adj[w - 1] != adj[k]) adj[w ++] = adj[k];
		uptr[i + 1] = w;
	}
	adj.resize(w);
	CQuotientGraph graph(n, uptr, adj);
	graph.Eliminate_All();
	graph.Get_Permutation(order);
}

} // namespace spp
