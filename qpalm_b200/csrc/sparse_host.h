// sparse_host.h -- host-side result of the one-time symbolic analysis (sparse_sym.cu) consumed by sparse.cu.
#pragma once
#include <vector>

namespace qb {

struct SymHost {
  int n = 0, nsuper = 0, nlevels = 0, max_ns = 0, max_nf = 0;
  long long nnzS = 0, nnzL = 0, upd_total = 0;
  double flops = 0;
  std::vector<int> perm, iperm;         // perm[new] = old, iperm[old] = new
  std::vector<int> sn_first;            // nsuper + 1: first (permuted) column of each supernode
  std::vector<int> sn_of_col;           // n
  std::vector<int> rows_off;            // nsuper + 1: offsets into rowidx / rel
  std::vector<int> rowidx;              // permuted row indices below each supernode, ascending
  std::vector<int> rel;                 // position of each of those rows in the parent's front [cols ; rows]
  std::vector<int> sn_parent;           // assembly tree
  std::vector<long long> panel_off, upd_off;   // nsuper + 1 (doubles)
  std::vector<int> child_ptr, child_idx;
  std::vector<int> lvl_ptr, lvl_sn;     // supernodes grouped by level (leaves = level 0)
  std::vector<int> lvl_max_ns, lvl_max_nf, lvl_max_child_nr;
};

// returns 0 ok, 1 union pattern too dense (and !force), < 0 internal error
int symbolic_analyze(int n, int m, const int *Acsc_p, const int *Acsc_i, const int *Acsr_p, const int *Acsr_j,
                     const long long *Qp, const long long *Qi, bool force, SymHost *S);

}  // namespace qb
