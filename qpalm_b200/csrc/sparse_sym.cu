// sparse_sym.cu -- host-side one-time symbolic analysis for the supernodal sparse Cholesky (sparse.cuh).
// Replaces cholmod_analyze as configured by /root/reference/src/solver_interface.c:523-541 (the reference uses the
// natural ordering; here a fill-reducing ordering is part of the design because the result only has to agree to 1e-8).
//
//   union pattern  S = pattern(Q) + pattern(A'A)          (row cliques of A, marker-based, capped)
//   ordering       approximate minimum degree on the quotient graph (element absorption, approximate external degree,
//                  aggressive absorption); written from the published algorithm (Amestoy/Davis/Duff 1996), no
//                  supervariables
//   etree          Liu's algorithm with path compression, then a depth-first postorder folded into the ordering
//   structure      row-subtree traversal for column counts and the row structure of each supernode's first column
//   supernodes     fundamental supernodes + relaxed amalgamation of a last child into its parent
//   assembly tree  parent = supernode of the first off-diagonal row; relative indices child -> parent front; levels
#include "sparse_host.h"

#include <algorithm>
#include <chrono>
#include <set>
#include <tuple>
#include <stdio.h>
#include <stdlib.h>

namespace qb {

// ------------------------------------------------------------------------------------------------
// approximate minimum degree
// ------------------------------------------------------------------------------------------------
static void min_degree_order(int n, const std::vector<std::vector<int>> &adj, std::vector<int> &order) {
  std::vector<std::vector<int>> vadj(adj), eadj(n), elem(n);
  std::vector<char> state(n, 0);            // 0 variable, 1 element, 2 dead (absorbed element)
  std::vector<int> degree(n), mark(n, -1), w(n, 0);
  // ties: QPALM_B200_MD_TIE = 0 lowest index, 1 most recently updated first (AMD's list heads), 2 least recently updated, 3 hashed
  const char *tb = getenv("QPALM_B200_MD_TIE");
  const int tie_mode = tb ? atoi(tb) : 1;
  std::vector<long long> tkey(n, 0);
  long long stamp = 0;
  auto tie = [&](int i) -> long long {
    if (tie_mode == 1) return -(++stamp);
    if (tie_mode == 2) return ++stamp;
    if (tie_mode == 3) return (long long)((unsigned)(i * 2654435761u) >> 4);
    return 0;
  };
  // Priority structure.  Default tie rule (most recently updated first) = degree buckets with LIFO insertion, O(1) per
  // operation: the same order as the (degree, -stamp) keys of the ordered set, which stays as the fallback for the
  // experimental tie rules (the set cost 2 log n per updated variable: 1.2 s of the 1.3 s analysis at n = 90 000).
  const bool buckets = (tie_mode == 1);
  std::vector<int> bhead, bnext, bprev;
  int mindeg = 0;
  auto b_insert = [&](int i, int d) {
    bnext[i] = bhead[d]; bprev[i] = -1;
    if (bhead[d] != -1) bprev[bhead[d]] = i;
    bhead[d] = i;
    if (d < mindeg) mindeg = d;
  };
  auto b_remove = [&](int i, int d) {
    if (bprev[i] != -1) bnext[bprev[i]] = bnext[i]; else bhead[d] = bnext[i];
    if (bnext[i] != -1) bprev[bnext[i]] = bprev[i];
  };
  std::set<std::tuple<int, long long, int>> heap;
  if (buckets) { bhead.assign(n + 1, -1); bnext.assign(n, -1); bprev.assign(n, -1); mindeg = n; }
  for (int i = 0; i < n; i++) {
    degree[i] = (int)vadj[i].size();
    if (buckets) b_insert(i, degree[i]);
    else { tkey[i] = tie(i); heap.insert({degree[i], tkey[i], i}); }
  }
  order.clear(); order.reserve(n);
  int wbase = 1;
  std::vector<int> Lp;
  for (int k = 0; k < n; k++) {
    int p;
    if (buckets) {
      while (bhead[mindeg] == -1) mindeg++;
      p = bhead[mindeg];
      b_remove(p, mindeg);
    } else {
      p = std::get<2>(*heap.begin());
      heap.erase(heap.begin());
    }
    order.push_back(p);
    // ---- form the new element L_p ----
    Lp.clear();
    mark[p] = k;
    for (int v : vadj[p]) if (state[v] == 0 && mark[v] != k) { mark[v] = k; Lp.push_back(v); }
    for (int e : eadj[p]) {
      if (state[e] != 1) continue;
      for (int v : elem[e]) if (state[v] == 0 && mark[v] != k) { mark[v] = k; Lp.push_back(v); }
      state[e] = 2; std::vector<int>().swap(elem[e]);
    }
    state[p] = 1;
    std::vector<int>().swap(vadj[p]); std::vector<int>().swap(eadj[p]);
    // ---- |L_e \ L_p| for every element adjacent to a member of L_p ----
    if (wbase > 0x3fffffff - n) { std::fill(w.begin(), w.end(), 0); wbase = 1; }
    for (int i : Lp)
      for (int e : eadj[i]) {
        if (state[e] != 1) continue;
        if (w[e] < wbase) w[e] = wbase + (int)elem[e].size();
        w[e]--;
      }
    // ---- prune the adjacency of the members, update their degrees ----
    const int lp = (int)Lp.size();
    for (int i : Lp) {
      if (buckets) b_remove(i, degree[i]); else heap.erase({degree[i], tkey[i], i});
      size_t o = 0;
      for (int v : vadj[i]) if (state[v] == 0 && mark[v] != k) vadj[i][o++] = v;   // drop p, members of L_p, eliminated
      vadj[i].resize(o);
      o = 0;
      long long ext = 0;
      for (int e : eadj[i]) {
        if (state[e] != 1) continue;
        const int d = w[e] - wbase;
        if (d <= 0) { continue; }              // aggressive absorption: L_e is inside L_p
        eadj[i][o++] = e; ext += d;
      }
      eadj[i].resize(o);
      eadj[i].push_back(p);
      long long d = (long long)vadj[i].size() + (lp - 1) + ext;
      const long long d2 = (long long)degree[i] + (lp - 1);
      if (d > d2) d = d2;
      if (d > n - k - 2) d = n - k - 2;
      if (d < 0) d = 0;
      degree[i] = (int)d;
      if (buckets) b_insert(i, degree[i]);
      else { tkey[i] = tie(i); heap.insert({degree[i], tkey[i], i}); }
    }
    // an element absorbed above (L_e inside L_p) is referenced by no variable any more: all its variables are in L_p and
    // each of them just dropped it; it simply becomes unreachable
    elem[p].swap(Lp);
    Lp.clear();
    wbase += n + 1;
  }
}

// ------------------------------------------------------------------------------------------------
int symbolic_analyze(int n, int m, const int *Acsc_p, const int *Acsc_i, const int *Acsr_p, const int *Acsr_j,
                     const long long *Qp, const long long *Qi, bool force, SymHost *S) {
  S->n = n;
  const bool timing = getenv("QPALM_B200_SYM_TIMING") != nullptr;
  auto t_last = std::chrono::steady_clock::now();
  auto lap = [&](const char *what) {
    if (!timing) return;
    const auto now = std::chrono::steady_clock::now();
    fprintf(stderr, "[qpalm_b200] symbolic: %-28s %8.1f ms\n", what, std::chrono::duration<double, std::milli>(now - t_last).count());
    t_last = now;
  };
  // ---- union pattern (full symmetric adjacency without the diagonal) ----
  const double cap = force ? 4.0e18 : 0.12 * (double)n * (double)n + 64.0 * n;
  std::vector<std::vector<int>> adj(n);
  {
    std::vector<std::vector<int>> qadj(n);
    for (int j = 0; j < n; j++)
      for (long long k = Qp[j]; k < Qp[j + 1]; k++) {
        const int i = (int)Qi[k];
        if (i > j) { qadj[j].push_back(i); qadj[i].push_back(j); }
      }
    std::vector<int> mark(n, -1);
    double total = 0;
    for (int j = 0; j < n; j++) {
      mark[j] = j;
      std::vector<int> &a = adj[j];
      for (int i : qadj[j]) if (mark[i] != j) { mark[i] = j; a.push_back(i); }
      if (m > 0)
        for (int k = Acsc_p[j]; k < Acsc_p[j + 1]; k++) {
          const int r = Acsc_i[k];
          for (int t = Acsr_p[r]; t < Acsr_p[r + 1]; t++) { const int c = Acsr_j[t]; if (mark[c] != j) { mark[c] = j; a.push_back(c); } }
        }
      total += (double)a.size();
      if (total > cap) return 1;   // too dense: the caller takes the dense path
    }
    S->nnzS = (long long)(total / 2) + n;
  }
  lap("union pattern");
  // ---- fill-reducing ordering ----
  std::vector<int> order;
  min_degree_order(n, adj, order);
  lap("minimum-degree ordering");
  std::vector<int> iperm(n);
  for (int k = 0; k < n; k++) iperm[order[k]] = k;
  // ---- elimination tree of the permuted pattern ----
  auto build_lowrows = [&](const std::vector<int> &ip, std::vector<int> &lp, std::vector<int> &li) {
    lp.assign(n + 1, 0);
    for (int v = 0; v < n; v++) { const int i = ip[v]; for (int u : adj[v]) if (ip[u] < i) lp[i + 1]++; }
    for (int i = 0; i < n; i++) lp[i + 1] += lp[i];
    li.resize(lp[n]);
    std::vector<int> fill(lp.begin(), lp.end() - 1);
    for (int v = 0; v < n; v++) { const int i = ip[v]; for (int u : adj[v]) if (ip[u] < i) li[fill[i]++] = ip[u]; }
  };
  std::vector<int> lp, li, parent(n, -1);
  build_lowrows(iperm, lp, li);
  {
    std::vector<int> anc(n, -1);
    for (int i = 0; i < n; i++)
      for (int t = lp[i]; t < lp[i + 1]; t++) {
        int r = li[t];
        while (anc[r] != -1 && anc[r] != i) { const int nx = anc[r]; anc[r] = i; r = nx; }
        if (anc[r] == -1) { anc[r] = i; parent[r] = i; }
      }
  }
  // ---- postorder (children visited in ascending order), folded into the permutation ----
  {
    std::vector<int> head(n, -1), next(n, -1), post(n), stack;
    for (int j = n - 1; j >= 0; j--) if (parent[j] >= 0) { next[j] = head[parent[j]]; head[parent[j]] = j; }
    int cnt = 0;
    for (int root = 0; root < n; root++) {
      if (parent[root] != -1) continue;
      stack.push_back(root);
      while (!stack.empty()) {
        const int v = stack.back();
        const int c = head[v];
        if (c == -1) { post[v] = cnt++; stack.pop_back(); }
        else { head[v] = next[c]; stack.push_back(c); }
      }
    }
    std::vector<int> np(n, -1);
    for (int j = 0; j < n; j++) if (parent[j] >= 0) np[post[j]] = post[parent[j]];
    parent.swap(np);
    for (int v = 0; v < n; v++) iperm[v] = post[iperm[v]];
  }
  lap("etree + postorder");
  S->iperm = iperm;
  S->perm.assign(n, 0);
  for (int v = 0; v < n; v++) S->perm[iperm[v]] = v;
  build_lowrows(iperm, lp, li);
  // ---- column counts ----
  std::vector<int> count(n, 0), mark(n, -1);
  for (int i = 0; i < n; i++) {
    mark[i] = i;
    for (int t = lp[i]; t < lp[i + 1]; t++)
      for (int j = li[t]; mark[j] != i; j = parent[j]) { mark[j] = i; count[j]++; }
  }
  // ---- fundamental supernodes ----
  std::vector<int> fs_first;   // first column of each fundamental supernode
  for (int j = 0; j < n; j++)
    if (j == 0 || parent[j - 1] != j || count[j - 1] != count[j] + 1) fs_first.push_back(j);
  const int nfs = (int)fs_first.size();
  fs_first.push_back(n);
  std::vector<int> fs_of(n);
  for (int s = 0; s < nfs; s++) for (int j = fs_first[s]; j < fs_first[s + 1]; j++) fs_of[j] = s;
  // row structure below each fundamental supernode (rows of its first column beyond its last column)
  std::vector<std::vector<int>> rows(nfs);
  for (int s = 0; s < nfs; s++) rows[s].reserve(std::max(0, count[fs_first[s]] - (fs_first[s + 1] - fs_first[s] - 1)));
  std::fill(mark.begin(), mark.end(), -1);
  for (int i = 0; i < n; i++) {
    mark[i] = i;
    for (int t = lp[i]; t < lp[i + 1]; t++)
      for (int j = li[t]; mark[j] != i; j = parent[j]) {
        mark[j] = i;
        const int s = fs_of[j];
        if (j == fs_first[s] && i >= fs_first[s + 1]) rows[s].push_back(i);
      }
  }
  lap("counts + row structure");
  // ---- relaxed amalgamation: a supernode may absorb the supernode immediately before it when that one is its child ----
  std::vector<int> first(fs_first.begin(), fs_first.end() - 1), ncol(nfs);
  std::vector<double> zeros(nfs, 0.0);
  std::vector<char> alive(nfs, 1);
  for (int s = 0; s < nfs; s++) ncol[s] = fs_first[s + 1] - fs_first[s];
  for (int s = 0; s + 1 < nfs; s++) {
    if (rows[s].empty()) continue;
    const int p = fs_of[rows[s][0]];
    if (p != s + 1 || rows[s][0] >= fs_first[p + 1]) continue;      // not contiguous with its parent
    const double ns_s = ncol[s], nr_s = (double)rows[s].size(), ns_p = ncol[p], nr_p = (double)rows[p].size();
    const double z_new = ns_s * (ns_p + nr_p - nr_s);                // explicit zeros added to the columns of s
    const double nsm = ns_s + ns_p;
    const double ztot = zeros[s] + zeros[p] + z_new;
    const double size = nsm * (nsm + 1) / 2 + nsm * nr_p;
    const double zf = ztot / size;
    bool merge = false;
    if (nsm <= 4) merge = true;
    else if (nsm <= 16) merge = zf < 0.8;
    else if (nsm <= 48) merge = zf < 0.1;
    else merge = zf < 0.05;
    if (!merge || getenv("QPALM_B200_NO_AMALG")) continue;
    first[p] = first[s]; ncol[p] = (int)nsm; zeros[p] = ztot; alive[s] = 0;
    std::vector<int>().swap(rows[s]);
  }
  // ---- final supernodes ----
  std::vector<int> ids;
  for (int s = 0; s < nfs; s++) if (alive[s]) ids.push_back(s);
  const int ns_ = (int)ids.size();
  S->nsuper = ns_;
  S->sn_first.assign(ns_ + 1, n);
  S->rows_off.assign(ns_ + 1, 0);
  S->sn_of_col.assign(n, 0);
  for (int t = 0; t < ns_; t++) {
    S->sn_first[t] = first[ids[t]];
    S->rows_off[t + 1] = S->rows_off[t] + (int)rows[ids[t]].size();
  }
  for (int t = 0; t < ns_; t++) for (int j = S->sn_first[t]; j < S->sn_first[t + 1]; j++) S->sn_of_col[j] = t;
  S->rowidx.resize(S->rows_off[ns_]);
  for (int t = 0; t < ns_; t++) std::copy(rows[ids[t]].begin(), rows[ids[t]].end(), S->rowidx.begin() + S->rows_off[t]);
  // ---- assembly tree, relative indices, levels, storage offsets ----
  S->sn_parent.assign(ns_, -1);
  S->rel.assign(S->rowidx.size(), 0);
  std::vector<int> level(ns_, 0);
  S->panel_off.assign(ns_ + 1, 0); S->upd_off.assign(ns_ + 1, 0);
  S->max_ns = 0; S->max_nf = 0; S->flops = 0;
  for (int t = 0; t < ns_; t++) {
    const int ns = S->sn_first[t + 1] - S->sn_first[t], ro = S->rows_off[t], nr = S->rows_off[t + 1] - ro, nf = ns + nr;
    S->panel_off[t + 1] = S->panel_off[t] + (long long)nf * ns;
    S->upd_off[t + 1] = S->upd_off[t] + (long long)nr * nr;
    S->max_ns = std::max(S->max_ns, ns); S->max_nf = std::max(S->max_nf, nf);
    for (int k = 0; k < ns; k++) { const double r = nf - k - 1; S->flops += r * r + 2 * r + 1; }
    if (nr == 0) continue;
    const int p = S->sn_of_col[S->rowidx[ro]];
    S->sn_parent[t] = p;
    level[p] = std::max(level[p], level[t] + 1);
    const int pf = S->sn_first[p], pns = S->sn_first[p + 1] - pf, pro = S->rows_off[p], pnr = S->rows_off[p + 1] - pro;
    int q = 0;
    for (int i = 0; i < nr; i++) {
      const int r = S->rowidx[ro + i];
      if (r < pf + pns) { S->rel[ro + i] = r - pf; continue; }
      while (q < pnr && S->rowidx[pro + q] < r) q++;
      if (q >= pnr || S->rowidx[pro + q] != r) { fprintf(stderr, "[qpalm_b200] symbolic analysis: row %d of supernode %d missing in its parent\n", r, t); return -1; }
      S->rel[ro + i] = pns + q;
    }
  }
  S->nnzL = S->panel_off[ns_]; S->upd_total = S->upd_off[ns_];
  // children lists (ascending child index => fixed extend-add order)
  S->child_ptr.assign(ns_ + 1, 0);
  for (int t = 0; t < ns_; t++) if (S->sn_parent[t] >= 0) S->child_ptr[S->sn_parent[t] + 1]++;
  for (int t = 0; t < ns_; t++) S->child_ptr[t + 1] += S->child_ptr[t];
  S->child_idx.resize(S->child_ptr[ns_]);
  {
    std::vector<int> fill(S->child_ptr.begin(), S->child_ptr.end() - 1);
    for (int t = 0; t < ns_; t++) if (S->sn_parent[t] >= 0) S->child_idx[fill[S->sn_parent[t]]++] = t;
  }
  int nl = 0;
  for (int t = 0; t < ns_; t++) nl = std::max(nl, level[t] + 1);
  S->nlevels = nl;
  S->lvl_ptr.assign(nl + 1, 0);
  for (int t = 0; t < ns_; t++) S->lvl_ptr[level[t] + 1]++;
  for (int l = 0; l < nl; l++) S->lvl_ptr[l + 1] += S->lvl_ptr[l];
  S->lvl_sn.resize(ns_);
  {
    std::vector<int> fill(S->lvl_ptr.begin(), S->lvl_ptr.end() - 1);
    for (int t = 0; t < ns_; t++) S->lvl_sn[fill[level[t]]++] = t;
  }
  lap("supernodes + assembly tree");
  S->lvl_max_ns.assign(nl, 0); S->lvl_max_nf.assign(nl, 0); S->lvl_max_child_nr.assign(nl, 0);
  for (int t = 0; t < ns_; t++) {
    const int ns = S->sn_first[t + 1] - S->sn_first[t], nf = ns + S->rows_off[t + 1] - S->rows_off[t];
    S->lvl_max_ns[level[t]] = std::max(S->lvl_max_ns[level[t]], ns);
    S->lvl_max_nf[level[t]] = std::max(S->lvl_max_nf[level[t]], nf);
  }
  return 0;
}

}  // namespace qb
