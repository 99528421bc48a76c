// qps.cu -- QPS front end (host code): problem file -> QPALMData in the CHOLMOD-CSC layout qpalm_setup uploads.
//
// Replaces interfaces/qps/src/qpalm_qps.c of the reference (get_sizes_and_check_format :69-214, read_data :216-575,
// read_settings :610-689, main :692-831) and its helpers index_hash.c / qps_conversion.c.  Same problem construction:
//   * constraint rows (L/G/E) keep their file order; the last N row is the objective,
//   * every column without an FR bound gets one extra row  0 <= x_j <= +inf  appended BELOW the constraint rows
//     (an identity entry stored last in its column), in column order,
//   * RHS on the objective row sets c = -value; RANGES widen L rows downwards and G rows upwards (E rows untouched);
//     UP/LO/FX set the bound row, every other bound type (MI, PL, BV, ...) is accepted and ignored,
//   * QUADOBJ entries are the lower triangle of Q (stype -1); matrix values are clipped to +-QPALM_INFTY.
// Designed differently from the reference: one pass over the file into hash maps and per-column entry lists instead
// of two passes with sscanf over fixed buffers, so name length and line length are unbounded; the old fixed-column
// format (names containing blanks) is parsed in place by field position instead of being rewritten to a `_copy.qps`
// file first.  The data then goes to the GPU through the ordinary qpalm_setup (api.cu).
#include "../../include/qpalm_b200.h"
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

namespace {

struct RowInfo { long index; char sign; };
struct Entry { long row; double val; };
struct QEntry { long col, row; double val; long seq; };
struct BoundOp { std::string type, col; double val; };

double clip_inf(double v) { return v > QPALM_INFTY ? QPALM_INFTY : (v < -QPALM_INFTY ? -QPALM_INFTY : v); }

bool read_line(FILE *fp, std::string &out) {
  out.clear();
  int ch;
  bool any = false;
  while ((ch = fgetc(fp)) != EOF) {
    any = true;
    if (ch == '\n') break;
    if (ch != '\r') out.push_back((char)ch);
  }
  return any;
}

std::vector<std::string> tokens_free(const std::string &line) {
  std::vector<std::string> t;
  size_t i = 0, n = line.size();
  while (i < n) {
    while (i < n && isspace((unsigned char)line[i])) i++;
    size_t j = i;
    while (j < n && !isspace((unsigned char)line[j])) j++;
    if (j > i) t.push_back(line.substr(i, j - i));
    i = j;
  }
  return t;
}

// fixed MPS fields (1-based columns 2-3, 5-12, 15-22, 25-36, 40-47, 50-61); blanks inside names are dropped,
// like qps_conversion.c:remove_spaces does when it rewrites an old-format file
std::vector<std::string> tokens_fixed(const std::string &line, bool first_field_is_type) {
  static const int lo[6] = {1, 4, 14, 24, 39, 49}, hi[6] = {3, 12, 22, 36, 47, 61};
  std::vector<std::string> t;
  for (int f = first_field_is_type ? 0 : 1; f < 6; f++) {
    if ((int)line.size() <= lo[f]) break;
    std::string s = line.substr(lo[f], std::min((size_t)(hi[f] - lo[f]), line.size() - lo[f]));
    std::string c;
    for (char ch : s) if (!isspace((unsigned char)ch)) c.push_back(ch);
    if (!c.empty()) t.push_back(c);
  }
  return t;
}

bool is_number(const std::string &s) {
  if (s.empty()) return false;
  char *end = nullptr;
  strtod(s.c_str(), &end);
  return end && *end == 0;
}

struct Problem {
  std::string name, objective;
  std::vector<std::string> n_rows;                    // every N row; the last one is the objective
  std::vector<std::string> row_names;                 // constraint rows only
  std::unordered_map<std::string, RowInfo> rows;
  std::vector<std::string> col_names;
  std::unordered_map<std::string, long> cols;
  std::vector<std::vector<Entry>> col_entries;
  std::vector<double> q, bmin0, bmax0;                // bmin0/bmax0: constraint rows
  std::vector<char> is_free;
  std::vector<BoundOp> bound_ops;
  std::vector<QEntry> qentries;
  double c = 0.0;
};

bool apply_rhs(Problem &P, const std::string &row, double v) {
  if (row == P.objective) { P.c = -v; return true; }
  auto it = P.rows.find(row);
  if (it == P.rows.end()) return std::find(P.n_rows.begin(), P.n_rows.end(), row) != P.n_rows.end();
  const long r = it->second.index;
  switch (it->second.sign) {
    case 'L': P.bmax0[r] = v; P.bmin0[r] = -QPALM_INFTY; break;
    case 'G': P.bmin0[r] = v; break;
    case 'E': P.bmin0[r] = v; P.bmax0[r] = v; break;
  }
  return true;
}

bool apply_range(Problem &P, const std::string &row, double v) {
  auto it = P.rows.find(row);
  if (it == P.rows.end()) return false;
  const long r = it->second.index;
  if (it->second.sign == 'L') P.bmin0[r] = P.bmax0[r] - v;
  else if (it->second.sign == 'G') P.bmax0[r] = P.bmin0[r] + v;
  return true;
}

// returns 0 ok, 1 cannot open, 2 format error
int parse(const char *path, Problem &P) {
  FILE *fp = fopen(path, "r");
  if (!fp) { fprintf(stderr, "Could not open file %s\n", path); return 1; }
  // whole file in one read, then split into lines (a per-character fgetc loop ran at 40 MB/s)
  std::vector<std::string> lines;
  {
    std::string buf;
    char chunk[1 << 16];
    size_t got;
    while ((got = fread(chunk, 1, sizeof chunk, fp)) > 0) buf.append(chunk, got);
    fclose(fp);
    size_t pos = 0;
    while (pos < buf.size()) {
      size_t nl = buf.find('\n', pos);
      if (nl == std::string::npos) nl = buf.size();
      size_t end = nl;
      if (end > pos && buf[end - 1] == '\r') end--;
      lines.emplace_back(buf, pos, end - pos);
      pos = nl + 1;
    }
  }
  if (lines.empty()) return 2;
  {
    std::vector<std::string> t = tokens_free(lines[0]);
    if (t.size() < 2 || t[0] != "NAME") {
      fprintf(stderr, "Wrong file format. Expected first line to contain NAME problem_name.\n");
      return 2;
    }
    P.name = t[1];
  }
  // old fixed-column format: a ROWS line with more than two blank-separated tokens (qpalm_qps.c:108-113)
  bool fixed = false;
  {
    bool in_rows = false;
    for (size_t li = 1; li < lines.size() && !fixed; li++) {
      const std::string &l = lines[li];
      if (l.empty() || l[0] == '*') continue;
      if (!isspace((unsigned char)l[0])) {
        const std::vector<std::string> t = tokens_free(l);
        if (!t.empty() && t[0] == "COLUMNS") break;
        in_rows = !t.empty() && t[0] == "ROWS";
        continue;
      }
      if (in_rows && tokens_free(l).size() > 2) fixed = true;
    }
  }
  std::string sec;
  std::string prev_col;
  for (size_t li = 1; li < lines.size(); li++) {
    const std::string &l = lines[li];
    if (l.empty() || l[0] == '*') continue;
    if (!isspace((unsigned char)l[0])) {
      std::vector<std::string> t = tokens_free(l);
      sec = t.empty() ? "" : t[0];
      if (sec == "ENDATA") break;
      continue;
    }
    const bool typed = (sec == "ROWS" || sec == "BOUNDS");
    std::vector<std::string> t = fixed ? tokens_fixed(l, typed) : tokens_free(l);
    if (t.empty()) continue;
    if (sec == "ROWS") {
      if (t.size() < 2) return 2;
      const char sign = t[0][0];
      if (sign == 'N') { P.objective = t[1]; P.n_rows.push_back(t[1]); continue; }
      if (sign != 'L' && sign != 'G' && sign != 'E') return 2;
      const long idx = (long)P.row_names.size();
      P.rows[t[1]] = RowInfo{idx, sign};
      P.row_names.push_back(t[1]);
      P.bmin0.push_back(sign == 'L' ? -QPALM_INFTY : 0.0);
      P.bmax0.push_back(sign == 'G' ? QPALM_INFTY : 0.0);
    } else if (sec == "COLUMNS") {
      if (t.size() >= 3 && t[1] == "'MARKER'") continue;      // integrality markers carry no data
      if (t.size() < 3) return 2;
      if (t[0] != prev_col) {
        auto it = P.cols.find(t[0]);
        if (it == P.cols.end()) {
          P.cols[t[0]] = (long)P.col_names.size();
          P.col_names.push_back(t[0]);
          P.col_entries.emplace_back();
          P.q.push_back(0.0);
        }
        prev_col = t[0];
      }
      const long col = P.cols[t[0]];
      for (size_t k = 1; k + 1 < t.size(); k += 2) {
        if (!is_number(t[k + 1])) return 2;
        const double v = strtod(t[k + 1].c_str(), nullptr);
        if (t[k] == P.objective) { P.q[col] = v; continue; }
        auto it = P.rows.find(t[k]);
        if (it == P.rows.end()) {
          if (std::find(P.n_rows.begin(), P.n_rows.end(), t[k]) != P.n_rows.end()) continue;   // unused extra N row
          fprintf(stderr, "QPS: unknown row %s in column %s\n", t[k].c_str(), t[0].c_str());
          return 2;
        }
        P.col_entries[col].push_back(Entry{it->second.index, clip_inf(v)});
      }
    } else if (sec == "RHS" || sec == "RANGES") {
      // "[set] row value [row value]": the set name is present iff the token count is odd (qpalm_qps.c:153-158)
      const size_t k0 = (t.size() % 2 == 1) ? 1 : 0;
      for (size_t k = k0; k + 1 < t.size(); k += 2) {
        if (!is_number(t[k + 1])) return 2;
        const double v = strtod(t[k + 1].c_str(), nullptr);
        if (!(sec == "RHS" ? apply_rhs(P, t[k], v) : apply_range(P, t[k], v))) {
          fprintf(stderr, "QPS: unknown row %s in %s\n", t[k].c_str(), sec.c_str());
          return 2;
        }
      }
    } else if (sec == "BOUNDS") {
      // "type [set] column [value]"
      const std::string &type = t[0];
      const bool has_value = (type == "UP" || type == "LO" || type == "FX");
      std::string col;
      double v = 0.0;
      if (has_value) {
        if (t.size() < 3 || !is_number(t.back())) return 2;
        col = t[t.size() - 2];
        v = strtod(t.back().c_str(), nullptr);
      } else {
        // FR / MI / PL / BV ...: "type [set] column", a trailing number (some writers emit one) is ignored
        if (t.size() < 2) return 2;
        size_t last = t.size() - 1;
        if (last >= 2 && is_number(t[last]) && P.cols.find(t[last]) == P.cols.end()) last--;
        col = t[last];
      }
      if (P.cols.find(col) == P.cols.end()) { fprintf(stderr, "QPS: unknown column %s in BOUNDS\n", col.c_str()); return 2; }
      P.bound_ops.push_back(BoundOp{type, col, v});
    } else if (sec == "QUADOBJ" || sec == "QMATRIX") {
      if (t.size() < 3 || !is_number(t[2])) return 2;
      auto ic = P.cols.find(t[0]), ir = P.cols.find(t[1]);
      if (ic == P.cols.end() || ir == P.cols.end()) return 2;
      const double v = strtod(t[2].c_str(), nullptr);
      if (sec == "QMATRIX" && ir->second < ic->second) continue;   // full symmetric listing: keep the lower triangle
      P.qentries.push_back(QEntry{ic->second, ir->second, clip_inf(v), (long)P.qentries.size()});
    }
    // other sections (OBJSENSE, ...) carry nothing the reference reads
  }
  P.is_free.assign(P.col_names.size(), 0);
  for (const BoundOp &b : P.bound_ops)
    if (b.type == "FR") { auto it = P.cols.find(b.col); if (it != P.cols.end()) P.is_free[it->second] = 1; }
  return 0;
}

solver_sparse *alloc_csc(size_t nrow, size_t ncol, size_t nnz, int stype) {
  solver_sparse *M = (solver_sparse *)calloc(1, sizeof(solver_sparse));
  M->nrow = nrow; M->ncol = ncol; M->nzmax = nnz ? nnz : 1;
  M->p = calloc(ncol + 1, sizeof(int64_t));
  M->i = calloc(M->nzmax, sizeof(int64_t));
  M->x = calloc(M->nzmax, sizeof(double));
  M->nz = nullptr; M->z = nullptr;
  M->stype = stype; M->itype = 2 /* CHOLMOD_LONG */; M->xtype = 1 /* CHOLMOD_REAL */; M->dtype = 0 /* CHOLMOD_DOUBLE */;
  M->sorted = 1; M->packed = 1;
  return M;
}

void free_csc(solver_sparse *M) {
  if (!M) return;
  free(M->p); free(M->i); free(M->x); free(M);
}

}  // namespace

extern "C" void qpalm_b200_qps_free(QPALMData *data) {
  if (!data) return;
  free_csc(data->A); free_csc(data->Q);
  free(data->q); free(data->bmin); free(data->bmax);
  free(data);
}

extern "C" int qpalm_b200_qps_read(const char *path, QPALMData **data_out, char *name_out, size_t name_len) {
  if (!path || !data_out) return 2;
  *data_out = nullptr;
  Problem P;
  if (int rc = parse(path, P)) return rc;
  const size_t n = P.col_names.size(), m0 = P.row_names.size();
  // bound row of column j: m0 + (number of non-free columns before j)
  std::vector<long> bound_row(n, -1);
  size_t n_bounds = 0;
  for (size_t j = 0; j < n; j++) if (!P.is_free[j]) bound_row[j] = (long)(m0 + n_bounds++);
  const size_t m = m0 + n_bounds;
  size_t annz = n_bounds;
  for (size_t j = 0; j < n; j++) annz += P.col_entries[j].size();

  QPALMData *d = (QPALMData *)calloc(1, sizeof(QPALMData));
  d->n = n; d->m = m; d->c = P.c;
  d->q = (c_float *)calloc(n ? n : 1, sizeof(c_float));
  d->bmin = (c_float *)calloc(m ? m : 1, sizeof(c_float));
  d->bmax = (c_float *)calloc(m ? m : 1, sizeof(c_float));
  for (size_t j = 0; j < n; j++) d->q[j] = P.q[j];
  for (size_t r = 0; r < m0; r++) { d->bmin[r] = P.bmin0[r]; d->bmax[r] = P.bmax0[r]; }
  for (size_t r = m0; r < m; r++) { d->bmin[r] = 0.0; d->bmax[r] = QPALM_INFTY; }
  for (const BoundOp &b : P.bound_ops) {
    auto it = P.cols.find(b.col);
    if (it == P.cols.end() || bound_row[it->second] < 0) continue;
    const long r = bound_row[it->second];
    if (b.type == "UP") d->bmax[r] = b.val;
    else if (b.type == "LO") d->bmin[r] = b.val;
    else if (b.type == "FX") { d->bmin[r] = b.val; d->bmax[r] = b.val; }
  }

  d->A = alloc_csc(m, n, annz, 0);
  {
    int64_t *Ap = (int64_t *)d->A->p, *Ai = (int64_t *)d->A->i;
    double *Ax = (double *)d->A->x;
    size_t k = 0;
    for (size_t j = 0; j < n; j++) {
      Ap[j] = (int64_t)k;
      for (const Entry &e : P.col_entries[j]) { Ai[k] = e.row; Ax[k] = e.val; k++; }
      if (bound_row[j] >= 0) { Ai[k] = bound_row[j]; Ax[k] = 1.0; k++; }
    }
    Ap[n] = (int64_t)k;
    d->A->nzmax = annz ? annz : 1;
  }
  std::stable_sort(P.qentries.begin(), P.qentries.end(), [](const QEntry &a, const QEntry &b) { return a.col < b.col; });
  d->Q = alloc_csc(n, n, P.qentries.size(), -1);
  {
    int64_t *Qp = (int64_t *)d->Q->p, *Qi = (int64_t *)d->Q->i;
    double *Qx = (double *)d->Q->x;
    size_t k = 0;
    for (size_t j = 0; j < n; j++) {
      Qp[j] = (int64_t)k;
      while (k < P.qentries.size() && (size_t)P.qentries[k].col == j) { Qi[k] = P.qentries[k].row; Qx[k] = P.qentries[k].val; k++; }
    }
    Qp[n] = (int64_t)k;
  }
  // Entries come out in FILE order (that is what the reference reader produces and what the parity tests pin), so a column
  // is sorted only if the file listed it that way: say so truthfully instead of claiming sorted = 1.
  for (solver_sparse *M : {d->A, d->Q}) {
    const int64_t *Mp = (const int64_t *)M->p, *Mi = (const int64_t *)M->i;
    int sorted = 1;
    for (size_t j = 0; j < M->ncol && sorted; j++)
      for (int64_t k = Mp[j] + 1; k < Mp[j + 1]; k++) if (Mi[k] <= Mi[k - 1]) { sorted = 0; break; }
    M->sorted = sorted;
  }
  if (name_out && name_len) { strncpy(name_out, P.name.c_str(), name_len - 1); name_out[name_len - 1] = 0; }
  *data_out = d;
  return 0;
}

// settings file: five header lines, then "name value" pairs (interfaces/qps/sample_settings.txt).
// Returns 0 ok, 1 cannot open, 3 stopped at an unrecognised setting (settings read so far are kept, as in the reference).
extern "C" int qpalm_b200_read_settings(const char *path, QPALMSettings *s) {
  if (!s) return 2;
  qpalm_set_default_settings(s);
  FILE *fp = path ? fopen(path, "r") : nullptr;
  if (!fp) return 1;
  std::string line;
  for (int i = 0; i < 5; i++) read_line(fp, line);
  char name[128];
  double v;
  int rc = 0;
  while (fscanf(fp, "%127s %le", name, &v) == 2) {
    const std::string k(name);
#define QB_SET_F(field) else if (k == #field) s->field = v
#define QB_SET_I(field) else if (k == #field) s->field = (c_int)v
    if (false) {}
    QB_SET_I(max_iter); QB_SET_I(inner_max_iter); QB_SET_F(eps_abs); QB_SET_F(eps_rel); QB_SET_F(eps_abs_in);
    QB_SET_F(eps_rel_in); QB_SET_F(rho); QB_SET_F(eps_prim_inf); QB_SET_F(eps_dual_inf); QB_SET_F(theta);
    QB_SET_F(delta); QB_SET_F(sigma_max); QB_SET_F(sigma_init); QB_SET_I(proximal); QB_SET_F(gamma_init);
    QB_SET_F(gamma_upd); QB_SET_F(gamma_max); QB_SET_I(scaling); QB_SET_I(nonconvex); QB_SET_I(verbose);
    QB_SET_I(print_iter); QB_SET_I(warm_start); QB_SET_I(reset_newton_iter); QB_SET_I(enable_dual_termination);
    QB_SET_F(dual_objective_limit); QB_SET_F(time_limit); QB_SET_I(ordering); QB_SET_I(factorization_method);
    QB_SET_I(max_rank_update); QB_SET_F(max_rank_update_fraction);
    else { printf("Unrecognised setting: %s\n", name); rc = 3; break; }
#undef QB_SET_F
#undef QB_SET_I
  }
  fclose(fp);
  return rc;
}

// Whole front end, the body of main() in qpalm_qps.c:692-831: read the problem (and optional settings), set the
// workspace up on the device, solve, hand the solution back.  x_out (n) / y_out (m) may be NULL; pass capacities in
// *n_io / *m_io (0 = just report the sizes).
extern "C" int qpalm_b200_qps_solve(const char *qps_path, const char *settings_path, QPALMInfo *info_out,
                                    c_float *x_out, c_float *y_out, size_t *n_io, size_t *m_io) {
  QPALMData *data = nullptr;
  if (int rc = qpalm_b200_qps_read(qps_path, &data, nullptr, 0)) return rc;
  QPALMSettings settings;
  if (settings_path) {
    if (qpalm_b200_read_settings(settings_path, &settings) == 1) {
      printf("Could not open file %s\nUsing default settings instead\n", settings_path);
      qpalm_set_default_settings(&settings);
    }
  } else {
    qpalm_set_default_settings(&settings);
  }
  QPALMWorkspace *work = qpalm_setup(data, &settings);
  if (!work) { qpalm_b200_qps_free(data); return 4; }
  qpalm_solve(work);
  if (info_out) *info_out = *work->info;
  if (x_out && n_io && *n_io >= data->n) memcpy(x_out, work->solution->x, sizeof(c_float) * data->n);
  if (y_out && m_io && *m_io >= data->m) memcpy(y_out, work->solution->y, sizeof(c_float) * data->m);
  if (n_io) *n_io = data->n;
  if (m_io) *m_io = data->m;
  qpalm_cleanup(work);
  qpalm_b200_qps_free(data);
  return 0;
}
