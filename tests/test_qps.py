"""QPS front end: qpalm_b200_qps_read / read_settings / qps_solve against the reference reader.

CPU tests: the reader is host code of the product library (no compute call).  It is pinned three ways:
  * committed reference outputs tests/golden/qps_ref_outputs.json (made by tests/golden/make_golden_qps.py from the
    unmodified interfaces/qps/src/qpalm_qps.c) -- bit-exact CSC arrays, bounds, q, c;
  * hand-computed expectations for what the reference cannot read (MI bounds with a set name crash it, and its
    old-format converter cuts numbers at fixed offsets; qpalm_qps.c:163-170, qps_conversion.c:88-96);
  * the live reference reader on freshly generated files when oracle/_ref is present.
GPU test: a QPS file goes through qpalm_b200_qps_solve and must match the reference solve of the same data.
"""
import json
import os

import numpy as np
import pytest
import scipy.sparse as sp

from qpalm_b200 import problems, qps
from oracle import refbind

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")
KEYS = ("A_p", "A_i", "A_x", "Q_p", "Q_i", "Q_x", "q", "bmin", "bmax")
with open(os.path.join(GOLD, "qps_ref_outputs.json")) as _f:
    REF = json.load(_f)


def _same(a, b):
    assert (a.n, a.m, a.c) == (b.n, b.m, b.c)
    for k in KEYS:
        assert np.array_equal(getattr(a, k), getattr(b, k)), k


@pytest.mark.parametrize("fname", sorted(k for k in REF if k.endswith(".qps")))
def test_reader_matches_committed_reference_outputs(fname):
    p = qps.read_qps(os.path.join(GOLD, "qps", fname))
    r = REF[fname]
    assert (p.n, p.m, p.c) == (r["n"], r["m"], r["c"])
    for k in KEYS:
        assert np.array_equal(getattr(p, k), np.array(r[k])), k


def test_hand1_layout_by_hand():
    """Spot-check of the construction rules on tests/golden/qps/hand1.qps, independent of any reader."""
    p = qps.read_qps(os.path.join(GOLD, "qps", "hand1.qps"))
    assert (p.name, p.n, p.m) == ("HAND1", 5, 9)            # 5 rows + 4 bound rows (X3 is FR)
    assert p.c == 2.5                                        # c = -RHS(objective)
    assert p.q.tolist() == [1.0, 2.0, -1.0, 0.0, 0.0]
    # LIM1 (L): (-inf, 4]; LIM2 (G): [1, inf); MYEQN (E): 7; RNG1 (G, R=2): [-3, -1]; RNG2 (L, R=4): [6, 10]
    assert p.bmin[:5].tolist() == [-1e20, 1.0, 7.0, -3.0, 6.0]
    assert p.bmax[:5].tolist() == [4.0, 1e20, 7.0, -1.0, 10.0]
    # bound rows, in column order without X3: X1 [0,4], X2 [-1,1], X4 fixed 0.5, X5 default
    assert p.bmin[5:].tolist() == [0.0, -1.0, 0.5, 0.0] and p.bmax[5:].tolist() == [4.0, 1.0, 0.5, 1e20]
    # the identity entry is the LAST entry of its column; the 1e25 coefficient is clipped to 1e20
    assert p.A_i[p.A_p[1] - 1] == 5 and p.A_i[p.A_p[2] - 1] == 6 and p.A_i[p.A_p[4] - 1] == 7 and p.A_i[p.A_p[5] - 1] == 8
    assert p.A_p[3] - p.A_p[2] == 2                          # X3: two constraint entries, no bound row
    assert 1e20 in p.A_x.tolist() and 1e25 not in p.A_x.tolist()
    assert p.Q_p.tolist() == [0, 2, 3, 5, 5, 6] and p.Q_i.tolist() == [0, 1, 1, 2, 4, 4]


def test_fixed_format_names_with_blanks_and_ignored_MI():
    p = qps.read_qps(os.path.join(GOLD, "qps", "mine_mi_fixed.qps"))
    assert (p.n, p.m) == (3, 4)                              # 2 rows + bound rows of COL A, COL B (MI is ignored like the reference does)
    assert p.A_p.tolist() == [0, 3, 6, 7] and p.A_i.tolist() == [0, 1, 2, 0, 1, 3, 1]
    assert p.A_x.tolist() == [2.0, -1.0, 1.0, 1.0, 1.0, 1.0, 3.0]
    assert p.bmin.tolist() == [-1e20, 4.0, 0.0, 0.0] and p.bmax.tolist() == [10.0, 4.0, 5.0, 1e20]
    assert p.q.tolist() == [1.5, 0.0, -2.0] and p.Q_i.tolist() == [0, 2, 2]


def test_errors(tmp_path):
    with pytest.raises(RuntimeError, match="code 1"):
        qps.read_qps(str(tmp_path / "missing.qps"))
    bad = tmp_path / "bad.qps"
    bad.write_text("ROWS\n N obj\nENDATA\n")
    with pytest.raises(RuntimeError, match="code 2"):
        qps.read_qps(str(bad))
    bad.write_text("NAME X\nROWS\n N obj\n L r1\nCOLUMNS\n    x  r9  1.0\nENDATA\n")
    with pytest.raises(RuntimeError, match="code 2"):
        qps.read_qps(str(bad))


def test_settings_file():
    s = qps.read_settings(os.path.join(GOLD, "qps_settings.txt"))
    assert s.pop("_rc") == 0
    assert s == REF["qps_settings.txt"]
    assert (s["eps_abs"], s["max_iter"], s["proximal"], s["scaling"], s["max_rank_update"]) == (1e-7, 5000, 0, 2, 80)
    assert s["sigma_init"] == 20.0 and s["gamma_upd"] == 10.0   # untouched defaults (constants.h:65-99)


def test_settings_unknown_name_stops_like_the_reference(tmp_path):
    f = tmp_path / "s.txt"
    f.write_text("#\n#\n#\n#\n#\neps_abs 1e-3\nnot_a_setting 1\neps_rel 1e-3\n")
    s = qps.read_settings(str(f))
    assert s["_rc"] == 3 and s["eps_abs"] == 1e-3 and s["eps_rel"] == 1e-4   # reading stops at the unknown name


def _random_problem_file(tmp_path, seed, n, m0, **kw):
    rng = np.random.default_rng(seed)
    A = sp.random(m0, n, density=0.3, random_state=np.random.RandomState(seed), format="csc")
    lo, up = -rng.random(m0), rng.random(m0)
    kind = rng.integers(0, 4, m0)
    lo[kind == 0] = -1e20; up[kind == 1] = 1e20; up[kind == 2] = lo[kind == 2]
    M = sp.random(n, n, density=0.2, random_state=np.random.RandomState(seed + 7), format="csc")
    Q = sp.tril(M @ M.T + sp.eye(n), format="csc")
    vlo, vup = np.zeros(n), np.full(n, 1e20)
    vk = rng.integers(0, 5, n)
    vlo[vk == 1] = -1e20
    vlo[vk == 2] = -rng.random((vk == 2).sum())
    vup[vk == 3] = 1 + rng.random((vk == 3).sum())
    vlo[vk == 4] = vup[vk == 4] = rng.random((vk == 4).sum())
    path = str(tmp_path / f"rnd{seed}.qps")
    qps.write_qps(path, f"RND{seed}", A, lo, up, rng.standard_normal(n), Q, c=0.5, var_lo=vlo, var_up=vup, **kw)
    return path


@pytest.mark.skipif(not refbind.reference_reader_available(), reason="oracle/_ref not built (no /root/reference here)")
@pytest.mark.parametrize("seed,n,m0,kw", [(0, 30, 20, {}), (1, 200, 150, dict(rhs_name=None, bnd_name=None)),
                                          (2, 64, 1, dict(two_per_line=False)), (3, 500, 700, {})])
def test_reader_matches_live_reference(tmp_path, seed, n, m0, kw):
    path = _random_problem_file(tmp_path, seed, n, m0, **kw)
    _same(qps.read_qps(path), refbind.read_qps_reference(path))


@pytest.mark.gpu
def test_qps_solve_matches_reference_solve(tmp_path):
    """grid QP (the C2 stand-in) written as QPS, solved through qpalm_b200_qps_solve, against the reference on the same data."""
    from qpalm_b200.interface import solve_qp
    from qpalm_b200.problems import CSC
    g = problems.grid_qp(14, seed=3)
    n = g.n
    Afull = sp.csc_matrix((g.A.x, g.A.i, g.A.p), shape=(g.m, n))
    m0 = g.m - n                                             # the last n rows of grid_qp are the identity box rows
    path = str(tmp_path / "grid.qps")
    qps.write_qps(path, "GRID14", Afull[:m0], g.bmin[:m0], g.bmax[:m0], g.q, sp.csc_matrix((g.Q.x, g.Q.i, g.Q.p), shape=(n, n)),
                  var_lo=g.bmin[m0:], var_up=g.bmax[m0:])
    st = tmp_path / "settings.txt"
    st.write_text("#\n#\n#\n#\n#\neps_abs 1e-6\neps_rel 1e-6\nverbose 0\n")
    info, x, y = qps.solve_qps(path, str(st))
    p = qps.read_qps(path)
    ref = solve_qp("reference" if refbind.have_reference() else "oracle",
                   CSC(n, n, p.Q_p, p.Q_i, p.Q_x, -1), CSC(p.m, n, p.A_p, p.A_i, p.A_x, 0), p.q, p.bmin, p.bmax, c=p.c,
                   eps_abs=1e-6, eps_rel=1e-6, verbose=0)
    assert info["status_val"] == ref.status_val == 1
    assert np.max(np.abs(x - ref.x)) / max(1.0, np.max(np.abs(ref.x))) < 1e-8
    assert np.max(np.abs(y - ref.y)) / max(1.0, np.max(np.abs(ref.y))) < 1e-8
    assert abs(info["iter"] - ref.iter) <= max(1, ref.iter // 20)


def test_reader_extensions_beyond_the_reference(tmp_path):
    """Inputs the reference reader does not accept: comment lines, an OBJSENSE section, QMATRIX (full symmetric listing:
    only the lower triangle is kept, as QUADOBJ would have given) and RHS / BOUNDS without set names mixed with comments."""
    f = tmp_path / "ext.qps"
    f.write_text(
        "NAME EXT\n* a comment\nOBJSENSE\n    MIN\nROWS\n N obj\n G r1\nCOLUMNS\n    x obj 1.0 r1 1.0\n    y obj 2.0 r1 1.0\n"
        "RHS\n    r1 1.0\n* another comment\nBOUNDS\n UP x 3.0\n PL y\nQMATRIX\n    x x 2.0\n    x y -1.0\n    y x -1.0\n    y y 4.0\nENDATA\n")
    p = qps.read_qps(str(f))
    assert (p.n, p.m) == (2, 3)
    assert p.Q_p.tolist() == [0, 2, 3] and p.Q_i.tolist() == [0, 1, 1] and p.Q_x.tolist() == [2.0, -1.0, 4.0]
    assert p.bmin.tolist() == [1.0, 0.0, 0.0] and p.bmax.tolist() == [1e20, 3.0, 1e20]
    assert p.A_i.tolist() == [0, 1, 0, 2]
