"""warm-started sequential solves with qpalm_update_q / qpalm_update_bounds / qpalm_warm_start on one
workspace (src/qpalm.c:322-399, 793-871; tests/src/test_update.c, test_basic_qp.c:202 `iter < 12` after a warm start).
Every step of the loop must match the reference doing the same loop: status, x / y to 1e-8, iterations within 5 %."""
import numpy as np
import pytest

from conftest import HAS_REF
from qpalm_b200 import mpc

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return np.max(np.abs(a - b)) / max(1.0, np.max(np.abs(b)))


@pytest.mark.parametrize("n,m0,kw", [(48, 80, {}), (48, 80, dict(proximal=1, scaling=10)), (240, 709, {})])
def test_warm_started_sequence_matches_reference(n, m0, kw):
    b, seq = mpc.mpc_sequence(6, n=n, m0=m0, seed=4)
    ref_impl = "reference" if HAS_REF else "oracle"
    g, _ = mpc.run_sequence("b200", b, seq, **kw)
    r, _ = mpc.run_sequence(ref_impl, b, seq, **kw)
    cold, _ = mpc.run_sequence(ref_impl, b, seq, warm=False, **kw)
    for k, (a, c) in enumerate(zip(g, r)):
        assert a.status_val == c.status_val == 1, (k, a.status, c.status)
        assert _rel(a.x, c.x) < 1e-8 and _rel(a.y, c.y) < 1e-8, (k, _rel(a.x, c.x), _rel(a.y, c.y))
        assert abs(a.iter - c.iter) <= max(1, int(np.ceil(0.05 * c.iter))), (k, a.iter, c.iter)
    if n == 48:   # the warm start pays off on the small chain (on the chain80w-sized one it does not, in the reference either)
        assert sum(a.iter for a in g[1:]) < sum(c.iter for c in cold[1:])
