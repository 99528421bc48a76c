"""The reference's OWN test-suite, relinked against the product.

oracle/Makefile (target `reftests`) compiles /root/reference/tests/run_all_tests.c and tests/src/*.c UNMODIFIED (CHOLMOD
build flags: -DDLONG -DUSE_CHOLMOD -DPROFILING -DPRINTING, -fcommon) and links them against qpalm_b200/libqpalm_b200.so:
every qpalm_* call, mat_vec / mat_tpose_vec / mat_inf_norm_* / ldlchol / ldlsolveLD_neg_dphi, the lin_alg.h kernels,
validate_* / update_status / print_final_message resolve to this repository's library (csrc/api.cu, csrc/compat.cu); only
cholmod_l_start / allocate_sparse / free_sparse / finish, which the tests use to allocate their INPUT matrices, come from
CHOLMOD Core objects.  SURVEY.md 8(c): against the reference's own library the suite reports 122 tests, 442 assertions,
0 failures; the same must hold here, on the GPU."""
import os
import re
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "oracle", "_ref", "run_all_tests_b200")


@pytest.mark.timeout(600)
def test_reference_test_suite_passes_against_the_product_library():
    if not os.path.exists(BIN):
        pytest.skip("oracle/_ref/run_all_tests_b200 not built (needs /root/reference at build time: make -C oracle reftests)")
    ldd = subprocess.run(["ldd", BIN], capture_output=True, text=True).stdout
    assert "libqpalm_b200.so" in ldd and "libqpalm_ref" not in ldd, ldd
    out = subprocess.run([BIN], capture_output=True, text=True, timeout=500)
    text = out.stdout + out.stderr
    m = re.search(r"(\d+) tests, (\d+) assertions, (\d+) failures", text)
    assert m, text[-3000:]
    tests, assertions, failures = (int(g) for g in m.groups())
    assert failures == 0 and out.returncode == 0, text[-4000:]
    assert tests == 122 and assertions == 442, (tests, assertions)      # the whole suite ran
