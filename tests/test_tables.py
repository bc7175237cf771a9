"""LUT parity: the product's tables (pb_luts.c) == the oracle's == the reference's generated table.c, bit for bit."""
import numpy as np
import pytest

import oracle_lib
import pandaseq_b200 as pb


def test_product_tables_equal_oracle_tables(built):
    a, b = pb.tables(), oracle_lib.tables("port")
    for k in a:
        assert np.array_equal(np.asarray(a[k]).view(np.uint64), np.asarray(b[k]).view(np.uint64)), k


@pytest.mark.skipif(not oracle_lib.have_ref(), reason="oracle/_ref not built (needs /root/reference)")
def test_tables_equal_reference_table_c(built):
    a, r = pb.tables(), oracle_lib.tables("ref")
    for k in a:
        assert np.array_equal(np.asarray(a[k]).view(np.uint64), np.asarray(r[k]).view(np.uint64)), k


def test_known_table_values(built):
    t = pb.tables()
    assert t["qual_nn"] == -1.38629                      # tablebuilder.c:124 prints log(0.25) with %g
    assert t["score"][0] == -2.0                         # mktable.c:68-70: PHRED 0 scores -2
    assert t["score_err"][0] == 0.0
    assert t["score"][40] == float("%g" % np.log(1 - 1e-4))
    assert np.allclose(t["match_sb"], t["match_sb"].T) and np.allclose(t["mismatch_pear"], t["mismatch_pear"].T)
    assert np.array_equal(t["mismatch_rdp"], t["mismatch_sb"])   # mktable.c:84-92 is the same expression as :33-41


def test_pear_test_table(built):
    """The table the device multiplies for pear_test (plugin_pear_test.c:31-35): partial sums of the plugin's own expression, in its
    order; compared with an independent evaluation through Python's math (CPython has its own lgamma, so to 1e-14, not bit for bit:
    the bit-level pin is the oracle against the compiled plugin, tests/test_filters_hang_oracle_vs_ref.py) and a binomial identity."""
    import ctypes as C
    import math
    rows, cols = pb.PB_MAX_LEN, pb.PB_MAX_LEN + 2
    buf = np.zeros(rows * cols, dtype=np.float64)
    fn = pb.lib().pb_build_pear_cdf
    fn.argtypes, fn.restype = [C.c_void_p], None
    fn(buf.ctypes.data)
    cdf = buf.reshape(rows, cols)
    assert np.all(cdf[:, 0] == 0.0)
    for i in (0, 1, 2, 7, 60, 149, 150, 299, 449):
        s = 0.0
        for k in range(i + 1):
            s += math.exp(math.lgamma(i + 1) - math.lgamma(k + 1) - math.lgamma(i - k + 1) + k * math.log(0.25) + (i - k) * math.log(0.75))
            assert abs(cdf[i, k + 1] - s) <= 1e-14 * max(1.0, s) * (k + 1), (i, k)
        assert abs(cdf[i, i + 1] - 1.0) < 1e-9                                   # the whole binomial distribution
        assert np.all(cdf[i, i + 1:] == cdf[i, i + 1])                           # limits past i + 1 add nothing
    assert np.all(np.diff(cdf, axis=1) >= 0)
