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
