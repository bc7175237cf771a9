"""The oracle port against the golden vectors produced by the unmodified reference: everything bit-identical."""
import numpy as np
import pytest

import golden_util
import oracle_lib


@pytest.mark.parametrize("name", golden_util.NAMES)
def test_oracle_port_matches_reference_golden(built, name):
    batch, cfg, want = golden_util.load(name)
    got = oracle_lib.assemble("port", cfg, batch)
    assert np.array_equal(got["status"], want["status"])
    assert np.array_equal(got["slow"], want["slow"])
    assert np.array_equal(got["counters"], want["counters"])
    ok = want["status"] == 0
    for k in ("overlap", "seq_len", "mismatches", "degenerates", "examined", "fwd_offset", "rev_offset"):
        assert np.array_equal(got[k][ok], want[k][ok]), k
    for k in ("quality", "est_prob"):       # doubles compared as bit patterns
        assert np.array_equal(got[k][ok].view(np.uint64), want[k][ok].view(np.uint64)), k
    w = want["seq_nt"].shape[1]
    assert np.array_equal(got["seq_nt"][ok][:, :w], want["seq_nt"][ok])
    assert np.array_equal(got["seq_p"][ok][:, :w].view(np.uint64), want["seq_p"][ok].view(np.uint64))


def test_golden_set_is_complete():
    assert len(golden_util.NAMES) >= 12
