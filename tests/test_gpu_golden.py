"""The CUDA path against the reference's golden vectors (no oracle in between)."""
import pytest

import golden_util
import pandaseq_b200 as pb
from parity import compare

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx(built):
    c = pb.Context(0)
    yield c
    c.close()


@pytest.mark.parametrize("name", golden_util.NAMES)
def test_gpu_matches_reference_golden(ctx, name):
    batch, cfg, want = golden_util.load(name)
    got = ctx.assemble_host(cfg, batch, want_nt=True, want_p=True)
    rep = compare(got, want, emitted_only_ok=True)
    assert rep["ok"], rep
