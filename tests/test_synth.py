"""The synthetic generator is deterministic and produces the shapes the BASELINE configs name."""
import numpy as np

from pandaseq_b200 import synth


def test_deterministic_per_seed_and_chunk():
    a = synth.generate_config(1, n=200).to_flat()
    b = synth.generate_config(1, n=200).to_flat()
    c = synth.generate_config(1, n=200, chunk_index=1).to_flat()
    assert np.array_equal(a.f_data, b.f_data) and np.array_equal(a.r_data, b.r_data)
    assert not np.array_equal(a.f_data, c.f_data)


def test_config_shapes():
    b = synth.generate_config(2, n=100).to_flat()
    fl, rl = b.lengths()
    assert (fl == 150).all() and (rl == 150).all()
    assert set(np.unique(b.f_data[:, 0])) <= {1, 2, 4, 8}
    assert b.f_data[:, 1].min() >= 2 and b.f_data[:, 1].max() <= 41
    m = synth.generate_config(5, n=300).to_flat()
    fl, rl = m.lengths()
    assert fl.min() >= 75 and fl.max() <= 300 and rl.min() >= 75 and rl.max() <= 300 and len(np.unique(fl)) > 50
    p = synth.generate_config(4, n=20).to_flat()
    f, r = p.pair(0)
    assert "".join("NACMGRSVTWYHKDBN"[x] for x in f[:17, 0]) == synth.FWD_PRIMER


def test_overlap_is_what_the_template_length_implies():
    # forward[F-o+i] and reverse[R-1-i] are the same template base when there are no errors
    rect = synth.generate(50, rl=(100, 100), tmpl=(150, 150), seed=9)
    b = rect.to_flat()
    o = 100 + 100 - 150
    agree = 0
    for i in range(b.n):
        f, r = b.pair(i)
        agree += int(((f[100 - o:, 0] & r[::-1][:o, 0]) != 0).sum())
    assert agree > 0.95 * b.n * o


def test_flat_slice_and_concat_roundtrip():
    b = synth.generate_config(5, n=64).to_flat()
    parts = [b.slice(0, 10), b.slice(10, 40), b.slice(40, 64)]
    c = synth.FlatBatch.concat(parts)
    assert np.array_equal(c.f_data, b.f_data) and np.array_equal(c.r_off, b.r_off)
