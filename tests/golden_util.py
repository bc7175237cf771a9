"""Load a tests/golden fixture (written by tests/golden/make_golden.py from the compiled reference)."""
import ast
import glob
import os

import numpy as np

import datasets
import pandaseq_b200 as pb
from pandaseq_b200.synth import FlatBatch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NAMES = sorted(n for n in (os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))
               if not n.startswith("io_"))          # io_*.npz: the FASTQ / text stages, see test_io_golden.py


def load(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    batch = FlatBatch(z["f_data"], z["f_off"], z["r_data"], z["r_off"])
    spec = ast.literal_eval(str(z["spec"]))
    kw = dict(spec)
    algo = kw.pop("algo")
    if kw.pop("primers", False):
        fwd, rev = datasets.primer_codes()
        kw.update(forward_primer=fwd, reverse_primer=rev)
    if kw.pop("hang", False):
        hf, hr = datasets.overhang_codes()
        kw.update(hang_forward=hf, hang_reverse=hr)
    cfg = pb.make_config(algo, **kw)
    want = {k: z[k] for k in ("status", "slow", "overlap", "seq_len", "mismatches", "degenerates", "examined", "fwd_offset",
                              "rev_offset", "quality", "est_prob", "seq_nt", "seq_p", "counters")}
    return batch, cfg, want
