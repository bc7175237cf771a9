"""CPU checks of bench.py's host-side helpers: workload labels and sizes follow the BASELINE configs, the byte model is SURVEY.md 8d's,
binding a rank next to its GPU is a no-op (not an error) where there is no GPU or no PCI locality to read."""
import argparse
import importlib.util
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
bench = importlib.util.module_from_spec(spec)
spec.loader.exec_module(bench)


def test_algorithmic_read_bytes():
    # 4-bit nt + 8-bit PHRED for both reads + 8 B of metadata: 458 B at 2x150
    assert bench.algorithmic_read_bytes(150, 150) == 458
    assert bench.algorithmic_read_bytes(151, 75) == 76 + 38 + 151 + 75 + 8
    a = bench.algorithmic_read_bytes(np.array([150, 300]), np.array([150, 300]))
    assert a.tolist() == [458, 908]


def test_labels_and_sizes_follow_the_config():
    from pandaseq_b200 import synth
    assert bench.shape_label(synth.CONFIGS[2]) == "2x150 bp"
    assert bench.shape_label(synth.CONFIGS[5]) == "2x(75-300) bp mixed lengths"
    args = argparse.Namespace(pairs=None)
    assert bench.pairs_per_gpu(synth.CONFIGS[2], args) == 10_000_000
    assert bench.pairs_per_gpu(synth.CONFIGS[5], args) == 12_500_000          # 100 M pairs over 8 GPUs
    assert bench.pairs_per_gpu(synth.CONFIGS[3], argparse.Namespace(pairs=1234)) == 1234


def test_binding_without_a_gpu_is_a_no_op():
    info = bench.bind_near_gpu(0)
    assert info["bound"] is False or info["cpus"]
