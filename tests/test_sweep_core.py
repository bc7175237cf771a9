"""The diagonal sweep's core (pandaseq_b200/csrc/pb_sweep.cuh: plane building, sweep, certificate) compiled for the HOST
(tests/c/sweep_host.cpp) against the oracle's table-based seeding (oracle/panda_oracle.c po_seed_bits, i.e. K1-K3 of
assembler.c:84-116): for every pair the sweep does not hand to the general kernel, the candidate-overlap masks must be
identical bit for bit."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import datasets
import oracle_lib
import pandaseq_b200 as pb
from pandaseq_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GENERAL = 1


@pytest.fixture(scope="module")
def host(built, tmp_path_factory):
    so = str(tmp_path_factory.mktemp("sweep") / "libsweep_host.so")
    subprocess.run(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-Wno-unknown-pragmas", "-I", os.path.join(ROOT, "pandaseq_b200", "csrc"),
                    "-I", os.path.join(ROOT, "include"), "-o", so, os.path.join(ROOT, "tests", "c", "sweep_host.cpp")], check=True)
    L = C.CDLL(so)
    L.sweep_host_run.restype = C.c_int
    L.sweep_host_run.argtypes = [C.c_int, C.c_size_t] + [C.c_void_p] * 4 + [C.c_int] + [C.c_void_p] * 3
    return L


def oracle_bits(cfg, b):
    L = oracle_lib._get("port")[0]
    L.po_seed_bits.restype = C.c_size_t
    L.po_seed_bits.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p]
    out = np.zeros((b.n, 32), np.uint32)
    nbits = np.zeros(b.n, np.int64)
    f_data, r_data = np.ascontiguousarray(b.f_data), np.ascontiguousarray(b.r_data)
    for i in range(b.n):
        fo, fe, ro, re = int(b.f_off[i]), int(b.f_off[i + 1]), int(b.r_off[i]), int(b.r_off[i + 1])
        nbits[i] = L.po_seed_bits(C.byref(cfg), f_data.ctypes.data + 2 * fo, fe - fo, r_data.ctypes.data + 2 * ro, re - ro,
                                  out[i].ctypes.data)
    return out, nbits


def sweep(host, nw, b, mo):
    cw = np.zeros((b.n, 16), np.uint32)
    flags = np.zeros(b.n, np.uint32)
    lowest = np.zeros(b.n, np.int32)
    f_data, r_data = np.ascontiguousarray(b.f_data), np.ascontiguousarray(b.r_data)
    f_off, r_off = np.ascontiguousarray(b.f_off, dtype=np.uint64), np.ascontiguousarray(b.r_off, dtype=np.uint64)
    rc = host.sweep_host_run(nw, b.n, f_data.ctypes.data, f_off.ctypes.data, r_data.ctypes.data, r_off.ctypes.data, mo,
                             cw.ctypes.data, flags.ctypes.data, lowest.ctypes.data)
    assert rc == 0
    return cw, flags, lowest


def check(host, nw, b, mo=2, max_general=None):
    cfg = pb.make_config("simple_bayesian", minoverlap=mo)
    want, nbits = oracle_bits(cfg, b)
    cw, flags, lowest = sweep(host, nw, b, mo)
    own = (flags & GENERAL) == 0
    assert np.array_equal(cw[own], want[own, :16]), np.nonzero((cw != want[:, :16]).any(axis=1) & own)[0][:10]
    # a pair the sweep keeps has at least one candidate, and `lowest` is its lowest bit
    for i in np.nonzero(own)[0][:2000]:
        bits = np.nonzero(np.unpackbits(cw[i].view(np.uint8), bitorder="little"))[0]
        assert len(bits) and bits[0] == lowest[i] and bits[-1] < nbits[i]
    # pairs without any candidate are the general kernel's (ALL_BITS_IF_NONE)
    none = ~want.any(axis=1)
    assert (flags[none] & GENERAL).all()
    if max_general is not None:
        assert (~own).sum() <= max_general, f"{(~own).sum()} of {b.n} pairs handed on"
    return int((~own).sum())


def test_config2_shape(host):
    b = synth.generate_config(2, n=30_000).to_flat()
    handed = check(host, 5, b, max_general=60)      # no N in this set: only lost k-mers and seedless pairs are handed on
    print("handed on:", handed)


def test_config3_shape(host):
    check(host, 8, synth.generate_config(3, n=6_000).to_flat(), max_general=20)


def test_mixed_lengths(host):
    b = synth.generate_config(5, n=6_000).to_flat()
    check(host, 10, b, max_general=30)
    check(host, 8, b)       # reads above 256 nt are handed on


def test_low_complexity(host):
    """Homopolymers and tandem repeats: identical 8-mers many times over, the lost k-mers of assembler.c:95-97."""
    b = datasets.low_complexity(600)
    handed = check(host, 5, b)
    assert handed < b.n      # some of them stay


def test_decorated_and_edge_sets(host):
    for b, nw in ((datasets.cfg1(4000), 5), (datasets.stress(3000), 5), (datasets.edge_cases(), 5), (datasets.edge_cases(), 10),
                  (datasets.long250(1500), 8)):
        check(host, nw, b)


@pytest.mark.parametrize("mo", [2, 9, 17, 40, 149])
def test_minoverlap(host, mo):
    check(host, 5, synth.generate_config(2, n=3_000).to_flat(), mo=mo)
