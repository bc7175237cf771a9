"""Host-side mirror of the reference's algorithm objects (pandaseq-algorithm.h): registry, parameters, per-base posterior."""
import ctypes as C
import math

import numpy as np

import pandaseq_b200 as pb


class Qual(C.Structure):
    _fields_ = [("nt", C.c_char), ("qual", C.c_char)]


def L():
    lib = pb.lib()
    for name in ("panda_algorithm_simple_bayes_new", "panda_algorithm_pear_new", "panda_algorithm_rdp_mle_new", "panda_algorithm_flash_new",
                 "panda_algorithm_ref", "panda_algorithm_class"):
        getattr(lib, name).restype = C.c_void_p
    lib.panda_algorithm_unref.argtypes = [C.c_void_p]
    lib.panda_algorithm_ref.argtypes = [C.c_void_p]
    lib.panda_algorithm_class.argtypes = [C.c_void_p]
    lib.panda_algorithm_is_a.argtypes = [C.c_void_p, C.c_void_p]
    lib.panda_algorithm_is_a.restype = C.c_bool
    lib.panda_algorithm_quality_compare.argtypes = [C.c_void_p, C.POINTER(Qual), C.POINTER(Qual)]
    lib.panda_algorithm_quality_compare.restype = C.c_double
    for name in ("panda_algorithm_simple_bayes_get_error_estimation", "panda_algorithm_pear_get_random_base_log_p"):
        getattr(lib, name).argtypes = [C.c_void_p]
        getattr(lib, name).restype = C.c_double
    lib.panda_algorithm_simple_bayes_set_error_estimation.argtypes = [C.c_void_p, C.c_double]
    lib.panda_algorithm_pear_set_random_base_log_p.argtypes = [C.c_void_p, C.c_double]
    return lib


def test_registry_sorted_by_name(built):
    lib = L()
    n = C.c_size_t.in_dll(lib, "panda_algorithms_length").value
    arr = C.POINTER(C.c_void_p).in_dll(lib, "panda_algorithms")

    class Klass(C.Structure):
        _fields_ = [("data_size", C.c_size_t), ("name", C.c_char_p), ("create", C.c_void_p), ("destroy", C.c_void_p),
                    ("overlap", C.c_void_p), ("match", C.c_void_p), ("prob_unpaired", C.c_double)]
    names = [C.cast(arr[i], C.POINTER(Klass)).contents.name.decode() for i in range(n)]
    assert names == sorted(names) == ["ea_util", "flash", "pear", "rdp_mle", "simple_bayesian", "stitch", "uparse"]   # algo.c:125-131
    for i in range(n):
        assert C.cast(arr[i], C.POINTER(Klass)).contents.prob_unpaired == -1.38629


def test_parameters_and_class_identity(built):
    lib = L()
    sb, pear = lib.panda_algorithm_simple_bayes_new(), lib.panda_algorithm_pear_new()
    sb_class = C.addressof(C.c_char.in_dll(lib, "panda_algorithm_simple_bayes_class"))
    assert lib.panda_algorithm_is_a(sb, sb_class) and not lib.panda_algorithm_is_a(pear, sb_class)
    assert lib.panda_algorithm_class(sb) == sb_class
    assert lib.panda_algorithm_simple_bayes_get_error_estimation(sb) == 0.36
    lib.panda_algorithm_simple_bayes_set_error_estimation(sb, 1.5)           # rejected: not in (0,1)
    assert lib.panda_algorithm_simple_bayes_get_error_estimation(sb) == 0.36
    lib.panda_algorithm_simple_bayes_set_error_estimation(sb, 0.2)
    assert lib.panda_algorithm_simple_bayes_get_error_estimation(sb) == 0.2
    assert lib.panda_algorithm_simple_bayes_get_error_estimation(pear) == -1  # wrong class
    assert lib.panda_algorithm_pear_get_random_base_log_p(pear) == math.log(0.25)
    assert lib.panda_algorithm_pear_get_random_base_log_p(sb) == 1
    lib.panda_algorithm_unref(lib.panda_algorithm_ref(sb))
    lib.panda_algorithm_unref(sb)
    lib.panda_algorithm_unref(pear)


def test_quality_compare_matches_tables(built):
    lib, t = L(), pb.tables()
    algos = dict(sb=lib.panda_algorithm_simple_bayes_new(), pear=lib.panda_algorithm_pear_new(),
                 rdp=lib.panda_algorithm_rdp_mle_new(), flash=lib.panda_algorithm_flash_new())

    def cmp(a, nt1, q1, nt2, q2):
        x, y = Qual(bytes([nt1]), bytes([q1 & 0xFF])), Qual(bytes([nt2]), bytes([q2 & 0xFF]))
        return lib.panda_algorithm_quality_compare(algos[a], C.byref(x), C.byref(y))

    assert cmp("sb", 1, 30, 1, 20) == t["match_sb"][30][20]
    assert cmp("sb", 1, 30, 2, 20) == t["mismatch_sb"][30][20]
    assert cmp("sb", 1, 100, 1, -5) == t["match_sb"][46][0]                  # PHREDCLAMP
    assert cmp("pear", 4, 11, 8, 40) == t["mismatch_pear"][11][40]
    assert cmp("rdp", 1, 12, 1, 33) == t["score"][33]                          # match: the higher of the two
    assert cmp("rdp", 1, 12, 2, 33) == t["mismatch_rdp_asm"][12][33]
    assert cmp("flash", 8, 40, 8, 10) == t["score"][40]
    assert cmp("flash", 8, 40, 4, 39) == t["score"][2]                         # |40-39| < 2 -> 2
    assert cmp("flash", 8, 10, 4, 40) == t["score"][30]
    assert cmp("sb", 15, 30, 2, 30) == t["match_sb"][30][30]                   # N matches everything
    for a in algos.values():
        lib.panda_algorithm_unref(a)


def test_layout_helper_matches_python(built):
    lib = pb.lib()
    rng = np.random.default_rng(1)
    fl, rl = rng.integers(0, 451, 500), rng.integers(0, 451, 500)
    f_off = np.concatenate([[0], np.cumsum(fl)]).astype(np.uint64)
    r_off = np.concatenate([[0], np.cumsum(rl)]).astype(np.uint64)
    rec = np.zeros(500, np.uint32)
    total = lib.pb_layout_host(500, f_off.ctypes.data, r_off.ctypes.data, rec.ctypes.data)
    sizes = pb.record_bytes(fl, rl)
    assert total == sizes.sum()
    assert np.array_equal(rec.astype(np.int64) * 16, np.concatenate([[0], np.cumsum(sizes)[:-1]]))
    assert (sizes % 16 == 0).all() and sizes[(fl == 150) & (rl == 150)].tolist() in ([], [464] * int(((fl == 150) & (rl == 150)).sum()))
    assert pb.record_bytes(150, 150) == 464 and pb.record_bytes(0, 0) == 0


def test_module_objects_without_a_device(built):
    """panda_module_new / ref / unref / get_name / get_api (pandaseq-module.h:47-57, 68-79, 83-113): plain host objects; the cleanup
    function runs exactly once, when the last reference goes."""
    lib = pb.lib()
    vp = C.c_void_p
    lib.panda_module_new.restype, lib.panda_module_new.argtypes = vp, [C.c_char_p, vp, vp, vp, vp]
    lib.panda_module_ref.restype, lib.panda_module_ref.argtypes = vp, [vp]
    lib.panda_module_unref.argtypes = [vp]
    lib.panda_module_get_name.restype, lib.panda_module_get_name.argtypes = C.c_char_p, [vp]
    lib.panda_module_get_api.restype, lib.panda_module_get_api.argtypes = C.c_int, [vp]
    CHECK = C.CFUNCTYPE(C.c_bool, vp, vp, vp)
    DESTROY = C.CFUNCTYPE(None, vp)
    cleaned = []
    chk, destroy = CHECK(lambda logger, seq, user: True), DESTROY(lambda user: cleaned.append(user))
    assert not lib.panda_module_new(b"no callbacks", None, None, None, None)          # module.c:258-260
    assert not lib.panda_module_new(None, C.cast(chk, vp), None, None, None)
    m = lib.panda_module_new(b"keeps everything", C.cast(chk, vp), None, 1234, C.cast(destroy, vp))
    assert m and lib.panda_module_get_name(m) == b"keeps everything" and lib.panda_module_get_api(m) == 3
    assert lib.panda_module_ref(m) == m
    lib.panda_module_unref(m)
    assert cleaned == []
    lib.panda_module_unref(m)
    assert cleaned == [1234]
    lib.panda_module_unref(None)
