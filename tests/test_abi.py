"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads, exports every symbol
include/th_gpu.h declares, and fails loudly (no CPU fallback) when no CUDA device is present."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header):
    txt = open(os.path.join(ROOT, header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(th_(?:gpu|host)_[a-z0-9_]+)\s*\(", txt)))


def test_gpu_abi_exports_every_declared_symbol():
    import tidehunter_b200 as T
    g = T.gpu_lib()
    names = _declared("include/th_gpu.h")
    assert len(names) >= 10
    for n in names:
        assert hasattr(g, n), n


def test_host_layer_exports_every_declared_symbol():
    import tidehunter_b200 as T
    h = T.host_lib()
    for n in _declared("host/th_host.h"):
        assert hasattr(h, n), n


def test_abi_structs_match_header_sizes():
    import tidehunter_b200 as T
    # th_gpu_params: 4 x i32, f64, 2 x i64, 6 x i32, 3 x i32 -> 16 + 8 + 16 + 36 (+4 pad) = 80
    assert C.sizeof(T.GpuParams) == 80
    assert C.sizeof(T.GpuStats) == 10 * 4 + 11 * 8
    p = T.GpuParams()
    T.gpu_lib().th_gpu_default_params(C.byref(p))
    assert (p.k, p.w, p.min_copy, p.min_p, p.max_p, p.simd_lanes16) == (8, 1, 2, 30, 10000, 16)
    assert (p.match, p.mismatch, p.gap_open1, p.gap_ext1, p.gap_open2, p.gap_ext2) == (2, 4, 4, 2, 24, 1)
    assert p.max_div == 0.25


def test_no_cpu_fallback():
    import tidehunter_b200 as T
    g = T.gpu_lib()
    if g.th_gpu_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(RuntimeError):
        T.GpuContext()
    with pytest.raises(RuntimeError):
        T.TideHunter()
    assert b"CUDA" in g.th_gpu_last_error()


def test_product_never_touches_the_oracle():
    bad = []
    for d in ("tidehunter_b200", "host", "include"):
        for dp, _, fs in os.walk(os.path.join(ROOT, d)):
            for f in fs:
                if f.endswith((".py", ".c", ".h", ".cu", ".cuh")):
                    txt = open(os.path.join(dp, f), errors="ignore").read()
                    if re.search(r"oracle_py|th_oracle|libth_oracle|oracle/", txt):
                        bad.append(os.path.join(dp, f))
    assert not bad, bad


def test_adapter_search_bitvector_equals_definition():
    """host/th_host.c: the Myers bit-vector infix search (adapter detection of -5/-3/-F/-s, restating edlib_align_HW)
    against the plain column DP on 30,000 random adapter/text pairs (1-300 bp adapters incl. the > 256 bp fallback,
    mixed case, planted noisy copies, thresholds).  Host code only, no GPU needed."""
    import ctypes as C
    import tidehunter_b200 as T
    h = T.host_lib()
    h.th_host_selftest_infix.argtypes = [C.c_int, C.c_uint]
    h.th_host_selftest_infix.restype = C.c_int
    assert h.th_host_selftest_infix(30000, 20260117) == 0


def test_options_reach_the_kernels():
    """th_gpu_debug_dev_params (no device needed): every option lands in the kernels' parameter block -- vector width and
    only_unit included (an edit once dropped them; only the GPU run noticed) -- and the affine and linear gap modes are flagged."""
    import ctypes as C
    import tidehunter_b200 as T
    g = T.gpu_lib()
    g.th_gpu_debug_dev_params.argtypes = [C.POINTER(T.GpuParams), C.c_int32, C.POINTER(C.c_int32)]
    g.th_gpu_debug_dev_params.restype = C.c_int
    names = ["k", "w", "hpc", "min_copy", "min_p", "max_p", "max_div_e6", "match", "mismatch", "o1", "e1", "o2", "e2", "affine", "o2_raw", "e2_raw", "pn", "only_unit", "linear"]

    def dev(**kw):
        p = T.GpuParams()
        g.th_gpu_default_params(C.byref(p))
        for k, v in kw.items():
            setattr(p, k, v)
        out = (C.c_int32 * 32)()
        n = g.th_gpu_debug_dev_params(C.byref(p), 32, out)
        assert n == len(names)
        return dict(zip(names, out[:n]))
    d = dev()
    assert d == dict(k=8, w=1, hpc=0, min_copy=2, min_p=30, max_p=10000, max_div_e6=250000, match=2, mismatch=4, o1=4, e1=2, o2=24, e2=1,
                     affine=0, o2_raw=24, e2_raw=1, pn=16, only_unit=0, linear=0)
    d = dev(k=12, w=5, hpc=1, min_copy=3, min_p=10, max_p=3000, max_div=0.1, simd_lanes16=8, only_unit=1, gap_open1=6, gap_ext1=3, gap_open2=40, gap_ext2=2)
    assert (d["k"], d["w"], d["hpc"], d["min_copy"], d["min_p"], d["max_p"], d["max_div_e6"], d["pn"], d["only_unit"]) == (12, 5, 1, 3, 10, 3000, 100000, 8, 1)
    assert (d["o1"], d["e1"], d["o2"], d["e2"], d["affine"]) == (6, 3, 40, 2, 0)
    d = dev(gap_open2=0)
    assert d["affine"] == 1 and d["linear"] == 0 and (d["o2_raw"], d["e2_raw"]) == (0, 1) and d["pn"] == 16
    d = dev(gap_open1=0, gap_open2=0)   # abpoa_set_gap_mode: linear wins
    assert d["linear"] == 1 and d["affine"] == 0 and d["o1"] == 0
