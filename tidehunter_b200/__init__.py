"""tidehunter_b200 -- B200-native replacement for TideHunter's per-read hot path.

Python is only the thin binding used by the tests and bench.py: every call goes through the C ABI
(include/th_gpu.h) or the host C layer above it (host/th_host.h).  There is no CPU fallback; creating
a context without the CUDA library or without a GPU raises.
"""
import ctypes as C
import os

from . import build as _build

PKG = os.path.dirname(os.path.abspath(__file__))


class GpuParams(C.Structure):
    """th_gpu_params (include/th_gpu.h) = numeric fields of mini_tandem_para (src/tidehunter.h:47-61)."""
    _fields_ = [
        ("k", C.c_int32), ("w", C.c_int32), ("hpc", C.c_int32), ("min_copy", C.c_int32),
        ("max_div", C.c_double), ("min_p", C.c_int64), ("max_p", C.c_int64),
        ("match", C.c_int32), ("mismatch", C.c_int32), ("gap_open1", C.c_int32), ("gap_open2", C.c_int32),
        ("gap_ext1", C.c_int32), ("gap_ext2", C.c_int32),
        ("only_unit", C.c_int32), ("need_cov", C.c_int32), ("simd_lanes16", C.c_int32),
    ]


class GpuStats(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("ms_h2d", "ms_pack", "ms_seed", "ms_chain", "ms_select", "ms_partition",
                                         "ms_poa", "ms_ksw", "ms_d2h", "ms_total")] + \
               [(n, C.c_int64) for n in ("n_bases", "n_hits", "n_chain_evals", "n_poa_cells", "n_poa_rows",
                                         "n_ksw_cells", "n_tasks", "n_launches", "h2d_bytes", "d2h_bytes", "n_ksw_cells_full")]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class GpuResult(C.Structure):
    _fields_ = [
        ("n_reads", C.c_int32), ("n_tasks", C.c_int32),
        ("read_task_off", C.POINTER(C.c_int32)), ("task_pos_off", C.POINTER(C.c_int32)), ("pos", C.POINTER(C.c_int32)),
        ("task_n_seqs", C.POINTER(C.c_int32)), ("task_cons_off", C.POINTER(C.c_int32)), ("cons_base", C.POINTER(C.c_uint8)),
        ("cons_cov", C.POINTER(C.c_int32)), ("iden_n", C.POINTER(C.c_int32)), ("ext", C.POINTER(C.c_int32)),
        ("task_status", C.POINTER(C.c_int32)), ("read_status", C.POINTER(C.c_int32)), ("stats", GpuStats),
    ]


class HostPara(C.Structure):
    """th_host_para (host/th_host.h)."""
    _fields_ = [
        ("gpu", GpuParams), ("out_fmt", C.c_int), ("min_len", C.c_int), ("min_cov", C.c_int), ("min_frac", C.c_double),
        ("only_longest", C.c_int), ("only_full_length", C.c_int), ("single_copy", C.c_int), ("ada_match_rat", C.c_float),
        ("five_seq", C.c_char_p), ("three_seq", C.c_char_p), ("chunk_reads", C.c_int), ("lanes", C.c_int),
    ]


_gpu = None
_host = None


def _load():
    global _gpu, _host
    if _gpu is None:
        gpu_so, host_so, _ = _build.build()
        if not os.path.exists(gpu_so):
            raise RuntimeError("libth_gpu.so is missing: the CUDA extension must be built (no CPU fallback)")
        _gpu = C.CDLL(gpu_so, mode=C.RTLD_GLOBAL)
        _host = C.CDLL(host_so)
        g, h = _gpu, _host
        g.th_gpu_default_params.argtypes = [C.POINTER(GpuParams)]
        g.th_gpu_create.argtypes = [C.POINTER(GpuParams), C.c_int]
        g.th_gpu_create.restype = C.c_void_p
        g.th_gpu_destroy.argtypes = [C.c_void_p]
        g.th_gpu_process_chunk.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_char_p), C.POINTER(C.c_int32), C.POINTER(GpuResult)]
        g.th_gpu_upload.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_char_p), C.POINTER(C.c_int32)]
        g.th_gpu_process_resident.argtypes = [C.c_void_p, C.POINTER(GpuResult)]
        g.th_gpu_debug_hits.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
        g.th_gpu_debug_chain_dp.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
        g.th_gpu_debug_chains.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
        g.th_gpu_debug_par_pos.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_int32)]
        g.th_gpu_ksw_batch.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_void_p), C.POINTER(C.c_int32),
                                       C.POINTER(C.c_void_p), C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
        g.th_gpu_mark.argtypes = [C.c_void_p, C.c_int32]
        g.th_gpu_mark_elapsed.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.POINTER(C.c_float)]
        g.th_gpu_debug_counters.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_int64)]
        g.th_gpu_last_error.restype = C.c_char_p
        g.th_gpu_device_count.restype = C.c_int
        h.th_host_default_para.argtypes = [C.POINTER(HostPara)]
        h.th_host_create.argtypes = [C.POINTER(HostPara), C.c_int]
        h.th_host_create.restype = C.c_void_p
        h.th_host_create_multi.argtypes = [C.POINTER(HostPara), C.c_int, C.POINTER(C.c_int)]
        h.th_host_create_multi.restype = C.c_void_p
        h.th_host_destroy.argtypes = [C.c_void_p]
        h.th_host_run.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_char_p), C.POINTER(C.c_int32), C.POINTER(C.c_size_t)]
        h.th_host_run.restype = C.c_void_p
        h.th_host_stats.argtypes = [C.c_void_p, C.POINTER(GpuStats)]
        h.th_host_gpu.argtypes = [C.c_void_p]
        h.th_host_gpu.restype = C.c_void_p
        h.th_host_last_error.restype = C.c_char_p
    return _gpu, _host


def gpu_lib():
    return _load()[0]


def host_lib():
    return _load()[1]


def default_host_para(**kw):
    """Reference defaults (mini_tandem_init_para, src/main.c:325-362); keyword overrides use the
    reference's field names (k, w, hpc, min_copy, max_div, min_p, max_p, match, ..., out_fmt, min_len, ...)."""
    _, h = _load()
    p = HostPara()
    h.th_host_default_para(C.byref(p))
    gpu_fields = {n for n, _ in GpuParams._fields_}
    for k, v in kw.items():
        if k in ("five_seq", "three_seq") and isinstance(v, str):
            v = v.encode()
        if k in gpu_fields:
            setattr(p.gpu, k, v)
        else:
            setattr(p, k, v)
    return p


def _arrays(seqs):
    bs = [s if isinstance(s, bytes) else s.encode() for s in seqs]
    n = len(bs)
    return bs, (C.c_char_p * n)(*bs), (C.c_int32 * n)(*[len(s) for s in bs])


class Batch:
    """Reads held the way the C entry point takes them (th_host_run: arrays of name / sequence pointers and lengths), built
    once from Python lists.  The command line front end gets these arrays straight from its reader; a Python caller that
    processes the same lists more than once (bench.py) builds them once instead of on every call."""

    def __init__(self, names, seqs):
        self.names = [x if isinstance(x, bytes) else x.encode() for x in names]
        self.seqs, self.seqs_a, self.lens_a = _arrays(seqs)
        self.names_a = (C.c_char_p * len(self.names))(*self.names)
        self.n = len(self.seqs)


class TideHunter:
    """Drop-in for the reference's per-chunk loop (src/main.c:402-425): reads in, output text out."""

    def __init__(self, device=0, devices=None, **para):
        """`devices`: list of CUDA devices driven by this one object (`lanes` contexts on each); default [device]."""
        _, h = _load()
        self.para = default_host_para(**para)
        devs = list(devices) if devices else [device]
        self._h = h.th_host_create_multi(C.byref(self.para), len(devs), (C.c_int * len(devs))(*devs))
        if not self._h:
            raise RuntimeError("th_host_create failed: %s" % h.th_host_last_error().decode())

    def run(self, names, seqs=None, first_index=None, copy=True):
        """Records of these reads as the reference prints them.  `names` may be a Batch (then `seqs` is not given).
        `first_index`: index of the first read in the whole input when this object only sees part of it (a rank of a sharded
        run): the reference's FASTQ quality slot is index % 4096.  copy=False returns a memoryview of the library's own
        output buffer (valid until the next call) instead of a bytes copy."""
        _, h = _load()
        if first_index is not None:
            h.th_host_set_read_index.argtypes = [C.c_void_p, C.c_longlong]
            h.th_host_set_read_index.restype = None
            h.th_host_set_read_index(self._h, int(first_index))
        b = names if isinstance(names, Batch) else Batch(names, seqs)
        out_len = C.c_size_t(0)
        ptr = h.th_host_run(self._h, b.n, b.names_a, b.seqs_a, b.lens_a, C.byref(out_len))
        if not ptr:
            raise RuntimeError("th_host_run failed: %s" % h.th_host_last_error().decode())
        if copy:
            return C.string_at(ptr, out_len.value)
        return memoryview((C.c_char * out_len.value).from_address(ptr)).cast("B") if out_len.value else memoryview(b"")

    def stats(self):
        s = GpuStats()
        host_lib().th_host_stats(self._h, C.byref(s))
        return s.as_dict()

    def failed_tasks(self):
        """Consensus tasks the GPU path reported as failed since construction (their records are missing; stderr says why)."""
        h = host_lib()
        h.th_host_failed_tasks.argtypes = [C.c_void_p]
        h.th_host_failed_tasks.restype = C.c_longlong
        return int(h.th_host_failed_tasks(self._h))

    @property
    def gpu_ctx(self):
        return host_lib().th_host_gpu(self._h)

    def close(self):
        if self._h:
            host_lib().th_host_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class GpuContext:
    """Raw C-ABI context (include/th_gpu.h) for the stage-level parity tests and the device-only bench leg."""

    def __init__(self, device=0, **para):
        g, _ = _load()
        self.params = GpuParams()
        g.th_gpu_default_params(C.byref(self.params))
        for k, v in para.items():
            setattr(self.params, k, v)
        self._c = g.th_gpu_create(C.byref(self.params), device)
        if not self._c:
            raise RuntimeError("th_gpu_create failed: %s" % g.th_gpu_last_error().decode())
        self._keep = None

    def _err(self):
        return gpu_lib().th_gpu_last_error().decode()

    def process(self, seqs):
        bs, seqs_a, lens_a = _arrays(seqs)
        self._keep = (bs, seqs_a, lens_a)
        r = GpuResult()
        if gpu_lib().th_gpu_process_chunk(self._c, len(bs), seqs_a, lens_a, C.byref(r)):
            raise RuntimeError(self._err())
        return r

    def upload(self, seqs):
        bs, seqs_a, lens_a = _arrays(seqs)
        self._keep = (bs, seqs_a, lens_a)
        if gpu_lib().th_gpu_upload(self._c, len(bs), seqs_a, lens_a):
            raise RuntimeError(self._err())

    def process_resident(self):
        r = GpuResult()
        if gpu_lib().th_gpu_process_resident(self._c, C.byref(r)):
            raise RuntimeError(self._err())
        return r

    def mark(self, slot):
        if gpu_lib().th_gpu_mark(self._c, slot):
            raise RuntimeError(self._err())

    def elapsed_ms(self, slot, other, other_slot):
        """device time from this context's mark `slot` to `other`'s mark `other_slot`"""
        ms = C.c_float(0)
        if gpu_lib().th_gpu_mark_elapsed(self._c, slot, other._c, other_slot, C.byref(ms)):
            raise RuntimeError(self._err())
        return ms.value

    def hits(self, read, cap):
        e = (C.c_int32 * cap)(); p = (C.c_int32 * cap)()
        n = gpu_lib().th_gpu_debug_hits(self._c, read, cap, e, p)
        return n, list(e[:max(n, 0)]), list(p[:max(n, 0)])

    def chain_dp(self, read, cap):
        s = (C.c_int32 * cap)(); f = (C.c_int32 * cap)()
        n = gpu_lib().th_gpu_debug_chain_dp(self._c, read, cap, s, f)
        return n, list(s[:max(n, 0)]), list(f[:max(n, 0)])

    def chains(self, read, cap):
        nch = C.c_int32(0); cl = (C.c_int32 * 1024)(); cells = (C.c_int32 * cap)()
        tot = gpu_lib().th_gpu_debug_chains(self._c, read, cap, C.byref(nch), cl, cells)
        out, o = [], 0
        for i in range(nch.value):
            out.append(list(cells[o:o + cl[i]])); o += cl[i]
        return out

    def par_pos(self, read, chain, cap=65536):
        a = (C.c_int32 * cap)()
        n = gpu_lib().th_gpu_debug_par_pos(self._c, read, chain, cap, a)
        return None if n < 0 else list(a[:n])

    def counters(self):
        a = (C.c_int64 * 32)()
        n = gpu_lib().th_gpu_debug_counters(self._c, 32, a)
        if n < 0:
            raise RuntimeError(self._err())
        return list(a[:n])

    def ksw_batch(self, mode, qs, ts, args=None):
        n = len(qs)
        qp = (C.c_void_p * n)(*[C.cast(C.c_char_p(q), C.c_void_p) for q in qs])
        tp = (C.c_void_p * n)(*[C.cast(C.c_char_p(t), C.c_void_p) for t in ts])
        ql = (C.c_int32 * n)(*[len(q) for q in qs]); tl = (C.c_int32 * n)(*[len(t) for t in ts])
        ar = (C.c_int32 * n)(*(args if args is not None else [0] * n))
        out = (C.c_int32 * (2 * n))()
        self._keep2 = (qs, ts)
        if gpu_lib().th_gpu_ksw_batch(self._c, n, mode, qp, ql, tp, tl, ar, out):
            raise RuntimeError(self._err())
        return [(out[2 * i], out[2 * i + 1]) for i in range(n)]

    def close(self):
        if self._c:
            gpu_lib().th_gpu_destroy(self._c)
            self._c = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
