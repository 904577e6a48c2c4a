"""ctypes binding of the CPU oracle (oracle/libth_oracle.so) and a runner for oracle/_ref/TideHunter.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  The product (tidehunter_b200/, host/) never imports this module.
"""
import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libth_oracle.so")
REF_BIN = os.path.join(HERE, "_ref", "TideHunter")


class Para(C.Structure):
    _fields_ = [
        ("k", C.c_int), ("w", C.c_int), ("hpc", C.c_int),
        ("min_copy", C.c_int), ("min_cov", C.c_int),
        ("max_div", C.c_double), ("min_frac", C.c_double),
        ("min_p", C.c_int64), ("max_p", C.c_int64),
        ("match", C.c_int), ("mismatch", C.c_int), ("gap_open1", C.c_int), ("gap_open2", C.c_int),
        ("gap_ext1", C.c_int), ("gap_ext2", C.c_int),
        ("out_fmt", C.c_int), ("min_len", C.c_int), ("only_unit", C.c_int), ("only_longest", C.c_int),
        ("only_full_length", C.c_int), ("single_copy", C.c_int),
        ("ada_match_rat", C.c_float),
        ("five_seq", C.c_char_p), ("three_seq", C.c_char_p),
        ("pn16", C.c_int),
    ]


class Chain(C.Structure):
    _fields_ = [
        ("n_cells", C.c_int),
        ("start", C.POINTER(C.c_int)), ("end", C.POINTER(C.c_int)), ("score", C.POINTER(C.c_int)),
        ("from_", C.POINTER(C.c_int)), ("row", C.POINTER(C.c_int)),
        ("n_chain", C.c_int),
        ("chain_off", C.POINTER(C.c_int)), ("cells", C.POINTER(C.c_int)),
        ("est_start", C.POINTER(C.c_int)), ("est_period", C.POINTER(C.c_int)),
        ("n_evals", C.c_int64),
    ]


_lib = None


def build(force=False):
    """Compile the C restatement (and, when /root/reference is present, oracle/_ref)."""
    if force or not os.path.exists(LIB_PATH):
        subprocess.check_call(["make", "-s", "-C", HERE, "port"])
    if os.path.isdir("/root/reference") and not os.path.exists(REF_BIN):
        subprocess.check_call(["make", "-s", "-j8", "-C", HERE, "ref"])


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        L.tho_default_para.argtypes = [C.POINTER(Para)]
        L.tho_get_bseq.argtypes = [C.c_char_p, C.c_int, C.c_void_p]
        L.tho_collect_hits.argtypes = [C.c_void_p, C.c_int, C.POINTER(Para), C.POINTER(C.POINTER(C.c_uint64))]
        L.tho_collect_hits.restype = C.c_int
        L.tho_tandem_chain.argtypes = [C.c_void_p, C.c_int, C.POINTER(Para), C.POINTER(Chain)]
        L.tho_tandem_chain.restype = C.c_int
        L.tho_chain_free.argtypes = [C.POINTER(Chain)]
        L.tho_partition.argtypes = [C.c_void_p, C.c_int, C.POINTER(Chain), C.c_int, C.POINTER(Para), C.POINTER(C.POINTER(C.c_int))]
        L.tho_partition.restype = C.c_int
        L.tho_ksw2_global.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.POINTER(C.c_uint32))]
        L.tho_ksw2_global.restype = C.c_int
        L.tho_ksw2_ext.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.tho_ksw2_left_ext.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.tho_ksw2_backtrack_left_end.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.tho_ksw2_backtrack_left_end.restype = C.c_int
        L.tho_abpoa_cons.argtypes = [C.POINTER(Para), C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.c_void_p, C.c_void_p, C.POINTER(C.c_int64)]
        L.tho_abpoa_cons.restype = C.c_int
        L.tho_run_batch.argtypes = [C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_char_p), C.POINTER(C.c_int),
                                    C.POINTER(Para), C.c_int, C.POINTER(C.c_size_t), C.POINTER(C.c_int64)]
        L.tho_run_batch.restype = C.c_void_p
        L.tho_free.argtypes = [C.c_void_p]
        _lib = L
    return _lib


def default_para(**kw):
    p = Para()
    lib().tho_default_para(C.byref(p))
    for k, v in kw.items():
        if k in ("five_seq", "three_seq") and isinstance(v, str):
            v = v.encode()
        setattr(p, k, v)
    return p


def run_batch(names, seqs, para=None, threads=1):
    """Run the oracle over reads; returns (output_text_bytes, counters dict)."""
    L = lib()
    para = para or default_para()
    n = len(seqs)
    bn = [x if isinstance(x, bytes) else x.encode() for x in names]
    bs = [x if isinstance(x, bytes) else x.encode() for x in seqs]
    names_a = (C.c_char_p * n)(*bn)
    seqs_a = (C.c_char_p * n)(*bs)
    lens_a = (C.c_int * n)(*[len(x) for x in bs])
    out_len = C.c_size_t(0)
    cnt = (C.c_int64 * 4)()
    ptr = L.tho_run_batch(n, names_a, seqs_a, lens_a, C.byref(para), threads, C.byref(out_len), cnt)
    text = C.string_at(ptr, out_len.value)
    L.tho_free(ptr)
    return text, {"hits": cnt[0], "chain_evals": cnt[1], "poa_cells": cnt[2], "ksw_cells": cnt[3]}


def read_fastx(path):
    """Minimal FASTA/FASTQ reader with kseq semantics: name = up to first whitespace."""
    names, seqs = [], []
    with open(path, "rb") as f:
        lines = f.read().split(b"\n")
    i = 0
    while i < len(lines):
        ln = lines[i]
        if ln.startswith(b">"):
            names.append(ln[1:].split()[0] if ln[1:].split() else b"")
            i += 1
            s = []
            while i < len(lines) and not lines[i].startswith(b">"):
                s.append(lines[i].strip())
                i += 1
            seqs.append(b"".join(s))
        elif ln.startswith(b"@"):
            names.append(ln[1:].split()[0] if ln[1:].split() else b"")
            seqs.append(lines[i + 1].strip())
            i += 4
        else:
            i += 1
    return names, seqs


def write_fasta(path, names, seqs):
    with open(path, "wb") as f:
        for n, s in zip(names, seqs):
            f.write(b">" + (n if isinstance(n, bytes) else n.encode()) + b"\n")
            f.write((s if isinstance(s, bytes) else s.encode()) + b"\n")


def run_ref(path, args=(), threads=1):
    """Run the unmodified reference binary (oracle/_ref/TideHunter) on a FASTA/FASTQ file."""
    cmd = [REF_BIN, "-t", str(threads)] + list(args) + [path]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, check=True)
    return r.stdout
