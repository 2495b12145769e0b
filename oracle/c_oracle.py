"""ctypes wrapper + build recipe for oracle/dsk_oracle.c — TEST INFRASTRUCTURE ONLY (see the C header)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(_HERE, "dsk_oracle.c")
OUT_DIR = os.path.join(_HERE, "_build")
LIB = os.path.join(OUT_DIR, "libdsk_oracle.so")
_lib = None


def build(force=False):
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    os.makedirs(OUT_DIR, exist_ok=True)
    cmd = ["gcc", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", LIB, SRC, "-lm"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode:
        raise RuntimeError("gcc failed: " + r.stderr)
    return LIB


def load():
    global _lib
    if _lib is None:
        try:
            build()
        except Exception:
            if not os.path.exists(LIB):
                raise
        _lib = ctypes.CDLL(LIB)
    return _lib


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def pairwise_distance(x1, x2):
    x1, x2 = _f(x1), _f(x2)
    B, D = x1.shape
    out = np.empty(B, np.float32)
    load().orc_pairwise_distance(_p(x1), _p(x2), B, D, _p(out))
    return out


def triplet_loss(a, p, n, margin):
    a, p, n = _f(a), _f(p), _f(n)
    B, D = a.shape
    loss = np.empty(1, np.float32)
    d_p = np.empty(B, np.float32)
    d_n = np.empty(B, np.float32)
    load().orc_triplet_loss(_p(a), _p(p), _p(n), B, D, ctypes.c_float(margin), _p(loss), _p(d_p), _p(d_n))
    return loss[0], d_p, d_n


def margin_select(d_p, d_n, margin):
    d_p, d_n = _f(d_p), _f(d_n)
    idx = np.empty(len(d_p), np.int64)
    lib = load()
    lib.orc_margin_select.restype = ctypes.c_int
    k = lib.orc_margin_select(_p(d_p), _p(d_n), len(d_p), ctypes.c_float(margin), _p(idx))
    return idx[:k].copy()


def allpairs_topk(E, labels, k):
    E = _f(E)
    labels = np.ascontiguousarray(labels, dtype=np.int64)
    N, D = E.shape
    idx = np.empty((N, k), np.int64)
    val = np.empty((N, k), np.float32)
    load().orc_allpairs_topk(_p(E), _p(labels), N, D, k, _p(idx), _p(val))
    return idx, val
