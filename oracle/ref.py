"""ORACLE helper (test infrastructure): ctypes wrapper over oracle/ssa_ref.c.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference arm may
import this.  The product (passport_zk_circuits_b200) never does."""
import ctypes
import json
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
SO = os.path.join(_HERE, "_build", "libpzkref.so")


def build(force=False):
    src = os.path.join(_HERE, "ssa_ref.c")
    hdr = os.path.join(_ROOT, "include", "pzk_program.h")
    if not force and os.path.exists(SO) and os.path.getmtime(SO) >= max(os.path.getmtime(src), os.path.getmtime(hdr)):
        return SO
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-I" + os.path.join(_ROOT, "include"), "-o", SO, src])
    return SO


_lib = None


def _L():
    global _lib
    if _lib is None:
        if not os.path.exists(SO):
            build()
        L = ctypes.CDLL(SO)
        L.pzk_ref_load.restype = ctypes.c_void_p
        L.pzk_ref_load.argtypes = [ctypes.c_char_p]
        L.pzk_ref_free.argtypes = [ctypes.c_void_p]
        for f in ("pzk_ref_n_wires", "pzk_ref_n_inputs", "pzk_ref_n_constraints", "pzk_ref_n_outputs"):
            getattr(L, f).restype = ctypes.c_uint32
            getattr(L, f).argtypes = [ctypes.c_void_p]
        L.pzk_ref_meta.restype = ctypes.c_void_p
        L.pzk_ref_meta.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_uint64)]
        L.pzk_ref_witness.restype = ctypes.c_uint32
        L.pzk_ref_witness.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                      ctypes.POINTER(ctypes.c_int64), ctypes.c_int]
        _lib = L
    return _lib


class RefProgram:
    """One compiled program evaluated on the CPU, one witness per call."""

    def __init__(self, program_path):
        L = _L()
        self.h = L.pzk_ref_load(os.fsencode(program_path))
        if not self.h:
            raise RuntimeError("ssa_ref: cannot load " + program_path)
        self.n_wires = L.pzk_ref_n_wires(self.h)
        self.n_inputs = L.pzk_ref_n_inputs(self.h)
        self.n_constraints = L.pzk_ref_n_constraints(self.h)
        self.n_outputs = L.pzk_ref_n_outputs(self.h)
        ln = ctypes.c_uint64()
        p = L.pzk_ref_meta(self.h, ctypes.byref(ln))
        self.meta = json.loads(ctypes.string_at(p, ln.value).decode())

    def witness(self, inputs_u64, want_witness=True, check_rows=True):
        """inputs_u64: uint64 [n_inputs, 4].  Returns (status, first_bad, witness uint64 [n_wires, 4] | None)."""
        inp = np.ascontiguousarray(inputs_u64, dtype=np.uint64)
        assert inp.shape == (self.n_inputs, 4)
        wit = np.zeros((self.n_wires, 4), dtype=np.uint64) if want_witness else None
        fb = ctypes.c_int64()
        st = _L().pzk_ref_witness(self.h, inp.ctypes.data, wit.ctypes.data if want_witness else None,
                                  ctypes.byref(fb), 1 if check_rows else 0)
        return st, fb.value, wit

    def close(self):
        if self.h:
            _L().pzk_ref_free(self.h)
            self.h = None
