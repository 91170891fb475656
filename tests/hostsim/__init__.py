"""Host simulation of the device headers (test infrastructure)."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SO = os.path.join(HERE, "libcore_host.so")
CSRC = os.path.join(ROOT, "ac_solver_b200", "csrc")


def build():
    srcs = [os.path.join(HERE, "core_host.cpp"), os.path.join(CSRC, "ac_core.cuh"), os.path.join(CSRC, "ac_pack.cuh")]
    if not os.path.exists(SO) or any(os.path.getmtime(s) > os.path.getmtime(SO) for s in srcs):
        subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-x", "c++",
                               "-I/usr/local/cuda/include", "-I" + CSRC, srcs[0], "-o", SO])
    return SO


_lib = None


def moves(states, actions, cyclical=True, trusted=False, use_words=True):
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.hostsim_moves.argtypes = [C.c_void_p] * 5 + [C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int]
    s = np.ascontiguousarray(states, np.int8)
    a = np.ascontiguousarray(actions, np.uint8)
    n, w = s.shape
    out = np.empty_like(s)
    lens = np.zeros((n, 2), np.uint8)
    status = np.zeros(n, np.uint8)
    rc = _lib.hostsim_moves(s.ctypes.data, a.ctypes.data, out.ctypes.data, lens.ctypes.data, status.ctypes.data,
                            n, w // 2, int(cyclical), int(trusted), int(use_words))
    assert rc == 0
    return out, lens, status
