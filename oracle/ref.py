"""TEST INFRASTRUCTURE ONLY: handle on the CPU oracle (oracle/libclrsdp_ref.so).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
import ctypes
import os
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
_PKG = os.path.join(os.path.dirname(_HERE), "clustered-low-rank-sdp-solver_b200")
if _PKG not in sys.path:
    sys.path.insert(0, _PKG)

from clrsdp import capi  # noqa: E402

LIB = os.path.join(_HERE, "libclrsdp_ref.so")


def build(force=False):
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(os.path.join(_HERE, "clrsdp_ref.cpp")):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return LIB


_lib = None


def oracle_handle(prec=256, nthreads=0) -> capi.Handle:
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
    return capi.Handle(_lib, "clrsdp_ref_", prec, nthreads)
