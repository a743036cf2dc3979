"""The C-ABI library loads without a GPU and exports every symbol include/clrsdp.h declares; the oracle
exports the same surface with the clrsdp_ref_ prefix; struct layouts match the ctypes mirror."""
import ctypes
import os
import re

import pytest

from clrsdp import capi
from oracle import ref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "clrsdp.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(clrsdp_[a-z_0-9]+)\s*\(", src)))


def test_header_declares_the_expected_entry_points():
    syms = declared_symbols()
    for must in ("clrsdp_create", "clrsdp_set_structure", "clrsdp_upload_cluster", "clrsdp_iterate", "clrsdp_solve",
                 "clrsdp_fetch", "clrsdp_op_gemm", "clrsdp_comm_init", "clrsdp_launch_count"):
        assert must in syms


def test_product_library_exports_every_declared_symbol():
    lib = capi.load_product_library()          # raises if the CUDA library was not built: no fallback
    for s in declared_symbols():
        assert hasattr(lib, s), s


def test_oracle_exports_the_solver_surface():
    lib = ctypes.CDLL(ref.build())
    for s in ("create", "destroy", "set_structure", "upload_cluster", "upload_objective", "set_params", "init_point",
              "upload_point", "download_point", "prepare", "iterate", "solve", "fetch", "op_gemm", "op_cholesky",
              "op_lambda_min", "op_elementwise"):
        assert hasattr(lib, "clrsdp_ref_" + s), s


def test_iter_info_layout():
    # 4 int32 + 16 doubles + 17 doubles
    assert ctypes.sizeof(capi.IterInfo) == 16 + 8 * 16 + 8 * 17
    assert capi.IterInfo.timings.offset == 16 + 8 * 16
    assert len(capi.TIMING_NAMES) == 17 and len(capi.SCALARS) == 13


def test_create_fails_loudly_without_a_gpu():
    import shutil
    import subprocess
    has_gpu = shutil.which("nvidia-smi") is not None and subprocess.run(["nvidia-smi", "-L"], capture_output=True).returncode == 0
    if has_gpu:
        pytest.skip("a GPU is present")
    with pytest.raises(capi.ClrsdpError) as e:
        capi.Handle(capi.load_product_library(), "clrsdp_", 256, 0)
    assert e.value.code == -2


def test_bad_precision_is_rejected():
    lib = capi.load_product_library()
    h = ctypes.c_void_p()
    f = lib.clrsdp_create
    f.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, ctypes.c_int]
    assert f(ctypes.byref(h), 100, 0) != 0
    assert f(None, 256, 0) == -1
