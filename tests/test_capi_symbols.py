"""CPU-side checks of the drop-in boundary: the shared library builds, loads, and exports every
symbol that include/lanczos_b200.h declares (no compute calls: there is no GPU here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "lanczos_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(lz_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    from lanczos_b200 import build as lzbuild
    path = lzbuild.build()
    return ctypes.CDLL(path)


def test_header_declares_the_expected_surface():
    syms = declared_symbols()
    for must in ("lz_ctx_create", "lz_op_stencil_create", "lz_op_csr_create", "lz_op_apply",
                 "lz_op_export_csr", "lz_lanczos_run", "lz_reorthogonalize", "lz_ritz_vectors"):
        assert must in syms


def test_library_exports_every_declared_symbol(lib):
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, missing


def test_binding_table_matches_header(lib):
    from lanczos_b200 import _capi
    assert sorted(_capi.SIGNATURES) == declared_symbols()


def _struct_fields(name):
    text = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (name, name), text, flags=re.S).group(1)
    return [(t, f) for t, f in re.findall(r"\b(int32_t|float|double)\s+(\w+)\s*;", body)]


def test_struct_layouts_match_header_and_integration_stub():
    """lz_run_opts / lz_run_info: the ctypes mirrors (and the stub shown in INTEGRATION.md) carry the
    header's fields in the header's order."""
    from lanczos_b200 import _capi
    ctype = {"int32_t": ctypes.c_int32, "float": ctypes.c_float, "double": ctypes.c_double}
    for cname, mirror in (("lz_run_opts", _capi.RunOpts), ("lz_run_info", _capi.RunInfo)):
        want = [(f, ctype[t]) for t, f in _struct_fields(cname)]
        assert [(f, t) for f, t in mirror._fields_] == want
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    for cname in ("lz_run_opts", "lz_run_info"):
        for _, f in _struct_fields(cname):
            assert '"%s"' % f in doc, f"INTEGRATION.md stub lacks field {f} of {cname}"


def test_abi_version_and_error_string(lib):
    lib.lz_abi_version.restype = ctypes.c_int
    assert lib.lz_abi_version() == 2
    lib.lz_last_error.restype = ctypes.c_char_p
    assert isinstance(lib.lz_last_error(), bytes)


def test_argument_validation_without_gpu(lib):
    # null-argument paths return LZ_ERR_INVALID before touching CUDA
    lib.lz_device_count.restype = ctypes.c_int
    assert lib.lz_device_count(None) == 1
    n = ctypes.c_int(-1)
    assert lib.lz_device_count(ctypes.byref(n)) == 0 and n.value >= 0
    lib.lz_op_rows.restype = ctypes.c_int
    assert lib.lz_op_rows(None, None) == 1
    assert b"null" in lib.lz_last_error()


def test_header_is_plain_c_and_links_from_c(lib, tmp_path):
    """examples/c_abi_demo.c: the header compiles as C99 with -Wall -Werror, the library links from C,
    and without a CUDA device the client stops after the boundary checks (no CPU path)."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    cuda_lib = "/usr/local/cuda/lib64"
    if not os.path.exists(os.path.join(cuda_lib, "libcudart.so")):
        pytest.skip("libcudart.so not found (the demo allocates device memory through the CUDA runtime)")
    exe = str(tmp_path / "c_abi_demo")
    libdir = os.path.join(ROOT, "lanczos_b200")
    cmd = ["gcc", "-std=c99", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"),
           os.path.join(ROOT, "examples", "c_abi_demo.c"), "-L" + libdir, "-llanczos_b200", "-L" + cuda_lib, "-lcudart",
           "-Wl,-rpath," + libdir, "-Wl,-rpath," + cuda_lib, "-o", exe]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    run = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert run.returncode == 0, run.stdout + run.stderr
    assert "lz_abi_version = 2" in run.stdout
    import torch
    if torch.cuda.is_available():
        assert "steps_done = 12" in run.stdout
    else:
        assert "no CUDA device" in run.stdout


def test_product_has_no_cpu_path():
    """Without a CUDA device the drop-in must fail loudly, never fall back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    import lanczos_b200 as lz
    import numpy as np
    op = lz.StencilOperator((8, 8), 4.0, -1.0)
    L = lz.Lanczos(op)
    with pytest.raises(RuntimeError, match="no CPU path"):
        L.execute_Lanczos(4)
    with pytest.raises(RuntimeError):
        op.matvec(np.ones(64))


def test_product_does_not_import_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "lanczos_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle|import_module\([\"']oracle", src, flags=re.M), f
