"""The C-ABI library loads and exports every symbol include/nfm.h declares,
with the argument counts the ctypes binding assumes.  No compute calls."""
import ctypes
import os
import re

import pytest

from conftest import ROOT
from nitorch_fastmath_b200 import _lib

HEADER = os.path.join(ROOT, "include", "nfm.h")


def _declarations():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    decls = {}
    for m in re.finditer(r"\b([A-Za-z_][\w \*]*?)\b(nfm_\w+)\s*\(([^)]*)\)\s*;", text):
        args = m.group(3).strip()
        nargs = 0 if args in ("", "void") else len(args.split(","))
        decls[m.group(2)] = nargs
    return decls


def test_library_is_built():
    assert os.path.isfile(_lib.LIB_PATH), "run `make -C nitorch_fastmath_b200/csrc -j8` or __graft_entry__.build()"


def test_every_declared_symbol_is_exported():
    decls = _declarations()
    assert len(decls) >= 18
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in decls:
        assert hasattr(lib, name), f"{name} declared in nfm.h but not exported"


def test_binding_matches_header():
    decls = _declarations()
    assert set(decls) == set(_lib.SIGNATURES), set(decls) ^ set(_lib.SIGNATURES)
    for name, nargs in decls.items():
        assert len(_lib.SIGNATURES[name][1]) == nargs, name


def test_version_and_error_string():
    lib = _lib.load()
    assert lib.nfm_version() == 100
    assert isinstance(lib.nfm_last_error_string(), bytes)


def test_argument_validation_without_a_gpu():
    """Bad arguments are rejected before any CUDA call."""
    lib = _lib.load()
    buf = ctypes.create_string_buffer(256)
    p = ctypes.addressof(buf)
    assert lib.nfm_sym_solve(_lib.F32, 11, _lib.LAYOUT_SYM, 0, 4, p, 66, p, 11, None, 0, p, 11, None) == -1
    assert lib.nfm_sym_solve(7, 3, _lib.LAYOUT_SYM, 0, 4, p, 6, p, 3, None, 0, p, 3, None) == -1
    assert lib.nfm_sym_solve(_lib.F32, 3, _lib.LAYOUT_SYM, 0, 4, None, 6, p, 3, None, 0, p, 3, None) == -2
    assert lib.nfm_sym_solve(_lib.F32, 3, _lib.LAYOUT_SYM, 0, -1, p, 6, p, 3, None, 0, p, 3, None) == -2
    assert lib.nfm_sym_matvec(_lib.F32, 3, _lib.LAYOUT_SYM, 4, p, 6, p, 3, p, 3, 0, p, 3, None) == -2
    assert b"sign" in lib.nfm_last_error_string()
    assert lib.nfm_batch_solve(_lib.F64, 4, 0, 0, 4, p, 16, p, 4, p, 4, None) == -2
    assert lib.nfm_sym_matmul(_lib.F32, 2, 3, 1, 4, p, 6, p, 3, p, 6, None) == -1
    assert lib.nfm_sym_matmul(_lib.F32, 11, 3, 0, 4, p, 33, p, 66, p, 6, None) == -1
    with pytest.raises(_lib.NfmError):
        _lib.check(-2, "demo")


def test_header_is_plain_c():
    """include/nfm.h is a C ABI: it must compile as C99 without torch / CUDA headers."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    r = subprocess.run(["gcc", "-x", "c", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-fsyntax-only", HEADER],
                       capture_output=True, text=True)
    assert r.returncode == 0 and r.stderr.strip() == "", r.stderr


def _build_c_caller(tmp_path):
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    if not os.path.isfile(os.path.join(cuda, "include", "cuda_runtime_api.h")):
        pytest.skip("no CUDA toolkit headers")
    exe = str(tmp_path / "abi_c_call")
    libdir = os.path.dirname(_lib.LIB_PATH)
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-O1", os.path.join(ROOT, "tests", "abi_c_call.c"), "-o", exe,
                        "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(cuda, "include"),
                        "-L" + libdir, "-l:libnfm_sm100a.so", "-L" + os.path.join(cuda, "lib64"), "-lcudart", "-lm",
                        "-Wl,-rpath," + libdir + ",-rpath," + os.path.join(cuda, "lib64")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_c_program_links_against_the_abi(tmp_path):
    """tests/abi_c_call.c (plain C99, no Python in between) compiles and links
    against include/nfm.h + libnfm_sm100a.so."""
    _build_c_caller(tmp_path)


@pytest.mark.gpu
def test_c_program_calls_the_abi(tmp_path):
    """... and, on the GPU box, runs: cudaMalloc buffers, nfm_sym_solve -> nfm_sym_matvec
    round trip checked on the host, error codes for bad arguments."""
    import subprocess
    exe = _build_c_caller(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.startswith("OK"), (r.returncode, r.stdout, r.stderr)
