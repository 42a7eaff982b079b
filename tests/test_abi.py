"""The C-ABI library loads, exports every symbol include/*.h declares, and refuses to compute
without a GPU (there is no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = set()
    inc = os.path.join(ROOT, "include")
    for fn in sorted(os.listdir(inc)):
        if not fn.endswith(".h"):
            continue
        text = open(os.path.join(inc, fn)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        for m in re.finditer(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\([^;{}]*\)\s*;", text):
            if m.group(1) not in ("defined",):
                names.add((fn, m.group(1)))
    return names


def test_every_declared_symbol_is_exported(pk):
    libs = {"b200_seqaij.h": pk.lib, "b200_mpiaij.h": pk.lib}
    host = os.path.join(ROOT, "petsc-openacc_b200", "libb200petsc.so")
    if os.path.exists(host):
        libs["b200_petsc_symbols.h"] = libs["b200_gamg.h"] = C.CDLL(host)
    decl = declared_symbols()
    assert len([1 for f, _ in decl if f == "b200_seqaij.h"]) >= 30
    for fn, name in sorted(decl):
        lib = libs.get(fn)
        assert lib is not None, f"no library built for {fn}"
        assert hasattr(lib, name), f"{name} declared in include/{fn} is not exported"
    assert set(pk.ABI_SYMBOLS) <= {n for _, n in decl}


def test_no_cpu_fallback(pk):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    ai = np.array([0, 1], np.int32)
    with pytest.raises(pk.B200Error) as e:
        pk.Csr(ai, np.array([0], np.int32), np.array([1.0]))
    assert e.value.code == 92 and "no CPU fallback" in str(e.value)
    out = np.zeros(1)
    rc = pk.lib.b200_vec_dot(None, None, C.c_int64(0), out.ctypes.data_as(C.c_void_p), None)
    assert rc == 92


def test_argument_errors_do_not_abort(pk):
    assert pk.lib.b200_csr_create(None, 1, 1, None, None, None) == 60
    assert pk.lib.b200_spmv(None, None, None, 0, None) == 60
    assert pk.lib.b200_csr_destroy(None) == 0
    assert b"b200" in pk.lib.b200_version()


def test_headers_are_plain_c_and_c_example_refuses_without_gpu():
    """include/*.h compile as C99 (-pedantic); the C example links against the library and, on a
    machine without a GPU, stops with the library's error instead of computing on the CPU."""
    import subprocess
    import torch
    for h in ("b200_seqaij.h", "b200_mpiaij.h", "b200_petsc_symbols.h"):
        src = f'#include "{h}"\nint main(void) {{ return 0; }}\n'
        r = subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-fsyntax-only", "-I", os.path.join(ROOT, "include"),
                            "-x", "c", "-"], input=src, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
    exe = os.path.join(ROOT, "petsc-openacc_b200", "bin", "spmv_from_c")
    assert os.path.exists(exe), "make -C petsc-openacc_b200/host builds the C example"
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    if torch.cuda.is_available():
        assert r.returncode == 0 and "mismatches 0" in r.stdout, r.stdout + r.stderr
    else:
        assert r.returncode == 2 and "no CPU fallback" in r.stderr
