"""CPU checks of the boundary: the shared library loads, exports every symbol include/conos_gpu.h declares, the
ctypes table covers them all, and nothing computes without a GPU (loud failure, no fallback)."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "conos_gpu.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(cg_[a-z0-9_]+)\s*\(", txt)))


def test_header_symbols_exported():
    import conos_b200
    lib = conos_b200.load()
    decl = declared_symbols()
    assert len(decl) >= 35
    out = subprocess.check_output(["nm", "-D", "--defined-only", conos_b200._lib.LIB_PATH], text=True)
    exported = set(re.findall(r" T (cg_[a-z0-9_]+)", out))
    missing = [s for s in decl if s not in exported]
    assert not missing, missing
    for s in decl:
        getattr(lib, s)


def test_ctypes_table_matches_header():
    import conos_b200
    decl = set(declared_symbols())
    bound = set(conos_b200._lib.SIGNATURES)
    assert decl == bound, (sorted(decl - bound), sorted(bound - decl))


def test_no_cpu_fallback():
    import conos_b200
    lib = conos_b200.load()
    if lib.cg_device_count() > 0:
        pytest.skip("a GPU is visible")
    with pytest.raises(conos_b200.ConosGpuError):
        conos_b200.Context(0)
    from conos_b200 import ops
    import numpy as np
    with pytest.raises(conos_b200.ConosGpuError):
        ops.crossKnn(np.zeros((4, 3)), np.zeros((4, 3)), 2)


def test_product_never_imports_oracle():
    """The product path must not route through the oracle (or any CPU fallback)."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "conos_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                bad = re.findall(r"import\s+conos_oracle|from\s+conos_oracle|libconos_oracle|\boracle_[a-z0-9_]+\s*\(|"
                                 r"_build/", txt)
                assert not bad, (os.path.join(dirpath, f), bad)
                if f == "comm.cu":
                    # the one run-time binding in the product: NCCL, only when a multi-GPU communicator is created
                    assert re.findall(r'const char\* names\[\] = \{([^}]*)\}', txt) == ['"libnccl.so.2", "libnccl.so"']
                    assert len(re.findall(r"\bdlopen\s*\(", txt)) == 1 and "oracle" not in txt
                else:
                    assert not re.findall(r"\bdlopen\b", txt), os.path.join(dirpath, f)


def test_only_sm100a_is_built():
    mk = open(os.path.join(ROOT, "conos_b200", "csrc", "Makefile")).read()
    assert "arch=compute_100a,code=sm_100a" in mk and "-lineinfo" in mk
