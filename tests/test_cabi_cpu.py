"""CPU-side checks of the product library: it loads, exports every symbol include/*.h declares, refuses to run
without a GPU (no CPU fallback), and never links the oracle."""
import ctypes as C
import os
import subprocess

import pytest


def test_every_declared_symbol_is_exported(zlib):
    protos = zlib.declared_prototypes()
    names = [n for n, _, _ in protos]
    assert len(names) == len(set(names)) and len(names) >= 60
    raw = C.CDLL(zlib.LIB_PATH)
    for n in names:
        assert hasattr(raw, n), f"{n} declared in include/*.h but not exported"
    for required in ("zb_ctx_create", "zb_mle_upload", "zb_mle_round_sums", "zb_mle_fold_inplace", "zb_mle_eval",
                     "zb_merkle_build", "zb_merkle_open", "zb_xxh3_rows", "zb_prod_round_coeffs", "zh_sumcheck_prove",
                     "zh_lasso_prove", "zh_commit_open", "zh_transcript_challenge"):
        assert required in names


def test_exported_symbols_are_all_declared(zlib):
    out = subprocess.run(["nm", "-D", "--defined-only", zlib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l and l.split()[-1].startswith(("zb_", "zh_"))}
    declared = {n for n, _, _ in zlib.declared_prototypes()}
    assert exported == declared


def test_library_does_not_reference_the_oracle(zlib):
    out = subprocess.run(["nm", "-D", zlib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    assert "zo_" not in out
    deps = subprocess.run(["ldd", zlib.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in deps
    # and the package sources never import it
    root = os.path.dirname(zlib.LIB_PATH)
    for dirpath, _, files in os.walk(root):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "pyoracle" not in src and "zigz_oracle" not in src and "zo_hash" not in src, f


def test_status_names_are_the_reference_error_names(zlib):
    L = zlib.lib()
    names = {-1: "EmptyEvaluations", -2: "LengthNotPowerOfTwo", -3: "WrongNumberOfVariables", -4: "NoVariables",
             -5: "EmptyValues", -6: "IndexOutOfBounds", -7: "PointDimensionMismatch", -8: "NoQueries",
             -9: "MappingLengthMismatch", -10: "InvalidMapping", -11: "QueryTableMismatch", -12: "WrongNumberOfChallenges",
             -13: "DifferentNumberOfVariables", -100: "OutOfMemory", -200: "NoCudaDevice"}
    for code, name in names.items():
        assert L.zb_status_name(code).decode() == name


def test_no_cpu_fallback_without_a_device(zlib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the no-device path is exercised on the CPU box")
    with pytest.raises(zlib.ZigzError) as e:
        zlib.Context(0)
    assert e.value.name == "NoCudaDevice"


def test_zig_binding_declares_the_whole_c_abi():
    """bindings/zigz_b200.zig (what a zigz maintainer links against) must declare exactly the functions the headers
    declare: hand-written externs plus the section tools/gen_zig_externs.py derives from include/*.h."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "gen_zig_externs.py"), "--check"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
