"""zb_ctx_create_mask: ONE context over several GPUs of one process (SURVEY.md §8b `device_mask`), no NCCL / IPC / torch in
the product path. The C++ test drives it through the C ABI only; it needs >= 2 GPUs and is skipped on a one-GPU box
(compiled on every box)."""
import ctypes as C
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "test_mask_ctx.cpp")
BIN = os.path.join(ROOT, "tests", "cpp", "_build", "test_mask_ctx")


def _build(zlib, po):
    os.makedirs(os.path.dirname(BIN), exist_ok=True)
    odir = os.path.join(ROOT, "oracle", "_build")
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", SRC, "-o", BIN, "-L" + os.path.dirname(zlib.LIB_PATH), "-lzigz_b200", "-L" + odir,
           "-lzigz_oracle", "-Wl,-rpath," + os.path.dirname(zlib.LIB_PATH), "-Wl,-rpath," + odir]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]


def test_mask_ctx_test_compiles_and_links(zlib, po):
    _build(zlib, po)
    assert os.path.exists(BIN)


def _gpu_count():
    r = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True)
    return sum(1 for ln in r.stdout.splitlines() if ln.startswith("GPU "))


@pytest.mark.gpu
def test_single_bit_mask_is_an_ordinary_context(zlib, po):
    with zlib.Context(device_mask=1) as c:
        assert c.n_devices == 1
        e = po.fill_synthetic(zlib.BABYBEAR_P, 5, 0, 1 << 12)
        assert zlib.SumcheckProver.prove(zlib.Multilinear.init(c, e)).to_bytes() == po.sumcheck_prove(zlib.BABYBEAR_P, e).to_bytes()
    h = C.c_void_p()
    assert zlib.lib().zb_ctx_create_mask(0, C.byref(h)) == -22 and zlib.lib().zb_ctx_create_mask(0b111, C.byref(h)) == -22  # BadArgument


@pytest.mark.gpu
def test_one_process_many_gpus_bit_equal_to_one_gpu(zlib, po):
    n = _gpu_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    _build(zlib, po)
    for gpus in [g for g in (2, 4, 8) if g <= n]:
        r = subprocess.run([BIN, str(gpus), "24"], capture_output=True, text=True, timeout=900)
        assert r.returncode == 0 and "all checks passed" in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]
