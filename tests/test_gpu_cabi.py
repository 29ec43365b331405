"""The drop-in boundary is a C ABI: a plain C99 program (tests/c/gae_abi_test.c) includes
include/b200ppo.h, links libb200ppo.so and the CUDA runtime, and runs `b200ppo_gae` on the reference's
own known-answer recipe (ppo_test.py:229-264; golden in tests/golden/gae_kat.npz).  No Python, torch or
ctypes between the caller and the kernel."""
import os
import struct
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "c", "gae_abi_test.c")


def _compile(tmp_path):
    from nnx_ppo_b200 import build
    build.build()
    cuda = os.path.dirname(os.path.dirname(build._nvcc()))
    exe = str(tmp_path / "gae_abi_test")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                    "-isystem", os.path.join(cuda, "include"), SRC,     # -isystem: CUDA's own headers are not -pedantic clean
                    "-L", build.LIB_DIR, "-lb200ppo",
                    "-L", os.path.join(cuda, "lib64"), "-lcudart", f"-Wl,-rpath,{build.LIB_DIR}",
                    f"-Wl,-rpath,{os.path.join(cuda, 'lib64')}", "-o", exe], check=True)
    return exe


def test_c99_caller_compiles_and_links(tmp_path):
    """CPU check: the header is enough for a C99 translation unit and the library resolves every symbol
    it uses (no compute call without a GPU)."""
    _compile(tmp_path)


@pytest.mark.gpu
def test_c99_caller_reproduces_the_reference_gae_kat(tmp_path, cuda_device):
    exe = _compile(tmp_path)
    k = np.load(os.path.join(ROOT, "tests", "golden", "gae_kat.npz"))
    rewards, values = k["rewards_f32"], k["values_f32"]
    T, B = rewards.shape
    done = np.unpackbits(k["done_bits"])[:T * B].reshape(T, B).astype(np.uint8)
    trunc = np.unpackbits(k["trunc_bits"])[:T * B].reshape(T, B).astype(np.uint8)
    inp, outp = tmp_path / "in.bin", tmp_path / "out.bin"
    with open(inp, "wb") as f:
        f.write(struct.pack("<iiff", T, B, 0.95, 0.8))           # lambda_, gamma of ppo_test.py:229-264
        for a in (rewards, values[:-1], values[-1], done, trunc):
            f.write(np.ascontiguousarray(a).tobytes())
    r = subprocess.run([exe, str(inp), str(outp)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    adv = np.fromfile(outp, np.float32).reshape(T, B)
    assert np.abs(adv - k["adv_f64"]).max() < 1e-6           # the reference test's own gate (ppo_test.py:264)
