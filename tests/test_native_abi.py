"""The C ABI from C: tests/native/abi_smoke.c is compiled with gcc against include/h263cu.h + include/h263synth.h,
linked with libh263cu.so, and run.  Without a GPU it covers the host entry points and checks that the device entry
points refuse loudly; on the GPU box the same binary decodes a stream and checks the fused RGBA."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "native", "abi_smoke.c")
LIBDIR = os.path.join(ROOT, "h263_rs_b200")


def build(tmp_path):
    exe = str(tmp_path / "abi_smoke")
    subprocess.check_call(["gcc", "-O1", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"), "-o", exe, SRC,
                           "-L", LIBDIR, "-lh263cu", "-Wl,-rpath," + LIBDIR, "-ldl"])
    return exe


def test_abi_from_c_host_part(tmp_path):
    out = subprocess.run([build(tmp_path), os.path.join(LIBDIR, "libh263synth.so")], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    assert "abi_smoke ok" in out.stdout


@pytest.mark.gpu
def test_abi_from_c_device_part(tmp_path):
    out = subprocess.run([build(tmp_path), os.path.join(LIBDIR, "libh263synth.so")], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    assert "host + device" in out.stdout
