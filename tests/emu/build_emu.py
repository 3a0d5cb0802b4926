"""Builds the CPU emulator of the CUDA kernels (TEST INFRASTRUCTURE ONLY).

g++ compiles scarplet_b200/csrc/sb_lib.cu with -DSB_EMU against tests/emu/sb_emu.h:
same kernel source, CUDA threads emulated as fibers.  The result is loaded only by
tests (``tests/conftest.py::emu_lib``); the product never opens it."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
OUT = os.path.join(HERE, "libscarplet_b200_emu.so")
SRC = os.path.join(ROOT, "scarplet_b200", "csrc")


def build(force=False):
    srcs = [os.path.join(SRC, f) for f in os.listdir(SRC)] + [os.path.join(HERE, "sb_emu.h"),
                                                             os.path.join(ROOT, "include", "scarplet_b200.h")]
    if not force and os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(s) for s in srcs):
        return OUT
    cmd = ["g++", "-O2", "-std=c++17", "-DSB_EMU", "-x", "c++", "-ffp-contract=off", "-fPIC", "-shared",
           "-I" + HERE, "-I" + SRC, "-o", OUT, os.path.join(SRC, "sb_lib.cu"), "-lpthread"]
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force=True))
