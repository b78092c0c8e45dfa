"""Compile the kernel sources for the fiber-based CUDA emulator (tests only; see cuda_emu.h)."""
from __future__ import annotations

import fcntl
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "intel_sigir2023_b200", "csrc")
OUT = os.path.join(HERE, "_build")
LIB = os.path.join(OUT, "libintel_b200_emu.so")


def _up_to_date() -> bool:
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "cuda_emu.h"),
                                                                  os.path.join(ROOT, "include", "intel_b200.h")]
    return os.path.exists(LIB) and all(os.path.getmtime(d) <= os.path.getmtime(LIB) for d in deps)


def build_emu() -> str:
    """Builds (once) and returns the emulator library.  Several test processes may ask at the same time (the gloo
    workers are spawned in parallel): the build runs under a file lock and the library is moved into place atomically,
    so nobody ever loads a half-written file."""
    from intel_sigir2023_b200.build import SOURCES
    os.makedirs(OUT, exist_ok=True)
    if _up_to_date():
        return LIB
    with open(os.path.join(OUT, ".lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if _up_to_date():           # somebody else built it while we were waiting
                return LIB
            unity = os.path.join(OUT, f"unity.{os.getpid()}.cpp")
            tmp = LIB + f".{os.getpid()}.tmp"
            with open(unity, "w") as f:
                for s in SOURCES:
                    f.write(f'#include "{os.path.join(CSRC, s)}"\n')
            cmd = ["g++", "-O1", "-g", "-std=c++17", "-fPIC", "-shared", "-DINTEL_EMU", "-Wno-unknown-pragmas",
                   "-fno-omit-frame-pointer", "-pthread", "-I", HERE, "-I", CSRC, "-I", os.path.join(ROOT, "include"),
                   "-x", "c++", unity, "-o", tmp]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError("emu build failed:\n" + r.stdout + r.stderr[-8000:])
            os.replace(tmp, LIB)
            os.remove(unity)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB


if __name__ == "__main__":
    print(build_emu())
