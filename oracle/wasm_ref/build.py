"""Builds oracle/wasm_ref/wasm_ref (test infrastructure: runs the reference's own wasm artefact)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
EXE = os.path.join(HERE, "wasm_ref")


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "wasm_ref.c")
    if force or not os.path.exists(EXE) or os.path.getmtime(EXE) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-msse2", "-mfpmath=sse",
                               "-w", src, "-lm", "-o", EXE])
    return EXE


if __name__ == "__main__":
    print(build(True))
