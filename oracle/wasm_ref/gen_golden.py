#!/usr/bin/env python
"""Generates tests/golden/* by running the reference ITSELF: pixo's committed WebAssembly build
(/root/reference/web/src/lib/pixo-wasm/pixo_bg.wasm, produced by pixo's authors from the same
crate) executed by oracle/wasm_ref.  Run in the build container only (needs /root/reference):

    python oracle/wasm_ref/gen_golden.py

Every fixture is real pixo output: complete JPEG files (pixo::jpeg::encode via wasm encodeJpeg)
and complete PNG files (pixo::png::encode via encodePng), plus a manifest describing how to
regenerate each input deterministically and the SHA-256 of that input.
"""
import hashlib
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle.wasm_ref import build as wb  # noqa: E402
from golden_inputs import make_input  # noqa: E402

WASM = "/root/reference/web/src/lib/pixo-wasm/pixo_bg.wasm"
OUT = os.path.join(ROOT, "tests", "golden")

JPEG_CASES = []
for (w, h) in [(1, 1), (7, 7), (8, 8), (9, 9), (15, 17), (16, 16), (1, 100), (100, 1), (70, 45), (253, 131)]:
    for kind in ("noise", "primaries"):
        for s420 in (0, 1):
            JPEG_CASES.append(dict(w=w, h=h, ct=2, kind=kind, seed=42, q=80, preset=0, s420=s420))
for q in (1, 50, 95, 100):
    for s420 in (0, 1):
        JPEG_CASES.append(dict(w=70, h=45, ct=2, kind="noise", seed=7, q=q, preset=0, s420=s420))
        JPEG_CASES.append(dict(w=70, h=45, ct=2, kind="smooth", seed=7, q=q, preset=1, s420=s420))
# BASELINE config C1: 256x256 RGB q=80 (4:2:0 and 4:4:4, standard and optimised Huffman)
for kind in ("gradient", "noise"):
    for s420 in (1, 0):
        for preset in (0, 1):
            JPEG_CASES.append(dict(w=256, h=256, ct=2, kind=kind, seed=42, q=80, preset=preset, s420=s420))
for (w, h) in [(33, 17), (64, 64), (100, 75)]:
    for preset in (0, 1):
        JPEG_CASES.append(dict(w=w, h=h, ct=0, kind="noise", seed=11, q=80, preset=preset, s420=0))
JPEG_CASES.append(dict(w=512, h=384, ct=2, kind="smooth", seed=3, q=85, preset=0, s420=1))

PNG_CASES = []
for (w, h) in [(65, 64), (200, 33), (80, 20), (120, 90), (3, 2)]:
    for kind in ("noise", "smooth", "vgrad", "mixed"):
        for preset in (0, 1, 2):
            if preset == 2 and w * h > 7000 and kind != "noise":
                continue  # Zopfli-style DEFLATE is slow under the interpreter
            PNG_CASES.append(dict(w=w, h=h, ct=3, kind=kind, seed=5, preset=preset))
for ct in (0, 1, 2):
    PNG_CASES.append(dict(w=77, h=70, ct=ct, kind="noise", seed=9, preset=1))
    PNG_CASES.append(dict(w=77, h=70, ct=ct, kind="mixed", seed=9, preset=0))


def run(args, data):
    with tempfile.TemporaryDirectory() as td:
        inp, outp = os.path.join(td, "in.raw"), os.path.join(td, "out.bin")
        open(inp, "wb").write(data.tobytes())
        a = [wb.EXE, WASM, args[0], inp] + [str(x) for x in args[1:]] + [outp]
        r = subprocess.run(a, capture_output=True, text=True)
        if r.returncode:
            raise RuntimeError(f"{a}: {r.stderr}")
        return open(outp, "rb").read()


def main():
    wb.build()
    os.makedirs(OUT, exist_ok=True)
    manifest = {"source": "pixo_bg.wasm from leerob/pixo @ 437bf63 (web/src/lib/pixo-wasm), sha256 " +
                hashlib.sha256(open(WASM, "rb").read()).hexdigest(),
                "runner": "oracle/wasm_ref/wasm_ref.c", "jpeg": [], "png": []}
    for i, c in enumerate(JPEG_CASES):
        img = make_input(c["kind"], c["w"], c["h"], 3 if c["ct"] == 2 else 1, c["seed"])
        out = run(["jpeg", c["w"], c["h"], c["ct"], c["q"], c["preset"], c["s420"]], img)
        name = f"j{i:03d}.jpg"
        open(os.path.join(OUT, name), "wb").write(out)
        manifest["jpeg"].append(dict(c, file=name, input_sha256=hashlib.sha256(img.tobytes()).hexdigest()))
    for i, c in enumerate(PNG_CASES):
        bpp = (1, 2, 3, 4)[c["ct"]]
        img = make_input(c["kind"], c["w"], c["h"], bpp, c["seed"])
        out = run(["png", c["w"], c["h"], c["ct"], c["preset"], 0], img)
        name = f"p{i:03d}.png"
        open(os.path.join(OUT, name), "wb").write(out)
        manifest["png"].append(dict(c, file=name, input_sha256=hashlib.sha256(img.tobytes()).hexdigest()))
    json.dump(manifest, open(os.path.join(OUT, "manifest.json"), "w"), indent=1)
    total = sum(os.path.getsize(os.path.join(OUT, f)) for f in os.listdir(OUT))
    print(f"{len(JPEG_CASES)} JPEG + {len(PNG_CASES)} PNG fixtures, {total / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
