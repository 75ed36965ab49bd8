"""Fixtures under tests/golden/ are REAL pixo output: complete JPEG / PNG files produced by
pixo's own WebAssembly build (web/src/lib/pixo-wasm/pixo_bg.wasm) executed by oracle/wasm_ref in
the build container (oracle/wasm_ref/gen_golden.py).  These tests pin
  * the CPU oracle to the reference (CPU, always run), and
  * the CUDA product to the reference directly (GPU), without the oracle in between.
"""
import hashlib
import json
import os
import struct
import zlib

import numpy as np
import pytest

from golden_inputs import make_input

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MANIFEST = json.load(open(os.path.join(GOLD, "manifest.json")))
PNG_STRATEGY = {0: 7, 1: 6, 2: 8}  # preset -> FilterStrategy (AdaptiveFast, Adaptive, Bigrams)


def _input(c, ch):
    img = make_input(c["kind"], c["w"], c["h"], ch, c["seed"])
    assert hashlib.sha256(img.tobytes()).hexdigest() == c["input_sha256"], "input generator drifted"
    return img


def _png_parts(png):
    assert png[:8] == b"\x89PNG\r\n\x1a\n"
    p, idat, ihdr = 8, b"", None
    while p < len(png):
        ln = struct.unpack(">I", png[p:p + 4])[0]
        typ, data = png[p + 4:p + 8], png[p + 8:p + 8 + ln]
        assert struct.unpack(">I", png[p + 8 + ln:p + 12 + ln])[0] == zlib.crc32(typ + data)
        if typ == b"IDAT":
            idat += data
        if typ == b"IHDR":
            ihdr = struct.unpack(">IIBBBBB", data)
        p += 12 + ln
    return ihdr, zlib.decompress(idat), struct.unpack(">I", idat[-4:])[0]


def _png_filter_input(c, img, po=None):
    """What pixo hands to filter::apply_filters for this fixture: presets 1/2 enable
    optimize_alpha (src/png/mod.rs:633-671); the inputs are chosen so no colour-type or palette
    reduction applies (IHDR is asserted to keep the input colour type).  With `po` the pre-pass
    is the oracle's (so the fixtures pin it too); without it the raw input is returned and the
    caller asks the product to fuse the pre-pass."""
    bpp = (1, 2, 3, 4)[c["ct"]]
    src = img.copy().reshape(-1)
    if po is not None and c["preset"] in (1, 2):
        src = po.optimize_alpha(src, c["ct"])
    return src, bpp


def test_manifest_is_complete():
    assert len(MANIFEST["jpeg"]) >= 60 and len(MANIFEST["png"]) >= 50
    assert "pixo_bg.wasm" in MANIFEST["source"]


@pytest.mark.parametrize("c", MANIFEST["jpeg"], ids=lambda c: c["file"])
def test_oracle_reproduces_pixo_jpeg_bytes(po, c):
    img = _input(c, 3 if c["ct"] == 2 else 1)
    want = open(os.path.join(GOLD, c["file"]), "rb").read()
    got = po.jpeg_encode(img, c["w"], c["h"], c["ct"], c["q"], c["s420"], 0, c["preset"] == 1)
    assert got == want


@pytest.mark.parametrize("c", MANIFEST["png"], ids=lambda c: c["file"])
def test_oracle_reproduces_pixo_png_filter_stream(po, c):
    img = _input(c, (1, 2, 3, 4)[c["ct"]])
    ihdr, raw, adler = _png_parts(open(os.path.join(GOLD, c["file"]), "rb").read())
    assert ihdr[:2] == (c["w"], c["h"])
    if (ihdr[2], ihdr[3]) != (8, (0, 4, 2, 6)[c["ct"]]):
        pytest.skip("pixo's lossless palette/colour-type reduction rewrote this tiny input (outside the filter path)")
    src, bpp = _png_filter_input(c, img, po)
    # the wasm build has no `parallel` feature: always the sequential loop (sticky AdaptiveFast)
    mine = po.apply_filters(src, c["w"], c["h"], bpp, PNG_STRATEGY[c["preset"]], parallel_feature=False)
    assert mine.tobytes() == raw
    assert po.adler32(mine) == adler == zlib.adler32(raw)


def test_golden_filter_choices_are_diverse():
    """The fixtures exercise every filter type and non-trivial ladders."""
    seen = set()
    for c in MANIFEST["png"]:
        ihdr, raw, _ = _png_parts(open(os.path.join(GOLD, c["file"]), "rb").read())
        if (ihdr[2], ihdr[3]) != (8, (0, 4, 2, 6)[c["ct"]]):
            continue
        bpp = (1, 2, 3, 4)[c["ct"]]
        seen |= set(np.frombuffer(raw, np.uint8).reshape(c["h"], c["w"] * bpp + 1)[:, 0].tolist())
    assert seen == {0, 1, 2, 3, 4}


# ---- the CUDA product against real pixo output, no oracle involved -------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("c", MANIFEST["jpeg"], ids=lambda c: c["file"])
def test_gpu_reproduces_pixo_jpeg_bytes(gpu_ctx, c):
    from pixo_b200 import ColorType, jpeg
    from pixo_b200.jpeg import JpegOptions, Subsampling
    img = _input(c, 3 if c["ct"] == 2 else 1)
    want = open(os.path.join(GOLD, c["file"]), "rb").read()
    o = JpegOptions.from_preset(c["w"], c["h"], c["q"], c["preset"])   # as src/wasm.rs:112-148 builds them
    o.color_type = ColorType(c["ct"])
    o.subsampling = Subsampling.S420 if c["s420"] else Subsampling.S444
    assert jpeg.encode(img, o, ctx=gpu_ctx) == want


@pytest.mark.gpu
@pytest.mark.parametrize("c", MANIFEST["png"], ids=lambda c: c["file"])
def test_gpu_reproduces_pixo_png_filter_stream(gpu_ctx, c):
    """pixo-wasm always runs the sequential loop; the product implements the default-feature
    (rayon) semantics, which coincide for Adaptive everywhere and for AdaptiveFast when
    height <= 32 or the image is tiny (forced Sub)."""
    from pixo_b200 import ColorType, png
    from pixo_b200.png import FilterStrategy, PngOptions
    if c["preset"] == 0 and c["h"] > 32 and c["w"] * c["h"] > 4096:
        pytest.skip("sticky AdaptiveFast beyond 32 rows exists only in no-`parallel` builds")
    img = _input(c, (1, 2, 3, 4)[c["ct"]])
    ihdr, raw, adler = _png_parts(open(os.path.join(GOLD, c["file"]), "rb").read())
    if (ihdr[2], ihdr[3]) != (8, (0, 4, 2, 6)[c["ct"]]):
        pytest.skip("pixo's lossless palette/colour-type reduction rewrote this tiny input")
    src, bpp = _png_filter_input(c, img)
    st = FilterStrategy(PNG_STRATEGY[c["preset"]])
    # presets 1/2 set optimize_alpha: the product applies it on the fly to the raw input
    got, ad = png.apply_filters(src, c["w"], c["h"], bpp,
                                PngOptions(c["w"], c["h"], ColorType(c["ct"]), st, c["preset"] in (1, 2)),
                                with_adler=True, ctx=gpu_ctx)
    assert got.tobytes() == raw
    assert ad == adler
