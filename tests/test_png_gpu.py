"""GPU parity for the PNG filter-selection + Adler-32 path (through the C ABI) vs the oracle."""
import zlib

import numpy as np
import pytest

import pixo_b200
from pixo_b200 import ColorType, png
from pixo_b200.png import FilterStrategy, PngOptions

pytestmark = pytest.mark.gpu

STRATS = [FilterStrategy.NoFilter, FilterStrategy.Sub, FilterStrategy.Up, FilterStrategy.Average,
          FilterStrategy.Paeth, FilterStrategy.MinSum, FilterStrategy.Adaptive, FilterStrategy.AdaptiveFast,
          FilterStrategy.Bigrams]


def _smooth(w, h, bpp, seed):
    rng = np.random.default_rng(seed)
    x = np.cumsum(rng.integers(-2, 3, (h, w * bpp)), axis=1) + np.cumsum(rng.integers(-1, 2, (h, 1)), axis=0)
    return (x & 255).astype(np.uint8).reshape(-1)


@pytest.mark.parametrize("bpp", [1, 2, 3, 4])
@pytest.mark.parametrize("w,h", [(1, 1), (3, 2), (64, 64), (65, 64), (100, 33), (257, 40), (1000, 7), (4099, 35)])
def test_filters_match_oracle(po, gpu_ctx, w, h, bpp):
    # tests/simd_fallback_equality.rs shape: seeded rows, bpp 1-4, assorted widths
    for name, img in (("noise", po.gen_noise(w, h, bpp, 42)), ("smooth", _smooth(w, h, bpp, 1)),
                      ("zeros", np.zeros(w * h * bpp, np.uint8))):
        for st in STRATS:
            opts = PngOptions(w, h, ColorType.Rgba, st)
            got, ad = png.apply_filters(img, w, h, bpp, opts, with_adler=True, ctx=gpu_ctx)
            ref = po.apply_filters(img, w, h, bpp, int(st))
            assert np.array_equal(got, ref), (name, st, w, h, bpp, np.flatnonzero(got != ref)[:5])
            assert ad == po.adler32(ref) == zlib.adler32(ref.tobytes())


@pytest.mark.parametrize("bpp", [3, 4])
@pytest.mark.parametrize("w,h", [(300, 96), (1024, 70), (2051, 53)])
def test_rows_whose_ladder_ends_early_next_to_rows_that_need_paeth(po, gpu_ctx, w, h, bpp):
    # The band kernel scores Paeth only while a row's ladder is open and switches between a one-pass
    # and a two-pass scoring of the row by what the row above needed: flat rows (Sub or Up wins at
    # once), noise rows (all five scored) and every order of the two, across band borders (16 rows).
    rng = np.random.default_rng(7)
    noise = po.gen_noise(w, h, bpp, 9).reshape(h, w * bpp)
    flat = np.repeat(rng.integers(0, 256, (h, 1, bpp), dtype=np.uint8), w, axis=1).reshape(h, w * bpp)
    ramp = ((np.arange(w * bpp) // bpp)[None, :] + np.arange(h)[:, None]).astype(np.uint8)
    patterns = {
        "alternate": np.arange(h) % 2 == 0,
        "runs of 5": (np.arange(h) // 5) % 2 == 0,
        "band border": (np.arange(h) // 16) % 2 == 0,
        "random": rng.integers(0, 2, h).astype(bool),
    }
    for name, pick in patterns.items():
        for calm in (flat, ramp):
            img = np.where(pick[:, None], calm, noise).reshape(-1)
            for st in (FilterStrategy.Adaptive, FilterStrategy.AdaptiveFast, FilterStrategy.MinSum):
                opts = PngOptions(w, h, ColorType.Rgba, st)
                got, ad = png.apply_filters(img, w, h, bpp, opts, with_adler=True, ctx=gpu_ctx)
                ref = po.apply_filters(img, w, h, bpp, int(st))
                assert np.array_equal(got, ref), (name, st, w, h, bpp, np.flatnonzero(got != ref)[:5])
                assert ad == zlib.adler32(ref.tobytes())
        # both kinds of row really occur: the winners include None/Sub/Up rows and Paeth or Average rows
        types = set(ref.reshape(h, w * bpp + 1)[:, 0].tolist())
        assert len(types) >= 2, types


@pytest.mark.parametrize("ct,bpp", [(3, 4), (1, 2), (2, 3)])
@pytest.mark.parametrize("w,h", [(3, 2), (65, 64), (100, 33), (300, 20), (1000, 70), (4099, 35)])
def test_optimize_alpha_fused(po, gpu_ctx, w, h, ct, bpp):
    """PngOptions::optimize_alpha (maybe_optimize_alpha, src/png/mod.rs:633-671) fused into the row
    reads: same stream as the oracle's pre-pass followed by the plain filter; Rgb is untouched."""
    rng = np.random.default_rng(w * 31 + h)
    img = rng.integers(0, 256, (h, w, bpp), dtype=np.uint8)
    if bpp in (2, 4):
        img[..., -1] = np.where(rng.random((h, w)) < 0.4, 0, img[..., -1])   # many transparent pixels
    img = img.reshape(-1)
    pre = po.optimize_alpha(img, ct)
    assert (bpp == 3) == np.array_equal(pre, img)
    for st in STRATS:
        opts = PngOptions(w, h, ColorType(ct), st, True)
        got, ad = png.apply_filters(img, w, h, bpp, opts, with_adler=True, ctx=gpu_ctx)
        ref = po.apply_filters(pre, w, h, bpp, int(st))
        assert np.array_equal(got, ref), (st, np.flatnonzero(got != ref)[:5])
        assert ad == zlib.adler32(ref.tobytes())


def test_sticky_adaptive_fast_small_height(po, gpu_ctx):
    # height <= 32 takes the sequential path where AdaptiveFast reuses row 0's winner
    for w, h in ((300, 32), (300, 20), (5000, 2)):
        for img in (po.gen_noise(w, h, 4, 3), _smooth(w, h, 4, 2)):
            got = png.apply_filters(img, w, h, 4, PngOptions(w, h, ColorType.Rgba, FilterStrategy.AdaptiveFast), ctx=gpu_ctx)
            ref = po.apply_filters(img, w, h, 4, po.F_ADAPTIVE_FAST)
            assert np.array_equal(got, ref)
            assert len(set(ref.reshape(h, -1)[:, 0])) == 1


def test_row_bytes_override_sub_byte_depth(po, gpu_ctx):
    # palette / low-bit-depth rows: row_bytes != width*bpp (src/png/mod.rs:556-560)
    w, h, row_bytes = 1001, 50, 126
    img = po.gen_noise(row_bytes, h, 1, 4)
    got = png.apply_filters_with_row_bytes(img, w, h, row_bytes, 1, PngOptions(w, h, ColorType.Gray, FilterStrategy.Adaptive), ctx=gpu_ctx)
    assert np.array_equal(got, po.apply_filters(img, w, h, 1, po.F_ADAPTIVE, row_bytes=row_bytes))


def test_c5_4k_rgba(po, gpu_ctx):
    """BASELINE config C5 geometry at full size: 3840x2160 RGBA, Adaptive and AdaptiveFast."""
    w, h = 3840, 2160
    rgb = po.gen_gradient_rgb(w, h).reshape(h, w, 3)
    grad = np.concatenate([rgb, np.full((h, w, 1), 255, np.uint8)], -1).reshape(-1)
    for img in (grad, po.gen_noise(w, h, 4, 42)):
        for st in (FilterStrategy.Adaptive, FilterStrategy.AdaptiveFast):
            got, ad = png.apply_filters(img, w, h, 4, PngOptions(w, h, ColorType.Rgba, st), with_adler=True, ctx=gpu_ctx)
            ref = po.apply_filters(img, w, h, 4, int(st))
            assert np.array_equal(got, ref)
            assert ad == zlib.adler32(ref.tobytes())


def test_long_rows_span_segments(po, gpu_ctx):
    # rows longer than the 32 KiB staging segment
    w, h, bpp = 30000, 40, 4
    img = _smooth(w, h, bpp, 5)
    for st in (FilterStrategy.Adaptive, FilterStrategy.Paeth, FilterStrategy.AdaptiveFast, FilterStrategy.Bigrams):
        got, ad = png.apply_filters(img, w, h, bpp, PngOptions(w, h, ColorType.Rgba, st), with_adler=True, ctx=gpu_ctx)
        ref = po.apply_filters(img, w, h, bpp, int(st))
        assert np.array_equal(got, ref)
        assert ad == zlib.adler32(ref.tobytes())


def test_adler32_known_answers_and_sizes(po, gpu_ctx):
    assert png.adler32(b"", ctx=gpu_ctx) == 1
    assert png.adler32(b"hello", ctx=gpu_ctx) == 0x062C0215
    assert png.adler32(b"Adler-32", ctx=gpu_ctx) == 0x0C34027B
    assert png.adler32(b"123456789", ctx=gpu_ctx) == 0x091E01DE
    rng = np.random.default_rng(0)
    for n in (1, 15, 16, 17, 5552, 5553, 65521, 1 << 20, (1 << 24) + 3):
        d = rng.integers(0, 256, n, dtype=np.uint8)
        assert png.adler32(d, ctx=gpu_ctx) == zlib.adler32(d.tobytes())
    d = np.full(33179760, 255, np.uint8)   # C5 filtered-stream size, worst-case sums
    assert png.adler32(d, ctx=gpu_ctx) == zlib.adler32(d.tobytes())


def test_invalid_arguments(gpu_ctx):
    img = np.zeros(70 * 70 * 4, np.uint8)
    with pytest.raises(pixo_b200.PixoError):
        png.apply_filters(img, 70, 70, 5, PngOptions(70, 70), ctx=gpu_ctx)
    with pytest.raises(pixo_b200.PixoError):
        png.apply_filters(img[:-1], 70, 70, 4, PngOptions(70, 70), ctx=gpu_ctx)
