"""Host half of the product (headers, Huffman tables, multi-threaded entropy coder, argument
validation) against the oracle.  CPU only."""
import numpy as np
import pytest

import pixo_b200
from pixo_b200 import ColorType
from pixo_b200.jpeg import JpegOptions, Subsampling, entropy_encode


@pytest.mark.parametrize("w,h", [(1, 1), (17, 15), (70, 45), (256, 256), (640, 483)])
@pytest.mark.parametrize("ct", [0, 2])
def test_entropy_stage_is_byte_identical(po, w, h, ct):
    ch = 3 if ct == 2 else 1
    img = po.gen_noise(w, h, ch, 7) if (w + h) % 2 else po.gen_noise(w, h, ch, 1) // 3
    for ss in (0, 1):
        for q in (1, 50, 80, 95, 100):
            y, cb, cr = po.jpeg_coefficients(img, w, h, ct, ss, q)
            for ri in (0, 1, 7, 100):
                for opt in (False, True):
                    ref = po.jpeg_encode(img, w, h, ct, q, ss, ri, opt)
                    o = JpegOptions(w, h, ColorType(ct), q, Subsampling(ss), ri or None, opt)
                    assert entropy_encode(y, cb, cr, o) == ref, (ss, q, ri, opt)


def test_entropy_thread_count_does_not_change_bytes(po):
    w, h = 1920, 1080
    img = po.gen_gradient_rgb(w, h)
    y, cb, cr = po.jpeg_coefficients(img, w, h, 2, 1, 80)
    ref = po.jpeg_encode_from_coefficients(y, cb, cr, w, h, 2, 80, 1)
    o = JpegOptions(w, h, ColorType.Rgb, 80, Subsampling.S420)
    assert entropy_encode(y, cb, cr, o) == ref
    o.restart_interval = 120
    assert entropy_encode(y, cb, cr, o) == po.jpeg_encode_from_coefficients(y, cb, cr, w, h, 2, 80, 1, 120)


def test_all_ff_stuffing(po):
    """Coefficients chosen to emit long runs of 1 bits (0xFF bytes) at every alignment."""
    w, h = 64, 64
    rng = np.random.default_rng(3)
    y = np.zeros((64, 64), np.int16); y[:, :] = rng.choice([-1023, 1023, 511, -511], size=(64, 64))
    cb = np.zeros((0, 64), np.int16)
    ref = po.jpeg_encode_from_coefficients(y, cb, cb, w, h, 0, 80, 0)
    assert ref.count(b"\xff\x00") > 10
    o = JpegOptions(w, h, ColorType.Gray, 80, Subsampling.S444)
    assert entropy_encode(y, cb, cb, o) == ref


def test_quant_tables_match_oracle(po):
    from pixo_b200 import jpeg
    for q in list(range(1, 101)):
        a = jpeg.quant_tables(q); b = po.quant_tables(q)
        for x, y in zip(a, b):
            assert np.array_equal(x, y)


def test_entropy_validation_errors():
    y = np.zeros((1, 64), np.int16)
    for q in (0, 101):
        with pytest.raises(pixo_b200.PixoError) as e:
            entropy_encode(y, y, y, JpegOptions(1, 1, ColorType.Rgb, q))
        assert e.value.code == pixo_b200._lib.ERR_INVALID_QUALITY
    with pytest.raises(pixo_b200.PixoError) as e:
        entropy_encode(y, y, y, JpegOptions(0, 1, ColorType.Rgb, 80))
    assert e.value.code == pixo_b200._lib.ERR_INVALID_DIMENSIONS
    with pytest.raises(pixo_b200.PixoError) as e:
        entropy_encode(y, y, y, JpegOptions(70000, 1, ColorType.Rgb, 80))
    assert e.value.code == pixo_b200._lib.ERR_IMAGE_TOO_LARGE
    with pytest.raises(pixo_b200.PixoError) as e:
        entropy_encode(y, y, y, JpegOptions(1, 1, ColorType.Rgba, 80))
    assert e.value.code == pixo_b200._lib.ERR_UNSUPPORTED_COLOR


def test_entropy_rejects_uncodable_coefficients_and_bad_restart():
    """the standard tables have no code for AC category > 10 / DC-difference category
    > 11, and the DRI field is 16 bits - reject instead of reading past the tables."""
    E = pixo_b200._lib
    o = JpegOptions(8, 8, ColorType.Gray, 80, Subsampling.S444)
    z = np.zeros((0, 64), np.int16)
    for pos, val in ((5, 1024), (5, -1024), (0, 2048), (0, -2048)):
        y = np.zeros((1, 64), np.int16); y[0, pos] = val
        with pytest.raises(pixo_b200.PixoError) as e:
            entropy_encode(y, z, z, o)
        assert e.value.code == E.ERR_INVALID_ARGUMENT
    y = np.zeros((1, 64), np.int16); y[0, 0] = 2047; y[0, 5] = -1023
    assert entropy_encode(y, z, z, o)[:2] == b"\xff\xd8"
    with pytest.raises(pixo_b200.PixoError) as e:
        entropy_encode(y, z, z, JpegOptions(8, 8, ColorType.Gray, 80, Subsampling.S444, 70000))
    assert e.value.code == E.ERR_INVALID_RESTART
    from pixo_b200 import jpeg
    for fn in (lambda: entropy_encode(y, z, z, JpegOptions(8, 8, ColorType.Gray, 80, Subsampling.S444, 0)),
               lambda: jpeg.encode_batch(np.zeros((1, 64), np.uint8), JpegOptions(8, 8, ColorType.Gray, 80, restart_interval=0))):
        with pytest.raises(pixo_b200.PixoError) as e:
            fn()
        assert e.value.code == E.ERR_INVALID_RESTART
