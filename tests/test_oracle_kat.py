"""Pins the CPU oracle against every known-answer test the reference holds for the hot path
(SURVEY.md §8c).  CPU only."""
import io

import numpy as np
import pytest


def test_adler32_known_answers(po):
    # src/compress/adler32.rs:54-63, src/simd/fallback.rs:166-170
    assert po.adler32(b"") == 1
    assert po.adler32(b"hello") == 0x062C0215
    assert po.adler32(b"Adler-32") == 0x0C34027B
    assert po.adler32(b"123456789") == 0x091E01DE


def test_adler32_matches_zlib(po):
    import zlib
    rng = np.random.default_rng(1)
    for n in (1, 5551, 5552, 5553, 11104, 200001, 1 << 20):
        d = rng.integers(0, 256, n, dtype=np.uint8)
        assert po.adler32(d) == zlib.adler32(d.tobytes())
    assert po.adler32(np.full(1 << 20, 255, np.uint8)) == zlib.adler32(b"\xff" * (1 << 20))


def test_crc32_known_answers(po):
    # src/simd/fallback.rs:174-175, src/png/chunk.rs:45
    assert po.crc32(b"123456789") == 0xCBF43926
    assert po.crc32(b"IEND") == 0xAE426082


def test_colour_known_answers(po):
    # src/color.rs:118-140
    assert po.rgb_to_ycbcr(0, 0, 0) == (0, 128, 128)
    assert po.rgb_to_ycbcr(255, 255, 255) == (255, 128, 128)
    y, cb, cr = po.rgb_to_ycbcr(255, 0, 0)
    assert 50 < y < 100 and cb < 128 and cr > 200
    # the clamp is live for pure blue / pure red
    assert po.rgb_to_ycbcr(0, 0, 255)[1] == 255
    assert po.rgb_to_ycbcr(255, 0, 0)[2] == 255


def test_quant_known_answers(po):
    # src/jpeg/quantize.rs:175-190,218-243,287-307
    lz, cz, ln, cn = po.quant_tables(50)
    assert ln[0] == 16 and lz[0] == 16
    for a, b in ((0, 1), (101, 100)):
        assert np.array_equal(po.quant_tables(a)[2], po.quant_tables(b)[2])
    q = np.full(64, 16, np.float32)
    d = np.zeros(64, np.float32); d[0] = 16.5; d[1] = -160
    out = po.quantize_block(d, q)
    assert out[0] == 1 and out[1] == -10
    # round half away from zero
    d[:4] = [8.0, -8.0, 24.0, -24.0]
    assert list(po.quantize_block(d, q)[:4]) == [1, -1, 2, -2]
    for ql in (1, 50, 80, 95, 100):
        lzq, czq, lnq, cnq = po.quant_tables(ql)
        assert lnq.min() >= 1 and lnq.max() <= 255
        assert np.array_equal(lzq, lnq[np.array(ZIGZAG)].astype(np.uint8))


ZIGZAG = [0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27,
          20, 13, 6, 7, 14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51,
          58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63]


def test_zigzag_known_answers(po):
    # src/jpeg/quantize.rs:139-146,257-266
    zz = po.zigzag_reorder(np.arange(64, dtype=np.int16))
    assert list(zz[:6]) == [0, 1, 8, 16, 9, 2]
    assert sorted(zz.tolist()) == list(range(64))
    assert zz.tolist() == ZIGZAG


def test_huffman_symbol_known_answers(po):
    # src/jpeg/huffman.rs:488-507: category(0,±1,±3,127,255)=0,1,2,7,8; encode_value
    def dc_sym(v):
        b = np.zeros(64, np.int16); b[0] = v
        rs, amp, nb, _ = po.block_symbols(b, 0)
        return int(rs[0]), int(amp[0]), int(nb[0])
    assert [dc_sym(v)[0] for v in (0, 1, -1, 3, -3, 127, 255)] == [0, 1, 1, 2, 2, 7, 8]
    assert dc_sym(-1)[1:] == (0, 1)
    assert dc_sym(-3)[1:] == (0, 2)
    assert dc_sym(3)[1:] == (3, 2)
    # EOB iff trailing zeros; ZRL for runs >= 16
    b = np.zeros(64, np.int16); b[ZIGZAG[20]] = 5
    rs, amp, nb, _ = po.block_symbols(b, 0)
    assert list(rs) == [0, 0xF0, (3 << 4) | 3, 0x00]
    b[63] = 1
    assert po.block_symbols(b, 0)[0][-1] != 0x00


def test_dct_loose_known_answers(po):
    # src/jpeg/dct.rs:806-851
    assert np.abs(po.dct_2d(np.zeros(64))).max() < 1e-3
    c = po.dct_2d(np.full(64, 100.0))
    assert abs(c[0]) > 100 and np.abs(c[1:]).max() < 5
    # against the textbook orthonormal DCT-II in float64
    rng = np.random.default_rng(0)
    blk = rng.integers(-128, 128, 64).astype(np.float32)
    k = np.arange(8)
    C = np.cos((2 * k[None, :] + 1) * k[:, None] * np.pi / 16) * np.where(k[:, None] == 0, np.sqrt(1 / 8), 0.5)
    ref = C @ blk.reshape(8, 8).astype(np.float64) @ C.T
    assert np.abs(po.dct_2d(blk).reshape(8, 8) - ref).max() < 1e-2


def test_png_filter_known_answers(po):
    # src/png/filter.rs:695-719,869-894; src/simd/fallback.rs:220-244
    assert po.filter_sub(bytes([10, 20, 30, 40, 50, 60]), 3).tolist() == [10, 20, 30, 30, 30, 30]
    assert po.filter_up(bytes([50, 60, 70]), bytes([10, 20, 30])).tolist() == [40, 40, 40]
    assert po.paeth_predictor(10, 20, 15) == 15
    assert po.paeth_predictor(10, 10, 10) == 10      # ties: a first
    assert po.paeth_predictor(5, 9, 7) == 7           # pa == pb == 2, pc == 0 -> c
    assert po.paeth_predictor(0, 0, 100) == 0
    assert po.filter_sub(bytes([10, 20, 30, 40]), 1).tolist() == [10, 10, 10, 10]
    assert po.filter_average(bytes([10, 20, 30, 40]), bytes([0, 0, 0, 0]), 1).tolist() == [10, 15, 20, 25]
    # score: tests/simd_fallback_equality.rs:466-484
    assert po.score_filter(bytes(1000)) == 0
    assert po.score_filter(b"\x80" * 1000) == 128000
    assert po.score_filter(b"\xff" * 1000) == 1000


def test_optimize_alpha_known_answer(po):
    # src/png/mod.rs:2053-2065 (test_optimize_alpha_zeroes_color) + the GrayAlpha arm :662-668
    assert po.optimize_alpha([10, 20, 30, 0, 1, 2, 3, 255], 3).tolist() == [0, 0, 0, 0, 1, 2, 3, 255]
    assert po.optimize_alpha([9, 0, 7, 1], 1).tolist() == [0, 0, 7, 1]
    assert po.optimize_alpha([9, 0, 7], 2).tolist() == [9, 0, 7]       # Rgb / Gray: untouched


def test_small_image_uses_sub(po):
    # src/png/filter.rs:1070 (area <= 4096 forces Sub for the adaptive strategies)
    img = po.gen_noise(64, 64, 4, 3)
    for strat in (po.F_ADAPTIVE, po.F_ADAPTIVE_FAST, po.F_BIGRAMS):
        out = po.apply_filters(img, 64, 64, 4, strat).reshape(64, 64 * 4 + 1)
        assert (out[:, 0] == 1).all()
    out = po.apply_filters(po.gen_noise(65, 64, 4, 3), 65, 64, 4, po.F_ADAPTIVE).reshape(64, -1)
    assert set(out[:, 0]) <= {0, 1, 2, 3, 4}


def test_png_filters_roundtrip_and_zlib(po):
    """The oracle's filtered stream unfilters back to the input (mirrors the reference's
    lossless round-trip checks) for every strategy."""
    w, h, bpp = 67, 41, 3
    img = po.gen_noise(w, h, bpp, 9).reshape(h, w * bpp)
    for strat in range(9):
        out = po.apply_filters(img, w, h, bpp, strat).reshape(h, w * bpp + 1)
        rec = np.zeros((h, w * bpp), np.uint8)
        for y in range(h):
            f = out[y, 0]; row = out[y, 1:].astype(np.int32)
            prev = rec[y - 1].astype(np.int32) if y else np.zeros(w * bpp, np.int32)
            cur = np.zeros(w * bpp, np.int32)
            for i in range(w * bpp):
                a = cur[i - bpp] if i >= bpp else 0
                b = prev[i]
                c = prev[i - bpp] if i >= bpp else 0
                if f == 0: p = 0
                elif f == 1: p = a
                elif f == 2: p = b
                elif f == 3: p = (a + b) // 2
                else: p = po.paeth_predictor(int(a), int(b), int(c))
                cur[i] = (row[i] + p) & 255
            rec[y] = cur
        assert np.array_equal(rec, img), strat


@pytest.mark.parametrize("w,h", [(70, 45), (33, 17), (256, 256)])
def test_oracle_jpeg_decodes_with_libjpeg(po, w, h):
    """Structure pins of tests/jpeg_conformance.rs (SOI/APP0/JFIF/4 DHT/EOI) plus a third-party
    decoder accepting the stream with sane PSNR."""
    from PIL import Image
    img = po.gen_gradient_rgb(w, h)
    for ss in (po.S444, po.S420):
        for q in (50, 80, 95):
            for opt in (False, True):
                j = po.jpeg_encode(img, w, h, po.RGB, q, ss, 0, opt)
                assert j[:2] == b"\xff\xd8" and j[-2:] == b"\xff\xd9"
                assert j[2:4] == b"\xff\xe0" and j[6:11] == b"JFIF\0"
                assert j.count(b"\xff\xc4") >= 4
                im = Image.open(io.BytesIO(j)); im.load()
                assert im.size == (w, h)
                a = np.asarray(im.convert("RGB")).astype(float)
                mse = ((a - img.reshape(h, w, 3).astype(float)) ** 2).mean()
                assert 10 * np.log10(255 ** 2 / max(mse, 1e-9)) > 30
    g = img[::3].copy()
    j = po.jpeg_encode(g, w, h, po.GRAY, 80, po.S444)
    im = Image.open(io.BytesIO(j)); im.load()
    assert im.mode == "L" and im.size == (w, h)


def test_oracle_restart_markers(po):
    # tests/jpeg_conformance.rs:595-654: DRI present, RSTn cycle, no trailing RST before EOI
    w, h = 64, 48
    img = po.gen_noise(w, h, 3, 5)
    j = po.jpeg_encode(img, w, h, po.RGB, 80, po.S420, 2)
    assert b"\xff\xdd\x00\x04\x00\x02" in j
    sos = j.index(b"\xff\xda")
    body = j[sos:]
    rst = [body[i + 1] for i in range(len(body) - 1) if body[i] == 0xFF and 0xD0 <= body[i + 1] <= 0xD7]
    assert rst == [0xD0 + (i & 7) for i in range(len(rst))] and len(rst) == 12 // 2 - 1
    assert not (0xD0 <= j[-3] <= 0xD7 and j[-4] == 0xFF)
    from PIL import Image
    Image.open(io.BytesIO(j)).load()


def test_oracle_errors(po):
    # tests/jpeg_conformance.rs:242-292
    img = bytes(3)
    for q in (0, 101):
        with pytest.raises(ValueError):
            po.jpeg_encode(img, 1, 1, po.RGB, q)
    with pytest.raises(ValueError):
        po.jpeg_encode(img, 0, 1, po.RGB, 80)
    with pytest.raises(ValueError):
        po.jpeg_encode(bytes(4), 1, 1, po.RGBA, 80)
    with pytest.raises(ValueError):
        po.jpeg_encode(bytes(5), 1, 1, po.RGB, 80)
