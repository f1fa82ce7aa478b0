"""The PNG deflate format the device encoder emits (csrc/png.cu), through its CPU twin png_codec: the streams must
be valid for the stock decoders (zlib, PIL) -- the reference writes its frames with PIL (render.py:4462-4467), so
"same file content after decoding" is the parity statement for this stage."""
import io
import zlib

import numpy as np
import pytest

from black_hole_renderer_b200 import png_codec as pc


def _frames():
    rng = np.random.default_rng(5)
    yield "noise", rng.integers(0, 256, (37, 53, 3), dtype=np.uint8)
    yield "black", np.zeros((40, 64, 3), np.uint8)
    yield "white", np.full((9, 700, 3), 255, np.uint8)                     # runs longer than one segment
    g = np.zeros((48, 96, 3), np.uint8)
    g[...] = (np.arange(96)[None, :, None] * 2 + np.arange(48)[:, None, None]) % 256
    yield "gradient", g
    s = np.zeros((64, 64, 3), np.uint8)                                     # sparse stars on black: short + long runs
    s[rng.integers(0, 64, 40), rng.integers(0, 64, 40)] = rng.integers(1, 256, (40, 3))
    yield "stars", s
    yield "one_pixel", np.array([[[7, 8, 9]]], np.uint8)
    yield "one_column", rng.integers(0, 256, (300, 1, 3), dtype=np.uint8)


@pytest.mark.parametrize("name,img", list(_frames()), ids=[n for n, _ in _frames()])
def test_stream_decodes_with_zlib_and_pil(name, img):
    from PIL import Image
    filtered = pc.sub_filter(img)
    stream, nbits = pc.encode_stream_reference(filtered)
    assert zlib.decompress(bytes(stream)) == filtered.tobytes()
    assert len(stream) <= pc.stream_capacity(img.shape[1], img.shape[0])
    png = pc.encode_frame_reference(img)
    assert np.array_equal(np.array(Image.open(io.BytesIO(png)).convert("RGB")), img)


def test_sub_filter_layout():
    img = np.arange(2 * 3 * 3, dtype=np.uint8).reshape(2, 3, 3) * 9
    f = pc.sub_filter(img).reshape(2, 10)
    assert np.all(f[:, 0] == 1)
    assert np.array_equal(f[:, 1:4], img[:, 0].reshape(2, 3))
    assert np.array_equal(f[:, 4:], (img[:, 1:].astype(int) - img[:, :-1].astype(int)).reshape(2, 6).astype(np.uint8))


def test_tokenizer_rules():
    """First byte of a run is a literal, >= 4 repeats after it become one (length, distance 1) match, runs never
    cross a 256-byte segment."""
    seg = np.array([5] * 5 + [6] * 4 + [7] * 300, np.uint8)
    toks = pc.tokenize_segment(seg[:pc.SEGMENT])
    assert toks[:2] == [("L", 5), ("M", 4)]
    assert toks[2:6] == [("L", 6)] * 4                                    # 3 repeats: literals
    assert toks[6:] == [("L", 7), ("M", pc.SEGMENT - 10)]


def test_static_code_is_a_complete_prefix_code():
    code = pc.static_code()
    lens = code.lit_len[code.lit_len > 0].astype(int)
    assert abs(sum(2.0 ** -l for l in lens) - 1.0) < 1e-12                  # Kraft equality: zlib rejects anything else
    assert code.lit_len.max() <= 15 and code.header_nbits < 64 * 8
    t = code.device_tables()
    assert t[0].shape == (256,) and t[2].shape == (259,) and np.all(t[3][pc.MIN_RUN:] > 0)


def test_black_frame_is_tiny():
    img = np.zeros((1080, 1920, 3), np.uint8)
    stream, _ = pc.encode_stream_reference(pc.sub_filter(img))
    assert len(stream) < img.size // 80             # ~23 bits per 256-byte segment
    assert zlib.decompress(bytes(stream)) == pc.sub_filter(img).tobytes()
