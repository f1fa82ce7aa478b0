"""Device-side PNG deflate streams (csrc/png.cu) against the CPU twin (png_codec, byte for byte) and the stock
decoders: PIL must read back exactly the 8-bit frame the renderer produced -- what the reference's PIL save of each
video frame (render.py:4462-4467) guarantees."""
import io
import zlib

import numpy as np
import pytest

import oracle as O
from util import RESOLUTIONS, synthetic_disk_texture, synthetic_skybox

pytestmark = pytest.mark.gpu


def _scene(W, H, **kw):
    from black_hole_renderer_b200 import Renderer
    pov, fov = [6, 0, 0.5], 90
    n_phi, n_r = O.disk_texture_resolution(W, H, pov, fov, 2.0, 15.0)
    return Renderer(W, H, synthetic_skybox(), synthetic_disk_texture(n_r, n_phi), **kw), pov, fov


@pytest.mark.parametrize("size", [(160, 90), (333, 187), (85, 3), (1, 1), (4, 300)])
def test_device_stream_equals_cpu_twin(size):
    from PIL import Image
    from black_hole_renderer_b200 import png_codec as pc
    W, H = size
    r, pov, fov = _scene(W, H)
    u8 = r.render_u8(pov, fov).copy()
    png = r.encode_png_current()
    assert png == pc.encode_frame_reference(u8)
    assert np.array_equal(np.array(Image.open(io.BytesIO(png)).convert("RGB")), u8)


def test_fhd_stream_decodes_and_async_path_matches():
    from PIL import Image
    from black_hole_renderer_b200 import png_codec as pc
    W, H = RESOLUTIONS["fhd"]
    r, pov, fov = _scene(W, H)
    u8 = r.render_u8(pov, fov).copy()
    png = r.encode_png_current()
    assert np.array_equal(np.array(Image.open(io.BytesIO(png)).convert("RGB")), u8)
    cap = r.png_stream_capacity()
    assert cap >= len(png)
    # pipelined path: full copy, then a deliberately short copy completed by the fetch
    buf = r.pinned_bytes(8 + cap)
    for slot, copy_bytes in ((3, None), (4, 1000), (5, 0)):
        buf[:] = 0
        copied = r.render_png_async(pov, fov, buf, slot, copy_bytes=copy_bytes)
        r.wait_frame(slot)
        stream = r.png_stream(buf, slot, copied)
        assert zlib.decompress(bytes(stream)) == pc.sub_filter(u8).tobytes(), slot
        assert int(buf[4:8].view(np.uint32)[0]) == zlib.adler32(pc.sub_filter(u8).tobytes())
        assert r.png_file_bytes(buf, slot, copied) == png
    print(f"fhd frame: raw {u8.size} B, device PNG {len(png)} B ({u8.size / len(png):.1f}x), "
          f"zlib-1 of the same filtered rows {len(zlib.compress(pc.sub_filter(u8).tobytes(), 1))} B, "
          f"zlib-6 {len(zlib.compress(pc.sub_filter(u8).tobytes(), 6))} B")


def test_noise_frame_longer_than_raw_is_still_exact():
    """An incompressible frame (stream longer than the pixels) through the encoder: upload noise as the 8-bit frame."""
    import ctypes as C
    from black_hole_renderer_b200 import _lib as L, png_codec as pc
    W, H = 320, 200
    r, pov, fov = _scene(W, H)
    r.render_u8(pov, fov)
    rng = np.random.default_rng(3)
    noise = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    import torch
    from black_hole_renderer_b200.dist import device_tensor
    device_tensor(r, L.BUF_FINAL_U8, (H, W, 3), torch.uint8).copy_(torch.from_numpy(noise).cuda())
    torch.cuda.synchronize()
    png = r.encode_png_current()
    assert png == pc.encode_frame_reference(noise)
    assert len(png) > noise.size


def test_video_loop_png_mode_writes_the_same_pixels():
    """driver.run_video_frames(png=True): the streams of an orbit's first frames decode to the u8 frames of the raw
    path, with the adaptive copy size in play."""
    from black_hole_renderer_b200 import Renderer, png_codec as pc
    from black_hole_renderer_b200.driver import run_video_frames
    W, H = 640, 360
    pov, fov = [6.0, 0.0, 0.5], 90.0
    n_phi, n_r = O.disk_texture_resolution(W, H, pov, fov, 2.0, 15.0)
    sky = synthetic_skybox()
    raw, streams = {}, {}
    r1 = Renderer(W, H, sky, np.zeros((n_r, n_phi, 4), np.float32))
    run_video_frames(r1, 40, fov, pov, True, 360.0, 0.1, sink=lambda f, img: raw.__setitem__(f, img.copy()))
    r2 = Renderer(W, H, sky, np.zeros((n_r, n_phi, 4), np.float32))
    timing = {}
    run_video_frames(r2, 40, fov, pov, True, 360.0, 0.1, png=True, timing=timing,
                     sink=lambda f, s: streams.__setitem__(f, bytes(s)))
    assert sorted(streams) == list(range(40))
    for f in range(40):
        assert zlib.decompress(streams[f]) == pc.sub_filter(raw[f]).tobytes(), f
    assert timing["png_stream_bytes"] == sum(len(s) for s in streams.values())
