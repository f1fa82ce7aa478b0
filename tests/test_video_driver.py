"""render_video's camera path, resume protocol and frame sharding with a fake renderer (the
approach of the reference's tests/unit/test_orbit_degrees.py: no rendering, no GPU)."""
import json
import os

import numpy as np


class FakeRenderer:
    def __init__(self, n_r=32, n_phi=128):
        self.dtex_h, self.dtex_w = n_r, n_phi
        self.r_disk_inner, self.r_disk_outer = 2.0, 15.0
        self.cams, self.stats_calls, self.bg_times = [], 0, []

    def init_background_layer(self, n_r, n_phi, seed=42): pass
    def generate_background(self, t): self.bg_times.append(t)
    def accumulate_entity_layer(self, factories, now): pass
    def recompute_interactive_stats(self): self.stats_calls += 1
    def compose_interactive_texture(self, solo_idx=-1): pass

    def pinned_frame(self, dtype=np.uint8):
        return np.zeros((4, 8, 3), dtype=dtype)

    def render_u8_async(self, cam_pos, fov, out, slot, frame=0):
        self.cams.append(list(cam_pos))
        self.slots = getattr(self, "slots", []) + [slot]
        out[...] = len(self.cams) % 251          # frame-dependent content: checks the buffer ring

    def wait_frame(self, slot):
        assert slot in self.slots


def _run(tmp_path, renderer, n_frames=12, resume=False, rank=0, world=1, orbit=True, degrees=90.0):
    from black_hole_renderer_b200.driver import render_video
    out = str(tmp_path / "v.mp4")
    render_video(renderer, 8, 4, n_frames=n_frames, fps=4, output_path=out, fov=90.0,
                 static_cam_pos=[6, 0, 0.5], orbit=orbit, resume=resume, disk_rotation_speed=0.1,
                 orbit_degrees=degrees, rank=rank, world_size=world)
    frames_dir = [d for d in os.listdir(tmp_path) if d.startswith(".frames_")][0]
    return tmp_path / frames_dir


def test_orbit_camera_path_and_files(tmp_path):
    r = FakeRenderer()
    d = _run(tmp_path, r, n_frames=12, degrees=90.0)
    assert len(r.cams) == 12
    radius = np.sqrt(36.25)
    for f, cam in enumerate(r.cams):
        a = np.radians(f * 90.0 / 12)
        np.testing.assert_allclose(cam, [radius * np.cos(a), radius * np.sin(a), 0.5], atol=1e-12)
    assert sorted(p for p in os.listdir(d) if p.endswith(".png")) == [f"frame_{f:04d}.png" for f in range(12)]
    prog = json.load(open(d / "progress.json"))
    assert sorted(prog["completed"]) == list(range(12))
    from PIL import Image
    for f in range(12):     # every PNG holds its own frame although the pinned buffers are recycled
        assert np.all(np.array(Image.open(d / f"frame_{f:04d}.png")) == (f + 1) % 251)
    assert prog["params"] == {"n_frames": 12, "fov": 90.0, "orbit": True, "disk_rotation_speed": 0.1,
                              "orbit_degrees": 90.0}
    # init (1) + frame 0 (frame % 60 == 0)
    assert r.stats_calls == 2
    # negative orbit_degrees reverses the direction (reference test_orbit_degrees.py)
    r2 = FakeRenderer()
    _run(tmp_path / "neg" if (tmp_path / "neg").mkdir() is None else tmp_path, r2, n_frames=4, degrees=-40.0)
    assert r2.cams[1][1] < 0


def test_resume_skips_completed_and_replays_lifecycle(tmp_path):
    r = FakeRenderer()
    d = _run(tmp_path, r, n_frames=10)
    prog = json.load(open(d / "progress.json"))
    prog["completed"] = [0, 1, 2, 3, 4]
    json.dump(prog, open(d / "progress.json", "w"))
    r2 = FakeRenderer()
    _run(tmp_path, r2, n_frames=10, resume=True)
    assert len(r2.cams) == 5                       # frames 5..9 only
    # init + the texture pass of the block's first frame (its statistics; frame 0 itself is on disk)
    # + the five new frames; frames 1..4 are replayed on the host only (factory ticks)
    assert len(r2.bg_times) == 1 + 1 + 5
    assert r2.stats_calls == 2
    # changed parameters restart from scratch
    r3 = FakeRenderer()
    _run(tmp_path, r3, n_frames=10, resume=True, degrees=180.0)
    assert len(r3.cams) == 10


def test_static_camera_and_sharding(tmp_path):
    r = FakeRenderer()
    _run(tmp_path, r, n_frames=5, orbit=False)
    assert all(c == [6, 0, 0.5] for c in r.cams)
    # two ranks, 150 frames: blocks of 60 -> rank 0 renders 0-59 and 120-149, rank 1 renders 60-119
    a, b = FakeRenderer(), FakeRenderer()
    (tmp_path / "s").mkdir()
    from black_hole_renderer_b200.driver import render_video
    out = str(tmp_path / "s" / "v.mp4")
    kw = dict(n_frames=150, fps=4, output_path=out, fov=90.0, static_cam_pos=[6, 0, 0.5], orbit=True,
              disk_rotation_speed=0.1, orbit_degrees=360.0, world_size=2)
    render_video(b, 8, 4, rank=1, **kw)
    render_video(a, 8, 4, rank=0, **kw)
    assert len(a.cams) == 90 and len(b.cams) == 60
    d = [x for x in os.listdir(tmp_path / "s") if x.startswith(".frames_")][0]
    prog = json.load(open(tmp_path / "s" / d / "progress.json"))
    assert sorted(prog["completed"]) == list(range(150))
    # every rank recomputes the statistics on the first frame of each block it owns
    assert a.stats_calls == 1 + 2 and b.stats_calls == 1 + 1


def test_sharded_resume_reads_the_per_rank_progress_files(tmp_path):
    """An interrupted 2-rank run leaves only progress.<rank>.json behind (the merged progress.json
    is written after the final barrier).  --resume must pick those up, render exactly the missing
    frames with the cameras and lifecycle times of an uninterrupted run, and list a frame only when
    its PNG exists."""
    from black_hole_renderer_b200.driver import load_progress, render_video
    out = str(tmp_path / "v.mp4")
    kw = dict(n_frames=150, fps=4, output_path=out, fov=90.0, static_cam_pos=[6, 0, 0.5], orbit=True,
              disk_rotation_speed=0.1, orbit_degrees=360.0, world_size=2)
    full = [FakeRenderer(), FakeRenderer()]
    for rank in (1, 0):
        render_video(full[rank], 8, 4, rank=rank, **kw)
    d = tmp_path / [x for x in os.listdir(tmp_path) if x.startswith(".frames_")][0]
    # simulate the crash: no merged file; rank 0 got through frames 0..39, rank 1 through 60..99,
    # and frame 70's PNG was never written although it is listed
    os.remove(d / "progress.json")
    params = json.load(open(d / "progress.0.json"))["params"]
    json.dump({"params": params, "completed": list(range(40))}, open(d / "progress.0.json", "w"))
    json.dump({"params": params, "completed": list(range(60, 100))}, open(d / "progress.1.json", "w"))
    os.remove(d / "frame_0070.png")
    done, match = load_progress(str(d), params)
    assert match and done == set(range(40)) | (set(range(60, 100)) - {70})
    res = [FakeRenderer(), FakeRenderer()]
    for rank in (1, 0):
        render_video(res[rank], 8, 4, rank=rank, resume=True, **kw)
    want0 = [c for f, c in zip(list(range(60)) + list(range(120, 150)), full[0].cams) if f >= 40]
    want1 = [c for f, c in zip(range(60, 120), full[1].cams) if f == 70 or f >= 100]
    assert res[0].cams == want0 and res[1].cams == want1
    # lifecycle times of the texture passes: each resumed block starts with its statistics frame
    assert res[0].bg_times[1:] == [0.0] + [f * 0.1 for f in range(40, 60)] + [f * 0.1 for f in range(120, 150)]
    assert res[1].bg_times[1:] == [6.0, 7.0] + [f * 0.1 for f in range(100, 120)]
    assert res[0].stats_calls == 1 + 2 and res[1].stats_calls == 1 + 1
    prog = json.load(open(d / "progress.json"))
    assert sorted(prog["completed"]) == list(range(150))
    assert all(os.path.isfile(d / f"frame_{f:04d}.png") for f in range(150))


def test_encode_png_round_trips_through_pil(tmp_path):
    """driver.encode_png (Sub-filtered scanlines + zlib, the frame files of a video run) decodes to the same
    pixels with an independent PNG reader: ragged sizes, every zlib level, extreme values."""
    from PIL import Image
    from black_hole_renderer_b200.driver import encode_png
    rng = np.random.default_rng(3)
    for (h, w), level in (((1, 1), 1), ((7, 13), 0), ((64, 48), 1), ((33, 257), 6), ((360, 640), 1)):
        img = rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
        img[0, 0] = (0, 255, 0)
        img[::3] = np.minimum(img[::3], 8)                 # smooth, compressible rows too
        path = tmp_path / f"t_{h}x{w}.png"
        path.write_bytes(encode_png(img, level))
        back = Image.open(path)
        assert back.mode == "RGB" and back.size == (w, h)
        assert np.array_equal(np.array(back), img)
    sliced = rng.integers(0, 256, size=(20, 30, 3), dtype=np.uint8)[::2, ::3]      # non-contiguous input
    (tmp_path / "s.png").write_bytes(encode_png(sliced))
    assert np.array_equal(np.array(Image.open(tmp_path / "s.png")), sliced)


class FakePngRenderer(FakeRenderer):
    """Stands in for the device PNG path: render_png_async fills the ring buffer with {bytes, adler} + the stream of
    the CPU twin encoder, honouring copy_bytes; png_stream completes short copies the way Renderer.png_stream does."""
    width, height = 8, 4

    def png_stream_capacity(self):
        from black_hole_renderer_b200 import png_codec
        return png_codec.stream_capacity(self.width, self.height)

    def pinned_bytes(self, n):
        return np.zeros(n, np.uint8)

    def render_png_async(self, cam_pos, fov, buf, slot, frame=0, copy_bytes=None):
        from black_hole_renderer_b200 import png_codec
        self.cams.append(list(cam_pos))
        self.slots = getattr(self, "slots", []) + [slot]
        img = np.full((self.height, self.width, 3), len(self.cams) % 251, np.uint8)
        img[0, 0, 0] = (7 * len(self.cams)) % 256
        stream, _ = png_codec.encode_stream_reference(png_codec.sub_filter(img))
        stream = np.frombuffer(bytes(stream), np.uint8)
        self.full = getattr(self, "full", {})
        self.full[slot] = stream
        room = buf.size - 8
        copy_bytes = room if copy_bytes is None else min(copy_bytes, room)
        if getattr(self, "starve", False):
            copy_bytes = min(copy_bytes, 5)     # force the second-copy path
        buf[:4] = np.array([stream.size], np.uint32).view(np.uint8)
        n = min(copy_bytes, stream.size)
        buf[8:8 + n] = stream[:n]
        self.copies = getattr(self, "copies", []) + [copy_bytes]
        return copy_bytes

    def png_stream(self, buf, slot, copied):
        n = int(buf[:4].view(np.uint32)[0])
        if n > copied:
            self.fetches = getattr(self, "fetches", 0) + 1
            buf[8 + copied:8 + n] = self.full[slot][copied:n]
        return buf[8:8 + n]


def test_device_png_streams_become_the_frame_files(tmp_path):
    from PIL import Image
    for starve in (False, True):
        r = FakePngRenderer()
        r.starve = starve
        sub = tmp_path / f"s{int(starve)}"
        sub.mkdir()
        d = _run(sub, r, n_frames=40, degrees=90.0)
        for f in range(40):
            img = np.array(Image.open(d / f"frame_{f:04d}.png"))
            assert img.shape == (4, 8, 3) and img[0, 0, 0] == (7 * (f + 1)) % 256 and np.all(img[1:] == (f + 1) % 251)
        assert (getattr(r, "fetches", 0) > 0) == starve
        # after the first frames retire, the copy size follows the streams instead of the buffer size
        assert starve or r.copies[-1] <= r.copies[0]


def test_video_file_is_a_lossless_movie_of_the_frame_files(tmp_path):
    """Without imageio the frame files become the samples of a QuickTime 'png ' movie (mov.py): the
    index lists every frame file byte for byte, and a stock demuxer + PNG decoder (OpenCV's ffmpeg)
    reads back exactly the rendered frames at the requested rate."""
    from black_hole_renderer_b200 import mov
    os.environ["BHR_MUX"] = "png"
    try:
        d = _run(tmp_path, FakeRenderer(), n_frames=12, degrees=90.0)
    finally:
        del os.environ["BHR_MUX"]
    out = tmp_path / "v.mp4"
    assert out.exists() and not os.path.exists(str(out) + ".part")
    w, h, fps, index = mov.read_png_movie_index(out)
    assert (w, h, fps, len(index)) == (8, 4, 4.0, 12)
    blob = open(out, "rb").read()
    for f, (off, size) in enumerate(index):
        assert blob[off:off + size] == open(d / f"frame_{f:04d}.png", "rb").read()
    try:
        import cv2
    except Exception:
        return
    cap = cv2.VideoCapture(str(out))
    assert cap.isOpened() and cap.get(cv2.CAP_PROP_FPS) == 4.0 and int(cap.get(cv2.CAP_PROP_FRAME_COUNT)) == 12
    for f in range(12):
        ok, bgr = cap.read()
        assert ok and bgr.shape == (4, 8, 3) and np.all(bgr == (f + 1) % 251)
    assert not cap.read()[0]


def test_png_movie_accepts_streams_and_rejects_garbage(tmp_path):
    from black_hole_renderer_b200 import mov, png_codec
    import pytest
    rng = np.random.default_rng(3)
    frames = [rng.integers(0, 256, (9, 16, 3), dtype=np.uint8) for _ in range(3)]
    streams = [png_codec.encode_stream_reference(png_codec.sub_filter(f))[0] for f in frames]
    parts = [png_codec.png_container_parts(16, 9, s) for s in streams]        # (head, stream view, tail) as the video loop writes them
    whole = [png_codec.png_container(16, 9, s) for s in streams]
    assert whole[0] == png_codec.encode_frame_reference(frames[0])
    assert mov.write_png_movie(str(tmp_path / "a.mov"), parts, 16, 9, 29.97) == 3
    assert mov.write_png_movie(str(tmp_path / "b.mov"), whole, 16, 9, 29.97) == 3
    assert open(tmp_path / "a.mov", "rb").read() == open(tmp_path / "b.mov", "rb").read()
    w, h, fps, index = mov.read_png_movie_index(tmp_path / "a.mov")
    assert (w, h, len(index)) == (16, 9, 3) and abs(fps - 29.97) < 1e-9
    bad = tmp_path / "bad.png"
    bad.write_bytes(b"not a png")
    with pytest.raises(ValueError):
        mov.write_png_movie(str(tmp_path / "c.mov"), [str(bad)], 16, 9, 30)
    with pytest.raises(ValueError):
        mov.write_png_movie(str(tmp_path / "d.mov"), [], 16, 9, 30)
    assert not (tmp_path / "c.mov").exists() and not (tmp_path / "d.mov").exists()


def test_mp4v_mux_option(tmp_path):
    """BHR_MUX=mp4v: a compact lossy file through OpenCV (when it is installed)."""
    import pytest
    cv2 = pytest.importorskip("cv2")
    from black_hole_renderer_b200.driver import encode_png, mux_video
    frames = []
    for f in range(6):
        img = np.zeros((48, 64, 3), np.uint8)
        img[8:40, 4 + 6 * f:30 + 6 * f] = (200, 120, 40)
        frames.append(img)
        with open(tmp_path / f"frame_{f:04d}.png", "wb") as fh:
            fh.write(encode_png(img))
    os.environ["BHR_MUX"] = "mp4v"
    try:
        mux_video(str(tmp_path), 6, 12, str(tmp_path / "v.mp4"))
    finally:
        del os.environ["BHR_MUX"]
    cap = cv2.VideoCapture(str(tmp_path / "v.mp4"))
    assert cap.isOpened() and int(cap.get(cv2.CAP_PROP_FRAME_COUNT)) == 6
    for f in range(6):
        ok, bgr = cap.read()
        assert ok and np.abs(bgr[..., ::-1].astype(int) - frames[f].astype(int)).mean() < 6.0
