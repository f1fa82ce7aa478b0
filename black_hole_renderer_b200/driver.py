"""Drivers above the renderer: single frame, orbit/static video with resume, frame sharding.

Reference: render_image (render.py:4031-4076), render_video (4356-4511), save_image (420-425),
compute_disk_texture_resolution (1128-1149), load_disk_texture (448-459).  Video frames are
independent given the lifecycle state, so with several GPUs (one process per GPU, torchrun) the
frames are dealt to ranks in blocks of 60 -- the cadence at which the normalisation statistics
are recomputed (render.py:4457) -- with no data-path collective.
"""
import hashlib
import json
import math
import os
import struct
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np
from PIL import Image

from .lifecycle import advance_lifecycle_frame, init_lifecycle_system
from .mov import write_png_movie
from .png_codec import png_container_parts
from .renderer import R_DISK_INNER_DEFAULT, R_DISK_OUTER_DEFAULT, Renderer, compute_edge_alpha
from .skybox import load_or_generate_skybox

STATS_PERIOD = 60   # frames between recompute_interactive_stats calls
FRAME_SLOTS = 32    # completion-event slots of bhr_render_async (csrc/common.cuh: BHR_FRAME_SLOTS)


def save_image(image, path):
    """PNG of trunc(clip(image, 0, 1) * 255) (render.py:420-425)."""
    os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
    Image.fromarray((np.clip(image, 0, 1) * 255).astype(np.uint8), "RGB").save(path)
    print(f"Saved: {path}")


def encode_png(img_u8, level=1):
    """8-bit RGB (H, W, 3) -> PNG bytes: filter type 1 (Sub) on every scanline, one zlib stream.

    The frame files of a video run are an intermediate for the muxer (render.py:4462-4467), so speed
    matters more than size.  PIL searches a filter per row; a fixed Sub filter (difference to the
    pixel on the left, one wrapping uint8 subtraction over the frame, 4 ms) makes a rendered frame
    both faster to deflate and a third smaller than unfiltered rows, and costs less than half of
    PIL's time at the same zlib level.  zlib / crc32 release the GIL, so a thread pool scales over
    the host cores.  Any PNG reader decodes the file to the same pixels."""
    import struct
    import zlib
    h, w, c = img_u8.shape
    assert c == 3 and img_u8.dtype == np.uint8
    rows = img_u8.reshape(h, w * 3)
    raw = np.empty((h, w * 3 + 1), np.uint8)
    raw[:, 0] = 1                                  # filter type of the scanline: Sub
    raw[:, 1:] = rows
    raw[:, 4:] -= rows[:, :-3]                     # uint8 arithmetic wraps modulo 256, as the filter requires
    data = zlib.compress(memoryview(raw).cast("B"), level)

    def chunk(tag, payload):
        return struct.pack(">I", len(payload)) + tag + payload + struct.pack(">I", zlib.crc32(payload, zlib.crc32(tag)))
    return (b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 2, 0, 0, 0))
            + chunk(b"IDAT", data) + chunk(b"IEND", b""))


def compute_disk_texture_resolution(width, height, cam_pos, fov, r_inner, r_outer, rs=1.0):
    """(n_phi, n_r): about one azimuthal texel per pixel across the disk's angular extent and
    half a radial texel per pixel, floored at 256 x 128 and rounded up to multiples of 16."""
    distance = math.sqrt(cam_pos[0] ** 2 + cam_pos[1] ** 2 + cam_pos[2] ** 2)
    half_angle = math.atan(r_outer / distance)
    fov_rad = fov * math.pi / 180.0
    n_phi = max(256, int(width * (2 * half_angle / fov_rad)))
    n_r = max(128, int(height * (half_angle / fov_rad) * 0.5))
    return n_phi + (16 - n_phi % 16) % 16, n_r + (16 - n_r % 16) % 16


def load_disk_texture(path):
    """RGB image -> (h, w, 4) float32 with the soft-edge alpha, or None when no file is given."""
    if path and os.path.isfile(path):
        print(f"Loading disk texture: {path}")
        rgb = np.array(Image.open(path).convert("RGB"), dtype=np.float32) / 255.0
        h, w = rgb.shape[:2]
        alpha = np.broadcast_to(compute_edge_alpha(h)[:, None, None].astype(np.float32), (h, w, 1))
        return np.concatenate([rgb, alpha], axis=2)
    return None


def make_renderer(width, height, cam_pos, fov, skybox_path=None, n_stars=6000, tex_w=2048,
                  tex_h=1024, disk_texture_path=None, **renderer_kw):
    """Skybox + (placeholder | loaded) disk texture + Renderer; returns (renderer, use_lifecycle)."""
    skybox, _, _ = load_or_generate_skybox(skybox_path, tex_w, tex_h, n_stars)
    disk_tex = load_disk_texture(disk_texture_path)
    use_lifecycle = disk_tex is None
    if use_lifecycle:
        n_phi, n_r = compute_disk_texture_resolution(
            width, height, cam_pos, fov, renderer_kw.get("r_disk_inner", R_DISK_INNER_DEFAULT),
            renderer_kw.get("r_disk_outer", R_DISK_OUTER_DEFAULT))
        disk_tex = np.zeros((n_r, n_phi, 4), dtype=np.float32)
    return Renderer(width, height, skybox, disk_tex, **renderer_kw), use_lifecycle


def render_image(width, height, cam_pos, fov, step_size, skybox_path=None, n_stars=6000,
                 tex_w=2048, tex_h=1024, r_max=10.0, device="gpu", disk_texture_path=None,
                 r_disk_inner=R_DISK_INNER_DEFAULT, r_disk_outer=R_DISK_OUTER_DEFAULT, disk_tilt=0.0,
                 lens_flare=False, anti_alias="disabled", aa_strength=1.0, disk_rotation_speed=0.1,
                 disk_generation_scale=2, force_regenerate_disk_texture=False,
                 ignore_taichi_cache=False):
    """One frame; the disk texture comes from the lifecycle system at t = 0 unless a file is
    given (render.py:4031-4076).  Deprecated arguments are accepted and ignored."""
    renderer, use_lifecycle = make_renderer(
        width, height, cam_pos, fov, skybox_path, n_stars, tex_w, tex_h, disk_texture_path,
        step_size=step_size, r_max=r_max, device=device, r_disk_inner=r_disk_inner,
        r_disk_outer=r_disk_outer, disk_tilt=disk_tilt, lens_flare=lens_flare,
        anti_alias=anti_alias, aa_strength=aa_strength, disk_rotation_speed=disk_rotation_speed)
    if use_lifecycle:
        factories = init_lifecycle_system(renderer, renderer.dtex_h, renderer.dtex_w, seed=42)
        advance_lifecycle_frame(renderer, factories, t=0.0, dt=0.0, recompute_stats=True)
    t0 = time.time()
    print(f"B200: {width}x{height}, cam_pos={list(cam_pos)}, fov={fov}°, step_size={step_size}")
    img = renderer.render(cam_pos, fov, frame=0)
    print(f"Done in {time.time() - t0:.3f}s")
    return img


def orbit_camera(static_cam_pos, frame, n_frames, orbit_degrees):
    """Camera of orbit frame `frame`: radius |pov| (3-D norm), height pov.z (render.py:4440-4446)."""
    radius = float(np.linalg.norm(static_cam_pos))
    angle = np.radians(frame * (orbit_degrees / n_frames))
    return [radius * np.cos(angle), radius * np.sin(angle), static_cam_pos[2]]


def frame_owner(frame, world_size, block=STATS_PERIOD):
    """Rank that renders `frame` when frames are dealt in `block`-frame blocks round-robin."""
    return (frame // block) % world_size


def load_progress(temp_dir, params):
    """Completed frames recorded under `temp_dir`: the union of progress.json (a finished or
    single-process run) and every progress.<rank>.json (what the ranks of an interrupted sharded
    run left behind).  Returns (completed set, params_match): a file written with other parameters
    invalidates the whole directory, like the reference's check (render.py:4392-4405)."""
    completed, match = set(), True
    if not os.path.isdir(temp_dir):
        return completed, match
    for name in sorted(os.listdir(temp_dir)):
        if not (name == "progress.json" or (name.startswith("progress.") and name.endswith(".json"))):
            continue
        try:
            with open(os.path.join(temp_dir, name)) as f:
                saved = json.load(f)
        except (OSError, ValueError):
            continue                        # a file cut off by the crash: its frames are re-rendered
        if saved.get("params", {}) != params:
            match = False
            continue
        completed |= {int(f) for f in saved.get("completed", [])
                      if os.path.isfile(os.path.join(temp_dir, f"frame_{int(f):04d}.png"))}
    return completed, match


def video_ring(renderer, ring):
    """The renderer's ring of page-locked 8-bit frames, allocated once (page-locking a 6 MB frame takes ~6 ms:
    48 of them cost more than rendering 250 frames) and reused by every later video loop of that renderer."""
    cache = getattr(renderer, "_video_ring", None)
    if cache is None or len(cache) < ring:
        cache = list(cache or []) + [renderer.pinned_frame(np.uint8) for _ in range(ring - len(cache or []))]
        try:
            renderer._video_ring = cache
        except AttributeError:
            pass
    return cache[:ring]


def png_ring(renderer, ring):
    """Ring of page-locked byte buffers for device-encoded PNG streams ({u32 bytes, u32 adler} + stream).  Sized for a
    stream as long as the raw frame (rendered frames compress several times; a longer stream -- noise -- is fetched
    into a temporary by Renderer.png_stream)."""
    cache = getattr(renderer, "_png_ring", None)
    if cache is None or len(cache) < ring:
        size = 8 + min(renderer.png_stream_capacity(), 3 * renderer.width * renderer.height + 65536)
        cache = list(cache or []) + [renderer.pinned_bytes(size) for _ in range(ring - len(cache or []))]
        renderer._png_ring = cache
    return cache[:ring]


def run_video_frames(renderer, n_frames, fov, static_cam_pos, orbit, orbit_degrees, dt, rank=0, world_size=1,
                     completed=(), sink=None, on_rendered=None, ring=None, depth=28, factories=None, timing=None,
                     png=False):
    """The frame loop of render_video (render.py:4437-4458) for the frames `rank` owns.

    Pipelined: frames are enqueued without waiting (texture kernels, render, D2H into one of `ring`
    pinned buffers) and the host runs up to `depth` frames ahead of the device: it does the
    lifecycle ticks / entity packing of the next frames -- and, with several ranks, the ticks of
    the frames other ranks own -- while the device works; a frame is retired (waited for, handed to
    `sink(frame, u8_array)`) when `depth` newer ones are in flight.  `depth` is sized for the sharded
    job: between two of its own 60-frame blocks a rank of an 8-GPU run ticks 420 foreign frames
    (~30 ms of host time); 28 queued frames (~30 ms of device work) keep the GPU busy meanwhile.  `sink` may return a future: the
    buffer is not reused before it resolves.

    Lifecycle state: every rank ticks the factories of EVERY frame (the RNG streams are the
    contract); the device stages are stateless functions of (t, factories) except the statistics,
    which frame 60 b fixes for block b.  A block that still has frames to render therefore starts
    with the full texture pass + statistics of its first frame even when that frame is already done
    (resume), so resumed frames equal those of an uninterrupted run.
    `png=True`: the frames leave the device as the deflate stream of their PNG file (csrc/png.cu) instead of raw
    pixels; `sink(frame, stream)` then receives the zlib stream (uint8 view of the ring buffer).  Only as many bytes
    as recent frames needed (+25 % + 64 KB) are copied per frame; a longer stream is completed by a second copy when
    the frame is retired.
    `timing`: a dict that receives the host-side seconds spent in foreign ticks, in the texture pass + render
    calls of own frames, and blocked on the device (frame retirement, buffer reuse).
    Returns the number of frames rendered."""
    clock = time.perf_counter
    t_tick = t_own = t_wait = 0.0
    if factories is None:
        factories = init_lifecycle_system(renderer, renderer.dtex_h, renderer.dtex_w, seed=42)
    completed = set(completed)
    if ring is None:                           # ~300 MB of page-locked frames: 48 at fhd, 12 at 4K
        frame_bytes = 3 * getattr(renderer, "width", 1920) * getattr(renderer, "height", 1080)
        ring = int(min(48, max(8, 300e6 // frame_bytes)))
    depth = max(1, min(depth, FRAME_SLOTS - 1, ring - 4))
    bufs = png_ring(renderer, ring) if png else video_ring(renderer, ring)
    copied = [0] * ring                        # png: bytes of the stream the enqueued copy covers
    png_need = None                            # png: longest stream among the recently retired frames
    png_bytes = 0
    busy = [None] * ring                       # future of the sink still reading the buffer
    in_flight = []                             # (frame, slot) enqueued, not yet waited for

    def retire(item):
        nonlocal t_wait, png_need, png_bytes
        frame_done, slot = item
        t0 = clock()
        renderer.wait_frame(slot % FRAME_SLOTS)    # (depth + 1 <= FRAME_SLOTS frames in flight)
        t_wait += clock() - t0
        data = bufs[slot]
        if png:
            data = renderer.png_stream(bufs[slot], slot % FRAME_SLOTS, copied[slot])
            png_need = data.size if png_need is None else max(data.size, int(0.9 * png_need))
            png_bytes += data.size
        if sink is not None:
            busy[slot] = sink(frame_done, data)

    def block_has_work(first):
        return any(f not in completed for f in range(first, min(first + STATS_PERIOD, n_frames)))

    rendered = 0
    for frame in range(n_frames):
        t = frame * dt
        mine = frame_owner(frame, world_size) == rank
        block_start = frame % STATS_PERIOD == 0
        if mine and frame not in completed:
            cam_pos = orbit_camera(static_cam_pos, frame, n_frames, orbit_degrees) if orbit else static_cam_pos
            slot = rendered % ring
            if busy[slot] is not None:
                t0 = clock()
                busy[slot].result()
                t_wait += clock() - t0
                busy[slot] = None
            t0 = clock()
            advance_lifecycle_frame(renderer, factories, t, dt, recompute_stats=block_start)
            if png:
                guess = None if png_need is None else int(1.25 * png_need) + 65536
                copied[slot] = renderer.render_png_async(cam_pos, fov, bufs[slot], slot % FRAME_SLOTS, frame=0,
                                                         copy_bytes=guess)
            else:
                renderer.render_u8_async(cam_pos, fov, bufs[slot], slot % FRAME_SLOTS, frame=0)
            t_own += clock() - t0
            in_flight.append((frame, slot))
            if len(in_flight) > depth:
                retire(in_flight.pop(0))
            rendered += 1
            if on_rendered is not None:
                on_rendered(rendered, frame)
        elif mine and block_start and block_has_work(frame):
            # resumed block: its first frame is on disk already, but its statistics are needed
            advance_lifecycle_frame(renderer, factories, t, dt, recompute_stats=True)
        else:
            t0 = clock()
            for f in factories.values():      # keep the RNG streams in step; no device work
                f.tick(now=t, dt=dt)
            t_tick += clock() - t0
    for item in in_flight:
        retire(item)
    for b in busy:
        if b is not None:
            b.result()
    if timing is not None:
        timing.update(host_foreign_ticks_s=t_tick, host_own_frames_s=t_own, host_blocked_on_device_s=t_wait)
        if png:
            timing.update(png_stream_bytes=png_bytes)
    return rendered


def render_video(renderer, width, height, n_frames, fps, output_path, fov, static_cam_pos,
                 orbit=False, resume=False, disk_rotation_speed=0.1, orbit_degrees=360.0,
                 rank=0, world_size=1, barrier=None, **_deprecated_kwargs):
    """Render `n_frames` frames to PNGs under .frames_<md5(output)>/ and mux them.

    Same temp-dir / progress.json protocol as the reference (render.py:4380-4405, 4469-4472).
    With world_size > 1 each rank renders the frames it owns (`frame_owner`), writes
    progress.<rank>.json as it goes, and rank 0 merges the per-rank lists and muxes after
    `barrier()`.  `--resume` reads the merged file AND the per-rank files of an interrupted run; a
    frame is listed only once its PNG is on disk.
    """
    out_dir = os.path.dirname(output_path)
    os.makedirs(out_dir or ".", exist_ok=True)
    temp_dir = os.path.join(out_dir, ".frames_" + hashlib.md5(output_path.encode()).hexdigest()[:16])
    progress_file = os.path.join(temp_dir, "progress.json")
    my_progress = progress_file if world_size == 1 else os.path.join(temp_dir, f"progress.{rank}.json")
    params = {"n_frames": n_frames, "fov": fov, "orbit": orbit,
              "disk_rotation_speed": disk_rotation_speed, "orbit_degrees": orbit_degrees}

    completed = set()
    wipe = os.path.isdir(temp_dir) and not resume
    if resume and os.path.isdir(temp_dir):
        completed, match = load_progress(temp_dir, params)
        if not match:
            print("Warning: parameters changed, starting over")
            completed, wipe = set(), True
        elif completed:
            print(f"Resuming: {len(completed)}/{n_frames} frames already rendered")
    if barrier:
        barrier()                              # every rank has read the old state
    if wipe:
        # stale progress files must not be merged into this run: every rank drops its own, rank 0
        # the merged one and those of ranks a larger, older run had
        for name in os.listdir(temp_dir):
            if not (name.startswith("progress.") and name.endswith(".json")):
                continue
            mid = name[len("progress."):-len(".json")].strip(".")
            owner = int(mid) if mid.isdigit() else -1
            if owner == rank or (rank == 0 and (owner < 0 or owner >= world_size)):
                os.remove(os.path.join(temp_dir, name))
    if barrier:
        barrier()
    os.makedirs(temp_dir, exist_ok=True)
    if hasattr(renderer, "set_option"):
        renderer.set_option("stage_timing", 0)      # no per-stage timing events in the frame loop

    # the deflate stream comes from the device (csrc/png.cu) unless BHR_PNG_DEVICE=0: the host then only frames it
    # (chunk lengths, CRC-32) and writes the file
    device_png = os.environ.get("BHR_PNG_DEVICE", "1") != "0" and hasattr(renderer, "render_png_async")

    # File writers.  With the streams from the device a writer only frames and writes (CRC-32 and write() release the GIL):
    # two to four threads keep up and more only contend with the frame loop for the GIL (measured 862 / 824 / 805 / 837
    # frames/s with 2 / 4 / 8 / 16).  The host-side encoder (tens of ms of zlib per 1080p frame and core) uses the cores.
    default_workers = min(3, os.cpu_count() or 2) if device_png else min(16, os.cpu_count() or 2)
    pool = ThreadPoolExecutor(max_workers=int(os.environ.get("BHR_PNG_WORKERS", str(default_workers))))
    png_level = int(os.environ.get("BHR_PNG_LEVEL", "1"))
    jobs = {}                                  # frame -> future of its PNG file
    mine_done = set()                          # frames of THIS run whose file is on disk

    def save_png(path, img_u8):
        tmp = path + ".part"
        with open(tmp, "wb") as f:
            f.write(encode_png(img_u8, png_level))
        os.replace(tmp, path)                  # a crash never leaves a truncated frame under the final name

    def save_stream(path, stream):
        tmp = path + ".part"
        with open(tmp, "wb") as f:
            for part in png_container_parts(width, height, stream):     # (no copy of the stream: CRC-32 and write release the GIL)
                f.write(part)
        os.replace(tmp, path)

    def sink(frame, data):
        jobs[frame] = pool.submit(save_stream if device_png else save_png,
                                  os.path.join(temp_dir, f"frame_{frame:04d}.png"), data)
        return jobs[frame]

    def harvest():
        for frame in [f for f, j in jobs.items() if j.done()]:
            jobs.pop(frame).result()           # (re-raises a failed write)
            mine_done.add(frame)

    def write_progress():
        harvest()
        mine = sorted(completed | mine_done)    # (what was on disk at start + this rank's frames; files are merged by union)
        tmp = my_progress + ".part"
        with open(tmp, "w") as f:
            json.dump({"params": params, "completed": mine}, f)
        os.replace(tmp, my_progress)

    t_start = time.time()

    def on_rendered(rendered, frame):
        if rendered % 10 == 0:
            write_progress()
        if rendered % 100 == 0:
            print(f"  [rank {rank}] frame {frame}/{n_frames}, {rendered / (time.time() - t_start):.1f} frames/s")

    rendered = run_video_frames(renderer, n_frames, fov, static_cam_pos, orbit, orbit_degrees,
                                disk_rotation_speed, rank, world_size, completed, sink, on_rendered, png=device_png)
    pool.shutdown(wait=True)
    write_progress()
    completed |= mine_done
    if barrier:
        barrier()
    if rank != 0:
        return
    if world_size > 1:
        merged, _ = load_progress(temp_dir, params)
        completed |= merged
        with open(progress_file, "w") as f:
            json.dump({"params": params, "completed": sorted(completed)}, f)
    if rendered:
        print(f"Session rendered {rendered} frames in {(time.time() - t_start) / 60:.2f} min")
    if len(completed) < n_frames:
        print(f"Warning: only {len(completed)}/{n_frames} frames completed. Run again to resume.")
        return
    mux_video(temp_dir, n_frames, fps, output_path, width, height)


def mux_video(temp_dir, n_frames, fps, output_path, width=None, height=None):
    """Frame files -> video file (render.py:4497-4503).  The reference re-encodes the PNGs with x264
    through imageio / pyav; that path is taken when imageio is installed (or BHR_MUX=x264 asks for
    it).  Otherwise -- and this image has neither imageio nor an H.264 encoder -- the frame files,
    whose deflate streams the GPU already produced, become the samples of a QuickTime 'png ' movie
    (mov.write_png_movie): a lossless file copy, no second encoder, readable by ffmpeg / OpenCV.
    BHR_MUX=mp4v re-encodes the frames with OpenCV's MPEG-4 encoder instead (small file, lossy, slow)."""
    mode = os.environ.get("BHR_MUX", "auto")
    files = [os.path.join(temp_dir, f"frame_{frame:04d}.png") for frame in range(n_frames)]
    if mode == "mp4v":
        # a compact lossy file without imageio: OpenCV's bundled ffmpeg has no H.264 encoder, MPEG-4 part 2 it has
        # (single-threaded, ~50 1080p frames/s: an option, not the default)
        import cv2
        first = cv2.imread(files[0], cv2.IMREAD_COLOR)
        writer = cv2.VideoWriter(output_path, cv2.VideoWriter_fourcc(*"mp4v"), float(fps), (first.shape[1], first.shape[0]))
        if not writer.isOpened():
            raise RuntimeError("OpenCV cannot open an mp4v writer for " + output_path)
        for path in files:
            writer.write(cv2.imread(path, cv2.IMREAD_COLOR))
        writer.release()
        print(f"Video saved: {output_path}")
        return
    iio = None
    if mode in ("auto", "x264"):
        try:
            import imageio.v3 as iio
        except Exception:
            if mode == "x264":
                raise
    if iio is not None:
        writer = iio.imopen(output_path, "w", plugin="pyav")
        writer.init_video_stream("libx264", fps=fps)
        for path in files:
            writer.write_frame(iio.imread(path))
        writer.close()
    else:
        if width is None or height is None:
            with open(files[0], "rb") as f:
                width, height = struct.unpack(">II", f.read(24)[16:24])       # IHDR
        t0 = time.time()
        write_png_movie(output_path, files, width, height, fps)
        print(f"Muxed {n_frames} PNG frames into a QuickTime 'png ' movie in {time.time() - t0:.2f} s "
              f"({os.path.getsize(output_path) / 1e6:.1f} MB, lossless; x264 instead: install imageio + av, or\n"
              f"  ffmpeg -i {output_path} -c:v libx264 -crf 18 -pix_fmt yuv420p out.mp4)")
    print(f"Video saved: {output_path}")
