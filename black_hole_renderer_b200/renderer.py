"""`Renderer`: the TaichiRenderer-shaped front end of libbhr.so.

Same constructor arguments, methods, attributes and error behaviour as the reference's
`TaichiRenderer` (render.py:2189-4028) for the render path, so callers (`render_image`,
`render_video`, the reference's kernel-level unit tests) can switch classes.  Every method is a
thin wrapper over one C-ABI call (include/bhr.h); all arithmetic runs in the sm_100a kernels.
There is no CPU path: construction fails without a CUDA device.
"""
import ctypes as C
import os

import numpy as np

from . import _lib as L
from .camera import build_camera_scalar

R_DISK_INNER_DEFAULT = 2.0
R_DISK_OUTER_DEFAULT = 15.0
DISK_COLOR_TEMPERATURE = 6000
NUM_MIP_LEVELS = 5


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


class _Field:
    """Read-only stand-in for a Taichi field: `.to_numpy()` returns the reference's layout."""

    def __init__(self, getter):
        self._getter = getter

    def to_numpy(self):
        return self._getter()


def numpy_quantile_neighbours(n, q):
    """(lo, hi, gamma) of numpy's method="linear" quantile of an n-element float32 array: the two
    sorted positions it reads and the weight it interpolates with.  `q` is the quantile as numpy
    has it at that point: a float32 scalar (np.percentile divides by float32(100), np.quantile
    casts a Python float to the array's dtype), so the virtual index -- and with it the weight --
    is a FLOAT32 quantity (numpy/lib/_function_base_impl.py: _QuantileMethods['linear'],
    _get_indexes, _get_gamma).  tests/test_host.py holds this to np.percentile / np.quantile bit for bit."""
    q = np.asanyarray(q)
    vi = np.asanyarray((n - 1) * q)         # _QuantileMethods['linear']['get_virtual_index']
    prev = np.floor(vi)
    nxt = prev + 1
    if vi >= n - 1:
        prev, nxt = np.float32(n - 1), np.float32(n - 1)
    if vi < 0:
        prev, nxt = np.float32(0), np.float32(0)
    lo, hi = int(prev), int(nxt)
    gamma = np.asanyarray(vi - np.intp(lo) if vi < n - 1 else vi - np.intp(-1), dtype=vi.dtype)
    return lo, hi, gamma


def numpy_linear_lerp(a, b, t):
    """numpy's _lerp on the two neighbours (float32 arithmetic for float32 inputs)."""
    a, b = np.asanyarray(a), np.asanyarray(b)
    diff = np.subtract(b, a)
    out = np.asanyarray(np.add(a, diff * t))
    np.subtract(b, diff * (1 - t), out=out, where=t >= 0.5, casting="unsafe", dtype=type(out.dtype))
    return out[()] if out.ndim == 0 else out


def compute_edge_alpha(height, inner_soft=0.1, outer_soft=0.3):
    """Soft radial edges of the disk alpha (reference: compute_edge_alpha, render.py:437-445):
    cubic ramp over the inner 10 % of rows, quadratic fall-off over the outer 30 %."""
    v = np.linspace(0, 1, height).astype(np.float32)
    alpha = np.ones_like(v)
    lo = v < inner_soft
    hi = v > (1 - outer_soft)
    alpha[lo] = (v[lo] / inner_soft) ** 3.0
    alpha[hi] = ((1 - v[hi]) / outer_soft) ** 2
    return alpha


class Renderer:
    def __init__(self, width, height, skybox, disk_tex, step_size=0.1, r_max=10.0, device="gpu",
                 r_disk_inner=R_DISK_INNER_DEFAULT, r_disk_outer=R_DISK_OUTER_DEFAULT,
                 disk_tilt=0.0, lens_flare=False, anti_alias="disabled", aa_strength=1.0,
                 disk_rotation_speed=0.1, ignore_taichi_cache=False, cuda_device=None):
        # `device` ("cpu"/"gpu") and `ignore_taichi_cache` are accepted for drop-in compatibility;
        # the kernels always run on the CUDA device `cuda_device` (default: LOCAL_RANK or 0).
        self.width = int(width)
        self.height = int(height)
        self.step_size = step_size
        self.r_max = r_max
        self.r_disk_inner = r_disk_inner
        self.r_disk_outer = r_disk_outer
        self.disk_tilt = disk_tilt
        self.anti_alias = anti_alias
        self.aa_strength = aa_strength
        self.disk_rotation_speed = disk_rotation_speed
        self.num_mip_levels = NUM_MIP_LEVELS
        if cuda_device is None:
            cuda_device = int(os.environ.get("LOCAL_RANK", "0"))
        self.cuda_device = cuda_device

        skybox = _f32(skybox)
        disk_tex = _f32(disk_tex)
        self.tex_h, self.tex_w = skybox.shape[:2]
        self.dtex_h, self.dtex_w = disk_tex.shape[:2]

        self._lib = L.load()
        cfg = L.BhrConfig(self.width, self.height, step_size, r_max, r_disk_inner, r_disk_outer,
                          disk_tilt, int(bool(lens_flare)), 0 if anti_alias == "disabled" else 1,
                          aa_strength, disk_rotation_speed, cuda_device)
        ctx = C.c_void_p()
        rc = self._lib.bhr_create(C.byref(cfg), C.byref(ctx))
        if rc != 0:
            raise L.BhrError(f"bhr_create failed ({rc}): {self._lib.bhr_last_error(None).decode()}")
        self._ctx = ctx
        self._lens_flare = bool(lens_flare)
        self._check(self._lib.bhr_upload_skybox(ctx, _fp(skybox), self.tex_h, self.tex_w))
        self._check(self._lib.bhr_upload_disk_texture(ctx, _fp(disk_tex), self.dtex_h, self.dtex_w))
        self._bg_ready = False
        self._last_flags = 0

        W, H = self.width, self.height
        self.image_field = _Field(lambda: self._planar(L.BUF_BG).transpose(2, 1, 0).copy())
        # disk_layer_field is what the reference's field holds at that point: the ray march's layer,
        # or -- after a frame that ran _bloom_kernel -- clamp(layer + 0.4 blur) (its in-place add,
        # render.py:3112-3114; render() itself composites the copy it read before, render.py:3909)
        self._disk_post_bloom = False
        self.disk_layer_field = _Field(lambda: self._planar(
            L.BUF_DISK_POST if self._disk_post_bloom else L.BUF_DISK).transpose(2, 1, 0).copy())
        self.blur_field = _Field(lambda: self._planar(L.BUF_BLUR).transpose(2, 1, 0).copy())
        # final_field[i, j] = frame[H - 1 - j, i]: (W, H, 3) with the y axis flipped for ti.GUI
        # (_compose_final_kernel, render.py:3285-3300); filled by render_to_field.  The device keeps
        # the frame row-major (H, W, 3) -- the layout a CUDA-GL / Vulkan interop blit wants -- and
        # this view applies the reference's index map on read-back
        self.final_field = _Field(lambda: self._download(L.BUF_FINAL, (H, W, 3), np.float32)[::-1]
                                  .transpose(1, 0, 2).copy())
        self.disk_texture_field = _Field(
            lambda: self._download(L.BUF_DISK_TEX, (self.dtex_h, self.dtex_w, 4), np.float32))
        self.disk_mips_field = _Field(self._padded_mips)

    # ------------------------------------------------------------------ plumbing
    def _check(self, rc):
        L.check(self._ctx, rc)

    def __del__(self):
        ctx = getattr(self, "_ctx", None)
        if ctx:
            self._lib.bhr_destroy(ctx)          # (synchronises the context's streams first)
            self._ctx = None
        # page-locked frames handed out by pinned_frame(): numpy views of them must not be used
        # after close() / garbage collection of the renderer
        for p in self.__dict__.pop("_pinned", []):
            self._lib.bhr_host_free(p)

    def close(self):
        self.__del__()

    @property
    def lens_flare(self):
        return self._lens_flare

    @lens_flare.setter
    def lens_flare(self, value):
        self._lens_flare = bool(value)
        self._check(self._lib.bhr_set_lens_flare(self._ctx, int(self._lens_flare)))

    def set_stream(self, cuda_stream):
        """Run all kernels on the given CUDA stream handle (e.g. torch's current stream)."""
        self._check(self._lib.bhr_set_stream(self._ctx, C.c_void_p(cuda_stream)))

    def synchronize(self):
        self._check(self._lib.bhr_synchronize(self._ctx))

    def set_option(self, key, value):
        self._check(self._lib.bhr_set_option(self._ctx, key.encode(), float(value)))

    def device_buffer(self, buf_id):
        """(device pointer, bytes) of one of the context's buffers (see bhr_buffer_id)."""
        p, n = C.c_void_p(), C.c_size_t()
        self._check(self._lib.bhr_buffer(self._ctx, buf_id, C.byref(p), C.byref(n)))
        return p.value, n.value

    def _download(self, buf_id, shape, dtype):
        out = np.empty(shape, dtype=dtype)
        self._check(self._lib.bhr_download(self._ctx, buf_id, out.ctypes.data, out.nbytes))
        return out

    def _planar(self, buf_id):
        return self._download(buf_id, (3, self.height, self.width), np.float32)

    def _padded_mips(self):
        """Mip pyramid in the reference's padded layout (levels, n_r, n_phi, 4)."""
        n_r, n_phi = self.dtex_h, self.dtex_w
        total = sum((n_r >> l) * (n_phi >> l) for l in range(NUM_MIP_LEVELS))
        flat = self._download(L.BUF_DISK_MIPS, (total, 4), np.float32)
        out = np.zeros((NUM_MIP_LEVELS, n_r, n_phi, 4), dtype=np.float32)
        off = 0
        for l in range(NUM_MIP_LEVELS):
            h, w = n_r >> l, n_phi >> l
            out[l, :h, :w] = flat[off:off + h * w].reshape(h, w, 4)
            off += h * w
        return out

    # ------------------------------------------------------------------ textures
    def update_disk_texture(self, new_disk_tex):
        """Replace the disk texture and rebuild its mip pyramid (render.py:2292-2312)."""
        new_disk_tex = _f32(new_disk_tex)
        dtex_h, dtex_w = new_disk_tex.shape[:2]
        assert dtex_h == self.dtex_h and dtex_w == self.dtex_w, \
            f"Texture size mismatch: expected {self.dtex_h}x{self.dtex_w}, got {dtex_h}x{dtex_w}"
        self._check(self._lib.bhr_upload_disk_texture(self._ctx, _fp(new_disk_tex), dtex_h, dtex_w))

    # ------------------------------------------------------------------ hot path
    def _camera(self, cam_pos, fov, frame):
        # (scalar float64 twin of build_camera: this runs on the latency path of every frame)
        pos, right, up, forward, pw, ph, norm = build_camera_scalar(cam_pos, fov, self.width, self.height)
        cam = L.BhrCamera()
        cam.pos[:] = pos            # ctypes rounds float64 -> float32 to nearest, like np.float32()
        cam.right[:] = right
        cam.up[:] = up
        cam.forward[:] = forward
        cam.pixel_w, cam.pixel_h, cam.r_escape = pw, ph, max(self.r_max, norm * 2)
        cam.t_offset = float(frame) * self.disk_rotation_speed
        return cam

    def _flags(self, skip_differentials, skip_bloom, aux=False):
        return ((L.BHR_SKIP_DIFFERENTIALS if skip_differentials else 0)
                | (L.BHR_SKIP_BLOOM if skip_bloom else 0) | (L.BHR_WANT_AUX if aux else 0))

    def render(self, cam_pos, fov, frame=0, skip_differentials=False, skip_bloom=False, out=None,
               aux=False):
        """Render one frame; returns (height, width, 3) float32 in [0, 1] (render.py:3865-3923).

        `out` (extension): a C-contiguous (H, W, 3) float32 array to fill instead of allocating;
        pass pinned memory (`pinned_frame()`) to avoid the driver's staging copy.
        """
        cam = self._camera(cam_pos, fov, frame)
        if out is None:
            out = np.empty((self.height, self.width, 3), dtype=np.float32)
        assert out.dtype == np.float32 and out.flags.c_contiguous \
            and out.shape == (self.height, self.width, 3)
        self._check(self._lib.bhr_render(self._ctx, C.byref(cam),
                                         self._flags(skip_differentials, skip_bloom, aux),
                                         out.ctypes.data, None))
        self._disk_post_bloom = not skip_bloom
        return out

    def render_u8(self, cam_pos, fov, frame=0, skip_differentials=False, skip_bloom=False, out=None):
        """As `render` but returns the 8-bit frame the drivers save: trunc(clip(img)*255)."""
        cam = self._camera(cam_pos, fov, frame)
        if out is None:
            out = np.empty((self.height, self.width, 3), dtype=np.uint8)
        assert out.dtype == np.uint8 and out.flags.c_contiguous \
            and out.shape == (self.height, self.width, 3)
        self._check(self._lib.bhr_render(self._ctx, C.byref(cam),
                                         self._flags(skip_differentials, skip_bloom), None,
                                         out.ctypes.data))
        self._disk_post_bloom = not skip_bloom
        return out

    def render_async(self, cam_pos, fov, out, slot, frame=0, skip_differentials=False, skip_bloom=False):
        """Enqueue a frame whose result lands in the pinned array `out` (`pinned_frame(np.float32)`
        or `pinned_frame(np.uint8)`: the float frame of `render` or the 8-bit frame of `render_u8`);
        returns at once.  `wait_frame(slot)` blocks until it is there.  Video loops use it to
        overlap the host's lifecycle work and the frame copies with the device (driver.py)."""
        cam = self._camera(cam_pos, fov, frame)
        assert out.dtype in (np.uint8, np.float32) and out.flags.c_contiguous \
            and out.shape == (self.height, self.width, 3)
        f32 = out.ctypes.data if out.dtype == np.float32 else None
        u8 = out.ctypes.data if out.dtype == np.uint8 else None
        self._check(self._lib.bhr_render_async(self._ctx, C.byref(cam),
                                               self._flags(skip_differentials, skip_bloom), f32, u8, int(slot)))
        self._disk_post_bloom = not skip_bloom

    def render_u8_async(self, cam_pos, fov, out, slot, frame=0, skip_differentials=False, skip_bloom=False):
        assert out.dtype == np.uint8
        self.render_async(cam_pos, fov, out, slot, frame, skip_differentials, skip_bloom)

    def wait_frame(self, slot):
        self._check(self._lib.bhr_wait_frame(self._ctx, int(slot)))

    # ---- frame files: the PNG's deflate stream from the device (csrc/png.cu, png_codec.py) ----
    def _png_setup(self):
        if getattr(self, "_png_capacity", None) is None:
            from . import png_codec
            t = png_codec.static_code().device_tables()
            u32 = lambda a: a.ctypes.data_as(C.POINTER(C.c_uint32))
            keep = [np.ascontiguousarray(a, np.uint32) for a in t[:4]]
            header = np.frombuffer(t[4], np.uint8).copy()
            self._check(self._lib.bhr_png_setup(self._ctx, *(u32(a) for a in keep),
                                                header.ctypes.data_as(C.POINTER(C.c_uint8)), int(t[5]),
                                                int(t[6]), int(t[7])))
            cap = C.c_size_t()
            self._check(self._lib.bhr_png_capacity(self._ctx, C.byref(cap)))
            self._png_capacity = int(cap.value)
        return self._png_capacity

    def png_stream_capacity(self):
        """Upper bound (bytes) of one frame's zlib stream; `pinned_bytes(8 + capacity)` holds any frame."""
        return self._png_setup()

    def pinned_bytes(self, n):
        """A pinned uint8 array of n bytes (freed with the renderer)."""
        p = C.c_void_p()
        if self._lib.bhr_host_alloc(int(n), C.byref(p)) != 0:
            raise L.BhrError("bhr_host_alloc failed")
        self.__dict__.setdefault("_pinned", []).append(p)
        return np.frombuffer((C.c_char * int(n)).from_address(p.value), dtype=np.uint8)

    def render_png_async(self, cam_pos, fov, buf, slot, frame=0, copy_bytes=None, skip_differentials=False,
                         skip_bloom=False):
        """Like render_u8_async, but what reaches the host is the frame's PNG deflate stream: `buf` (pinned uint8,
        from pinned_bytes) receives {u32 stream_bytes, u32 adler32} + the first copy_bytes of the stream
        (default: all that fits buf).  After wait_frame(slot), `png_file_bytes(buf, slot)` is the file."""
        self._png_setup()
        cam = self._camera(cam_pos, fov, frame)
        assert buf.dtype == np.uint8 and buf.flags.c_contiguous and buf.ndim == 1 and buf.size > 8
        room = buf.size - 8
        copy_bytes = room if copy_bytes is None else max(0, min(int(copy_bytes), room))
        self._check(self._lib.bhr_render_async_png(self._ctx, C.byref(cam), self._flags(skip_differentials, skip_bloom),
                                                   buf.ctypes.data, copy_bytes, int(slot)))
        self._disk_post_bloom = not skip_bloom
        return copy_bytes

    def png_stream(self, buf, slot, copied):
        """The complete zlib stream of slot's frame as a uint8 view/array (after wait_frame): the bytes already in
        `buf`, plus a synchronous fetch of the remainder when the stream was longer than `copied`."""
        n = int(buf[:4].view(np.uint32)[0])
        if n <= copied:
            return buf[8:8 + n]
        out = buf[8:]
        if n > out.size:                        # longer than the ring buffer (incompressible frame): a temporary
            out = np.empty(n, np.uint8)
            out[:copied] = buf[8:8 + copied]
        self._check(self._lib.bhr_png_fetch(self._ctx, int(slot), copied, n - copied, out.ctypes.data + copied))
        return out[:n]

    def png_file_bytes(self, buf, slot, copied):
        from . import png_codec
        return png_codec.png_container(self.width, self.height, self.png_stream(buf, slot, copied))

    def encode_png_current(self):
        """PNG file bytes of the frame currently on the device (after render / render_u8 / render_device)."""
        from . import png_codec
        cap = self._png_setup()
        out = np.empty(cap, np.uint8)
        n = C.c_uint32()
        self._check(self._lib.bhr_png_encode_current(self._ctx, out.ctypes.data, cap, C.byref(n)))
        return png_codec.png_container(self.width, self.height, out[:n.value])

    def render_device(self, cam_pos, fov, frame=0, skip_differentials=False, skip_bloom=False,
                      aux=False):
        """Enqueue one frame and leave the results in device buffers (no host copy, no sync)."""
        cam = self._camera(cam_pos, fov, frame)
        self._check(self._lib.bhr_render(self._ctx, C.byref(cam),
                                         self._flags(skip_differentials, skip_bloom, aux), None, None))
        self._disk_post_bloom = not skip_bloom

    def render_to_field(self, cam_pos, fov, frame=0, skip_differentials=False, skip_bloom=False):
        """render_to_field (render.py:3819-3863): the frame stays on the device (no host copy, no
        synchronisation, no lens flare); `final_field.to_numpy()` reads it back in the reference's
        (W, H, 3), y-flipped layout.  Unlike render(), the reference composites the disk layer AFTER
        _bloom_kernel's in-place `disk = clamp(disk + 0.4 blur)` (render.py:3112-3114, 3857-3863):
        final = clamp(bg + clamp(disk + 0.4 blur) + blur) -- flag BHR_FIELD_COMPOSITE."""
        cam = self._camera(cam_pos, fov, frame)
        self._check(self._lib.bhr_render(self._ctx, C.byref(cam),
                                         self._flags(skip_differentials, skip_bloom) | L.SKIP_FLARE
                                         | L.FIELD_COMPOSITE, None, None))
        self._disk_post_bloom = not skip_bloom

    def pinned_frame(self, dtype=np.float32):
        """A page-locked (H, W, 3) array for `render(..., out=)` / `render_u8(..., out=)`; owned by
        the renderer and released by close() (copy what has to outlive it)."""
        n = self.height * self.width * 3 * np.dtype(dtype).itemsize
        p = C.c_void_p()
        if self._lib.bhr_host_alloc(n, C.byref(p)) != 0:
            raise L.BhrError("bhr_host_alloc failed")
        buf = (C.c_char * n).from_address(p.value)
        arr = np.frombuffer(buf, dtype=dtype).reshape(self.height, self.width, 3)
        self.__dict__.setdefault("_pinned", []).append(p)
        return arr

    def last_aux(self):
        """(class map u8 (H, W), RK4 evaluation counts i32 (H, W)) of the last render(aux=True)."""
        return (self._download(L.BUF_CLASS, (self.height, self.width), np.uint8),
                self._download(L.BUF_STEPS, (self.height, self.width), np.int32))

    def last_total_steps(self):
        v = C.c_uint64()
        self._check(self._lib.bhr_last_total_steps(self._ctx, C.byref(v)))
        return v.value

    def launch_count(self):
        """Kernels this context has launched so far (render path, texture pipeline, peer flags)."""
        v = C.c_uint64()
        self._check(self._lib.bhr_launch_count(self._ctx, C.byref(v)))
        return v.value

    def last_retrace_count(self):
        v = C.c_uint32()
        self._check(self._lib.bhr_last_retrace_count(self._ctx, C.byref(v)))
        return v.value

    def last_stage_ms(self):
        arr = (C.c_float * 5)()
        self._check(self._lib.bhr_last_stage_ms(self._ctx, arr))
        return dict(zip(("ray_march", "bloom_h", "bloom_v_composite", "flare", "total"), list(arr)))

    # ------------------------------------------------------------------ disk-texture pipeline
    def init_background_layer(self, n_r, n_phi, seed=42):
        """Allocate the 13-plane component field and draw the azimuthal-hotspot parameters
        (render.py:3491-3547; the two draws from default_rng(seed) keep their order)."""
        rng = np.random.default_rng(seed)
        self._bg_az_freq = int(rng.integers(2, 5))
        self._bg_az_shear = float(rng.uniform(2.0, 4.0))
        edge = compute_edge_alpha(n_r).astype(np.float32)
        r_norm = np.linspace(0, 1, n_r)
        r_vals = self.r_disk_inner + (self.r_disk_outer - self.r_disk_inner) * r_norm
        omega_rows = np.sqrt(0.5 / (r_vals ** 3 + 1e-6)).astype(np.float32)
        self._bg_omega_all_np = omega_rows
        self._bg_r_norm_all = r_norm
        self._bg_edge_np = edge
        self._bg_n_r, self._bg_n_phi = n_r, n_phi
        self._check(self._lib.bhr_init_background(self._ctx, n_r, n_phi, self._bg_az_freq,
                                                  self._bg_az_shear, _fp(edge), _fp(omega_rows)))
        # initial, deliberately loose stats (render.py:3532-3542)
        tb_init = np.clip(1.0 - np.linspace(0, 1, n_r), 0, 1) ** 1.3 * 0.25
        self._stats = np.array([0.5, 0.5], dtype=np.float32)
        self._row_stats = np.column_stack([np.maximum(tb_init, 0.25).astype(np.float32),
                                           np.maximum(tb_init * 0.8, 0.10).astype(np.float32)])
        self._push_stats()
        self._param_enable_rt = 1
        self._param_color_temp = float(DISK_COLOR_TEMPERATURE)
        self._bg_ready = True
        self._comp_field = _Field(lambda: self._download(L.BUF_COMP, (13, n_r, n_phi), np.float32))
        self._edge_field = _Field(lambda: self._bg_edge_np.copy())
        self._omega_rows_field = _Field(lambda: self._bg_omega_all_np.copy())
        self._param_stats_field = _Field(lambda: self._stats.copy())
        self._param_row_stats_field = _Field(lambda: self._row_stats.copy())

    def _push_stats(self):
        rows = _f32(self._row_stats)
        self._check(self._lib.bhr_set_stats(self._ctx, float(self._stats[0]), float(self._stats[1]),
                                            _fp(rows)))

    def generate_background(self, t):
        """Time-evolved simplex-FBM background planes [0,1,2,3,4,11,12] (render.py:3549-3562)."""
        assert self._bg_ready, "Must call init_background_layer() first"
        self._check(self._lib.bhr_generate_background(self._ctx, float(t)))

    def accumulate_entity_layer(self, factories, now):
        """Sum every alive entity into comp[5:11] on the device (render.py:3564-3653)."""
        from .lifecycle import EntityFactory, pack_entities_array, pack_foreign_entities
        assert self._bg_ready, "Must call init_background_layer() first"
        if all(isinstance(f, EntityFactory) for f in factories.values() if f is not None):
            ents = pack_entities_array(factories, now, self._bg_n_r)
        else:
            # the caller's own factory objects (e.g. the reference's EntityFactory with tabulated profiles)
            cache = self.__dict__.setdefault("_foreign_tables", {})
            ents, tables = pack_foreign_entities(factories, now, self._bg_n_r, self._bg_n_phi, cache)
            if tables is not None:
                self._check(self._lib.bhr_upload_entity_tables(self._ctx, _fp(tables) if len(tables) else None, len(tables)))
        ptr = ents.ctypes.data_as(C.POINTER(L.BhrEntity)) if len(ents) else None
        self._check(self._lib.bhr_accumulate_entities(self._ctx, ptr, len(ents)))

    def recompute_interactive_stats(self):
        """Normalisation statistics from the current component field (render.py:3655-3712):
        98th percentile of the density mix, 95th percentile of the positive structure
        temperature, per-row max / 70 % quantile of the scaled structure temperature.

        The reference reads the whole field back and calls np.percentile / np.quantile; here the
        device returns the exact order statistics on both sides of each quantile's virtual index
        (radix select / per-row sort, csrc/stats.cu) and `numpy_linear_lerp` applies numpy's own
        float32 interpolation to them -- the same numbers without the 63 MB read-back."""
        p98, scale, row_max, row_p70, tb_max = self._device_stats()
        p98, scale = max(p98, 0.01), max(scale, 0.01)
        row_max = np.maximum(row_max, tb_max)
        row_p70 = np.maximum(row_p70, tb_max * 0.8)
        self._stats = np.array([p98, scale], dtype=np.float32)
        self._row_stats = np.column_stack([row_max, row_p70]).astype(np.float32)
        self._push_stats()

    def _device_stats(self, floor_scale=True):
        """(p98 of the density mix, p95 of the positive structure temperature, per-row max and
        70 % quantile of the scaled structure, per-row max of the base temperature) of the current
        component field, numpy's numbers from device order statistics (csrc/stats.cu)."""
        n_tot, n_pos = C.c_uint64(), C.c_uint64()
        self._check(self._lib.bhr_stats_prepare(self._ctx, int(self._param_enable_rt),
                                                C.byref(n_tot), C.byref(n_pos)))
        n_tot, n_pos = int(n_tot.value), int(n_pos.value)
        lo_d, hi_d, g_d = numpy_quantile_neighbours(n_tot, np.true_divide(98, np.float32(100)))
        lo_s, hi_s, g_s = numpy_quantile_neighbours(max(n_pos, 1), np.true_divide(95, np.float32(100)))
        vals = np.zeros(4, dtype=np.float32)
        self._check(self._lib.bhr_stats_select(self._ctx, lo_d, lo_s, _fp(vals)))
        d_hi = vals[1] if hi_d != lo_d else vals[0]
        s_hi = vals[3] if hi_s != lo_s else vals[2]
        p98 = float(numpy_linear_lerp(vals[0], d_hi, g_d))
        scale = float(numpy_linear_lerp(vals[2], s_hi, g_s)) if n_pos > 0 else 1.0
        # struct / (scale + 1e-6) * 0.8: the Python-float divisor becomes a float32 operand
        denom = np.float32((max(scale, 0.01) if floor_scale else scale) + 1e-6)
        lo_r, hi_r, g_r = numpy_quantile_neighbours(self.dtex_w, np.asanyarray(0.7, dtype=np.float32))
        rows = np.zeros((self.dtex_h, 4), dtype=np.float32)
        self._check(self._lib.bhr_stats_rows(self._ctx, float(denom), lo_r, hi_r, _fp(rows)))
        row_p70 = numpy_linear_lerp(rows[:, 1], rows[:, 2], g_r).astype(np.float32)
        return p98, scale, rows[:, 0].copy(), row_p70, rows[:, 3].copy()

    def upload_parametric_state(self, state):
        """Legacy parametric rotation path (render.py:2314-2387): upload the 13 precomputed
        component planes of a DiskTextureRotatingState (any object with its attributes) and the
        statistics of the unrotated state; `update_disk_texture_gpu(t_offset)` then composes the
        Kepler-rotated texture on the device."""
        n_r, n_phi = int(state.n_r), int(state.n_phi)
        packed = np.stack([state.temp_base, state.spiral, state.spiral_temp, state.turbulence,
                           state.turb_temp, state.arcs, state.arcs_temp, state.rt_spikes, state.rt_temp,
                           state.hotspot, state.hotspot_temp, state.az_hotspot, state.disturb_mod],
                          axis=0).astype(np.float32)
        edge, omega = _f32(state.edge), _f32(state.omega_rows)
        self._bg_az_freq, self._bg_az_shear = 2, 2.0        # (unused: no background kernel on this path)
        self._check(self._lib.bhr_init_background(self._ctx, n_r, n_phi, self._bg_az_freq,
                                                  self._bg_az_shear, _fp(edge), _fp(omega)))
        self._bg_edge_np, self._bg_omega_all_np = edge, omega
        self._bg_n_r, self._bg_n_phi = n_r, n_phi
        self._bg_ready = True
        self._check(self._lib.bhr_upload_comp(self._ctx, _fp(_f32(packed))))
        self._param_enable_rt = 1 if state.enable_rt else 0
        self._param_color_temp = float(state.color_temp)
        # render.py:2363-2379: raw percentiles, no floors, no base-temperature term
        p98, scale, row_max, row_p70, _ = self._device_stats(floor_scale=False)
        self._stats = np.array([p98, scale], dtype=np.float32)
        self._row_stats = np.stack([row_max, row_p70], axis=1).astype(np.float32)
        self._push_stats()
        self._comp_field = _Field(lambda: self._download(L.BUF_COMP, (13, n_r, n_phi), np.float32))
        self._edge_field = _Field(lambda: self._bg_edge_np.copy())
        self._omega_rows_field = _Field(lambda: self._bg_omega_all_np.copy())
        self._param_stats_field = _Field(lambda: self._stats.copy())
        self._param_row_stats_field = _Field(lambda: self._row_stats.copy())
        self._parametric_gpu_ready = True

    def update_disk_texture_gpu(self, t_offset):
        """Compose the texture rotated by `t_offset` (each row by its Keplerian angle) and rebuild
        the mips on the device (render.py:3792-3817)."""
        assert getattr(self, "_parametric_gpu_ready", False), \
            "Must call upload_parametric_state() before update_disk_texture_gpu()"
        self._check(self._lib.bhr_compose_texture(self._ctx, float(t_offset), int(self._param_enable_rt),
                                                  float(self._param_color_temp)))

    _SOLO_PAIRS = {0: [], 1: [2], 2: [1], 3: [4], 4: [3], 5: [6], 6: [5], 7: [8], 8: [7],
                   9: [10], 10: [9], 11: [], 12: []}

    def compose_interactive_texture(self, solo_idx=-1):
        """Compose the RGBA disk texture from the component field and rebuild the mips
        (render.py:3714-3767).  solo_idx >= 0 isolates one component (debug aid)."""
        if solo_idx >= 0:
            comp = self._comp_field.to_numpy()
            keep = {solo_idx} | set(self._SOLO_PAIRS.get(solo_idx, []))
            for i in range(13):
                if i not in keep:
                    comp[i] = 1.0 if i == 12 else 0.0
            self._check(self._lib.bhr_upload_comp(self._ctx, _fp(_f32(comp))))
            self.recompute_interactive_stats()
        self._check(self._lib.bhr_compose_texture(self._ctx, 0.0, int(self._param_enable_rt),
                                                  float(self._param_color_temp)))

    def eval_noise(self, coords, mode="simplex", octaves=4, persistence=0.5, lacunarity=2.0):
        """Evaluate the device simplex / FBM noise at (N, 3) points (render.py:3769-3790)."""
        coords = _f32(coords)
        out = np.empty(coords.shape[0], dtype=np.float32)
        self._check(self._lib.bhr_eval_noise(self._ctx, _fp(coords), coords.shape[0],
                                             {"simplex": 0, "fbm": 1, "simplex_packed": 2}[mode], int(octaves),
                                             float(persistence), float(lacunarity), _fp(out)))
        return out


# drop-in alias: code written against the reference can keep the class name
TaichiRenderer = Renderer
