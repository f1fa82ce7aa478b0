"""ctypes binding of libbhr.so (include/bhr.h).  No CPU fallback: if the library is missing or no
CUDA device is usable, loading / context creation raises."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BHR_LIB") or os.path.join(_HERE, "libbhr.so")      # (BHR_LIB: A/B of two builds, tools/lib_ab.py)

BHR_SKIP_DIFFERENTIALS = 1
BHR_SKIP_BLOOM = 2
BHR_WANT_AUX = 4
SKIP_FLARE = BHR_SKIP_FLARE = 8
FIELD_COMPOSITE = BHR_FIELD_COMPOSITE = 16
FLARE_FROM_DEVICE = BHR_FLARE_FROM_DEVICE = 32

(BUF_BG, BUF_DISK, BUF_HBLUR, BUF_FINAL, BUF_FINAL_U8, BUF_CLASS, BUF_STEPS, BUF_DISK_TEX,
 BUF_DISK_MIPS, BUF_COMP, BUF_BLUR, BUF_DISK_POST, BUF_FLARE_SUMS) = range(13)


class BhrConfig(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("step_size", C.c_float),
                ("r_max", C.c_float), ("r_disk_inner", C.c_float), ("r_disk_outer", C.c_float),
                ("disk_tilt_deg", C.c_float), ("lens_flare", C.c_int32), ("anti_alias", C.c_int32),
                ("aa_strength", C.c_float), ("disk_rotation_speed", C.c_float),
                ("device", C.c_int32)]


class BhrCamera(C.Structure):
    _fields_ = [("pos", C.c_float * 3), ("right", C.c_float * 3), ("up", C.c_float * 3),
                ("forward", C.c_float * 3), ("pixel_w", C.c_float), ("pixel_h", C.c_float),
                ("r_escape", C.c_float), ("t_offset", C.c_float)]


class BhrEntity(C.Structure):
    _fields_ = [("kind", C.c_int32), ("row_begin", C.c_int32), ("row_end", C.c_int32),
                ("age", C.c_double), ("scale", C.c_double), ("p", C.c_double * 8)]


class BhrError(RuntimeError):
    pass


# name -> (restype, argtypes); every symbol include/bhr.h declares
_P = C.c_void_p
_FP = C.POINTER(C.c_float)
SIGNATURES = {
    "bhr_create": (C.c_int, [C.POINTER(BhrConfig), C.POINTER(_P)]),
    "bhr_destroy": (None, [_P]),
    "bhr_last_error": (C.c_char_p, [_P]),
    "bhr_set_stream": (C.c_int, [_P, _P]),
    "bhr_synchronize": (C.c_int, [_P]),
    "bhr_set_lens_flare": (C.c_int, [_P, C.c_int]),
    "bhr_version": (C.c_int, []),
    "bhr_set_option": (C.c_int, [_P, C.c_char_p, C.c_double]),
    "bhr_host_alloc": (C.c_int, [C.c_size_t, C.POINTER(_P)]),
    "bhr_host_free": (C.c_int, [_P]),
    "bhr_upload_skybox": (C.c_int, [_P, _FP, C.c_int, C.c_int]),
    "bhr_upload_disk_texture": (C.c_int, [_P, _FP, C.c_int, C.c_int]),
    "bhr_render": (C.c_int, [_P, C.POINTER(BhrCamera), C.c_uint32, _P, _P]),
    "bhr_device_pci_bus_id": (C.c_int, [C.c_int, C.c_char_p, C.c_int]),
    "bhr_host_register": (C.c_int, [_P, C.c_size_t]),
    "bhr_host_unregister": (C.c_int, [_P]),
    "bhr_peer_export": (C.c_int, [_P, _P]),
    "bhr_peer_attach": (C.c_int, [_P, C.c_int, C.c_int, _P]),
    "bhr_render_tiled_peer": (C.c_int, [_P, C.POINTER(BhrCamera), C.c_uint32, _P, _P]),
    "bhr_peer_detach": (C.c_int, [_P]),
    "bhr_render_tiled_peer_async": (C.c_int, [_P, C.POINTER(BhrCamera), C.c_uint32, _P, _P]),
    "bhr_peer_wait_frame": (C.c_int, [_P, C.c_int]),
    "bhr_peer_set_distributed_egress": (C.c_int, [_P, C.c_int]),
    "bhr_peer_probe_read": (C.c_int, [_P, C.c_int, C.c_int, C.POINTER(C.c_double)]),
    "bhr_peer_set_tiles": (C.c_int, [_P, C.POINTER(C.c_int)]),
    "bhr_row_costs": (C.c_int, [_P, C.c_int, C.c_int, C.POINTER(C.c_uint64)]),
    "bhr_render_async": (C.c_int, [_P, C.POINTER(BhrCamera), C.c_uint32, _P, _P, C.c_int]),
    "bhr_wait_frame": (C.c_int, [_P, C.c_int]),
    "bhr_png_setup": (C.c_int, [_P, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32),
                                C.POINTER(C.c_uint32), C.POINTER(C.c_uint8), C.c_uint32, C.c_uint32, C.c_uint32]),
    "bhr_png_capacity": (C.c_int, [_P, C.POINTER(C.c_size_t)]),
    "bhr_render_async_png": (C.c_int, [_P, C.POINTER(BhrCamera), C.c_uint32, _P, C.c_size_t, C.c_int]),
    "bhr_png_fetch": (C.c_int, [_P, C.c_int, C.c_size_t, C.c_size_t, _P]),
    "bhr_png_encode_current": (C.c_int, [_P, _P, C.c_size_t, C.POINTER(C.c_uint32)]),
    "bhr_render_rows_stage1": (C.c_int, [_P, C.POINTER(BhrCamera), C.c_uint32, C.c_int, C.c_int]),
    "bhr_render_rows_stage2": (C.c_int, [_P, C.c_uint32, C.c_int, C.c_int, C.POINTER(C.c_double)]),
    "bhr_flare_sums": (C.c_int, [_P, C.c_int, C.c_int, C.POINTER(C.c_double)]),
    "bhr_flare_sums_device": (C.c_int, [_P, C.c_int, C.c_int]),
    "bhr_bloom_radius": (C.c_int, [_P]),
    "bhr_buffer": (C.c_int, [_P, C.c_int, C.POINTER(_P), C.POINTER(C.c_size_t)]),
    "bhr_download": (C.c_int, [_P, C.c_int, _P, C.c_size_t]),
    "bhr_last_total_steps": (C.c_int, [_P, C.POINTER(C.c_uint64)]),
    "bhr_launch_count": (C.c_int, [_P, C.POINTER(C.c_uint64)]),
    "bhr_last_raymarch_timeline": (C.c_int, [_P, C.POINTER(C.c_uint64), C.c_int]),
    "bhr_last_retrace_count": (C.c_int, [_P, C.POINTER(C.c_uint32)]),
    "bhr_last_stage_ms": (C.c_int, [_P, C.POINTER(C.c_float)]),
    "bhr_init_background": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_float, _FP, _FP]),
    "bhr_generate_background": (C.c_int, [_P, C.c_float]),
    "bhr_accumulate_entities": (C.c_int, [_P, C.POINTER(BhrEntity), C.c_int]),
    "bhr_upload_entity_tables": (C.c_int, [_P, _FP, C.c_size_t]),
    "bhr_set_stats": (C.c_int, [_P, C.c_float, C.c_float, _FP]),
    "bhr_upload_comp": (C.c_int, [_P, _FP]),
    "bhr_compose_texture": (C.c_int, [_P, C.c_float, C.c_int, C.c_float]),
    "bhr_eval_noise": (C.c_int, [_P, _FP, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, _FP]),
    "bhr_stats_prepare": (C.c_int, [_P, C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "bhr_stats_select": (C.c_int, [_P, C.c_uint64, C.c_uint64, _FP]),
    "bhr_stats_rows": (C.c_int, [_P, C.c_float, C.c_int, C.c_int, _FP]),
    "bhr_measure_fp32_peak": (C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_double)]),
    "bhr_selftest_div6": (C.c_int, [C.c_int, C.POINTER(C.c_ulonglong), C.POINTER(C.c_ulonglong)]),
}

_lib = None


def load():
    """dlopen libbhr.so and declare every prototype.  Raises if the library has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise BhrError(
                f"{LIB_PATH} not found: build it with `python -m black_hole_renderer_b200.build` "
                "(there is no CPU fallback)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(ctx, rc):
    if rc != 0:
        msg = load().bhr_last_error(ctx)
        raise BhrError(f"libbhr error {rc}: {msg.decode() if msg else '?'}")
