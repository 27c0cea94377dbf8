"""ctypes binding of include/gpis_b200.h.  No CPU fallback: if the shared library is
missing this raises, and gpis_create fails with GPIS_ERR_NO_DEVICE without a GPU."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libgpis_b200.so")

GPIS_OK = 0
ERR_INVALID, ERR_CUDA, ERR_NOMEM, ERR_STATE, ERR_NOT_PD, ERR_NO_DEVICE, ERR_TIMEOUT = -1, -2, -3, -4, -5, -6, -7
BETA_FIXED, BETA_GOTOVOS, BETA_SELECTCHOL = 0, 1, 2
VAR_LATENT, VAR_PREDICTIVE = 0, 1
PREC_FP64, PREC_TF32 = 0, 1
LOOP_AUTO, LOOP_STEPS = 0, 1
APPEND_AUTO, APPEND_TWO_KERNELS, APPEND_TWO_PASS = 0, 1, 2
BATCH_AUTO, BATCH_PLAIN, BATCH_DMMA, BATCH_DFMA = 0, 1, 2, 3
CONFIG_SIZE_V1 = 152

# every symbol include/gpis_b200.h declares
SYMBOLS = [
    "gpis_version", "gpis_status_string", "gpis_last_error", "gpis_create", "gpis_destroy",
    "gpis_set_candidates", "gpis_set_grid_candidates", "gpis_select", "gpis_select_begin", "gpis_step_append", "gpis_step_update",
    "gpis_select_end", "gpis_exchange_buffers", "gpis_get_active", "gpis_get_factor", "gpis_fit",
    "gpis_predict", "gpis_get_levelset", "gpis_get_candidate_posterior", "gpis_get_stats", "gpis_get_timing",
    "gpis_set_profiling", "gpis_get_history", "gpis_peer_export", "gpis_peer_import", "gpis_get_phase_timing",
    "gpis_posterior_cov", "gpis_sample_shapes", "gpis_get_sampling_timing", "gpis_batch_select",
    "gpis_get_persist_timing", "gpis_get_persist_balance", "gpis_build_tsdf", "gpis_set_candidates_f32", "gpis_set_grid_candidates_f32",
    "gpis_get_active_f32", "gpis_predict_f32", "gpis_get_candidate_posterior_f32",
]


class GpisConfig(C.Structure):
    _fields_ = [
        ("struct_size", C.c_int32), ("dim", C.c_int32),
        ("num_candidates", C.c_int64), ("total_candidates", C.c_int64), ("id_base", C.c_int64),
        ("max_active", C.c_int32), ("use_gradients", C.c_int32), ("precision", C.c_int32),
        ("beta_mode", C.c_int32), ("variance_mode", C.c_int32),
        ("rank", C.c_int32), ("world_size", C.c_int32), ("device", C.c_int32),
        ("ell", C.c_double), ("sf2", C.c_double),
        ("mean_w", C.c_double * 3), ("mean_c", C.c_double),
        ("noise", C.c_double), ("level", C.c_double), ("eps", C.c_double), ("beta", C.c_double),
        ("beta_delta", C.c_double),
        ("loop_mode", C.c_int32), ("append_mode", C.c_int32), ("batch_kernel", C.c_int32),
        ("spin_timeout_ms", C.c_int32),
    ]


class GpisTsdfParams(C.Structure):
    """gpis_tsdf_params (include/gpis_b200.h): varParams of matlab/core/auto_tsdf.m."""
    _fields_ = [
        ("struct_size", C.c_int32), ("edge_win", C.c_int32), ("specular_noise", C.c_int32), ("noise_grad_mode", C.c_int32),
        ("noise_scale", C.c_double), ("occlusion_scale", C.c_double), ("interior_rate", C.c_double),
        ("sparsity_rate", C.c_double), ("horiz_scale", C.c_double), ("vert_scale", C.c_double),
        ("occ_y_low", C.c_double * 3), ("occ_y_high", C.c_double * 3), ("occ_x_low", C.c_double * 3), ("occ_x_high", C.c_double * 3),
    ]


class GpisError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__(f"gpis_b200 status {status}: {msg}")
        self.status = status


_lib = None


def load():
    """Load the C-ABI library (building nothing: run __graft_entry__.build() or
    python -m gpis_b200.build first)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        try:                      # building is not a fallback: the product is still the CUDA library
            from . import build
            build.build()
        except Exception:
            pass
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build it with `python -m gpis_b200.build` "
                          "(there is no CPU fallback for this path)")
    lib = C.CDLL(LIB_PATH)
    vp, i64p, i32p, dp, u8p = C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int32), C.c_void_p, C.c_void_p
    lib.gpis_version.restype = C.c_char_p
    lib.gpis_status_string.restype = C.c_char_p
    lib.gpis_status_string.argtypes = [C.c_int32]
    lib.gpis_last_error.restype = C.c_char_p
    lib.gpis_last_error.argtypes = [vp]
    sig = {
        "gpis_create": [C.POINTER(vp), C.POINTER(GpisConfig)],
        "gpis_destroy": [vp],
        "gpis_set_candidates": [vp, dp, dp, dp, dp, vp, vp],
        "gpis_set_grid_candidates": [vp, vp, vp, vp, C.c_int32, C.c_int32, dp, dp, dp, vp],
        "gpis_select": [vp, C.c_int64, C.c_int32, vp],
        "gpis_select_begin": [vp, C.c_int64, vp],
        "gpis_step_append": [vp, vp],
        "gpis_step_update": [vp, vp],
        "gpis_select_end": [vp, vp],
        "gpis_exchange_buffers": [vp, C.POINTER(vp), C.POINTER(vp), i64p],
        "gpis_get_active": [vp, vp, i32p, i32p, vp],
        "gpis_get_factor": [vp, dp, C.c_int64, dp, i32p, vp],
        "gpis_fit": [vp, dp, dp, dp, dp, C.c_int32, vp],
        "gpis_predict": [vp, dp, C.c_int64, dp, dp, dp, vp],
        "gpis_get_levelset": [vp, dp, u8p, u8p, u8p, vp],
        "gpis_get_candidate_posterior": [vp, dp, dp, vp],
        "gpis_get_stats": [vp, i64p, i64p, i64p, i64p],
        "gpis_get_timing": [vp, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double), i64p],
        "gpis_set_profiling": [vp, C.c_int32],
        "gpis_get_persist_timing": [vp, vp],
        "gpis_get_persist_balance": [vp, vp, C.c_int32, i32p],
        "gpis_get_phase_timing": [vp, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)],
        "gpis_peer_export": [vp, vp],
        "gpis_peer_import": [vp, vp, C.c_int32],
        "gpis_get_history": [vp, vp, vp, vp, i32p],
        "gpis_posterior_cov": [vp, dp, C.c_int64, C.c_int32, dp, dp, C.c_int64, vp],
        "gpis_sample_shapes": [vp, dp, C.c_int64, C.c_int32, C.c_int32, dp, C.c_double, dp, dp, dp, C.c_int64, vp],
        "gpis_get_sampling_timing": [vp, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)],
        "gpis_batch_select": [C.POINTER(GpisConfig), vp, vp, vp, C.c_int32, dp, C.c_int64, C.c_int32, vp, vp, vp,
                              dp, dp, u8p, C.POINTER(C.c_double), vp],
        "gpis_set_candidates_f32": [vp, dp, dp, dp, dp, vp, vp],
        "gpis_set_grid_candidates_f32": [vp, vp, vp, vp, C.c_int32, C.c_int32, dp, dp, dp, vp],
        "gpis_get_active_f32": [vp, C.c_int32, vp, vp, i32p, i32p, vp],
        "gpis_predict_f32": [vp, dp, C.c_int64, dp, dp, dp, vp],
        "gpis_get_candidate_posterior_f32": [vp, dp, dp, vp],
        "gpis_build_tsdf": [u8p, C.c_int32, C.c_int32, C.POINTER(GpisTsdfParams), dp, dp, dp, dp, vp, i32p, C.c_int32, vp],
    }
    for name, argtypes in sig.items():
        fn = getattr(lib, name)
        fn.restype = C.c_int32
        fn.argtypes = argtypes
    _lib = lib
    return lib


def check(lib, handle, status):
    if status != GPIS_OK:
        msg = lib.gpis_status_string(status).decode()
        detail = lib.gpis_last_error(handle if handle else None).decode()
        if detail:
            msg += ": " + detail
        raise GpisError(status, msg)
