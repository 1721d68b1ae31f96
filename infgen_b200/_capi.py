"""ctypes binding of libinfgen_b200.so (the C ABI declared in include/infgen_b200.h).

This is the stub a reference maintainer would add next to `infgen/modules/agent_decoder.py`; INTEGRATION.md shows
the few lines that call it.  There is deliberately no fallback: if the shared library is missing or no CUDA device
is present, loading / engine creation raises.
"""
import ctypes as C
import os
from typing import Optional

_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'lib', 'libinfgen_b200.so')
ABI_VERSION = 8
HOST, DEVICE = 0, 1

c_f32p = C.POINTER(C.c_float)
c_i32p = C.POINTER(C.c_int32)
c_u8p = C.POINTER(C.c_uint8)


class Config(C.Structure):
    _fields_ = [
        ('abi_version', C.c_int32), ('device', C.c_int32), ('num_layers', C.c_int32), ('hist_cols', C.c_int32),
        ('window', C.c_int32), ('shift', C.c_int32), ('num_historical_steps', C.c_int32), ('token_size', C.c_int32),
        ('grid_size', C.c_int32), ('num_seed_feature', C.c_int32), ('max_pl2a_neighbors', C.c_int32),
        ('max_a2a_neighbors', C.c_int32), ('pl2a_radius', C.c_float), ('a2a_radius', C.c_float),
        ('use_state_token', C.c_int32), ('disable_insertion', C.c_int32), ('motion_beam_size', C.c_int32),
        ('seed', C.c_uint32), ('use_cuda_graph', C.c_int32), ('trace', C.c_int32),
        ('insert_beam_size', C.c_int32), ('debug_force_enter', C.c_int32), ('pl2seed_radius', C.c_float),
        ('a2sa_radius', C.c_float), ('pl2sa_radius', C.c_float), ('angle_interval', C.c_float),
        ('teacher_forced', C.c_int32),
    ]


class SceneBatch(C.Structure):
    _fields_ = [
        ('n_scenes', C.c_int32), ('row_capacity', C.c_int32), ('n_cols', C.c_int32), ('n_iters', C.c_int32),
        ('n_rows', c_i32p), ('ego_row', c_i32p), ('scene_id', c_i32p),
        ('pos_hist', c_f32p), ('head_hist', c_f32p), ('state_hist', c_i32p), ('token_hist', c_i32p),
        ('grid_hist', c_i32p), ('tsrc_hist', c_u8p), ('interact_hist', c_u8p), ('type', c_i32p), ('shape', c_f32p),
        ('pt_ptr', c_i32p), ('pt_pos', c_f32p), ('pt_ori', c_f32p), ('x_pt', c_f32p),
    ]


class MapBatch(C.Structure):
    _fields_ = [
        ('n_scenes', C.c_int32), ('pt_ptr', c_i32p), ('pt_pos', c_f32p), ('pt_ori', c_f32p), ('type', c_i32p),
        ('pl_type', c_i32p), ('light_type', c_i32p), ('token_idx', c_i32p), ('pl2pl_radius', C.c_float),
    ]


c_i64p = C.POINTER(C.c_int64)


class PrepIn(C.Structure):
    _fields_ = [
        ('n_agents', C.c_int32), ('n_steps', C.c_int32), ('av_index', C.c_int32), ('n_pt', C.c_int32),
        ('valid_mask', c_u8p), ('heading', c_f32p), ('position', c_f32p), ('velocity', c_f32p), ('type', c_u8p),
        ('pt_position', c_f32p),
    ]


class PrepOut(C.Structure):
    _fields_ = [
        ('token_idx', c_i64p), ('state_idx', c_i64p), ('token_contour', c_f32p), ('token_pos', c_f32p),
        ('token_heading', c_f32p), ('raw_agent_valid_mask', c_u8p), ('agent_valid_mask', c_u8p),
        ('grid_token_idx', c_i64p), ('grid_offset_xy', c_f32p), ('heading_token_idx', c_i64p), ('pos_xy', c_f32p),
        ('heading_theta', c_f32p), ('sort_indices', c_i64p), ('inrange_mask', c_u8p), ('bos_mask', c_u8p),
        ('pt_grid_token_idx', c_i64p),
    ]


class MapMatchIn(C.Structure):
    _fields_ = [
        ('n_tokens', C.c_int32), ('n_vocab', C.c_int32), ('n_polygons', C.c_int32),
        ('traj_pos', c_f32p), ('traj_theta', c_f32p), ('pl_rank', c_i32p), ('side', c_u8p), ('sample_pt', c_f32p),
    ]


class MapMatchOut(C.Structure):
    _fields_ = [
        ('token_idx', c_i64p), ('position', c_f32p), ('orientation', c_f32p), ('side_counts', c_i32p),
        ('best_distance', c_f32p),
    ]


class Outputs(C.Structure):
    _fields_ = [
        ('pos', c_f32p), ('head', c_f32p), ('pred_traj', c_f32p), ('pred_head', c_f32p), ('pred_state', c_f32p),
        ('next_token', c_i32p), ('next_state', c_i32p), ('hist_traj', c_f32p), ('hist_head', c_f32p),
        ('n_rows_final', c_i32p), ('pred_type', c_i32p), ('pred_shape', c_f32p), ('rec_meta', c_i32p),
        ('rec_state_prob', c_f32p), ('rec_pos_prob', c_f32p), ('rec_agent_occ', c_f32p), ('rec_pt_occ', c_f32p),
        ('rec_occ_gt', c_f32p),
    ]


# every symbol include/infgen_b200.h declares: name -> (restype, argtypes)
SYMBOLS = {
    'infgen_abi_version': (C.c_int32, []),
    'infgen_last_error': (C.c_char_p, []),
    'infgen_weight_count': (C.c_int32, []),
    'infgen_weight_name': (C.c_char_p, [C.c_int32]),
    'infgen_weight_offset': (C.c_int64, [C.c_char_p]),
    'infgen_weight_numel': (C.c_int64, [C.c_char_p]),
    'infgen_weight_blob_floats': (C.c_int64, []),
    'infgen_create': (C.c_int32, [C.POINTER(Config), c_f32p, C.c_int64, c_f32p, c_f32p, C.POINTER(C.c_void_p)]),
    'infgen_destroy': (C.c_int32, [C.c_void_p]),
    'infgen_set_stream': (C.c_int32, [C.c_void_p, C.c_void_p]),
    'infgen_set_sampler': (C.c_int32, [C.c_void_p, C.c_int32, C.c_uint32]),
    'infgen_synchronize': (C.c_int32, [C.c_void_p]),
    'infgen_load_scenes': (C.c_int32, [C.c_void_p, C.POINTER(SceneBatch), C.c_int32]),
    'infgen_set_forcing': (C.c_int32, [C.c_void_p, c_i32p, c_i32p, C.c_int32]),
    'infgen_prefill': (C.c_int32, [C.c_void_p]),
    'infgen_step': (C.c_int32, [C.c_void_p, C.c_int32]),
    'infgen_rollout': (C.c_int32, [C.c_void_p]),
    'infgen_read': (C.c_int32, [C.c_void_p, C.POINTER(Outputs), C.c_int32]),
    'infgen_iterations_done': (C.c_int32, [C.c_void_p]),
    'infgen_forward': (C.c_int32, [C.c_void_p, c_f32p, c_f32p, c_f32p, C.c_int32]),
    'infgen_map_setup': (C.c_int32, [C.c_void_p, c_f32p, C.c_int32]),
    'infgen_map_encode': (C.c_int32, [C.c_void_p, C.POINTER(MapBatch), C.c_int32, c_f32p, c_f32p]),
    'infgen_prepare_scene': (C.c_int32, [C.c_void_p, C.POINTER(PrepIn), C.POINTER(PrepOut)]),
    'infgen_match_map_tokens': (C.c_int32, [C.c_void_p, C.POINTER(MapMatchIn), C.POINTER(MapMatchOut)]),
    'infgen_kernel_launches': (C.c_int64, [C.c_void_p]),
    'infgen_set_profile': (C.c_int32, [C.c_void_p, C.c_int32]),
    'infgen_profile_class_count': (C.c_int32, []),
    'infgen_profile_class_name': (C.c_char_p, [C.c_int32]),
    'infgen_profile_read': (C.c_int32, [C.c_void_p, C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    'infgen_debug_read': (C.c_int64, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int64]),
    'infgen_op_attention_layer': (C.c_int32, [C.c_void_p, C.c_char_p, c_f32p, C.c_int32, c_f32p, C.c_int32, c_f32p,
                                              c_i32p, c_i32p, c_f32p]),
    'infgen_op_fourier_embedding': (C.c_int32, [C.c_void_p, C.c_char_p, c_f32p, C.c_int32, C.c_int32, c_f32p, c_f32p]),
    'infgen_op_mlp_embedding': (C.c_int32, [C.c_void_p, C.c_char_p, c_f32p, C.c_int32, C.c_int32, c_f32p]),
    'infgen_op_mlp_layer': (C.c_int32, [C.c_void_p, C.c_char_p, c_f32p, C.c_int32, c_f32p]),
}

_lib: Optional[C.CDLL] = None


def lib_path() -> str:
    return _LIB_PATH


def load() -> C.CDLL:
    """Load the shared library (built by `__graft_entry__.build()` / `python -m infgen_b200.build`)."""
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise RuntimeError(f'{_LIB_PATH} is missing: build it with `python -m infgen_b200.build` '
                               '(the decode path is CUDA-only, there is no fallback)')
        lib = C.CDLL(_LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        if lib.infgen_abi_version() != ABI_VERSION:
            raise RuntimeError(f'ABI mismatch: library {lib.infgen_abi_version()} vs binding {ABI_VERSION}')
        _lib = lib
    return _lib


ERR_CAPACITY = -2


class InfgenError(RuntimeError):
    """A call into libinfgen_b200.so returned a negative `infgen_status` (`code`)."""

    def __init__(self, code: int, message: str):
        super().__init__(f'infgen_b200 error {code}: {message}')
        self.code = code


class CapacityError(InfgenError):
    """INFGEN_ERR_CAPACITY: the batch does not fit the row space it was loaded with."""


def check(rc: int) -> None:
    if rc != 0:
        msg = load().infgen_last_error().decode()
        raise (CapacityError if rc == ERR_CAPACITY else InfgenError)(rc, msg)


def f32p(t):
    """Pointer to a contiguous float32 torch tensor / numpy array (None -> NULL)."""
    if t is None:
        return None
    return C.cast(_ptr(t), c_f32p)


def i32p(t):
    if t is None:
        return None
    return C.cast(_ptr(t), c_i32p)


def i64p(t):
    if t is None:
        return None
    return C.cast(_ptr(t), c_i64p)


def u8p(t):
    if t is None:
        return None
    return C.cast(_ptr(t), c_u8p)


def _ptr(t) -> int:
    if hasattr(t, 'data_ptr'):
        assert t.is_contiguous()
        return t.data_ptr()
    assert t.flags['C_CONTIGUOUS']
    return t.ctypes.data
