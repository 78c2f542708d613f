"""ctypes loader for libibk.so (the C ABI of include/ibk.h).

The library is mandatory: there is no CPU or PyTorch fallback anywhere in this package.  If the
shared library is missing or fails to load, importing the compute API raises immediately.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("IBK_LIB") or os.path.join(HERE, "libibk.so")  # IBK_LIB: an experiment build (scripts/build_variant.sh)

IBK_MAX_DIM = 3


class ArrayDesc(C.Structure):
    _fields_ = [
        ("ndim", C.c_int),
        ("depth", C.c_int),
        ("dx", C.c_double * 3),
        ("x_lower", C.c_double * 3),
        ("x_upper", C.c_double * 3),
        ("ilower", C.c_int * 3),
        ("iupper", C.c_int * 3),
        ("nugc", C.c_int * 3),
        ("axis", C.c_int),
    ]


class PatchDesc(C.Structure):
    _fields_ = [
        ("ndim", C.c_int),
        ("lower", C.c_int * 3),
        ("upper", C.c_int * 3),
        ("gcw", C.c_int * 3),
        ("x_lower", C.c_double * 3),
        ("x_upper", C.c_double * 3),
        ("dx", C.c_double * 3),
        ("touches_physical_bdry", C.c_int),
    ]


class LevelDesc(C.Structure):
    _fields_ = [
        ("ndim", C.c_int),
        ("n_patches", C.c_int),
        ("domain_lower", C.c_int * 3),
        ("domain_upper", C.c_int * 3),
        ("x_lower", C.c_double * 3),
        ("x_upper", C.c_double * 3),
        ("periodic", C.c_int * 3),
        ("gcw", C.c_int * 3),
        ("patch_lower", C.POINTER(C.c_int)),
        ("patch_upper", C.POINTER(C.c_int)),
    ]


# every symbol include/ibk.h declares: name -> (restype, argtypes)
_vp, _i, _d = C.c_void_p, C.c_int, C.c_double
_pi, _pd = C.POINTER(C.c_int), C.POINTER(C.c_double)
_ppd = C.POINTER(C.POINTER(C.c_double))
_s = C.c_char_p
_ll = C.c_longlong
SYMBOLS = {
    "ibk_kernel_from_string": (_i, [_s]),
    "ibk_is_known_kernel": (_i, [_s]),
    "ibk_get_stencil_size": (_i, [_s]),
    "ibk_get_minimum_ghost_width": (_i, [_s]),
    "ibk_ctx_create": (_i, [_i, C.POINTER(_vp)]),
    "ibk_ctx_destroy": (_i, [_vp]),
    "ibk_last_error": (_s, [_vp]),
    "ibk_ctx_set_stream": (_i, [_vp, _vp]),
    "ibk_ctx_synchronize": (_i, [_vp]),
    "ibk_ctx_launch_count": (_ll, [_vp]),
    "ibk_ctx_enable_timing": (_i, [_vp, _i]),
    "ibk_ctx_last_ms": (_i, [_vp, _i, C.POINTER(C.c_float)]),
    "ibk_raw_interp": (_i, [_vp, _i, C.POINTER(ArrayDesc), _vp, _vp, _vp, _i, _vp, _i, _vp]),
    "ibk_raw_spread": (_i, [_vp, _i, C.POINTER(ArrayDesc), _vp, _vp, _i, _vp, _i, _vp, _vp]),
    "ibk_raw_interp_host": (_i, [_vp, _i, C.POINTER(ArrayDesc), _pd, _pi, _pd, _i, _pd, _i, _pd]),
    "ibk_raw_spread_host": (_i, [_vp, _i, C.POINTER(ArrayDesc), _pi, _pd, _i, _pd, _i, _pd, _pd]),
    "ibk_side_interpolate_host": (_i, [_vp, _s, C.POINTER(PatchDesc), _ppd, _i, _pi, _pi, _pd, _i, _i, _pd, _i, _i]),
    "ibk_side_spread_host": (_i, [_vp, _s, C.POINTER(PatchDesc), _ppd, _i, _pi, _pi, _pd, _i, _i, _pd, _i, _i]),
    "ibk_cell_interpolate_host": (_i, [_vp, _s, C.POINTER(PatchDesc), _pd, _i, _pi, _pi, _pd, _i, _i, _pd, _i, _i]),
    "ibk_cell_spread_host": (_i, [_vp, _s, C.POINTER(PatchDesc), _pd, _i, _pi, _pi, _pd, _i, _i, _pd, _i, _i]),
    "ibk_node_interpolate_host": (_i, [_vp, _s, C.POINTER(PatchDesc), _pd, _i, _pi, _pi, _pd, _i, _i, _pd, _i, _i]),
    "ibk_node_spread_host": (_i, [_vp, _s, C.POINTER(PatchDesc), _pd, _i, _pi, _pi, _pd, _i, _i, _pd, _i, _i]),
    "ibk_edge_interpolate_host": (_i, [_vp, _s, C.POINTER(PatchDesc), _ppd, _i, _pi, _pi, _pd, _i, _i, _pd, _i, _i]),
    "ibk_edge_spread_host": (_i, [_vp, _s, C.POINTER(PatchDesc), _ppd, _i, _pi, _pi, _pd, _i, _i, _pd, _i, _i]),
    "ibk_side_interpolate_indexed_host": (_i, [_vp, _s, C.POINTER(PatchDesc), _ppd, _pi, _pd, _i, _pd, _i, _pd]),
    "ibk_side_spread_indexed_host": (_i, [_vp, _s, C.POINTER(PatchDesc), _ppd, _pi, _pd, _i, _pd, _i, _pd]),
    "ibk_level_create": (_i, [_vp, C.POINTER(LevelDesc)]),
    "ibk_level_destroy": (_i, [_vp]),
    "ibk_grid_upload": (_i, [_vp, _i, _i, _i, _pd]),
    "ibk_grid_download": (_i, [_vp, _i, _i, _i, _pd]),
    "ibk_grid_fill": (_i, [_vp, _i, _d]),
    "ibk_grid_upload_async": (_i, [_vp, _i, _i, _i, _pd]),
    "ibk_grid_download_async": (_i, [_vp, _i, _i, _i, _pd]),
    "ibk_transfers_wait": (_i, [_vp]),
    "ibk_markers_set_positions": (_i, [_vp, _pd, _i]),
    "ibk_markers_upload": (_i, [_vp, _i, _pd]),
    "ibk_markers_download": (_i, [_vp, _i, _pd]),
    "ibk_markers_count": (_i, [_vp]),
    "ibk_rebin": (_i, [_vp, _i]),
    "ibk_markers_owned_count": (_i, [_vp, _pi]),
    "ibk_halo_pack_many": (_i, [_vp, _i, _i, _pi, _pi, _pi, _pi, C.POINTER(C.c_longlong), _vp]),
    "ibk_halo_unpack_many": (_i, [_vp, _i, _i, _pi, _pi, _pi, _pi, C.POINTER(C.c_longlong), _vp, _i]),
    "ibk_spread_force_part": (_i, [_vp, _s, _i]),
    "ibk_interpolate_velocity_part": (_i, [_vp, _s, _i]),
    "ibk_markers_set_ids": (_i, [_vp, C.POINTER(C.c_uint), C.c_uint]),
    "ibk_markers_get_ids": (_i, [_vp, C.POINTER(C.c_uint)]),
    "ibk_migrate_plan": (_i, [_vp, _i, _pi, _pi, _pi, _i, _i, _pi]),
    "ibk_migrate_pack": (_i, [_vp, _vp]),
    "ibk_migrate_unpack": (_i, [_vp, _vp, _i, C.c_uint]),
    "ibk_markers_lincomb": (_i, [_vp, _i, _d, _i, _d, _i]),
    "ibk_markers_zero_rows": (_i, [_vp, _i, _pi, _i]),
    "ibk_markers_scale_rows": (_i, [_vp, _i, _i, _pd]),
    "ibk_force_set_springs": (_i, [_vp, _i, _pi, _pi, _pd, _pd]),
    "ibk_force_set_beams": (_i, [_vp, _i, _pi, _pi, _pi, _pd, _pd]),
    "ibk_force_set_target_points": (_i, [_vp, _i, _pi, _pd, _pd, _pd]),
    "ibk_force_clear": (_i, [_vp]),
    "ibk_compute_lagrangian_force": (_i, [_vp, _i, _i, _i]),
    "ibk_io_last_error": (_s, []),
    "ibk_io_read_vertex_file": (_i, [_s, _i, _pd, _i, _pi]),
    "ibk_io_read_spring_file": (_i, [_s, _i, _i, _pi, _pi, _pd, _pd, _pi, _i, _pi]),
    "ibk_io_read_beam_file": (_i, [_s, _i, _i, _i, _pi, _pi, _pi, _pd, _pd, _i, _pi]),
    "ibk_io_read_target_file": (_i, [_s, _i, _i, _pi, _pd, _pd, _i, _pi]),
    "ibk_io_read_anchor_file": (_i, [_s, _i, _i, _pi, _i, _pi]),
    "ibk_bin_get_cells": (_i, [_vp, _pi, _pi]),
    "ibk_bin_get_order": (_i, [_vp, _pi]),
    "ibk_bin_get_patch_lists": (_i, [_vp, _i, _pi, _pi, _pd, _pi]),
    "ibk_spread_force": (_i, [_vp, _s, _i]),
    "ibk_spread_begin": (_i, [_vp]),
    "ibk_spread_end": (_i, [_vp]),
    "ibk_level_set_wall_bc": (_i, [_vp, _pd, _pd]),
    "ibk_spread_fold_walls": (_i, [_vp]),
    "ibk_set_user_kernel": (_i, [C.CFUNCTYPE(C.c_double, C.c_double), _i]),
    "ibk_construct_sc_interp_op": (_i, [_vp, _i, C.POINTER(C.POINTER(C.c_int)), _pi, _pd, _pi]),
    "ibk_amr_refine_side": (_i, [_vp, _vp, _i, C.POINTER(C.c_int), C.POINTER(C.c_longlong)]),
    "ibk_amr_coarsen_side": (_i, [_vp, _vp, _i, C.POINTER(C.c_int), C.POINTER(C.c_longlong)]),
    "ibk_interpolate_velocity": (_i, [_vp, _s, _i]),
    "ibk_halo_local": (_i, [_vp, _i]),
    "ibk_halo_pack": (_i, [_vp, _i, _i, _i, _pi, _pi, _vp]),
    "ibk_halo_unpack": (_i, [_vp, _i, _i, _i, _pi, _pi, _vp, _i]),
    "ibk_markers_device_ptr": (_i, [_vp, _i, C.POINTER(_vp), C.POINTER(_ll)]),
    "ibk_grid_device_ptr": (_i, [_vp, _i, _i, _i, C.POINTER(_vp), C.POINTER(_ll), _pi]),
    "ibk_count_touched_dofs": (_i, [_vp, _s, C.POINTER(_ll)]),
    "ibk_halo_plan_create": (_i, [_i, _i, _pi, _pi, _pi, _pi, _pi, _pi, _i, C.POINTER(_vp)]),
    "ibk_halo_plan_destroy": (None, [_vp]),
    "ibk_halo_plan_messages": (_i, [_vp, _i]),
    "ibk_halo_plan_message": (_i, [_vp, _i, _i, _pi, _pi, _pi, C.POINTER(_ll)]),
    "ibk_halo_plan_items": (_i, [_vp, _i, _i, _pi, _pi, _pi, _pi, _pi, _pi, _pi]),
    "ibk_comm_unique_id": (_i, [_vp]),
    "ibk_comm_init": (_i, [_vp, _vp, _i, _i]),
    "ibk_comm_init_loopback": (_i, [C.POINTER(_vp), _i]),
    "ibk_comm_set_patches": (_i, [_vp, _i, _pi, _pi, _pi]),
    "ibk_comm_destroy": (_i, [_vp]),
    "ibk_comm_set_reserved_sms": (_i, [_vp, _i]),
    "ibk_halo_fill_post": (_i, [_vp]),
    "ibk_halo_fill_finish": (_i, [_vp]),
    "ibk_halo_accumulate_post": (_i, [_vp]),
    "ibk_halo_accumulate_finish": (_i, [_vp]),
    "ibk_halo_bytes": (_ll, [_vp, _i]),
    "ibk_migrate": (_i, [_vp, C.c_uint, _pi, _pi]),
    "ibk_migrate_loopback": (_i, [C.POINTER(_vp), _i, C.c_uint, _pi]),
}

_LIB = None


def load():
    """Loads libibk.so and binds every declared symbol.  Raises if the library is missing."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m ibamr_b200.build` (nvcc, sm_100a). "
            "ibamr_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _LIB = lib
    return lib
