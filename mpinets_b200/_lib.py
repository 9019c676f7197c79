"""ctypes binding of ``libmpinets_b200.so`` (the C ABI in ``include/mpinets_b200.h``).

There is no CPU fallback: if the shared library is missing or a call fails, this raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmpinets_b200.so")

METRICS_COLS = 8
STAGES = ("fps1", "sa1", "fps2", "sa2", "sa3", "fc", "heads", "update", "sample_robot", "sweep", "build_cloud", "other")
PREC_FP32, PREC_BF16, PREC_BF16X3 = 0, 1, 2
PRECISIONS = {"fp32": PREC_FP32, "bf16": PREC_BF16, "bf16x3": PREC_BF16X3}
EVAL_COLS = 16
# columns of mpn_evaluate's table; names follow Evaluator.evaluate_trajectory's add_metric keys (metrics.py:470-523)
EVAL_COLUMNS = ("collision", "joint_limit_violation", "self_collision", "physical_violations", "position_error",
                "orientation_error", "eff_position_path_length", "eff_orientation_path_length", "correct_final_region",
                "success", "num_steps", "first_collision_step", "config_path_length", "max_collision_depth")

EXPORTS = (
    "mpn_last_error", "mpn_version", "mpn_ctx_create", "mpn_ctx_destroy", "mpn_reserve", "mpn_set_robot_tables",
    "mpn_load_weight", "mpn_weights_finalize", "mpn_fps", "mpn_ball_query", "mpn_gather_points", "mpn_group_points",
    "mpn_sa_forward", "mpn_fk", "mpn_sample_robot", "mpn_sample_end_effector", "mpn_compute_spheres", "mpn_normalize_joints",
    "mpn_unnormalize_joints", "mpn_sdf_points", "mpn_build_cloud", "mpn_build_cloud_from_points", "mpn_build_cloud_ids", "mpn_augment_joints", "mpn_clean_point_cloud", "mpn_render_depth_cloud", "mpn_sweep_flags", "mpn_evaluate", "mpn_sparc", "mpn_collision_loss", "mpn_point_match_loss", "mpn_bc_collision_losses", "mpn_encoder_forward",
    "mpn_policy_forward", "mpn_rollout", "mpn_param_count", "mpn_param_info", "mpn_get_params", "mpn_set_params", "mpn_weights_sync",
    "mpn_train_step_grads", "mpn_train_tc_gemm", "mpn_train_tc_wgrad", "mpn_train_pooled_rows", "mpn_adam_step", "mpn_launch_count", "mpn_profile", "mpn_profile_read", "mpn_tc_selftest", "mpn_tc_error", "mpn_tc_gemm_selftest", "mpn_sa_tile_counts",
)


class MpnConfig(C.Structure):
    _fields_ = [("n_robot", C.c_int32), ("n_obstacle", C.c_int32), ("n_target", C.c_int32), ("max_cuboids", C.c_int32),
                ("max_cylinders", C.c_int32), ("quirk_frames", C.c_int32), ("seed", C.c_uint64)]


class MpnScene(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("cuboid_centers", "cuboid_dims", "cuboid_quats", "cylinder_centers",
                                          "cylinder_radii", "cylinder_heights", "cylinder_quats")]


class MpnError(RuntimeError):
    pass


_lib = None


def load():
    """Loads the library (once). Raises if it has not been built (run ``python -m mpinets_b200.build``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MpnError(f"{LIB_PATH} not found: build it with `python mpinets_b200/build.py` "
                       "(there is no CPU fallback for the CUDA path)")
    lib = C.CDLL(LIB_PATH)
    lib.mpn_last_error.restype = C.c_char_p
    lib.mpn_version.restype = C.c_char_p
    lib.mpn_launch_count.restype = C.c_int64
    lib.mpn_param_count.restype = C.c_int64
    P, I, F, U32 = C.c_void_p, C.c_int, C.c_float, C.c_uint32
    SC = C.POINTER(MpnScene)
    sigs = {
        "mpn_ctx_create": [I, C.POINTER(MpnConfig), C.POINTER(P)],
        "mpn_ctx_destroy": [P],
        "mpn_reserve": [P, I],
        "mpn_set_robot_tables": [P, P, I, P, P, I, P, I, P, P, P, F],
        "mpn_load_weight": [P, C.c_char_p, P, C.POINTER(C.c_int64), I],
        "mpn_weights_finalize": [P],
        "mpn_fps": [P, P, P, I, I, I, I, P, P],
        "mpn_ball_query": [P, P, F, I, P, I, I, I, P, I, P],
        "mpn_gather_points": [P, P, P, I, I, I, P, I, P],
        "mpn_group_points": [P, P, P, I, I, I, P, I, I, P],
        "mpn_sa_forward": [P, P, I, I, P, I, P, I, I, I, P, P, P, P],
        "mpn_fk": [P, P, P, I, P, P],
        "mpn_sample_robot": [P, P, P, I, I, U32, P, I],
        "mpn_compute_spheres": [P, P, P, I, P],
        "mpn_sample_end_effector": [P, P, P, I, I, U32, P],
        "mpn_normalize_joints": [P, P, P, I, P],
        "mpn_unnormalize_joints": [P, P, P, I, P],
        "mpn_sdf_points": [P, P, SC, I, P, I, I, P],
        "mpn_build_cloud": [P, P, SC, I, P, P, U32, P],
        "mpn_build_cloud_from_points": [P, P, I, P, P, P, P, I, C.c_uint32, P],
        "mpn_build_cloud_ids": [P, P, SC, I, P, P, P, P],
        "mpn_augment_joints": [P, P, P, I, F, P, U32, P, P],
        "mpn_clean_point_cloud": [P, P, P, P, I, I, U32, P, P, P, P],
        "mpn_sweep_flags": [P, P, SC, I, P, I, I, I, P, P],
        "mpn_render_depth_cloud": [P, P, SC, I, P, I, I, I, F, F, F, F, P, P],
        "mpn_evaluate": [P, P, SC, I, P, I, P, P, SC, I, I, SC, I, I, P],
        "mpn_sparc": [P, P, I, I, P, P, C.c_float, I, C.c_float, C.c_float, P],
        "mpn_collision_loss": [P, P, SC, I, I, P, C.c_float, P, P],
        "mpn_point_match_loss": [P, P, C.c_int64, P, P, P, P],
        "mpn_bc_collision_losses": [P, P, SC, I, P, P, I, C.c_float, C.c_float, C.c_float, P, P],
        "mpn_encoder_forward": [P, P, I, P, I, I, P],
        "mpn_policy_forward": [P, P, I, P, P, I, I, P],
        "mpn_rollout": [P, P, I, SC, I, I, P, P, P, I, I, I, P, P],
        "mpn_launch_count": [P],
        "mpn_param_count": [P],
        "mpn_param_info": [P, I, C.c_char_p, I, C.POINTER(C.c_int64), C.POINTER(C.c_int64)],
        "mpn_get_params": [P, P, P],
        "mpn_set_params": [P, P, P],
        "mpn_weights_sync": [P],
        "mpn_train_step_grads": [P, P, SC, I, I, P, P, P, I, F, F, F, P, P, P, I],
        "mpn_adam_step": [P, P, P, F, F, F, F, F, I, P],
        "mpn_train_pooled_rows": [P, P, I, I, P],
        "mpn_train_tc_gemm": [P, P, I, P, P, P, P, C.c_int64, I, P],
        "mpn_train_tc_wgrad": [P, P, P, P, C.c_int64, P, C.c_int64, C.POINTER(C.c_int), I],
        "mpn_profile": [P, I],
        "mpn_tc_error": [P, C.POINTER(C.c_int)],
        "mpn_sa_tile_counts": [P, C.POINTER(C.c_uint64), I],
        "mpn_tc_selftest": [P, P, P, P, P, I, I, I, P],
        "mpn_tc_gemm_selftest": [P, P, P, P, P, I, I, I, P, I],
        "mpn_profile_read": [P, C.POINTER(C.c_float), C.POINTER(C.c_int64)],
    }
    for name, argtypes in sigs.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        if name not in ("mpn_launch_count", "mpn_param_count"):
            fn.restype = C.c_int
    _lib = lib
    return lib


def check(status: int):
    if status != 0:
        raise MpnError(f"mpinets_b200 error {status}: {load().mpn_last_error().decode()}")
