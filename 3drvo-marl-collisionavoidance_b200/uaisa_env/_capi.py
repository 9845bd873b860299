"""ctypes binding of libuaisa_b200.so (include/uaisa_b200.h).  No CPU fallback: if the library is
missing or the device is not a B200 the calls raise."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("UAISA_B200_LIB") or os.path.join(os.path.dirname(_HERE), "lib", "libuaisa_b200.so")

SYMBOLS = (
    "uaisa_last_error", "uaisa_version", "uaisa_default_config", "uaisa_create", "uaisa_destroy",
    "uaisa_set_scenarios", "uaisa_step", "uaisa_reset", "uaisa_observe", "uaisa_autoreset", "uaisa_rollout",
    "uaisa_get_state", "uaisa_set_state", "uaisa_cal_des", "uaisa_indicators", "uaisa_get_counters",
    "uaisa_launch_count", "uaisa_step_host", "uaisa_get_region_clocks", "uaisa_action_transform", "uaisa_gae", "uaisa_step_host_sparse", "uaisa_policy_features", "uaisa_policy_heads",
    "uaisa_opt_partials_len", "uaisa_opt_prepare", "uaisa_clip_adam", "uaisa_rvo_select", "uaisa_reader_fwd", "uaisa_reader_bwd", "uaisa_policy_pack_floats", "uaisa_policy_pack", "uaisa_policy_forward", "uaisa_policy_sample",
    "uaisa_mlp_gemm", "uaisa_mlp_out_bwd", "uaisa_colsum", "uaisa_step_autoreset",
    "uaisa_ppo_wsum", "uaisa_ppo_pi_loss", "uaisa_ppo_v_loss",
)


class UaisaConfig(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "n_envs", "n_drones", "max_waypoints", "max_buildings", "n_scenarios", "nm", "env_train",
        "threads_per_block", "count_pairs", "neighbor_slices")] + [(n, C.c_double) for n in (
        "goal_threshold", "max_acc", "max_angle_change", "dt", "vel_max", "neighbor_range", "building_range",
        "building_zgate", "ctime_threshold", "delta_t", "exp_radius")]


class UaisaStateView(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("pos_yaw", "vel_pitch", "real_len", "max_dev", "extra_len", "wp_flags")]


class UaisaRowRec(C.Structure):
    _fields_ = [("v", C.c_float * 9), ("agent", C.c_int32), ("slot", C.c_int32), ("nbr", C.c_int32)]


class UaisaHostStep(C.Structure):
    _fields_ = [("reward", C.c_void_p), ("self_obs", C.c_void_p), ("done", C.c_void_p), ("info", C.c_void_p), ("finish", C.c_void_p),
                ("count", C.c_void_p), ("rows", C.c_void_p), ("n_rows", C.c_int32), ("reserved", C.c_int32)]


class UaisaError(RuntimeError):
    pass


_lib = None


def lib():
    """Load the shared library (build it first with __graft_entry__.build() / make -C csrc)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise UaisaError("libuaisa_b200.so not built (%s); run `make -C %s`" % (LIB_PATH, os.path.join(os.path.dirname(_HERE), "csrc")))
        L = C.CDLL(LIB_PATH)
        vp, i32 = C.c_void_p, C.c_int32
        L.uaisa_last_error.restype = C.c_char_p
        L.uaisa_version.restype = C.c_int
        L.uaisa_default_config.argtypes = [C.POINTER(UaisaConfig)]
        L.uaisa_create.argtypes = [C.POINTER(UaisaConfig), C.POINTER(vp)]
        L.uaisa_destroy.argtypes = [vp]
        L.uaisa_set_scenarios.argtypes = [vp] * 9
        L.uaisa_step.argtypes = [vp] * 10
        L.uaisa_reset.argtypes = [vp] * 6
        L.uaisa_observe.argtypes = [vp] * 6
        L.uaisa_autoreset.argtypes = [vp, vp, vp, i32, i32, vp, vp, vp, vp, vp, vp]
        L.uaisa_rollout.argtypes = [vp, i32] + [vp] * 8 + [i32, i32, vp]
        L.uaisa_get_state.argtypes = [vp, C.POINTER(UaisaStateView), vp]
        L.uaisa_set_state.argtypes = [vp, C.POINTER(UaisaStateView), vp]
        L.uaisa_cal_des.argtypes = [vp, vp, vp]
        L.uaisa_indicators.argtypes = [vp, vp, vp, vp]
        L.uaisa_get_counters.argtypes = [vp, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64), vp]
        L.uaisa_launch_count.argtypes = [vp]
        L.uaisa_launch_count.restype = C.c_int64
        L.uaisa_get_region_clocks.argtypes = [vp, vp, vp]
        L.uaisa_action_transform.argtypes = [vp, vp, C.c_float, vp, vp, vp]
        L.uaisa_step_host_sparse.argtypes = [vp, vp, i32, i32, C.POINTER(UaisaHostStep)]
        L.uaisa_policy_features.argtypes = [vp, i32, i32, vp, vp, vp, vp, vp, vp]
        L.uaisa_policy_heads.argtypes = [vp, i32, vp, vp, vp, vp]
        L.uaisa_gae.argtypes = [vp, vp, vp, i32, i32, C.c_float, C.c_float, vp, vp, vp]
        L.uaisa_step_host.argtypes = [vp] * 9 + [i32, i32]
        L.uaisa_rvo_select.argtypes = [vp, C.c_double, C.c_double, vp, vp, vp]
        L.uaisa_reader_fwd.argtypes = [vp, vp, i32, i32, vp, vp, vp, vp, vp]
        L.uaisa_reader_bwd.argtypes = [vp, vp, i32, i32, vp, vp, vp, vp, vp, vp]
        L.uaisa_policy_pack_floats.restype = C.c_int
        L.uaisa_policy_pack.argtypes = [vp, vp, vp]
        L.uaisa_policy_forward.argtypes = [vp, vp, i32, i32, vp, vp, vp, vp, vp, vp, vp]
        L.uaisa_policy_sample.argtypes = [vp, vp, vp, C.c_float, i32, vp, vp, vp]
        L.uaisa_mlp_gemm.argtypes = [vp, i32, i32, vp, i32, i32, vp, i32, i32, i32, i32, i32, i32, vp, vp, i32, vp, vp, vp, i32, vp]
        L.uaisa_step_autoreset.argtypes = [vp] * 9 + [i32, i32, vp, vp, vp]
        f32 = C.c_float
        L.uaisa_ppo_wsum.argtypes = [vp, vp, i32, vp, vp]
        L.uaisa_ppo_pi_loss.argtypes = [vp, vp, f32, vp, vp, vp, vp, vp, i32, f32, f32, vp, vp, vp, vp, vp]
        L.uaisa_ppo_v_loss.argtypes = [vp, vp, vp, vp, i32, f32, vp, vp, vp, vp]
        L.uaisa_mlp_out_bwd.argtypes = [vp, vp, vp, i32, i32, vp, vp, vp, vp, vp]
        L.uaisa_colsum.argtypes = [vp, i32, i32, vp, vp]
        L.uaisa_opt_partials_len.restype = C.c_int
        L.uaisa_opt_prepare.argtypes = [vp, C.c_int64, i32, vp, vp, C.c_float, vp, vp, vp]
        L.uaisa_clip_adam.argtypes = [vp, vp, vp, vp, C.c_int64, C.c_int64, C.c_int64, vp] + [C.c_float] * 5 + [vp, vp, vp]
        for name in SYMBOLS:
            fn = getattr(L, name)
            if fn.restype is C.c_int and name not in ("uaisa_version",):
                fn.restype = C.c_int
        _lib = L
    return _lib


def check(rc: int):
    if rc != 0:
        raise UaisaError("libuaisa_b200 error %d: %s" % (rc, lib().uaisa_last_error().decode("utf-8", "replace")))


def default_config() -> UaisaConfig:
    c = UaisaConfig()
    check(lib().uaisa_default_config(C.byref(c)))
    return c
