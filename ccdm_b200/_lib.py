"""ctypes binding of libccdm_b200.so (the C ABI declared in include/ccdm_b200.h).

There is deliberately no fallback: if the shared library is missing or a call fails, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libccdm_b200.so")

MAX_SRC, MAX_Z, STEP_NCOEF = 4, 4, 12

EPI_BIAS, EPI_ROWSCALE, EPI_RMSNORM, EPI_SS = 0x1, 0x2, 0x4, 0x8
EPI_SILU, EPI_RESID, EPI_QSOFTMAX, EPI_SUMSQ_OUT, EPI_OUT_F32, EPI_KEXP = 0x10, 0x20, 0x40, 0x80, 0x100, 0x200
EPI_RELU, EPI_TANH, EPI_HEAD, EPI_RESACC = 0x400, 0x800, 0x1000, 0x2000
ACT_NONE, ACT_RELU, ACT_GELU, ACT_SILU = 0, 1, 2, 3
OBJ = {"pred_noise": 0, "pred_x0": 1, "pred_v": 2}

vp, i32, i64, f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float


class View(C.Structure):
    _fields_ = [("ptr", vp), ("C", i32), ("W", i32), ("H", i32), ("B", i32), ("sW", i64), ("sH", i64), ("sB", i64)]


class TapGemmArgs(C.Structure):
    _fields_ = [
        ("n_src", i32), ("src", View * MAX_SRC),
        ("gW", i32), ("gH", i32), ("gB", i32), ("tw", i32), ("th", i32), ("tb", i32),
        ("nz", i32), ("ngroups", i32), ("R", i32), ("sched", vp), ("wpacked", vp),
        ("n_rows", i32), ("w_batch_rows", i32), ("N", i32), ("n_tile", i32), ("flags", C.c_uint32),
        ("bias", vp), ("rowss", vp), ("gain", vp), ("gain_mul", f32), ("scale_shift", vp), ("ss_ld", i32), ("ss_off", i32),
        ("resid", vp), ("rsW", i64), ("rsH", i64), ("rsB", i64),
        ("out", vp), ("osW", i64), ("osH", i64), ("osB", i64), ("ooff", i64 * MAX_Z),
        ("out_rowss", vp), ("q_scale", f32), ("q_cols", i32),
        ("halo", i32), ("n_res", i32), ("res_bias", vp),
        ("head_n", i32), ("head_w", vp), ("head_b", vp), ("head_out", vp), ("hsC", i64), ("hsB", i64),
        ("ksteps", vp),
    ]


class StepArgs(C.Structure):
    _fields_ = [
        ("out_cond", vp), ("out_null", vp), ("x", vp), ("noise", vp), ("pred_noise", vp), ("pred_x0", vp),
        ("B", i32), ("chw", i32), ("cond_scale", f32), ("rescaled_phi", f32), ("keep_parallel_frac", f32),
        ("remove_parallel", i32), ("objective", i32), ("clip_x0", i32), ("cfg_plus_plus", i32), ("sampler", i32),
        ("coef", vp), ("step_counter", vp), ("advance", i32), ("t_rows", vp),
    ]


class QSampleArgs(C.Structure):
    _fields_ = [
        ("img01", vp), ("noise", vp), ("noise2", vp), ("cov", vp), ("keep", vp), ("t", vp), ("sqrt_acp", vp),
        ("sqrt_1m_acp", vp), ("x0", vp), ("noise_out", vp), ("x_t", vp), ("B", i32), ("chw", i32), ("normalize", i32),
    ]


class LossArgs(C.Structure):
    _fields_ = [
        ("model_out", vp), ("x0", vp), ("noise", vp), ("cov", vp), ("keep", vp), ("t", vp), ("sqrt_acp", vp),
        ("sqrt_1m_acp", vp), ("loss_weight", vp), ("row_weight", vp), ("per_sample", vp), ("loss", vp),
        ("grad_out", vp), ("B", i32), ("chw", i32), ("objective", i32),
    ]


class PackJob(C.Structure):
    _fields_ = [("w", vp), ("out", vp), ("psched", vp), ("mode", i32), ("cout", i32), ("cin_total", i32), ("ntaps", i32),
                ("nkb", i32), ("n_rows", i32), ("n_off", i32), ("n_count", i32), ("total", i64)]


class WgradArgs(C.Structure):
    _fields_ = [
        ("n_src", i32), ("src", View * MAX_SRC), ("dz", vp), ("dsW", i64), ("dsH", i64), ("dsB", i64),
        ("doff", i64 * MAX_Z), ("gW", i32), ("gH", i32), ("gB", i32), ("tw", i32), ("th", i32), ("tb", i32),
        ("nz", i32), ("ngroups", i32), ("R", i32), ("sched", vp), ("N", i32), ("n_rows", i32), ("wgrad_packed", vp),
        ("ksplit", i32), ("slots", i32), ("slot_stride", i64),
    ]


# name -> (restype, argtypes); every symbol include/ccdm_b200.h declares
SIGNATURES = {
    "ccdm_version": (C.c_int, []),
    "ccdm_last_error": (C.c_char_p, []),
    "ccdm_launch_count": (i64, []),
    "ccdm_struct_size": (C.c_int, [C.c_int]),
    "ccdm_tapgemm": (C.c_int, [C.POINTER(TapGemmArgs), vp]),
    "ccdm_pack_weights": (C.c_int, [vp, i32, i32, i32, vp, i32, i32, i32, vp, f32, vp, vp]),
    "ccdm_pack_multi": (C.c_int, [vp, i32, vp]),
    "ccdm_pack_weights_at": (C.c_int, [vp, i32, i32, i32, vp, i32, i32, i32, vp, f32, vp, i32, i32, vp]),
    "ccdm_rmsnorm_act": (C.c_int, [vp, vp, i64, i32, i32, vp, f32, vp, i32, i32, vp, vp, C.c_uint32, vp]),
    "ccdm_stem_im2row": (C.c_int, [vp, vp, i32, i32, i32, i32, vp]),
    "ccdm_stem_pack": (C.c_int, [vp, vp, i32, i32, i32, vp]),
    "ccdm_head_conv1": (C.c_int, [vp, vp, vp, vp, i32, i32, i32, i32, i32, vp]),
    "ccdm_linattn_context": (C.c_int, [vp, vp, vp, i32, i32, i32, vp, vp, i32, i32, vp]),
    "ccdm_kexp_bound": (C.c_int, [vp, i32, i32, i32, i32, vp, vp]),
    "ccdm_linattn_fold": (C.c_int, [vp, vp, vp, i32, i32, i32, i32, vp]),
    "ccdm_attention_small": (C.c_int, [vp, vp, i32, i32, i32, i32, f32, vp]),
    "ccdm_linattn_fused_units": (C.c_int, [i32]),
    "ccdm_linattn_kv_partials": (C.c_int, [vp, i32, i32, i32, vp, vp, vp, vp, vp, vp]),
    "ccdm_linattn_fold_partials": (C.c_int, [vp, vp, i32, i32, vp, i32, i32, vp, vp]),
    "ccdm_linattn_q_out": (C.c_int, [vp, i32, i32, i32, vp, vp, vp, i32, vp, vp, f32, f32, vp, vp]),
    "ccdm_linear_small": (C.c_int, [vp, i32, i32, vp, vp, i32, vp, vp, vp, vp, i32, i32, vp, i64, vp]),
    "ccdm_groupnorm_rows": (C.c_int, [vp, i32, i32, i32, vp, vp, C.c_float, i32, vp]),
    "ccdm_time_features": (C.c_int, [vp, i32, i32, vp, vp]),
    "ccdm_select_null": (C.c_int, [vp, vp, i32, vp, i32, i32, vp]),
    "ccdm_silu_concat_bf16": (C.c_int, [vp, i32, vp, i32, i32, vp, vp]),
    "ccdm_sampler_step": (C.c_int, [C.POINTER(StepArgs), vp]),
    "ccdm_broadcast_step_i64": (C.c_int, [vp, vp, vp, i32, vp]),
    "ccdm_cfg_combine": (C.c_int, [vp, vp, vp, i32, i32, f32, f32, i32, f32, vp]),
    "ccdm_q_sample": (C.c_int, [C.POINTER(QSampleArgs), vp]),
    "ccdm_vicinal_loss": (C.c_int, [C.POINTER(LossArgs), vp]),
    "ccdm_vicinal_weights": (C.c_int, [vp, i32, i32, i32, i32, vp, f32, vp, vp, vp]),
    "ccdm_pack_weights_t": (C.c_int, [vp, i32, i32, i32, vp, i32, i32, i32, i32, i32, vp, vp]),
    "ccdm_conv_wgrad": (C.c_int, [C.POINTER(WgradArgs), vp]),
    "ccdm_unpack_wgrad": (C.c_int, [vp, vp, i32, i32, i32, vp, i32, i32, i32, vp, f32, i32, vp]),
    "ccdm_unpack_wgrad_slots": (C.c_int, [vp, i32, i64, vp, i32, i32, i32, vp, i32, i32, i32, vp, f32, i32, vp]),
    "ccdm_block_bwd": (C.c_int, [vp, vp, vp, i64, i32, i32, vp, f32, vp, i32, i32, vp, C.c_uint32, vp]),
    "ccdm_block_bwd_finish": (C.c_int, [vp, i32, i32, vp, f32, vp, i32, i32, vp, vp, vp, vp]),
    "ccdm_colsum_bf16": (C.c_int, [vp, i64, i32, vp, vp]),
    "ccdm_linattn_prep": (C.c_int, [vp, i32, i32, vp, f32, vp]),
    "ccdm_linattn_pack_blockdiag": (C.c_int, [vp, vp, i32, vp, i32, vp]),
    "ccdm_linattn_dcontext": (C.c_int, [vp, vp, vp, i32, i32, vp]),
    "ccdm_linattn_bwd_rowdot": (C.c_int, [vp, vp, vp, vp, i32, vp]),
    "ccdm_linattn_bwd_finish": (C.c_int, [vp, vp, i32, i32, vp, f32, vp]),
    "ccdm_attention_small_bwd": (C.c_int, [vp, vp, vp, i32, i32, i32, i32, f32, vp]),
    "ccdm_head_conv1_bwd": (C.c_int, [vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, vp]),
    "ccdm_stem_unpack_wgrad": (C.c_int, [vp, vp, i32, i32, i32, vp]),
    "ccdm_gather_augment_u8": (C.c_int, [vp, i64, vp, vp, vp, i32, i32, i32, i32, vp]),
    "ccdm_condbn_coef": (C.c_int, [vp, vp, vp, vp, vp, vp, f32, i32, i32, vp, vp]),
    "ccdm_affine_act": (C.c_int, [vp, vp, i64, i32, i32, vp, i32, i32, i32, vp]),
    "ccdm_channel_stats": (C.c_int, [vp, i32, i32, i32, vp, i32, i32, i32, vp]),
    "ccdm_groupnorm_coef": (C.c_int, [vp, i32, i32, i32, i64, f32, vp, vp, vp, i32, i32, i32, vp, vp]),
    "ccdm_norm_bwd_stats": (C.c_int, [vp, vp, i32, i32, i32, vp, i32, i32, i32, vp, i32, i32, i32, vp]),
    "ccdm_groupnorm_bwd_coef": (C.c_int, [vp, vp, i32, i32, i32, i64, f32, vp, vp, vp, i32, i32, i32, vp, vp, vp, vp, vp]),
    "ccdm_norm_bwd_apply": (C.c_int, [vp, vp, vp, i64, i32, i32, vp, i32, i32, vp, i32, i32, i32, vp]),
    "ccdm_attention_tokens_bwd": (C.c_int, [vp, vp, vp, i32, i32, i32, i32, f32, i32, vp]),
    "ccdm_time_features_adm": (C.c_int, [vp, i32, i32, f32, vp, vp]),
    "ccdm_attention_tokens": (C.c_int, [vp, vp, i32, i32, i32, i32, f32, i32, vp]),
    "ccdm_fused_adam": (C.c_int, [vp, vp, vp, i32, vp, vp, vp, i64, vp, vp, f32, f32, f32, f32, f32, f32, vp]),
    "ccdm_multi_lerp": (C.c_int, [vp, vp, vp, i32, vp, vp]),
}

_libs = {}
# Precision tiers = builds of the same sources (csrc/ptx.cuh): "bf16" stores activations / packed weights as bfloat16,
# "fp16" as IEEE binary16 (TF32's 10-bit mantissa); both accumulate in fp32 on tcgen05.mma kind::f16.
LIB_PATHS = {"bf16": LIB_PATH, "fp16": os.path.join(_HERE, "libccdm_b200_f16.so")}


def lib(precision: str = "bf16"):
    """The loaded library of a precision tier; raises if it has not been built (python -c 'import __graft_entry__ as g; g.build()')."""
    handle = _libs.get(precision)
    if handle is None:
        path = LIB_PATHS[precision]
        if not os.path.exists(path):
            raise RuntimeError(
                f"{path} is missing: build it with `make -C ccdm_b200/csrc` (there is no CPU or PyTorch fallback)")
        handle = C.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)          # AttributeError if the header and the library disagree
            fn.restype, fn.argtypes = res, args
        for which, struct in enumerate((TapGemmArgs, View, StepArgs, QSampleArgs, LossArgs, WgradArgs)):
            if handle.ccdm_struct_size(which) != C.sizeof(struct):
                raise RuntimeError(f"ABI mismatch: {struct.__name__} is {C.sizeof(struct)} bytes here, "
                                   f"{handle.ccdm_struct_size(which)} in {path}")
        _libs[precision] = handle
    return handle


def check(rc: int, what: str = "", precision: str = "bf16"):
    if rc != 0:
        msg = lib(precision).ccdm_last_error().decode(errors="replace")
        raise RuntimeError(f"ccdm_b200 {what} failed ({rc}): {msg}")


def ptr(t):
    """Device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()
