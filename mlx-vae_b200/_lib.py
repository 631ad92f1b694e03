"""ctypes binding of libarcvae_sm100.so (the C ABI declared in include/arcvae_b200.h).

There is NO fallback: if the shared library is missing or a call fails, this raises.  torch is used only for
device memory and streams; every computation on the hot path happens inside the library's sm_100a kernels.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

MAX_LAYERS = 8
PREC_FP32, PREC_BF16 = 0, 1
LOSS_KEYS = ("total_loss", "recon_loss", "kl_loss", "weighted_kl", "collapse_penalty", "prop_loss",
             "weighted_prop_loss", "mutual_info", "mi_penalty")

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libarcvae_sm100.so")

fp = C.POINTER(C.c_float)


class Dims(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("V", "E", "H", "L", "C", "NL", "pad_token", "end_token")]


class EncoderParams(C.Structure):
    _fields_ = [("embedding", C.c_void_p), ("Wx", C.c_void_p * MAX_LAYERS), ("Wh", C.c_void_p * MAX_LAYERS),
                ("bias", C.c_void_p * MAX_LAYERS), ("condition_fc_w", C.c_void_p), ("condition_fc_b", C.c_void_p),
                ("fc_mu_w", C.c_void_p), ("fc_mu_b", C.c_void_p), ("fc_logvar_hidden_w", C.c_void_p),
                ("fc_logvar_hidden_b", C.c_void_p), ("fc_logvar_w", C.c_void_p), ("fc_logvar_b", C.c_void_p)]


class DecoderParams(C.Structure):
    _fields_ = [("z_to_hidden_w", C.c_void_p), ("z_to_hidden_b", C.c_void_p), ("condition_to_hidden_w", C.c_void_p),
                ("condition_to_hidden_b", C.c_void_p), ("embedding", C.c_void_p), ("Wx", C.c_void_p * MAX_LAYERS),
                ("Wh", C.c_void_p * MAX_LAYERS), ("bias", C.c_void_p * MAX_LAYERS), ("fc_out_w", C.c_void_p),
                ("fc_out_b", C.c_void_p)]


class LossHyper(C.Structure):
    _fields_ = [("beta", C.c_float), ("lambda_prop", C.c_float), ("lambda_collapse", C.c_float),
                ("free_bits", C.c_float), ("lambda_mi", C.c_float), ("target_mi", C.c_float),
                ("collapse_target_mi", C.c_float), ("pad_mask", C.c_int32)]


class ArcvaeError(RuntimeError):
    pass


_lib = None

# name -> (restype, argtypes); every symbol include/arcvae_b200.h declares
SIGNATURES = {
    "arcvae_last_error": (C.c_char_p, []),
    "arcvae_abi_version": (C.c_int, []),
    "arcvae_launch_count": (C.c_uint64, []),
    "arcvae_flop_count": (C.c_double, [C.c_int]),
    "arcvae_zero": (C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p]),
    "arcvae_timing_enable": (C.c_int, [C.c_int]),
    "arcvae_timing_read": (C.c_int, [C.POINTER(C.c_double), C.POINTER(C.c_int), C.c_int]),
    "arcvae_encoder_tape_bytes": (C.c_size_t, [C.POINTER(Dims), C.c_int, C.c_int]),
    "arcvae_encoder_scratch_bytes": (C.c_size_t, [C.POINTER(Dims), C.c_int, C.c_int]),
    "arcvae_encoder_forward": (C.c_int, [C.POINTER(Dims), C.POINTER(EncoderParams), C.c_void_p, C.c_void_p, C.c_int,
                                         C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]),
    "arcvae_encoder_backward": (C.c_int, [C.POINTER(Dims), C.POINTER(EncoderParams), C.c_void_p, C.c_int, C.c_int,
                                          C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(EncoderParams),
                                          C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]),
    "arcvae_encoder_check": (C.c_int, [C.POINTER(Dims), C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]),
    "arcvae_debug_set_rc_stamps": (C.c_int, [C.c_void_p]),
    "arcvae_recurrence_is_persistent": (C.c_int, [C.c_int]),
    "arcvae_device_error_read": (C.c_int, [C.POINTER(C.c_int), C.c_void_p]),
    "arcvae_device_error_clear": (C.c_int, [C.c_void_p]),
    "arcvae_debug_raise_device_error": (C.c_int, [C.c_void_p]),
    "arcvae_reparameterize": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_uint64, C.c_uint64,
                                        C.c_void_p, C.c_void_p]),
    "arcvae_decoder_tape_bytes": (C.c_size_t, [C.POINTER(Dims), C.c_int, C.c_int]),
    "arcvae_decoder_scratch_bytes": (C.c_size_t, [C.POINTER(Dims), C.c_int, C.c_int]),
    "arcvae_decoder_forward": (C.c_int, [C.POINTER(Dims), C.POINTER(DecoderParams), C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int,
                                         C.c_void_p]),
    "arcvae_decoder_ce_supported": (C.c_int, [C.POINTER(Dims), C.c_int, C.c_int]),
    "arcvae_decoder_forward_ce": (C.c_int, [C.POINTER(Dims), C.POINTER(DecoderParams), C.c_void_p, C.c_void_p, C.c_void_p,
                                            C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                                            C.c_int, C.c_void_p]),
    "arcvae_decoder_backward": (C.c_int, [C.POINTER(Dims), C.POINTER(DecoderParams), C.c_void_p, C.c_int, C.c_int,
                                          C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(DecoderParams), C.c_void_p,
                                          C.c_size_t, C.c_int, C.c_void_p]),
    "arcvae_decoder_cs_tape_bytes": (C.c_size_t, [C.POINTER(Dims), C.c_int, C.c_int]),
    "arcvae_decoder_cs_scratch_bytes": (C.c_size_t, [C.POINTER(Dims), C.c_int, C.c_int]),
    "arcvae_decoder_cs_forward": (C.c_int, [C.POINTER(Dims), C.POINTER(DecoderParams), C.c_void_p, C.c_void_p, C.c_void_p,
                                            C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                                            C.c_int, C.c_void_p]),
    "arcvae_decoder_cs_backward": (C.c_int, [C.POINTER(Dims), C.POINTER(DecoderParams), C.c_void_p, C.c_void_p, C.c_int,
                                             C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(DecoderParams),
                                             C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]),
    "arcvae_reparam_backward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                          C.c_void_p]),
    "arcvae_loss_fwd_bwd": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int64, C.c_int64, C.c_int,
                                      C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                      C.POINTER(LossHyper), C.c_uint64, C.c_uint64, C.c_int, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "arcvae_sampler_workspace_bytes": (C.c_size_t, [C.POINTER(Dims), C.c_int, C.c_int]),
    "arcvae_sample": (C.c_int, [C.POINTER(Dims), C.POINTER(DecoderParams), C.c_void_p, C.c_int, C.c_int, C.c_float,
                                C.c_int, C.c_int, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int,
                                C.c_void_p]),
    "arcvae_adam_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_float, C.c_float,
                                   C.c_float, C.c_float, C.c_float, C.c_void_p]),
    "arcvae_sumsq": (C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]),
    "arcvae_gemm_bf16": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                                   C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "arcvae_gemm_f32": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p,
                                  C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]),
}


def load():
    """Load the shared library (once).  Raises if it has not been built: there is no CPU path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ArcvaeError(f"{LIB_PATH} is missing — run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(nvcc, sm_100a).  This package has no CPU or PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int):
    if rc != 0:
        msg = load().arcvae_last_error()
        raise ArcvaeError(f"libarcvae_sm100 error {rc}: {msg.decode() if msg else '?'}")


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def ptr(t) -> int | None:
    return None if t is None else t.data_ptr()


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise ArcvaeError("mlx_vae_b200 runs on CUDA tensors only (no CPU fallback); got a CPU tensor")


TIME_CATEGORIES = ("gemm_f32", "recurrence", "loss", "adam", "gemm_tc", "pointwise", "sampler", "other")


def timing_enable(on: bool):
    check(load().arcvae_timing_enable(1 if on else 0))


def timing_read():
    ms = (C.c_double * 8)()
    cnt = (C.c_int * 8)()
    check(load().arcvae_timing_read(ms, cnt, 8))
    return {k: (ms[i], cnt[i]) for i, k in enumerate(TIME_CATEGORIES)}


def device_error_flag() -> int:
    """The sticky per-device error flag of the persistent kernels (0 = healthy).  Synchronises the current stream."""
    flag = C.c_int(0)
    check(load().arcvae_device_error_read(C.byref(flag), stream_ptr()))
    return int(flag.value)


def check_device_error():
    if device_error_flag() != 0:
        raise ArcvaeError("a persistent cluster kernel reported a barrier time-out on this device: the results of that "
                          "step are invalid and Adam updates are refused until arcvae_device_error_clear()")


def clear_device_error():
    check(load().arcvae_device_error_clear(stream_ptr()))


def flop_count(category: str) -> float:
    return float(load().arcvae_flop_count(TIME_CATEGORIES.index(category)))


def launch_count() -> int:
    return int(load().arcvae_launch_count())
