"""ctypes binding of `libmggan_b200.so` (the C ABI declared in include/mggan_b200.h).

There is no fallback: if the library is missing, cannot be loaded, or the current device is
not an sm_100 part, every kernel call raises.  Build it with `python mg-gan_b200/build_ext.py`
(or `__graft_entry__.build()`); nvcc cross-compiles sm_100a without a GPU.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_C", "libmggan_b200.so")

TABLE_MAX = 64


class TensorTable(ctypes.Structure):
    _fields_ = [
        ("p", ctypes.c_void_p * TABLE_MAX),
        ("g", ctypes.c_void_p * TABLE_MAX),
        ("m", ctypes.c_void_p * TABLE_MAX),
        ("v", ctypes.c_void_p * TABLE_MAX),
        ("n", ctypes.c_int * TABLE_MAX),
        ("bc1", ctypes.c_float * TABLE_MAX),
        ("bc2_sqrt", ctypes.c_float * TABLE_MAX),
        ("dyn", ctypes.c_void_p),
    ]


PEER_MAX = 16


class PeerTable(ctypes.Structure):
    _fields_ = [("region", ctypes.c_void_p * PEER_MAX), ("flags", ctypes.c_void_p * PEER_MAX),
                ("rank", ctypes.c_int), ("world", ctypes.c_int)]


_CT = {"T": ctypes.POINTER(PeerTable), "q": ctypes.c_longlong, "p": ctypes.c_void_p, "i": ctypes.c_int, "f": ctypes.c_float, "d": ctypes.c_double,
       "Q": ctypes.c_ulonglong, "s": ctypes.c_void_p, "t": ctypes.POINTER(TensorTable)}

# name -> argument codes (p pointer, i int, q long long, f float, d double, Q uint64, s stream, t tensor table*, T peer table*)
SIGNATURES = {
    "mggan_lstm_seq_fwd": "piiippppps",
    "mggan_lstm_seq_bwd": "piiipppppps",
    "mggan_linear_fwd": "piippiifps",
    "mggan_linear_bwd": "piipiifppppps",
    "mggan_disc_heads_fwd": "piiipppppppipps",
    "mggan_disc_heads_bwd": "piiipppppipppppps",
    "mggan_gumbel_sample": "piiiQQpps",
    "mggan_selection_build": "piiiippppppppps",
    "mggan_selection_all": "iiipppps",
    "mggan_decoder_fwd": "ipppppppppipppppppppiippppps",
    "mggan_decoder_fwd_tc": "ipppppppppipppppppppiippppps",
    "mggan_tc_selftest": "ppps",
    "mggan_decoder_bwd": "ipppppppipppppppppiippppppppppppppppps",
    "mggan_social_attn_fwd": "ppipppipppppps",
    "mggan_social_attn_bwd": "ppipppippppppppppppps",
    "mggan_scene_patch_stats": "ppipps",
    "mggan_scene_bn1_from_patches": "ppdipppppppffipps",
    "mggan_scene_fused12_fwd": "ppiippppppppps",
    "mggan_scene_bn_finalize": "pdipppppffipps",
    "mggan_scene_attn_fwd": "piipppppps",
    "mggan_scene_attn_bwd": "piipppppppppppppps",
    "mggan_scene_bn_bwd_finalize": "pdippps",
    "mggan_scene_fused12_bwd": "ppiippppppppppppppps",
    "mggan_scene_bn1_bwd_finalize": "ppdipppppppppps",
    "mggan_l2_scene_min": "ppiiipifippps",
    "mggan_bce_scalar_label": "pifppfpps",
    "mggan_mse_scalar_label": "pifppfpps",
    "mggan_ce_generators": "piippfpps",
    "mggan_pm_ml_loss": "ppiiiipfffppps",
    "mggan_scene_crop": "ppppippips",
    "mggan_tube_inside": "pippipps",
    "mggan_min_ade_fde": "ppiiipipfppps",
    "mggan_grad_sqnorm": "tips",
    "mggan_clip_adamw": "tipfffffffs",
    "mggan_multi_copy": "tis",
    "mggan_peer_allreduce": "Tipqpps",
    "mggan_mlp2_fwd": "pqippiifppiifps",
    "mggan_mlp2_bwd": "pqippiifpiifppppppps",
}
# plain (non status-returning) helpers
_PLAIN = {"mggan_version": ("", ctypes.c_int), "mggan_device_check": ("", ctypes.c_int),
          "mggan_selection_tiles": ("ii", ctypes.c_int), "mggan_last_error": ("", ctypes.c_char_p),
          "mggan_set_gemm_variant": ("i", ctypes.c_int)}

EXPORTED = sorted(list(SIGNATURES) + list(_PLAIN))


class MgganCudaError(RuntimeError):
    pass


_lib = None
_device_ok = False


def load():
    """dlopen the library (no device needed) and declare every prototype."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: the MG-GAN B200 path has no CPU/PyTorch fallback. "
            "Build it with `python mg-gan_b200/build_ext.py`.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, sig in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = [_CT[c] for c in sig]
        fn.restype = ctypes.c_int
    for name, (sig, res) in _PLAIN.items():
        fn = getattr(lib, name)
        fn.argtypes = [_CT[c] for c in sig]
        fn.restype = res
    _lib = lib
    if os.environ.get("MGGAN_GEMM", "1") in ("2", "3"):
        lib.mggan_set_gemm_variant(int(os.environ["MGGAN_GEMM"]))
    return lib


def _ensure_device():
    global _device_ok
    if _device_ok:
        return
    if not torch.cuda.is_available():
        raise MgganCudaError("no CUDA device: the MG-GAN B200 path cannot run (there is no CPU fallback)")
    lib = load()
    if lib.mggan_device_check() != 0:
        raise MgganCudaError(lib.mggan_last_error().decode())
    _device_ok = True


def ptr(t):
    """Device pointer of a contiguous tensor (None -> NULL)."""
    if t is None:
        return None
    if not (t.is_cuda and t.is_contiguous()):
        if not t.is_cuda:
            raise MgganCudaError(
                f"tensor on {t.device}: the MG-GAN B200 path only runs on CUDA (there is no CPU fallback)")
        raise AssertionError(("non-contiguous tensor", t.shape, t.stride()))
    return t.data_ptr()


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def stream():
    """cudaStream_t of torch's current stream on the current device (raw handle; no Stream object is built)."""
    if _raw_stream is not None:
        return _raw_stream(torch.cuda.current_device())
    return torch.cuda.current_stream().cuda_stream


# kernels launched per entry point (for the launch counter; memsets not counted)
KERNELS_PER_CALL = {"mggan_selection_build": 4, "mggan_linear_bwd": 2}
launch_count = 0          # kernels launched through this binding since import
_profile = None           # None | {"only": set or None, "events": [(name, start, end), ...]}


def profile_start(only=None, detail=False):
    """Record a CUDA-event pair around every call (or only the named entry points) on the launching stream.
    detail=True keys the result by entry point AND its integer arguments (shapes)."""
    global _profile
    _profile = {"only": set(only) if only else None, "events": [], "detail": detail}


def profile_stop():
    """-> {name: (calls, total_ms)}; synchronises the device."""
    global _profile
    prof, _profile = _profile, None
    torch.cuda.synchronize()
    out = {}
    for name, s, e in prof["events"]:
        c, t = out.get(name, (0, 0.0))
        out[name] = (c + 1, t + s.elapsed_time(e))
    return out


def call(name, *args):
    """Invoke a status-returning entry point on the current stream; raise on a non-zero status."""
    global launch_count
    _ensure_device()
    lib = _lib
    prof = _profile
    if prof is not None and (prof["only"] is None or name in prof["only"]):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        rc = getattr(lib, name)(*args, stream())
        e.record()
        key = name + str(tuple(a for a in args if isinstance(a, int) and not isinstance(a, bool) and abs(a) < (1 << 31))) \
            if prof.get("detail") else name
        prof["events"].append((key, s, e))
    else:
        rc = getattr(lib, name)(*args, stream())
    if rc != 0:
        raise MgganCudaError(f"{name} failed ({rc}): {lib.mggan_last_error().decode()}")
    launch_count += KERNELS_PER_CALL.get(name, 1)


def set_gemm_variant(variant):
    """GEMM kernel behind mggan_linear_*: 1 (default), 2 (FP32 128 x 64 tile, register prefetch) or 3 (tcgen05 tensor
    cores, 3 x TF32); returns the previous one.  MGGAN_GEMM=2 / 3 in the environment selects a variant at import."""
    prev = load().mggan_set_gemm_variant(int(variant))
    if prev < 0:
        raise ValueError(f"unknown GEMM variant {variant}")
    return prev


def selection_tiles(n_seq, num_gens):
    return load().mggan_selection_tiles(int(n_seq), int(num_gens))
