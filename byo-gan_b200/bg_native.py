"""ctypes binding of libbg_b200.so (include/bg_b200.h).

This is the only place the Python host touches the C ABI.  There is no fallback: if the library is
missing or a kernel reports an error the call raises.  Tensors are passed as raw device pointers;
the stream is torch's current CUDA stream.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libbg_b200.so")
_lib = None

_P = ctypes.c_void_p
_I = ctypes.c_int
_F = ctypes.c_float
_Z = ctypes.c_size_t

# name -> argtypes, mirrors include/bg_b200.h one to one
SIGNATURES = {
    "bg_pack_weight": [_P, _P, _P, _I, _I, _I, _I, _F, _P],
    "bg_pack_weight_grouped": [_P, _P, _P, _P, _P, _P, _P, _P, _I, _P],
    "bg_unpack_wgrad": [_P, _P, _I, _I, _I, _I, _F, _I, _P],
    "bg_conv_fprop": [_P, _P, _P, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P, _I, _F, _P],
    "bg_conv_fprop_stats": [_P, _P, _P, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P, _I, _F, _P, _I, _P],
    "bg_style_modulate": [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _F, _F, _P],
    "bg_conv_style_fprop": [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P, _P, _F, _P, _P],
    "bg_to_rgb_adain": [_P, _P, _P, _P, _P, _P, _I, _I, _I, _F, _F, _P],
    "bg_conv_pool_fprop": [_P, _P, _P, _I, _I, _I, _I, _I, _P, _P, _I, _F, _P],
    "bg_pack_weight_pool4": [_P, _P, _I, _I, _F, _P],
    "bg_conv_pool4_fprop": [_P, _P, _P, _I, _I, _I, _I, _I, _P, _P, _I, _F, _P],
    "bg_pack_weight_tconv4": [_P, _P, _I, _I, _F, _P],
    "bg_conv_pool4_dgrad": [_P, _P, _P, _I, _I, _I, _I, _I, _P, _F, _P, _P],
    "bg_conv_pool4_wgrad": [_P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "bg_unpack_wgrad_pool4": [_P, _P, _I, _I, _F, _I, _P],
    "bg_conv_pool4_supported": [_I, _I, _I, _I, _I],          # pure host-side predicate: no stream argument
    "bg_conv_fprop_tapwise": [_P, _P, _P, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P, _I, _F, _P],
    "bg_conv_wgrad": [_P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "bg_conv_wgrad_tapwise": [_P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "bg_act_gate": [_P, _P, _P, _Z, _F, _P],
    "bg_axpby": [_P, _P, _P, _Z, _F, _F, _P],
    "bg_pool_act_fwd": [_P, _P, _P, _I, _I, _I, _I, _F, _I, _P],
    "bg_pool_act_bwd": [_P, _P, _P, _I, _I, _I, _I, _F, _P, _P],
    "bg_upsample2x_fwd": [_P, _P, _I, _I, _I, _I, _P],
    "bg_upsample2x_bwd": [_P, _P, _I, _I, _I, _I, _P],
    "bg_channel_wsum": [_P, _P, _P, _Z, _I, _I, _Z, _Z, _I, _P],
    "bg_planes3_to_nhwc": [_P, _P, _P, _P, _P, _Z, _I, _I, _I, _I, _F, _I, _F, _P],
    "bg_nhwc_to_planes3": [_P, _P, _P, _P, _Z, _I, _I, _I, _I, _F, _P],
    "bg_in_stats": [_P, _P, _I, _I, _I, _P],
    "bg_adain_apply": [_P, _P, _P, _P, _I, _I, _I, _F, _P],
    "bg_adain_bwd_reduce": [_P, _P, _P, _P, _I, _I, _I, _F, _P],
    "bg_adain_bwd_apply": [_P, _P, _P, _P, _P, _P, _I, _I, _I, _F, _F, _I, _P, _P, _P],
    "bg_linear_fwd": [_P, _P, _P, _P, _I, _I, _I, _F, _I, _F, _P],
    "bg_linear_bwd_weight": [_P, _P, _P, _P, _I, _I, _I, _F, _I, _P],
    "bg_linear_bwd_input": [_P, _P, _P, _I, _I, _I, _F, _P],
    "bg_linear_fwd_grouped": [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _F, _P],
    "bg_linear_bwd_weight_grouped": [_P, _P, _P, _P, _P, _P, _I, _I, _I, _P],
    "bg_linear_bwd_input_grouped": [_P, _P, _P, _P, _I, _I, _I, _P, _P],
    "bg_transpose_f32": [_P, _P, _I, _I, _P],
    "bg_act_gate_f32": [_P, _P, _P, _Z, _F, _P],
    "bg_axpby_f32": [_P, _P, _P, _Z, _F, _F, _P],
    "bg_const_noise_act": [_P, _P, _P, _P, _I, _I, _I, _F, _P],
    "bg_const_bwd": [_P, _P, _I, _I, _I, _P],
    "bg_img_avgpool2": [_P, _P, _I, _I, _I, _P],
    "bg_img_avgpool2_bwd": [_P, _P, _I, _I, _I, _F, _I, _P],
    "bg_img_up2_lerp": [_P, _P, _P, _I, _I, _I, _F, _P],
    "bg_img_up2_bwd": [_P, _P, _I, _I, _I, _F, _P],
    "bg_plane_sums": [_P, _P, _I, _I, _P],
    "bg_nhwc_to_nchw_f32": [_P, _P, _I, _I, _I, _P],
    "bg_nchw_f32_to_nhwc": [_P, _P, _P, _I, _I, _I, _F, _P],
    "bg_mbstd_fwd": [_P, _P, _P, _P, _I, _I, _I, _I, _I, _F, _P],
    "bg_mbstd_bwd": [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _F, _P],
    "bg_logistic_loss": [_P, _I, _F, _P, _P, _F, _P],
    "bg_sumsq": [_P, _Z, _F, _P, _P],
    "bg_gp_rows": [_P, _I, _Z, _F, _F, _P, _P, _P],
    "bg_image_feed_u8": [_P, _P, _P, _I, _I, _I, _P],
}

# host-side switches / predicates called directly (no stream argument, no launch)
HOST_FUNCS = {"bg_set_deterministic": [_I], "bg_get_deterministic": [], "bg_set_sm_reserve": [_I],
              "bg_set_prezeroed_range": [_P, _Z]}

launch_count = 0  # C-ABI calls made through this binding; each launches >= 1 kernel (bench.py: gpu_launches)
_timing = None    # when a list: (name, args, start_event, end_event) per call, for bench.py's roofline leg


def start_timing():
    """Record a CUDA-event pair around every subsequent call (on the launching stream)."""
    global _timing
    _timing = []


def stop_timing():
    """Returns [(name, args, milliseconds)] for the calls since start_timing()."""
    global _timing
    rec, _timing = _timing or [], None
    torch.cuda.synchronize()
    return [(n, a, s.elapsed_time(e)) for (n, a, s, e) in rec]


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise RuntimeError(
                f"{_LIB_PATH} is missing: build it with `python byo-gan_b200/build.py` "
                "(there is no CPU or eager fallback for the hot path)"
            )
        l = ctypes.CDLL(_LIB_PATH)
        l.bg_last_error.restype = ctypes.c_char_p
        l.bg_abi_version.restype = _I
        for name, argtypes in list(SIGNATURES.items()) + list(HOST_FUNCS.items()):
            fn = getattr(l, name)
            fn.argtypes = argtypes
            fn.restype = _I
        _lib = l
    return _lib


def set_deterministic(on: bool) -> bool:
    """Chain-deterministic reductions on/off (include/bg_b200.h: bg_set_deterministic).  Returns the previous setting."""
    prev = bool(lib().bg_get_deterministic())
    lib().bg_set_deterministic(1 if on else 0)
    return prev


_fns = {}
_Tensor = torch.Tensor
_raw_stream = torch._C._cuda_getCurrentRawStream      # (device index) -> cudaStream_t of torch's current stream


def call(name, *args):
    """Invoke a C-ABI entry point; tensors -> device pointers, torch's current stream appended.
    This is the host-side hot loop (~700 calls per training iteration), so it does the minimum: one dict lookup, one
    pass over the arguments (CUDA + contiguity checks stay: a wrong pointer would be silent corruption)."""
    global launch_count
    fn = _fns.get(name)
    if fn is None:
        fn = _fns[name] = getattr(lib(), name)
    conv = []
    dev = -1
    for a in args:
        if a is None:
            conv.append(None)
        elif isinstance(a, (list, tuple)):
            # host array argument of a grouped entry point: device pointers, ints or floats
            if a and isinstance(a[0], float):
                conv.append((ctypes.c_float * len(a))(*a))
            elif a and isinstance(a[0], int):
                conv.append((ctypes.c_int * len(a))(*a))
            else:
                ptrs = []
                for t in a:
                    if t is None:
                        ptrs.append(None)
                    else:
                        if not (t.is_cuda and t.is_contiguous()):
                            raise RuntimeError("bg_b200 kernels need contiguous CUDA tensors")
                        if dev < 0:
                            dev = t.device.index
                        ptrs.append(t.data_ptr())
                conv.append((ctypes.c_void_p * len(a))(*ptrs))
        elif isinstance(a, _Tensor):
            if not a.is_cuda:
                raise RuntimeError("bg_b200 kernels need CUDA tensors (no CPU fallback exists)")
            if not a.is_contiguous():
                raise RuntimeError("bg_b200 kernels need contiguous tensors")
            if dev < 0:
                dev = a.device.index
            conv.append(a.data_ptr())
        else:
            conv.append(a)
    stream = _raw_stream(dev if dev >= 0 else torch.cuda.current_device())
    if _timing is not None:
        s_ev, e_ev = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s_ev.record()
        rc = fn(*conv, stream)
        e_ev.record()
        _timing.append((name, tuple(a for a in args if not isinstance(a, (_Tensor, list, tuple)) and a is not None), s_ev, e_ev))
    else:
        rc = fn(*conv, stream)
    launch_count += 1
    if rc != 0:
        raise RuntimeError(f"{name} failed ({rc}): {lib().bg_last_error().decode()}")
