"""ctypes binding of the C-ABI CUDA library (include/movenet_b200.h).

There is no CPU implementation behind this module: if the shared library is
missing or a launcher reports an error, callers get an exception.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libmovenet_b200.so")

F32, BF16 = 0, 1
DECODE_CAUSAL, DECODE_REFERENCE = 0, 1


class Shape(C.Structure):
    """mirror of mvn_shape_t"""
    _fields_ = [(n, C.c_int) for n in (
        "layer_size", "stack_size", "input_channels", "residual_channels", "skip_channels",
        "context_in_channels", "batch", "frames", "has_video", "act_dtype", "remove_last", "output_logits", "no_grad")]

    def key(self):
        return tuple(getattr(self, n) for n, _ in self._fields_)


_P, _I, _SZ, _I64, _D = C.c_void_p, C.c_int, C.c_size_t, C.c_int64, C.c_double
_SP = C.POINTER(Shape)

# name -> (restype, argtypes); every int-returning launcher is error-checked
SIGNATURES = {
    "mvn_last_error": (C.c_char_p, []),
    "mvn_version": (_I, []),
    "mvn_launch_count": (C.c_ulonglong, []),
    "mvn_kernel_path": (_I, [_SP]),
    "mvn_receptive_fields": (_I, [_I, _I]),
    "mvn_output_size": (_I, [_I, _I, _I]),
    "mvn_packed_bytes": (_SZ, [_SP]),
    "mvn_acts_bytes": (_SZ, [_SP]),
    "mvn_scratch_bytes": (_SZ, [_SP]),
    "mvn_mulaw_encode": (_I, [_P, _I, _P, _I, _P, _I64, _P]),
    "mvn_mulaw_decode": (_I, [_P, _P, _I, _P, _I64, _P]),
    "mvn_one_hot": (_I, [_P, _P, _I, _I, _I, _P]),
    "mvn_pack_weights": (_I, [_SP, _P, _P, _P]),
    "mvn_unpack_grads": (_I, [_SP, _P, _P, _P, C.c_float, _P]),
    "mvn_peer_layout": (_I, [_SP, _P, _P]),
    "mvn_peer_alloc": (_I, [_SZ, _P, _P]),
    "mvn_peer_open": (_I, [_P, _P]),
    "mvn_peer_close": (_I, [_P]),
    "mvn_peer_free": (_I, [_P]),
    "mvn_peer_reduce_unpack": (_I, [_SP, _P, _I, _I, C.c_uint, _P, _P, _P, C.c_float, _P]),
    "mvn_codes_input": (_I, [_SP, _P, _P, _P]),
    "mvn_wavenet_forward": (_I, [_SP, _P, _P, _P, _P, _P, _P, _P]),
    "mvn_wavenet_backward": (_I, [_SP, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "mvn_fused_loss_supported": (_I, [_SP]),
    "mvn_wavenet_backward_loss": (_I, [_SP, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "mvn_adamw_segment_bytes": (_SZ, []),
    "mvn_adamw_chunk_elems": (_I, []),
    "mvn_adamw_step": (_I, [_P, _P, _I, _D, _D, _D, _D, _D, _D, _D, _D, _P, _P, _P]),
    "mvn_softmax_ce_partials": (_SZ, [_I, _I]),
    "mvn_softmax_ce_fwd": (_I, [_P, _P, _I, _I, _I, _P, _P, _P]),
    "mvn_softmax_ce_bwd": (_I, [_P, _P, _P, _I, _I, _I, _P, _P]),
    "mvn_onehot_to_codes": (_I, [_P, _I, _I, _I, _P, _P, _P]),
    "mvn_input_fwd": (_I, [_SP, _P, _P, _P, _P]),
    "mvn_video_fwd": (_I, [_SP, _P, _P, _P, _P]),
    "mvn_layer_fwd": (_I, [_SP, _P, _I, _P, _P, _P]),
    "mvn_head_fwd": (_I, [_SP, _P, _P, _P, _P, _P]),
    "mvn_layer_bwd": (_I, [_SP, _P, _I, _P, _P, _P, _P]),
    "mvn_read_activation": (_I, [_SP, _P, _I, _I, _P, _P]),
    "mvn_acts_offset": (_SZ, [_SP, _I, _I]),
    "mvn_decode_state_bytes": (_SZ, [_SP, _I]),
    "mvn_decode_prefill": (_I, [_SP, _P, _P, _I, _I, _P]),
    "mvn_decode_steps": (_I, [_SP, _P, _P, _P, _I, _I, _I, _P, _P, C.c_float, C.c_uint, _P]),
    "mvn_decode_tc_supported": (_I, [_SP]),
    "mvn_decode_tc_state_bytes": (_SZ, [_SP]),
    "mvn_decode_tc_prefill": (_I, [_SP, _P, _P, _P, _P]),
    "mvn_decode_tc_steps": (_I, [_SP, _P, _P, _I, _I, _P, _P, _P, C.c_float, C.c_uint, _P]),
}

_lib = None
#: bumped by every writer that updates parameters through raw pointers (movenet_b200.optim.AdamW): WaveNet._pack compares it
weights_epoch = [0]


class MovenetB200Error(RuntimeError):
    pass


def load():
    """Load the library once; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MovenetB200Error(
            f"{LIB_PATH} is missing: build it with `python -m movenet_b200.build` "
            "(movenet_b200 has no CPU or PyTorch fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def call(name, *args):
    """Call an int-returning launcher; non-zero -> MovenetB200Error(mvn_last_error())."""
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        raise MovenetB200Error(f"{name} failed ({rc}): {lib.mvn_last_error().decode()}")


def size(name, shape, *args):
    return int(getattr(load(), name)(C.byref(shape), *args))
