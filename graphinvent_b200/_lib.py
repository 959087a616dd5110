"""
ctypes binding of libgib200.so (C-ABI declared in include/gib200.h).

There is NO fallback: if the library cannot be built or loaded the import fails, and every
entry point raises on a non-zero return code.
"""
import ctypes
import os

from . import build as _build

c_p = ctypes.c_void_p
c_i = ctypes.c_int
c_ll = ctypes.c_longlong
c_f = ctypes.c_float
c_d = ctypes.c_double
c_sz = ctypes.c_size_t


class Dims(ctypes.Structure):
    """mirror of `gib_dims` (include/gib200.h)"""
    _fields_ = [(n, c_i) for n in (
        "model", "B", "N", "F", "Ef", "H", "M", "T", "msg_hidden", "msg_depth", "att_hidden", "att_depth",
        "eemb_hidden", "eemb_depth", "gather_width", "gatt_hidden", "gatt_depth", "gemb_hidden", "gemb_depth",
        "mlp1_hidden", "mlp1_depth", "mlp2_hidden", "mlp2_depth", "f_add", "f_conn")] + [("big", c_f), ("in_dtype", c_i)]


MODEL_ID = {"GGNN": 0, "MNN": 1, "AttGGNN": 2, "EMN": 3}
HDR_INTS = 16
HDR_E, HDR_P, HDR_TYPE_COUNT, HDR_TYPE_BASE, HDR_FLAGS, HDR_CAPACITY = 0, 1, 2, 6, 11, 12
FLAG_MULTITYPE, FLAG_NONBINARY, FLAG_OVERFLOW = 1, 2, 4
ABI_VERSION = 201      # must equal gib_version() of the loaded library (include/gib200.h)

_PROTOS = {
    "gib_last_error": (ctypes.c_char_p, []),
    "gib_version": (c_i, []),
    "gib_set_tensor_cores": (None, [c_i]),
    "gib_get_tensor_cores": (c_i, []),
    "gib_tc_debug": (None, [c_i]),
    "gib_device_sm_count": (c_i, []),
    "gib_scatter_variant": (None, [c_i]),
    "gib_tc_trace": (None, [c_p, c_i]),
    "gib_graph_count_ws_bytes": (c_sz, [c_p]),
    "gib_graph_count": (c_i, [c_p, c_p, c_p, c_p]),
    "gib_graph_header_capacity": (c_i, [c_p, c_i, c_p, c_p]),
    "gib_graph_bytes": (c_sz, [c_p, c_p]),
    "gib_graph_fill": (c_i, [c_p, c_p, c_p, c_p, c_p, c_p]),
    "gib_graph_array": (c_p, [c_p, c_p, c_p, c_i]),
    "gib_model_num_params": (c_i, [c_p]),
    "gib_model_param_numel": (c_ll, [c_p, c_i]),
    "gib_model_packed_bytes": (c_sz, [c_p]),
    "gib_model_pack": (c_i, [c_p, c_p, c_p, c_p]),
    "gib_model_workspace_bytes": (c_sz, [c_p, c_p]),
    "gib_model_forward": (c_i, [c_p] * 9),
    "gib_model_bwd_scratch_bytes": (c_sz, [c_p, c_p]),
    "gib_model_backward": (c_i, [c_p] * 12),
    "gib_model_backward_part": (c_i, [c_p] * 11 + [c_i, c_p]),
    "gib_kl_loss_fwd_bwd": (c_i, [c_p, c_p, c_i, c_i, c_f, c_p, c_p, c_p]),
    "gib_linear_fwd": (c_i, [c_p, c_i, c_p, c_i, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_p]),
    "gib_linear_fwd_tc": (c_i, [c_p, c_i, c_p, c_i, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_p]),
    "gib_linear_fwd_tc_planes": (c_i, [c_p, c_i, c_p, c_p, c_i, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_p, c_p, c_p]),
    "gib_split_planes": (c_i, [c_p, c_p, c_p, c_ll, c_p]),
    "gib_dw_scratch_bytes": (c_sz, [c_i, c_i, c_i]),
    "gib_linear_bwd_dw": (c_i, [c_p, c_i, c_i, c_p, c_i, c_i, c_i, c_p, c_p, c_i, c_i, c_p, c_p, c_p, c_p]),
    "gib_scatter_sum": (c_i, [c_p, c_p, c_i, c_p, c_p, c_p, c_ll, c_p]),
    "gib_seg_softmax": (c_i, [c_p, c_p, c_p, c_i, c_p, c_p, c_p, c_ll, c_p]),
    "gib_gru_gates": (c_i, [c_p, c_p, c_p, c_p, c_i, c_p, c_ll, c_p]),
    "gib_graph_gather": (c_i, [c_p, c_p, c_p, c_p, c_i, c_p, c_i, c_i, c_f, c_p]),
    "gib_validation_nll": (c_i, [c_p, c_p, c_i, c_i, c_p, c_p]),
    "gib_sum_scaled": (c_i, [c_p, c_i, c_f, c_p, c_p]),
    "gib_fill_zero": (c_i, [c_p, c_sz, c_p]),
    "gib_adam_step": (c_i, [c_p, c_p, c_p, c_p, c_ll, c_ll, c_d, c_d, c_d, c_d, c_d, c_d, c_p]),
    "gib_sample_actions": (c_i, [c_p, c_i, c_i, c_p, c_p, c_p, c_p]),
    "gib_generation_scratch_bytes": (c_sz, [c_i]),
    "gib_generation_round": (c_i, [c_i] * 7 + [c_p] * 11 + [c_i, c_p, c_p, c_p]),
    "gib_profile_enable": (None, [c_i]),
    "gib_launch_count": (c_ll, []),
    "gib_profile_collect": (c_i, [c_p, c_p, c_p]),
    "gib_profile_records": (c_i, [c_p, c_p, c_p, c_i]),
}


def exported_symbols():
    """every symbol include/gib200.h declares (checked by the CPU test-suite)"""
    return sorted(_PROTOS)


def _load():
    path = _build.LIB
    try:
        path = _build.build()
    except Exception as e:  # no nvcc on this box (the GPU box runs the library built in the container)
        if not os.path.exists(path):
            raise ImportError(f"libgib200.so is missing and could not be built: {e}") from e
        if _build.is_stale():
            import warnings
            warnings.warn(f"libgib200.so is older than its sources and could not be rebuilt ({e}); loading it anyway "
                          "-- the ABI version check below decides")
    lib = ctypes.CDLL(path)
    for name, (res, args) in _PROTOS.items():
        fn = getattr(lib, name)  # AttributeError here == header/library mismatch: fail loudly
        fn.restype = res
        fn.argtypes = args
    got = lib.gib_version()
    if got != ABI_VERSION:
        raise ImportError(f"{path} implements ABI version {got}, this package binds version {ABI_VERSION}: "
                          "rebuild with `python -m graphinvent_b200.build --force`")
    return lib


lib = _load()
LIB_PATH = _build.LIB
# A/B measurements only: GIB_TC_DEBUG=1 routes the dense GEMMs to the first-generation tcgen05 kernel for the whole
# process (include/gib200.h, gib_tc_debug) so that any test or tool can be re-run against it
if os.environ.get("GIB_TC_DEBUG"):
    lib.gib_tc_debug(int(os.environ["GIB_TC_DEBUG"], 0))


def check(rc, what=""):
    if rc == 0:
        return
    msg = lib.gib_last_error().decode(errors="replace")
    if rc > 0:
        raise RuntimeError(f"{what}: CUDA error {rc} ({msg})")
    raise RuntimeError(f"{what}: invalid argument ({rc}): {msg}")
