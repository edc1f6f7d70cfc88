"""ctypes binding of ``libciaosr_b200.so`` (the C ABI in include/ciaosr_b200.h).

There is no fallback: if the shared library has not been built (run
``python -c "import __graft_entry__ as g; g.build()"``) or cannot be loaded,
importing the hot path fails with an ImportError that says so.
"""
import ctypes
import os
from ctypes import (POINTER, Structure, c_char_p, c_float, c_int, c_int32, c_longlong,
                    c_size_t, c_void_p)

ABI_VERSION = 1
MAX_LAYERS = 8
MAX_SCALES = 4
N_STAGES = 7
STAGE_NAMES = ("layout", "cross_scale_attn", "lr_precompute", "pair_mlp", "query_mlp", "rdn_encoder",
               "encoder_linear")

ENGINE_AUTO, ENGINE_SIMT, ENGINE_TCGEN05 = 0, 1, 2
ENGINES = {"auto": ENGINE_AUTO, "simt": ENGINE_SIMT, "tcgen05": ENGINE_TCGEN05}

E_INVALID, E_WORKSPACE, E_CUDA, E_NO_DEVICE = -1, -2, -3, -4

LIB_PATH = os.environ.get(
    "CIAOSR_LIB", os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", "libciaosr_b200.so"))

# every symbol include/ciaosr_b200.h declares
EXPORTS = (
    "ciaosr_last_error", "ciaosr_abi_version", "ciaosr_launch_count",
    "ciaosr_engine_supported", "ciaosr_profile_enable", "ciaosr_profile_read",
    "ciaosr_plan_bytes", "ciaosr_plan_init", "ciaosr_workspace_bytes",
    "ciaosr_cross_scale_attn_workspace_bytes", "ciaosr_cross_scale_attn_forward", "ciaosr_query_rgb_forward",
    "ciaosr_tile_blend_accumulate", "ciaosr_tile_blend_finish",
    "ciaosr_rdn_plan_bytes", "ciaosr_rdn_plan_init", "ciaosr_rdn_workspace_bytes", "ciaosr_rdn_forward",
    "ciaosr_linear_plan_bytes", "ciaosr_linear_plan_init", "ciaosr_linear_forward",
    "ciaosr_window_attention_forward", "ciaosr_layernorm_forward", "ciaosr_linear_forward_res",
    "ciaosr_conv3x3_plan_bytes", "ciaosr_conv3x3_plan_init", "ciaosr_conv3x3_nhwc_forward",
    "ciaosr_linear_forward_split", "ciaosr_layernorm_split_forward", "ciaosr_window_attention_split_forward",
)


class MlpDesc(Structure):
    _fields_ = [("n_layers", c_int32),
                ("dims", c_int32 * (MAX_LAYERS + 1)),
                ("weight", c_void_p * MAX_LAYERS),
                ("bias", c_void_p * MAX_LAYERS)]


class CsAttnDesc(Structure):
    _fields_ = [("channels", c_int32), ("n_scales", c_int32),
                ("scales", c_int32 * MAX_SCALES), ("softmax_scale", c_float),
                ("match1_w", c_void_p), ("match1_b", c_void_p), ("match1_slope", c_void_p),
                ("match2_w", c_void_p), ("match2_b", c_void_p), ("match2_slope", c_void_p),
                ("assembly_w", c_void_p), ("assembly_b", c_void_p), ("assembly_slope", c_void_p),
                ("down_w", c_void_p), ("down_b", c_void_p), ("escape_nan", c_void_p)]


class HeadDesc(Structure):
    _fields_ = [("abi_version", c_int32), ("channels", c_int32), ("feat_unfold", c_int32),
                ("local_size", c_int32), ("non_local_attn", c_int32), ("softmax_scale", c_float),
                ("imnet_q", MlpDesc), ("imnet_k", MlpDesc), ("imnet_v", MlpDesc),
                ("cs_attn", CsAttnDesc)]


class RdnDesc(Structure):
    _fields_ = [("abi_version", c_int32), ("mid_channels", c_int32), ("channel_growth", c_int32),
                ("num_blocks", c_int32), ("num_layers", c_int32),
                ("sfe1_w", c_void_p), ("sfe1_b", c_void_p), ("sfe2_w", c_void_p), ("sfe2_b", c_void_p),
                ("dense_w", POINTER(c_void_p)), ("dense_b", POINTER(c_void_p)),
                ("lff_w", POINTER(c_void_p)), ("lff_b", POINTER(c_void_p)),
                ("gff0_w", c_void_p), ("gff0_b", c_void_p), ("gff1_w", c_void_p), ("gff1_b", c_void_p)]


class LinearDesc(Structure):
    _fields_ = [("abi_version", c_int32), ("in_features", c_int32), ("out_features", c_int32),
                ("weight", c_void_p), ("bias", c_void_p)]


class Conv3x3Desc(Structure):
    _fields_ = [("abi_version", c_int32), ("in_channels", c_int32), ("out_channels", c_int32),
                ("weight", c_void_p), ("bias", c_void_p)]


class CiaoSRNativeError(RuntimeError):
    """A C-ABI call returned a CIAOSR_E_* code."""

    def __init__(self, code, message):
        super().__init__(f"libciaosr_b200: {message} (code {code})")
        self.code = code


_lib = None


def load():
    """Load the shared library once and declare the prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: the CUDA library has not been built. Run "
            "`python -c \"import __graft_entry__ as g; g.build()\"` from the repo root. "
            "ciaosr_b200 has no CPU or PyTorch fallback for the head.")
    try:
        lib = ctypes.CDLL(LIB_PATH)
    except OSError as exc:  # pragma: no cover - depends on the box
        raise ImportError(f"cannot load {LIB_PATH}: {exc}") from exc
    missing = [s for s in EXPORTS if not hasattr(lib, s)]
    if missing:
        raise ImportError(f"{LIB_PATH} does not export {missing}; rebuild it")

    lib.ciaosr_last_error.restype = c_char_p
    lib.ciaosr_last_error.argtypes = []
    lib.ciaosr_abi_version.restype = c_int
    lib.ciaosr_launch_count.restype = c_longlong
    lib.ciaosr_engine_supported.argtypes = [POINTER(HeadDesc), c_int]
    lib.ciaosr_profile_enable.argtypes = [c_int]
    lib.ciaosr_profile_read.argtypes = [POINTER(c_float), POINTER(c_int), c_int]
    lib.ciaosr_plan_bytes.argtypes = [POINTER(HeadDesc), POINTER(c_size_t)]
    lib.ciaosr_plan_init.argtypes = [POINTER(HeadDesc), c_void_p, c_size_t, c_void_p]
    lib.ciaosr_workspace_bytes.argtypes = [POINTER(HeadDesc), c_int, c_int, c_int, c_int, c_int,
                                           POINTER(c_size_t)]
    lib.ciaosr_cross_scale_attn_workspace_bytes.argtypes = [POINTER(HeadDesc), c_int, c_int, c_int, c_int,
                                                            POINTER(c_size_t)]
    lib.ciaosr_cross_scale_attn_forward.argtypes = [
        POINTER(HeadDesc), c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_size_t,
        c_void_p]
    lib.ciaosr_query_rgb_forward.argtypes = [
        POINTER(HeadDesc), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
        c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_size_t, c_void_p]
    lib.ciaosr_tile_blend_accumulate.argtypes = [
        c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]
    lib.ciaosr_tile_blend_finish.argtypes = [
        c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p]
    lib.ciaosr_rdn_plan_bytes.argtypes = [POINTER(RdnDesc), POINTER(c_size_t)]
    lib.ciaosr_rdn_plan_init.argtypes = [POINTER(RdnDesc), c_void_p, c_size_t, c_void_p]
    lib.ciaosr_rdn_workspace_bytes.argtypes = [POINTER(RdnDesc), c_int, c_int, c_int, POINTER(c_size_t)]
    lib.ciaosr_rdn_forward.argtypes = [POINTER(RdnDesc), c_void_p, c_void_p, c_int, c_int, c_int, c_void_p,
                                       c_void_p, c_size_t, c_void_p]
    lib.ciaosr_linear_plan_bytes.argtypes = [POINTER(LinearDesc), POINTER(c_size_t)]
    lib.ciaosr_linear_plan_init.argtypes = [POINTER(LinearDesc), c_void_p, c_size_t, c_void_p]
    lib.ciaosr_linear_forward.argtypes = [POINTER(LinearDesc), c_void_p, c_void_p, c_longlong, c_int, c_void_p,
                                          c_void_p]
    lib.ciaosr_window_attention_forward.argtypes = [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                                                    c_int, c_float, c_void_p, c_void_p]
    lib.ciaosr_linear_forward_res.argtypes = [POINTER(LinearDesc), c_void_p, c_void_p, c_longlong, c_int, c_void_p,
                                              c_void_p, c_void_p]
    lib.ciaosr_conv3x3_plan_bytes.argtypes = [POINTER(Conv3x3Desc), POINTER(c_size_t)]
    lib.ciaosr_conv3x3_plan_init.argtypes = [POINTER(Conv3x3Desc), c_void_p, c_size_t, c_void_p]
    lib.ciaosr_conv3x3_nhwc_forward.argtypes = [POINTER(Conv3x3Desc), c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                                c_void_p, c_void_p, c_void_p]
    lib.ciaosr_linear_forward_split.argtypes = [POINTER(LinearDesc), c_void_p, c_void_p, c_void_p, c_int, c_longlong,
                                                c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p]
    lib.ciaosr_layernorm_split_forward.argtypes = [c_void_p, c_void_p, c_void_p, c_float, c_longlong, c_int, c_void_p,
                                                   c_void_p, c_int, c_void_p]
    lib.ciaosr_window_attention_split_forward.argtypes = [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                                                          c_int, c_float, c_void_p, c_void_p, c_int, c_void_p]
    lib.ciaosr_layernorm_forward.argtypes = [c_void_p, c_void_p, c_void_p, c_float, c_longlong, c_int, c_void_p,
                                             c_void_p]
    for name in EXPORTS[3:]:  # everything after the three non-int getters
        getattr(lib, name).restype = c_int
    if lib.ciaosr_abi_version() != ABI_VERSION:
        raise ImportError(f"{LIB_PATH} has ABI {lib.ciaosr_abi_version()}, bindings expect {ABI_VERSION}")
    _lib = lib
    return lib


def check(code):
    if code != 0:
        raise CiaoSRNativeError(code, load().ciaosr_last_error().decode("utf-8", "replace"))
