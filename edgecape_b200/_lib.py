"""ctypes binding of the C ABI declared in include/edgecape_b200.h.

There is no fallback: if the shared library is missing, or no CUDA device is present when a
compute entry point is called, this module raises.  The oracle under oracle/ is never imported
from here (or from anywhere in this package).
"""
import ctypes
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libedgecape_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(HERE), "include", "edgecape_b200.h")

c_fp = ctypes.c_void_p       # every device pointer travels as void*
c_int = ctypes.c_int
c_ll = ctypes.c_longlong
c_f = ctypes.c_float
c_d = ctypes.c_double
c_sz = ctypes.c_size_t

# name -> (restype, argtypes)   -- must mirror include/edgecape_b200.h
SIGNATURES = {
    "ec_version": (c_int, []),
    "ec_last_error_string": (ctypes.c_char_p, []),
    "ec_launch_count": (c_ll, []),
    "ec_gemm": (c_int, [c_fp, c_fp, c_fp, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_ll, c_ll,
                        c_ll, c_fp, c_int, c_fp, c_fp, c_int, c_ll, c_int, c_fp]),
    "ec_split_f16": (c_int, [c_fp, c_fp, c_int, c_int, c_int, c_int, c_ll, c_int, c_f, c_fp]),
    "ec_gemm_f16x3": (c_int, [c_fp, c_fp, c_fp, c_int, c_int, c_int, c_int, c_int, c_ll, c_f, c_fp, c_int, c_fp, c_fp, c_int,
                              c_int, c_int, c_fp, c_int, c_f, c_fp]),
    "ec_tc_set_tile_n": (c_int, [c_int]),
    "ec_tc_mode_launches": (c_ll, [c_int]),
    "ec_tc_set_cta_limit": (c_int, [c_int]),
    "ec_tc_set_split_tma": (c_int, [c_int]),
    "ec_tc_set_trace": (c_int, [c_fp]),
    "ec_tc_set_dynamic": (c_int, [c_int]),
    "ec_tc_set_ksplit": (c_int, [c_int]),
    "ec_tc_ksplit_launches": (c_ll, []),
    "ec_set_pdl": (c_int, [c_int]),
    "ec_tc_set_debug": (c_int, [c_int]),
    "ec_attention_tc_set_trace": (c_int, [c_fp, c_int]),
    "ec_attention_tc_set_variant": (c_int, [c_int]),
    "ec_layernorm": (c_int, [c_fp, c_int, c_int, c_ll, c_fp, c_int, c_fp, c_int, c_fp, c_int, c_fp, c_fp, c_f,
                             c_int, c_int, c_fp, c_int, c_int, c_fp]),
    "ec_add_rows": (c_int, [c_fp, c_fp, c_int, c_int, c_int, c_int, c_fp]),
    "ec_copy_rows": (c_int, [c_fp, c_int, c_int, c_ll, c_fp, c_int, c_int, c_ll, c_int, c_int, c_int, c_fp]),
    "ec_gather_blocks": (c_int, [c_fp, c_ll, c_fp, c_fp, c_ll, c_int, c_ll, c_fp]),
    "ec_axpby": (c_int, [c_fp, c_fp, c_fp, c_f, c_f, c_f, c_ll, c_fp]),
    "ec_attention": (c_int, [c_fp, c_fp, c_fp, c_fp, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                             c_int, c_ll, c_ll, c_ll, c_ll, c_f, c_fp, c_fp, c_fp, c_int, c_fp]),
    "ec_attention_tc": (c_int, [c_fp, c_fp, c_fp, c_fp, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                c_int, c_ll, c_ll, c_ll, c_ll, c_f, c_fp, c_fp, c_fp, c_int, c_fp]),
    "ec_attention_tc_split": (c_int, [c_fp, c_int, c_int, c_int, c_int, c_fp, c_int, c_int, c_int, c_fp, c_int, c_int,
                                      c_int, c_int, c_fp, c_int, c_int, c_int, c_int, c_int, c_ll, c_f, c_int, c_fp, c_fp,
                                      c_fp, c_int, c_fp]),
    "ec_attention_hop_bias_next": (c_int, [c_fp, c_int, c_int, c_fp, c_fp, c_fp, c_fp]),
    "ec_attention_split_fmt_next": (c_int, [c_int]),
    "ec_attention_set_cta_limit": (c_int, [c_int]),
    "ec_hop_bias": (c_int, [c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_int, c_int, c_int, c_int, c_int, c_fp]),
    "ec_mask_accumulate": (c_int, [c_fp, c_fp, c_int, c_int, c_fp]),
    "ec_kp_masks": (c_int, [c_fp, c_fp, c_fp, c_int, c_int, c_fp]),
    "ec_adj_from_edges": (c_int, [c_fp, c_fp, c_fp, c_fp, c_fp, c_int, c_int, c_fp]),
    "ec_soft_normalize_adj": (c_int, [c_fp, c_fp, c_fp, c_int, c_int, c_fp]),
    "ec_l2_normalize": (c_int, [c_fp, c_fp, c_int, c_int, c_f, c_fp]),
    "ec_edge_weights": (c_int, [c_fp, c_fp, c_fp, c_f, c_f, c_int, c_fp, c_fp, c_fp, c_fp, c_int, c_int, c_fp]),
    "ec_markov_powers": (c_int, [c_fp, c_int, c_int, c_int, c_fp]),
    "ec_gcn_pack_weights": (c_int, [c_fp, c_fp, c_fp, c_int, c_int, c_fp]),
    "ec_gcn": (c_int, [c_fp, c_fp, c_fp, c_fp, c_int, c_int, c_int, c_int, c_fp, c_sz, c_fp]),
    "ec_gcn_aggregate_split": (c_int, [c_fp, c_fp, c_fp, c_int, c_int, c_int, c_int, c_fp]),
    "ec_workspace_bytes_gcn": (c_sz, [c_int, c_int, c_int, c_int]),
    "ec_gcn_fused_slice": (c_int, [c_int, c_int, c_int]),
    "ec_gcn_fused_set_trace": (c_int, [c_fp, c_int]),
    "ec_gcn_fused_set_debug": (c_int, [c_int]),
    "ec_split_f16f8": (c_int, [c_fp, c_fp, c_int, c_int, c_int, c_int, c_ll, c_int, c_f, c_int, c_fp]),
    "ec_gemm_f16f8": (c_int, [c_fp, c_fp, c_fp, c_int, c_int, c_int, c_int, c_int, c_ll, c_f, c_fp, c_int, c_fp, c_fp, c_int,
                              c_int, c_int, c_fp, c_int, c_f, c_int, c_fp]),
    "ec_overflow_count": (c_int, [c_fp, c_int]),
    "ec_gcn_fused": (c_int, [c_fp, c_fp, c_fp, c_fp, c_int, c_f, c_fp, c_fp, c_int, c_int, c_int, c_int, c_int, c_fp]),
    "ec_gcn_fused2": (c_int, [c_fp, c_fp, c_fp, c_fp, c_int, c_f, c_fp, c_fp, c_int, c_int, c_int, c_int, c_int, c_fp]),
    "ec_gcn_fused2_slice": (c_int, [c_int, c_int, c_int]),
    "ec_gcn_fused2_set_trace": (c_int, [c_fp, c_int]),
    "ec_gcn_fused2_set_debug": (c_int, [c_int]),
    "ec_gcn_fused2_set_cta_limit": (c_int, [c_int]),
    "ec_support_weights": (c_int, [c_fp, c_fp, c_fp, c_int, c_int, c_int, c_int, c_int, c_int, c_fp]),
    "ec_sine_pe_coords": (c_int, [c_fp, c_fp, c_int, c_int, c_int, c_f, c_f, c_fp]),
    "ec_proposal": (c_int, [c_fp, c_fp, c_fp, c_fp, c_int, c_int, c_int, c_fp]),
    "ec_point_update": (c_int, [c_fp, c_fp, c_int, c_fp, c_int, c_fp]),
    "ec_decode_preds": (c_int, [c_fp, c_fp, c_fp, c_int, c_int, c_f, c_f, c_int, c_fp]),
    "ec_im2col_patches": (c_int, [c_fp, c_fp, c_int, c_int, c_int, c_int, c_int, c_fp, c_int, c_fp]),
    "ec_interp_pos_embed": (c_int, [c_fp, c_fp, c_int, c_int, c_int, c_int, c_d, c_fp]),
    "ec_write_cls": (c_int, [c_fp, c_fp, c_fp, c_int, c_ll, c_int, c_fp]),
    "ec_warp_affine_normalize_u8": (c_int, [c_fp, c_int, c_int, c_ll, c_fp, c_fp, c_int, c_int, c_fp, c_fp, c_fp]),
    "ec_msra_targets": (c_int, [c_fp, c_int, c_fp, c_int, c_fp, c_fp, c_int, c_int, c_int, c_int, c_int, c_f, c_fp]),
    "ec_pck_accumulate": (c_int, [c_fp, c_fp, c_fp, c_fp, c_fp, c_int, c_fp, c_int, c_int, c_fp]),
    "ec_metrics_accumulate": (c_int, [c_fp, c_fp, c_fp, c_fp, c_fp, c_int, c_int, c_fp, c_int, c_int, c_fp]),
}


class EdgeCapeLibraryError(RuntimeError):
    pass


def declared_symbols(header_path=HEADER_PATH):
    """Function names declared in the public header (used by the symbol-export test)."""
    text = open(header_path).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ec_[a-z0-9_]+)\s*\(", text)))


_lib = None


def load(build_if_missing=True):
    """dlopen the in-tree library (building it with nvcc first when absent)."""
    global _lib
    if _lib is not None:
        return _lib
    if build_if_missing:
        # cheap digest check of csrc/ + include/: rebuilds a missing or stale library when nvcc exists
        from .build import build
        try:
            build()
        except Exception as e:
            if not os.path.exists(LIB_PATH):
                raise EdgeCapeLibraryError(f"cannot build {LIB_PATH}: {e}") from e
            if "nvcc not found" not in str(e):
                raise
    if not os.path.exists(LIB_PATH):
        raise EdgeCapeLibraryError(f"{LIB_PATH} is missing; run `python -m edgecape_b200.build`")
    try:
        lib = ctypes.CDLL(LIB_PATH)
    except OSError as e:  # pragma: no cover
        raise EdgeCapeLibraryError(f"cannot load {LIB_PATH}: {e}") from e
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise EdgeCapeLibraryError(f"{LIB_PATH} does not export {name}; rebuild it") from e
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def call(name, *args):
    """Invoke an int-returning entry point and raise on a non-zero status."""
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        msg = lib.ec_last_error_string().decode(errors="replace")
        raise EdgeCapeLibraryError(f"{name} failed with status {rc}: {msg}")


def launch_count():
    return int(load().ec_launch_count())
