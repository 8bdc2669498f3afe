"""ctypes binding of libmagicmirror.so (include/magicmirror.h).

The product path has NO fallback: if the shared library is missing or fails to
load, importing any compute entry point raises.  Nothing under oracle/ is ever
imported from here.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmagicmirror.so")

_lib = None

c_float_p = ctypes.c_void_p   # device/host pointers are passed as raw addresses
c_int = ctypes.c_int
c_float = ctypes.c_float
c_void_p = ctypes.c_void_p

# name -> (restype, argtypes); mirrors include/magicmirror.h one to one
c_size_t = ctypes.c_size_t
_P = c_void_p

# name -> (restype, argtypes); mirrors include/magicmirror.h one to one
_RENDER_IN = [_P] * 6 + [c_int, c_int, c_int] + [_P] * 2 + [c_int]      # vertices..tex, Ht, Wt, tex_mirror, lights, bg, no_mask
SIGNATURES = {
    "mm_abi_version": (c_int, []),
    "mm_last_error": (ctypes.c_char_p, []),
    "mm_ctx_create": (c_int, [ctypes.POINTER(c_void_p), c_int, c_int, c_int, _P, _P, c_int, c_int,
                              c_float, c_float, c_float, c_float, c_int, c_float, c_float]),
    "mm_ctx_destroy": (c_int, [_P]),
    "mm_workspace_bytes": (c_size_t, [_P, c_int]),
    "mm_ctx_get_int": (c_int, [_P, ctypes.c_char_p]),
    "mm_render_forward": (c_int, [_P, c_int] + _RENDER_IN + [_P] * 4 + [_P, c_size_t, _P]),
    "mm_render_backward": (c_int, [_P, c_int] + _RENDER_IN + [_P] * 3 + [_P, c_float, c_float, c_float, _P] + [_P] * 8 +
                           [_P, c_size_t, _P]),
    "mm_recon_data_forward": (c_int, [_P, c_int, _P, _P, c_float, c_float, _P, _P, _P, c_size_t, _P]),
    "mm_recon_data_backward": (c_int, [_P, c_int, _P, _P, c_float, c_float, c_float, _P, _P, _P, c_size_t, _P]),
    "mm_render_compare_fwd_bwd": (c_int, [_P, c_int] + _RENDER_IN + [_P, c_float, c_float, c_float] + [_P] * 2 + [_P] * 3 +
                                  [_P] * 8 + [_P, c_size_t, _P]),
    "mm_debug_export_faces": (c_int, [_P, c_int, _P, c_size_t, _P, _P, _P, _P]),
    "mm_debug_workspace_offset": (c_size_t, [_P, c_int, ctypes.c_char_p]),
    "mm_face_normals_forward": (c_int, [_P, c_int] + [_P] * 6 + [_P, c_size_t, _P]),
    "mm_face_normals_backward": (c_int, [_P, c_int] + [_P] * 11 + [_P, c_size_t, _P]),
    "mm_ctx_set_regularizer_topology": (c_int, [_P, c_int, _P, _P, _P, _P, c_int, _P, _P, _P, c_float]),
    "mm_mesh_reg_forward": (c_int, [_P, c_int, _P, _P, _P, c_float, c_float, c_int, ctypes.c_uint, _P, _P, _P]),
    "mm_mesh_reg_backward": (c_int, [_P, c_int, _P, _P, _P, c_float, c_float, c_int, ctypes.c_uint, _P, _P, _P, _P, _P]),
    "mm_template_features_forward": (c_int, [_P, c_int, c_int, c_int, _P, _P, _P, _P, _P]),
    "mm_template_features_backward": (c_int, [_P, c_int, c_int, c_int, _P, _P, _P, _P, _P]),
    "mm_texture_flow_forward": (c_int, [_P] + [c_int] * 7 + [_P, _P, _P, _P]),
    "mm_texture_flow_backward": (c_int, [_P] + [c_int] * 7 + [_P, _P, _P, _P, _P, _P]),
    "mm_ctx_set_timing": (c_int, [_P, c_int]),
    "mm_ctx_get_timing": (c_int, [_P, _P, c_int]),
}
ABI_VERSION = 3


class MagicMirrorError(RuntimeError):
    pass


def lib():
    """Loads libmagicmirror.so (built by __graft_entry__.build()); raises loudly if absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise MagicMirrorError(
                "libmagicmirror.so not found at %s -- run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). There is no CPU or PyTorch fallback for the render path." % LIB_PATH)
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)     # AttributeError here == header/library mismatch
            fn.restype = res
            fn.argtypes = args
        if handle.mm_abi_version() != ABI_VERSION:
            raise MagicMirrorError("libmagicmirror ABI version mismatch")
        _lib = handle
    return _lib


def check(rc, what):
    if rc != 0:
        msg = lib().mm_last_error()
        raise MagicMirrorError("%s failed (rc=%d): %s" % (what, rc, msg.decode() if msg else ""))
