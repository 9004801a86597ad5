"""ctypes binding of kliff_b200/csrc/libkliff_b200.so (the C-ABI of include/kliff_b200.h).

There is no CPU fallback: if the shared library is missing or no CUDA device is present, the
first call raises.  Build the library with ``python -c "import __graft_entry__ as g; g.build()"``
or ``kliff_b200/csrc/build.sh``.
"""
import ctypes as C
import os

KB_MAX_SPECIES = 8
KB_MAX_TWO = 64
KB_MAX_GROUPS = 32
KB_GROUP_WIDTH = 8

KB_OK, KB_ERR_REFERENCE, KB_ERR_INVALID, KB_ERR_LIMIT, KB_ERR_UNSUPPORTED = 0, 1, 2, 3, 4
KB_SYMFUN_AUTO, KB_SYMFUN_GENERAL, KB_SYMFUN_TILED, KB_SYMFUN_WARP, KB_SYMFUN_SPLIT, KB_SYMFUN_QUEUE = 0, 1, 2, 3, 4, 6
KB_F64, KB_F32 = 0, 1

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", "libkliff_b200.so")
if os.environ.get("KB_LIB_PATH"):  # tuning aid: a differently compiled build of the same library
    LIB_PATH = os.environ["KB_LIB_PATH"]


class kb_symfun_params(C.Structure):
    _fields_ = [
        ("n_species", C.c_int32),
        ("n_desc", C.c_int32),
        ("n_two", C.c_int32),
        ("n_groups", C.c_int32),
        ("rcut", C.c_double * (KB_MAX_SPECIES * KB_MAX_SPECIES)),
        ("two_kind", C.c_int32 * KB_MAX_TWO),
        ("two_out", C.c_int32 * KB_MAX_TWO),
        ("two_p0", C.c_double * KB_MAX_TWO),
        ("two_p1", C.c_double * KB_MAX_TWO),
        ("grp_kind", C.c_int32 * KB_MAX_GROUPS),
        ("grp_n", C.c_int32 * KB_MAX_GROUPS),
        ("grp_eta", C.c_double * KB_MAX_GROUPS),
        ("grp_zeta", (C.c_double * KB_GROUP_WIDTH) * KB_MAX_GROUPS),
        ("grp_lambda", (C.c_double * KB_GROUP_WIDTH) * KB_MAX_GROUPS),
        ("grp_out", (C.c_int32 * KB_GROUP_WIDTH) * KB_MAX_GROUPS),
        ("canonical_zl", C.c_int32),
        ("reserved", C.c_int32),
    ]


KB_MLP_MAX_IN, KB_MLP_MAX_HIDDEN, KB_MLP_MAX_WIDTH = 128, 4, 32


class kb_mlp_desc(C.Structure):
    _fields_ = [("n_in", C.c_int32), ("n_hidden", C.c_int32), ("width", C.c_int32 * KB_MLP_MAX_HIDDEN),
                ("dtype", C.c_int32), ("reserved", C.c_int32)]


# every symbol include/kliff_b200.h declares (tests check that the library exports all of them)
EXPORTS = [
    "kb_version", "kb_error_string", "kb_device_info",
    "kb_pad_begin", "kb_pad_finish", "kb_pad_destroy",
    "kb_nl_begin", "kb_nl_finish", "kb_nl_destroy",
    "kb_batch_begin", "kb_batch_neighbors", "kb_batch_finish", "kb_batch_destroy",
    "kb_symfun_block_offsets", "kb_symfun_forward",
    "kb_force_scatter_fwd", "kb_force_scatter_bwd", "kb_dense_fold",
    "kb_probe_fp64_peak", "kb_probe_hbm_copy", "kb_pool_trim", "kb_launch_count", "kb_kernel_timing",
    "kb_moments", "kb_normalize",
    "kb_scatter_plan_begin", "kb_scatter_plan_finish", "kb_force_scatter_fwd_seg",
    "kb_mlp_num_params", "kb_mlp_energy_grad", "kb_train_residual", "kb_mlp_param_grad", "kb_adam_step",
    "kb_train_fast_ok", "kb_train_forward", "kb_train_residual_seg", "kb_train_backward", "kb_train_update",
    "kb_train_adam",
    "kb_peer_ws_bytes", "kb_peer_alloc", "kb_peer_open", "kb_peer_close", "kb_peer_free", "kb_peer_error",
    "kb_train_update_peer",
]

_lib = None


class KliffB200Error(RuntimeError):
    def __init__(self, code, where):
        self.code = code
        msg = lib().kb_error_string(code).decode() if _lib is not None else str(code)
        super().__init__(f"{where}: [{code}] {msg}")


def lib():
    """Load (once) and return the C-ABI library; raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: the CUDA extension has not been built "
            "(run `python -c 'import __graft_entry__ as g; g.build()'`). "
            "kliff_b200 has no CPU fallback."
        )
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_double
    L.kb_version.restype = C.c_char_p
    L.kb_error_string.restype = C.c_char_p
    L.kb_error_string.argtypes = [C.c_int]
    L.kb_device_info.argtypes = [C.POINTER(i32)] * 3
    L.kb_pad_begin.argtypes = [i32, dbl, C.POINTER(dbl), C.POINTER(i32), vp, C.POINTER(i64),
                               C.POINTER(vp), vp]
    L.kb_pad_finish.argtypes = [vp, vp, vp, vp, vp, vp]
    L.kb_pad_destroy.argtypes = [vp, vp]
    L.kb_pad_destroy.restype = None
    L.kb_nl_begin.argtypes = [i32, i32, vp, vp, dbl, dbl, i32, vp, vp, C.POINTER(i64),
                              C.POINTER(i32), C.POINTER(vp), vp]
    L.kb_nl_finish.argtypes = [vp, vp, vp]
    L.kb_nl_destroy.argtypes = [vp, vp]
    L.kb_nl_destroy.restype = None
    L.kb_batch_begin.argtypes = [i32, vp, vp, vp, dbl, vp, C.POINTER(i64), C.POINTER(vp), vp]
    L.kb_batch_neighbors.argtypes = [vp, vp, vp, vp, vp, vp, vp, C.POINTER(i64), C.POINTER(i32), vp]
    L.kb_batch_finish.argtypes = [vp, vp, vp, vp, vp]
    L.kb_batch_destroy.argtypes = [vp, vp]
    L.kb_batch_destroy.restype = None
    L.kb_symfun_block_offsets.argtypes = [i32, vp, vp, i32, vp, C.POINTER(i64), vp]
    L.kb_symfun_forward.argtypes = [C.POINTER(kb_symfun_params), i32, vp, vp, vp, vp, vp, i32, vp,
                                    vp, i32, vp, i32, vp]
    L.kb_force_scatter_fwd.argtypes = [i32, vp, vp, vp, vp, i32, vp, vp, vp, i32, vp, vp, vp]
    L.kb_force_scatter_bwd.argtypes = [i32, vp, vp, vp, vp, i32, vp, vp, vp, i32, vp, vp, vp]
    L.kb_dense_fold.argtypes = [i32, vp, vp, vp, vp, vp, i32, i32, vp, i32, vp, vp, vp, vp]
    L.kb_probe_fp64_peak.argtypes = [C.POINTER(dbl), vp]
    L.kb_probe_hbm_copy.argtypes = [C.c_size_t, C.POINTER(dbl), vp]
    L.kb_kernel_timing.argtypes = [i32, C.POINTER(dbl), C.POINTER(i32)]
    L.kb_pool_trim.argtypes = [C.c_size_t]
    L.kb_launch_count.argtypes = [i32]
    L.kb_launch_count.restype = i64
    L.kb_scatter_plan_begin.argtypes = [i32, vp, vp, vp, C.POINTER(i64), vp]
    L.kb_scatter_plan_finish.argtypes = [i32, vp, vp, vp, vp, i32, vp, i64, vp, vp, vp]
    L.kb_force_scatter_fwd_seg.argtypes = [i32, vp, vp, i32, vp, vp, vp, i32, vp, vp, i64, i64, i32, vp, vp, vp, vp, vp]
    L.kb_moments.argtypes = [i64, i32, vp, vp, vp, vp]
    L.kb_normalize.argtypes = [i64, i32, vp, vp, vp, vp, vp]
    md = C.POINTER(kb_mlp_desc)
    L.kb_mlp_num_params.argtypes = [md, C.POINTER(i32), C.POINTER(i32)]
    L.kb_mlp_energy_grad.argtypes = [md, i32, vp, vp, vp, vp, vp]
    L.kb_train_residual.argtypes = [i32, vp, i64, vp, vp, vp, vp, vp, vp, dbl, i32, i32, vp, vp, vp, vp]
    L.kb_mlp_param_grad.argtypes = [md, i32, vp, vp, vp, vp, vp, vp, vp, vp]
    L.kb_adam_step.argtypes = [i32, vp, vp, dbl, vp, vp, vp, dbl, dbl, dbl, dbl, i32, vp]
    L.kb_train_fast_ok.argtypes = [md, i32]
    L.kb_train_forward.argtypes = [md, i32, vp, vp, vp, vp, i32, vp, vp, vp, vp, vp, vp]
    L.kb_train_residual_seg.argtypes = [i32, vp, i64, vp, vp, vp, i64, vp, vp, vp, vp, vp, dbl, i32, i32,
                                        vp, vp, vp, vp, vp, vp]
    L.kb_train_backward.argtypes = [md, i32, vp, vp, vp, vp, i32, vp, vp, vp, vp, vp, vp, i32,
                                    C.POINTER(i32), vp]
    L.kb_train_update.argtypes = [md, i32, vp, i32, vp, vp, i32, vp, vp, vp, vp, dbl, dbl, dbl, dbl, vp, vp]
    L.kb_train_adam.argtypes = [i32, vp, vp, vp, vp, vp, dbl, dbl, dbl, dbl, i32, vp, vp]
    L.kb_peer_ws_bytes.argtypes = [i32, i32, C.POINTER(i64), C.POINTER(i32)]
    L.kb_peer_alloc.argtypes = [i64, C.POINTER(vp), vp]
    L.kb_peer_open.argtypes = [vp, C.POINTER(vp)]
    L.kb_peer_close.argtypes = [vp]
    L.kb_peer_free.argtypes = [vp]
    L.kb_peer_error.argtypes = [vp, C.POINTER(i32), vp]
    L.kb_train_update_peer.argtypes = [i32, i32, vp, i32, vp, vp, i32, vp, vp, vp, vp, dbl, dbl, dbl, dbl, i32, vp,
                                       C.POINTER(vp), i32, i32, vp]
    for name in EXPORTS:
        getattr(L, name)  # raises AttributeError when a declared symbol is missing
    _lib = L
    return L


def check(code, where):
    if code != KB_OK:
        raise KliffB200Error(code, where)
