"""ctypes binding of libvibo_b200.so (C ABI declared in include/vibo_b200.h).

The shared library is built in-tree by ``__graft_entry__.build()`` (nvcc,
sm_100a).  There is no fallback: if the library is missing or a symbol is
absent, importing/using the kernels raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvibo_b200.so")

MISSING_PRIOR, MISSING_DROP = 0, 1
ELBO_KL, ELBO_SAMPLE = 0, 1
MAX_ABILITY_DIM = 8


class Desc(C.Structure):
    """``vibo_desc`` of include/vibo_b200.h."""
    _fields_ = [
        ("num_person", C.c_int64),
        ("num_item", C.c_int32),
        ("ability_dim", C.c_int32),
        ("irt_model", C.c_int32),
        ("conditional", C.c_int32),
        ("missing_policy", C.c_int32),
        ("elbo_form", C.c_int32),
        ("person_offset", C.c_int64),
    ]


_p = C.c_void_p
_PD = C.POINTER(Desc)

# name -> (restype, argtypes); must list every symbol include/vibo_b200.h declares
SIGNATURES = {
    "vibo_version": (C.c_int, []),
    "vibo_last_error": (C.c_char_p, []),
    "vibo_workspace_bytes": (C.c_size_t, [_PD]),
    "vibo_fused_elbo": (C.c_int, [_PD, _p, _p, _p, _p, _p, C.c_uint64, C.c_float, _p, _p, _p, _p, _p,
                                  _p, _p, C.c_size_t, _p]),
    "vibo_fused_elbo_graph": (C.c_int, [_PD, _p, _p, _p, _p, _p, C.c_float, _p, _p, _p, _p, _p, _p, _p,
                                        C.c_size_t, _p]),
    "vibo_philox_normal": (C.c_int, [_PD, C.c_uint64, _p, _p, _p]),
    "vibo_host_staging_bytes": (C.c_size_t, [_PD, C.c_int64]),
    "vibo_fused_elbo_host": (C.c_int, [_PD, _p, _p, _p, _p, _p, C.c_uint64, C.c_float, _p, _p, _p, _p,
                                       C.c_int64, _p, C.c_size_t, _p, C.c_size_t, _p]),
    "vibo_fused_elbo_host_packed": (C.c_int, [_PD, _p, _p, _p, _p, C.c_uint64, C.c_float, _p, _p, _p, _p,
                                              C.c_int64, _p, C.c_size_t, _p, C.c_size_t, _p]),
    "vibo_pack": (C.c_int, [_PD, _p, _p, _p, _p]),
    "vibo_unpack": (C.c_int, [_PD, _p, _p, _p, _p]),
    "vibo_encode": (C.c_int, [_PD, _p, _p, _p, _p, _p, _p, _p]),
    "vibo_planar_params_forward": (C.c_int, [C.c_int, C.c_int, _p, _p, _p, _p, _p, _p, _p]),
    "vibo_planar_params_backward": (C.c_int, [C.c_int, C.c_int, _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    "vibo_person_counts": (C.c_int, [_PD, _p, _p, _p, _p]),
    "vibo_encode_counts": (C.c_int, [_PD, _p, _p, _p, _p, _p, _p, _p, _p]),
    "vibo_encode_backward_counts": (C.c_int, [_PD, _p, _p, _p, _p, _p, _p, _p, _p, C.c_size_t, _p]),
    "vibo_encode_backward": (C.c_int, [_PD, _p, _p, _p, _p, _p, _p, _p, _p, _p, C.c_size_t, _p]),
    "vibo_link_loglik": (C.c_int, [_PD, _p, _p, _p, _p, _p, _p, _p, _p, C.c_size_t, _p]),
    "vibo_decode": (C.c_int, [_PD, _p, _p, _p, _p]),
    "vibo_bernoulli_loglik": (C.c_int, [_PD, _p, _p, _p, _p, _p, _p, C.c_size_t, _p]),
    "vibo_param_forward": (C.c_int, [_PD, C.c_int] + [_p] * 14),
    "vibo_param_backward": (C.c_int, [_PD, C.c_int] + [_p] * 18),
    "vibo_percell_mlp": (C.c_int, [C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, _p, _p, _p, _p, _p, _p, _p,
                                   C.c_float, _p, _p]),
    "vibo_log_marginal_workspace_bytes": (C.c_size_t, [C.c_int]),
    "vibo_log_marginal": (C.c_int, [_PD, _p, _p, _p, _p, _p, C.c_int, _p, _p, C.c_uint64, _p, _p, _p, _p,
                                    C.c_size_t, _p]),
    "vibo_predictive_mean": (C.c_int, [_PD, _p, _p, _p, _p, C.c_int, C.c_uint64, _p, _p, _p]),
    "vibo_param_forward_draw": (C.c_int, [_PD, C.c_int] + [_p] * 14 + [_p]),
    "vibo_step_tail": (C.c_int, [_PD, C.c_int, C.c_float, C.c_float] + [_p] * 21 + [_p]),
    "vibo_adam_step": (C.c_int, [C.c_int64, _p, _p, _p, _p, _p, C.c_float, C.c_float, C.c_float, C.c_float, _p]),
    "vibo_flow_person_forward": (C.c_int, [_PD, C.c_int] + [_p] * 9 + [_p, C.c_size_t, _p]),
    "vibo_flow_person_backward": (C.c_int, [_PD, C.c_int] + [_p] * 13 + [_p, C.c_size_t, _p]),
    "vibo_comm_create": (C.c_int, [C.c_int, C.c_int, C.c_size_t, C.POINTER(C.c_void_p), _p]),
    "vibo_comm_connect": (C.c_int, [_p, _p]),
    "vibo_pack_host": (C.c_int, [_p, _p, _p, _p]),
    "vibo_host_threads": (C.c_int, []),
    "vibo_host_pack_share": (C.c_double, [_p, C.c_int64]),
    "vibo_comm_allreduce": (C.c_int, [_p, _p, C.c_size_t, _p]),
    "vibo_comm_allreduce_adam": (C.c_int, [_p, _p, C.c_size_t, C.c_size_t, _p, _p, _p, _p, C.c_float, C.c_float,
                                           C.c_float, C.c_float, _p]),
    "vibo_comm_status": (C.c_int, [_p]),
    "vibo_comm_destroy": (C.c_int, [_p]),
    "vibo_comm_last_error": (C.c_char_p, []),
    "vibo_single_pass": (C.c_int, [_PD]),
    "vibo_max_items": (C.c_int, [_PD]),
    "vibo_launch_count": (C.c_uint64, []),
    "vibo_profile_begin": (C.c_int, []),
    "vibo_profile_end": (C.c_int, [C.POINTER(C.c_int), C.POINTER(C.c_double)]),
}

_lib = None


def load():
    """Load the library once; raise loudly if it (or any symbol) is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: the CUDA extension is not built. Run "
            "`python -c 'import __graft_entry__ as g; g.build()'` at the repo root. "
            "There is no CPU or PyTorch fallback for the VIBO kernels.")
    # one process per GPU: share the host cores between the ranks of this node (the library's host
    # thread pool, used by the host-buffer entry points, reads VIBO_HOST_THREADS when it starts)
    if "VIBO_HOST_THREADS" not in os.environ:
        try:
            cores = len(os.sched_getaffinity(0))
        except AttributeError:
            cores = os.cpu_count() or 1
        local = max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1") or 1))
        os.environ["VIBO_HOST_THREADS"] = str(max(1, cores // local))
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class ViboError(RuntimeError):
    pass


def check(rc: int, what: str):
    if rc != 0:
        msg = load().vibo_last_error()
        raise ViboError(f"{what} failed with status {rc}: {msg.decode() if msg else ''}")
