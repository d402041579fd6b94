"""ctypes binding of libidgrec_sm100.so (the C ABI declared in include/idgrec.h).

There is no CPU fallback: every entry point raises if the library is missing or a
call fails.  ``build()`` compiles csrc/*.cu for sm_100a with nvcc (in-tree, so the
.so travels with the repo snapshot to the GPU box).
"""
from __future__ import annotations

import ctypes as C
import glob
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(os.path.dirname(_HERE), "csrc")
INCLUDE = os.path.join(os.path.dirname(os.path.dirname(_HERE)), "include")
LIB_PATH = os.path.join(_HERE, "libidgrec_sm100.so")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared", "-diag-suppress", "550"]


class IdgError(RuntimeError):
    pass


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(INCLUDE, "*.h"))
    return any(os.path.getmtime(p) > t for p in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """nvcc -gencode arch=compute_100a,code=sm_100a ... -shared -> idgrec/libidgrec_sm100.so"""
    if not force and not needs_build():
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    if not os.path.exists(nvcc):
        nvcc = "nvcc"
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + sources() + ["-o", LIB_PATH + ".tmp", "-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise IdgError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    os.replace(LIB_PATH + ".tmp", LIB_PATH)
    if verbose:
        print(r.stderr)
    return LIB_PATH


_p = C.c_void_p
_i32, _i64, _f32 = C.c_int32, C.c_int64, C.c_float

# name -> (restype, argtypes); must list every symbol declared in include/idgrec.h
SIGNATURES = {
    "idg_version": (C.c_int, []),
    "idg_last_error": (C.c_char_p, []),
    "idg_launch_count": (_i64, []),
    "idg_accumulate_f64": (C.c_int, [_p, _p, _i32, _p]),
    "idg_csr_structure": (C.c_int, [_p, _p, _i64, _i32, _i32, C.c_int, _p, _p, _p, _p, C.POINTER(_i64), _p]),
    "idg_csr_normalise": (C.c_int, [_p, _p, _p, _i32, _i64, _p, _p, _p, _p]),
    "idg_graph_create": (C.c_int, [_p, _p, _p, _i32, _i32, _i64, _i32, C.POINTER(_p), _p]),
    "idg_graph_destroy": (None, [_p]),
    "idg_graph_nnz": (_i64, [_p]),
    "idg_graph_rows": (_i32, [_p]),
    "idg_graph_classes": (_i32, [_p]),
    "idg_spmm_layer": (C.c_int, [_p, _p, _p, _p, _p, _f32, _p, _p, _f32, _i32, _p]),
    "idg_propagate_fwd": (C.c_int, [_p, _p, _i32, _i32, C.c_int, _p, _f32, _i32, _p, _p, _p, _p]),
    "idg_propagate_bwd": (C.c_int, [_p, _p, _p, _i32, _i32, C.c_int, _i32, _p, _p, _p]),
    "idg_batch_rows": (C.c_int, [_p, _p, _p, _i32, _i32, _p, _p, _p, _p]),
    "idg_batch_rows_unique": (C.c_int, [_p, _p, _p, _i32, _i32, _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    "idg_batch_rows_clear": (C.c_int, [_p, _p, _i32, _p, _p]),
    "idg_closure_bitmap": (C.c_int, [_p, _p, _p, _p]),
    "idg_graph_set_closure": (C.c_int, [_p, _p]),
    "idg_closure_from_rows": (C.c_int, [_p, _p, _p, _i32, _p, _p, _p]),
    "idg_spmm_layer_masked": (C.c_int, [_p, _p, _p, _p, _f32, _p, _p, _f32, _i32, _p, _p]),
    "idg_graph_worklist_ints": (_i64, [_p, _i32]),
    "idg_spmm_layer_rows": (C.c_int, [_p, _p, _p, _p, _f32, _p, _p, _p, _p, _f32, _i32, _p, _p, _i32, _p, _p]),
    "idg_spmm_layer_sparse_in": (C.c_int, [_p, _p, _p, _p, _p, _p, _f32, _i32, _p, C.c_int, _p]),
    "idg_propagate_fwd_ex": (C.c_int, [_p, _p, _i32, _i32, C.c_int, _p, _f32, _i32, _p, _p, _p, _p, _p, _i32, _p, _p]),
    "idg_propagate_bwd_ex": (C.c_int, [_p, _p, _p, _i32, _i32, C.c_int, _i32, _p, _p, _p, _p]),
    "idg_bpr_workspace_bytes": (_i64, [_i32]),
    "idg_bpr_forward": (C.c_int, [_p, _p, _p, _p, _p, _i32, _i32, _i32, _i32, _f32, C.c_int, _p, _p, _p]),
    "idg_bpr_forward_tail": (C.c_int, [_p, _p, _p, _p, _p, _i32, _i32, _i32, _i32, _f32, C.c_int, _p, _p, _p, _p]),
    "idg_bpr_finish_clear": (C.c_int, [_p, _p, _p, _i32, _i32, _f32, _p, _p, _p, _p, _p]),
    "idg_bpr_backward": (C.c_int, [_p, _i32, _i32, C.c_int, _p, _p, _f32, _p, _p, _p]),
    "idg_bpr_finish": (C.c_int, [_p, _p, _p, _i32, _i32, _f32, _p, _p, _p, _p]),
    "idg_axpby": (C.c_int, [_p, _f32, _p, _f32, _p, _i64, _p]),
    "idg_zero_rows": (C.c_int, [_p, _p, _i32, _i32, _p]),
    "idg_zero_rows_bitmap": (C.c_int, [_p, _p, _i32, _i32, _p]),
    "idg_infonce_workspace_bytes": (_i64, [_i32, _i32]),
    "idg_infonce_fwd_bwd": (C.c_int, [_p, _p, _p, _i32, _i32, _f32, _f32, _p, _p, _p, _p, _p]),
    "idg_unique_rows": (C.c_int, [_p, _i32, _i64, _p, _p, _p]),
    "idg_infonce_fwd_bwd_dev": (C.c_int, [_p, _p, _p, _p, _i32, _i32, _f32, _f32, _p, _p, _p, _p, _p]),
    "idg_ngcf_dense_fwd": (C.c_int, [_p, _p, _p, _p, _p, _p, _p, _f32, _i32, _p, _p, _p, _i32, _p]),
    "idg_ngcf_workspace_bytes": (_i64, []),
    "idg_ngcf_keep_masks": (C.c_int, [_p, _i64, _i32, _p, C.c_uint64, _p, _p]),
    "idg_copy_rows_strided": (C.c_int, [_p, _i32, _p, _i32, _i32, _i32, _p, _i32, _p]),
    "idg_ngcf_dense_bwd": (C.c_int, [_p, _p, _p, _p, _p, _f32, _p, _p, _p, _i32, _p, _i32, _p, _p, _p, _p, _p, _p, _p]),
    "idg_ngcf_keep_bits": (C.c_int, [_p, _i32, _i32, _p, C.c_uint64, _p, _p]),
    "idg_ngcf_dense_fwd_bits": (C.c_int, [_p, _p, _p, _p, _p, _p, _p, _f32, _i32, _p, _p, _p, _i32, _p]),
    "idg_ngcf_dense_bwd_bits": (C.c_int, [_p, _p, _p, _p, _p, _f32, _p, _p, _i32, _p, _i32, _p, _p, _p, _p, _p, _p, _p]),
    "idg_eval_workspace_bytes": (_i64, [_i32, _i32, _i32, _i32]),
    "idg_eval_topk": (C.c_int, [_p, _p, _i32, _i32, _i32, _p, _p, _p, _i32, _i32, _p, _p, _p, _p]),
    "idg_rating_matrix": (C.c_int, [_p, _p, _p, _i32, _i32, _i32, _p, _p]),
    "idg_eval_tc_bounds": (C.c_int, [_p, _p, _i32, _i32, _i32, _p, _p, _p, _i32, _p, _p, _p]),
    "idg_eval_metrics": (C.c_int, [_p, _p, _i32, _i32, _p, _p, C.POINTER(_i32), _i32, _p, _p, _p]),
    "idg_adam_step": (C.c_int, [_p, _p, _p, _p, _i64, _f32, _f32, _f32, _f32, _i32, _p]),
    "idg_adam_prepare": (C.c_int, [_p, _p, _f32, _f32, _f32, _p]),
    "idg_propagate_bwd_adam": (C.c_int, [_p, _p, _p, _i32, _i32, C.c_int, _i32, _p, _p, _p, _p]),
    "idg_spmm_layer_adam": (C.c_int, [_p, _p, _p, _f32, _i32, _p, _p]),
    "idg_spmm_layer_sparse_in_masked": (C.c_int, [_p, _p, _p, _p, _i32, _p, _p, C.c_int, _p]),
    "idg_spmm_layer_views": (C.c_int, [_p, _p, _p, _p, _p, _p, _p, _f32, _i32, _p]),
    "idg_spmm_layer_add2": (C.c_int, [_p, _p, _p, _p, _p, _f32, _i32, _p, C.c_int, _p]),
    "idg_adam_step_dev": (C.c_int, [_p, _p, _p, _p, _i64, _f32, _f32, _f32, _f32, _p, _p]),
    "idg_device_alloc": (C.c_int, [_i64, C.POINTER(_p)]),
    "idg_device_free": (C.c_int, [_p]),
    "idg_ipc_get_handle": (C.c_int, [_p, _p]),
    "idg_ipc_open": (C.c_int, [_p, C.POINTER(_p)]),
    "idg_ipc_close": (C.c_int, [_p]),
    "idg_peers_create": (C.c_int, [_p, _i64, _i32, _i32, C.POINTER(_p), C.POINTER(_p)]),
    "idg_peers_destroy": (None, [_p]),
    "idg_peers_set_multicast": (C.c_int, [_p, _p]),
    "idg_graph_set_peers": (C.c_int, [_p, _p]),
    "idg_peers_push": (C.c_int, [_p, _p, _i64, _p]),
    "idg_peers_push_ctas": (C.c_int, [_p, _p, _i64, _i32, _p]),
    "idg_peers_barrier": (C.c_int, [_p, _p, _p]),
    "idg_peers_set_timeout_ms": (C.c_int, [_p, _i64]),
    "idg_peers_status": (C.c_int, [_p, _p, _p]),
    "idg_neg_sample_replay": (C.c_int, [_p, _i64, _p, _p, _p, _i64, _p, C.POINTER(_i64)]),
    "idg_neg_sample_walk": (C.c_int, [_p, _i64, _i64, _p, _p, _p, _i64, _p, C.POINTER(_i64), C.POINTER(_i64)]),
    "idg_permute3": (C.c_int, [_p, _p, _p, _p, _i64, _p, _p]),
    "idg_batch_fetch": (C.c_int, [_p, _p, _i32, _i32, _p, _i32, _p]),
    "idg_tanh_fwd": (C.c_int, [_p, _p, _i64, _p]),
    "idg_tanh_bwd": (C.c_int, [_p, _p, _p, _i64, _p]),
    "idg_parse_ratings": (C.c_int, [C.c_char_p, _p, _p, _i64, C.POINTER(_i64), _p, _p, _i64, C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_i64)]),
    "idg_pair_loss_workspace_bytes": (_i64, [_i32, _i32]),
    "idg_pair_loss": (C.c_int, [_i32, _p, _p, _i32, _i32, _f32, _f32, _p, _p, _p, _p, _p]),
    "idg_pair_loss_ex": (C.c_int, [_i32, _p, _p, _i32, _i32, _f32, _f32, _f32, _p, _p, _p, _p, _p, _p, _p]),
    "idg_gather_rows": (C.c_int, [_p, _p, _i32, _i32, _p, _p]),
    "idg_scatter_add_rows": (C.c_int, [_p, _p, _i32, _i32, _p, _p]),
}



class AdamArgs(C.Structure):
    """idg_adam_args of include/idgrec.h."""
    _fields_ = [("p", _p), ("m", _p), ("v", _p), ("regc", _p), ("d_scalars", _p), ("beta1", _f32), ("beta2", _f32), ("eps", _f32)]


class StepTail(C.Structure):
    """idg_step_tail of include/idgrec.h."""
    _fields_ = [("d_loss_acc", _p), ("d_step", _p), ("d_scalars", _p), ("lr", _f32), ("beta1", _f32), ("beta2", _f32)]


_lib = None


def lib() -> C.CDLL:
    """Load (once) and type the library.  Fails loudly when it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            try:  # a fresh checkout: compile the extension (nvcc, sm_100a) -- this is a build, not a fallback
                build()
            except Exception:  # noqa: BLE001
                pass
        if not os.path.exists(LIB_PATH):
            raise IdgError(
                "libidgrec_sm100.so is not built (%s). Run `python -c 'import __graft_entry__ as g; g.build()'` "
                "at the repo root; there is no CPU fallback for the ID-GRec hot path." % LIB_PATH)
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)  # AttributeError if the symbol is missing
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(rc: int, what: str = ""):
    if rc != 0:
        raise IdgError("%s failed (rc=%d): %s" % (what or "idgrec call", rc, lib().idg_last_error().decode()))


def ptr(t):
    """Device/host pointer of a contiguous torch tensor (or None)."""
    if t is None:
        return None
    assert t.is_contiguous(), "idgrec: tensor must be contiguous"
    return t.data_ptr()


def cur_stream():
    import torch
    return torch.cuda.current_stream().cuda_stream
