"""ctypes binding of efgh_b200/lib/libefgh_b200.so (the C ABI declared in include/efgh_b200.h).

There is no CPU fallback: if the library is missing or a call fails this module raises.
"""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libefgh_b200.so")
_lib = None

vp, i64, i32, f32, sz = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_float, ctypes.c_size_t

# name -> (restype, argtypes); mirrors include/efgh_b200.h one to one (checked by tests/test_capi.py)
SIGNATURES = {
    "efgh_last_error": (ctypes.c_char_p, []),
    "efgh_version": (i32, []),
    "efgh_device_sm_count": (i32, []),
    "efgh_launch_count": (i64, []),
    "efgh_copy_matrix_async": (i32, [vp, i64, vp, i64, i64, i64, i32, vp]),
    "efgh_lattice_workspace_bytes": (sz, [i64]),
    "efgh_lattice_points": (i32, [vp, i64, i64, vp, f32, vp, vp, i64, i64, vp, vp, sz, vp]),
    "efgh_lattice_vertices": (i32, [i64, vp, vp, i64, vp, i32, i64, vp, vp, i64, vp, i64, f32, vp, vp, sz, vp]),
    "efgh_lattice_table_entries": (i64, [i64, i64]),
    "efgh_lattice_batch_workspace_bytes": (sz, [i32, i64, i64]),
    "efgh_lattice_batch_info_ints": (i64, [i32]),
    "efgh_lattice_vertex_offsets_ints": (i64, [i64]),
    "efgh_lattice_points_batch": (i32, [vp, i64, i64, vp, i32, i64, f32, vp, vp, i64, i64, vp, vp, vp, vp, vp, sz, vp]),
    "efgh_lattice_vertices_batch": (i32, [i64, vp, i32, i64, vp, vp, i64, vp, i32, i64, vp, vp, i64, vp, i64, f32, vp, vp, vp, vp, sz, vp]),
    "efgh_bcl_scatter": (i32, [vp, i64, i64, i32, vp, i64, i64, i32, i64, vp, vp, i64, vp, i32, i64, i32, vp, i64, vp, vp]),
    "efgh_bcl_stem_weight_floats": (i64, [i32, i32, i32, i32]),
    "efgh_bcl_scatter_stem": (i32, [vp, i64, i64, i32, vp, i64, i32, i32, i32, i32, vp, f32, i64, vp, vp, i64, vp, i32, i64, i32, vp, i64, vp, vp]),
    "efgh_bcl_stem_rows": (i32, [vp, i64, i32, i32, i32, i32, vp, f32, i64, vp, vp, i64, vp]),
    "efgh_bcl_splat_gather": (i32, [vp, vp, i64, i32, vp, vp, i64, vp, i32, vp, i64, vp, vp]),
    "efgh_bcl_zero": (i32, [vp, i64, i32, vp, vp, i64, i32, i64, vp, i32, vp]),
    "efgh_bcl_inv_norm": (i32, [vp, vp, i64, vp, i32, vp]),
    "efgh_bcl_gather": (i32, [vp, i64, i32, vp, i64, vp, vp, i64, vp, i32, i64, i32, vp, vp, i64, i64, vp]),
    "efgh_bcl_conv": (i32, [vp, i64, i32, vp, vp, i32, i64, i32, i64, vp, vp, vp, i32, i32, vp, i64, i32, vp]),
    "efgh_bcl_conv_tc_supported": (i32, [i32, i32, i32, i32]),
    "efgh_bcl_packed_weight_bytes": (sz, [i32, i32, i32]),
    "efgh_bcl_pack_weights": (i32, [vp, i32, i32, i32, vp, vp]),
    "efgh_bcl_conv_tc_groups": (i32, [i32, i32]),
    "efgh_bcl_conv_tc": (i32, [vp, i64, i32, vp, i32, vp, i32, i64, i32, i64, vp, vp, vp, i32, i32, vp, i64, i32, i32, vp]),
    "efgh_bcl_normalize": (i32, [vp, i64, i32, vp, vp, i64, vp, i32, vp]),
    "efgh_bcl_bias_act": (i32, [vp, i64, i32, i64, vp, vp, i32, vp]),
    "efgh_bcl_conv_dgrad": (i32, [vp, i64, vp, i64, i32, i32, vp, i32, i64, i32, i64, vp, vp, i32, vp, i64, vp]),
    "efgh_project_range_image": (i32, [vp, i64, i64, i64, i32, i32, i32, ctypes.c_double, ctypes.c_double, vp, vp, vp]),
    "efgh_project_depth_image": (i32, [vp, i64, i64, i64, i32, vp, i32, i32, vp, vp, vp]),
    "efgh_preproc_workspace_bytes": (sz, [i64]),
    "efgh_preproc_cloud": (i32, [vp, i64, i32, f32, vp, i64, i64, vp, vp, vp, vp, vp, sz, vp]),
    "efgh_bcl_conv_wgrad_tc_supported": (i32, [i32, i32, i32]),
    "efgh_bcl_conv_wgrad_tc": (i32, [vp, i64, i32, vp, i32, i64, i32, i64, vp, vp, i64, i32, vp, vp, vp]),
    "efgh_bcl_act_bwd": (i32, [vp, i64, vp, i64, i32, i32, i64, vp, vp]),
    "efgh_bcl_loss_half_mean_square": (i32, [vp, i64, i32, vp, i32, vp, i64, vp, i64, vp]),
    "efgh_bcl_conv_wgrad": (i32, [vp, i64, i32, vp, vp, i32, i64, i32, i64, vp, vp, i64, vp, i64, i32, i32, vp, vp, vp]),
}


class EfghError(RuntimeError):
    pass


def build(verbose=False):
    """Compile the library in-tree with nvcc for sm_100a (works without a GPU)."""
    cmd = ["make", "-s", "-C", os.path.join(_HERE, "csrc"), "-j4", "all"]
    subprocess.check_call(cmd, stdout=None if verbose else subprocess.DEVNULL)
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise EfghError("%s is missing - run `python -c 'import __graft_entry__ as g; g.build()'` "
                            "(there is no CPU fallback for this path)" % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(status, what):
    if status != 0:
        raise EfghError("%s failed (%d): %s" % (what, status, lib().efgh_last_error().decode()))


def ptr(t):
    """Device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


_raw_stream = None


def stream_ptr():
    """cudaStream_t of torch's current stream on the current device.  torch.cuda.current_stream() goes through device-index
    and availability checks (an os.getenv each time) - measured a quarter of the drop-in module path's host time at ~28
    calls per scan - so the raw accessor is used when this torch has it."""
    global _raw_stream
    import torch
    if _raw_stream is None:
        _raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", False)
    if _raw_stream:
        return _raw_stream(torch.cuda.current_device())
    return torch.cuda.current_stream().cuda_stream
