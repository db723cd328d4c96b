"""ctypes binding of oracle/liboracle.so (+ oracle/_ref/liboracle_ref.so) - TEST INFRASTRUCTURE.

`generate(pc, scales_filter_map)` restates reference nets/generate_data.py:117-193 (GenerateData.__call__)
on top of the C level builder; `blur_offsets` restates reference nets/transforms.py:95-122 (Traverse).
Parity status: PINNED against the live reference in the build container (tests/golden/*.npz, produced
by oracle/make_golden.py from the unmodified reference files) - the reference has no tests of its own.
"""
import ctypes
import itertools
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_libs = {}

c_f32p = ctypes.POINTER(ctypes.c_float)
c_i64p = ctypes.POINTER(ctypes.c_int64)


def build(force=False):
    """Compile the C oracle (and the _ref variant when the reference is present)."""
    so = os.path.join(HERE, "liboracle.so")
    src_m = max(os.path.getmtime(os.path.join(HERE, f)) for f in ("lattice_oracle.c", "i2i_map.h"))
    if force or not os.path.exists(so) or os.path.getmtime(so) < src_m:
        subprocess.check_call(["make", "-s", "-C", HERE, "all"], stdout=subprocess.DEVNULL)
    return so


def lib(variant="port"):
    """variant 'port': own int64 map; 'ref': compiled against the reference's khash headers."""
    if variant in _libs:
        return _libs[variant]
    if variant == "port":
        path = build()
    else:
        path = os.path.join(HERE, "_ref", "liboracle_ref.so")
        if not os.path.exists(path):
            build()
        if not os.path.exists(path):
            raise FileNotFoundError(path)
    L = ctypes.CDLL(path)
    L.efgh_oracle_expected_std.restype = ctypes.c_double
    L.efgh_oracle_elevate_matrix.argtypes = [c_f32p]
    L.efgh_oracle_keys.argtypes = [c_f32p, ctypes.c_int64, ctypes.c_int64, c_i64p, c_f32p, c_f32p]
    L.efgh_oracle_level_build.restype = ctypes.c_void_p
    L.efgh_oracle_level_build.argtypes = [c_f32p, ctypes.c_int64, ctypes.c_int64, ctypes.c_double,
                                          c_i64p, ctypes.c_int64, ctypes.c_int]
    L.efgh_oracle_level_hash_cnt.restype = ctypes.c_int64
    L.efgh_oracle_level_hash_cnt.argtypes = [ctypes.c_void_p]
    L.efgh_oracle_level_key_box.argtypes = [ctypes.c_void_p, c_i64p, c_i64p]
    L.efgh_oracle_level_export.argtypes = [ctypes.c_void_p, c_f32p, c_f32p, c_f32p, c_i64p, c_i64p, c_f32p]
    L.efgh_oracle_level_free.argtypes = [ctypes.c_void_p]
    _libs[variant] = L
    return L


def _fp(a):
    return a.ctypes.data_as(c_f32p)


def _ip(a):
    return a.ctypes.data_as(c_i64p)


def has_ref():
    return os.path.exists(os.path.join(HERE, "_ref", "liboracle_ref.so"))


def elevate_matrix():
    out = np.empty((4, 3), np.float32)
    lib().efgh_oracle_elevate_matrix(_fp(out))
    return out


def expected_std():
    return lib().efgh_oracle_expected_std()


def blur_offsets(radius, d=3):
    """(F, d+1) int64 neighbour offsets in the reference's traversal order.

    reference nets/transforms.py:95-122: a depth-first walk over step counts (i_0..i_d), i_k in
    [0, radius], last index fastest, keeping the tuples that contain at least one zero; one step
    in dimension k adds d+1 to coordinate k and subtracts 1 from every coordinate
    (advance_in_dimension, transforms.py:81-87).  F = (radius+1)^(d+1) - radius^(d+1).
    """
    d1 = d + 1
    rows = []
    for steps in itertools.product(range(radius + 1), repeat=d1):
        if 0 not in steps:
            continue
        row = np.full((d1,), -sum(steps), dtype=np.int64)
        row += d1 * np.asarray(steps, dtype=np.int64)
        rows.append(row)
    return np.stack(rows)


def keys_and_barycentric(pc, variant="port"):
    """reference nets/generate_data.py:56-112.  pc (3,N) f32 -> keys (4,N,4) i64, bary (4,N), elmgr (4,N)."""
    pc = np.ascontiguousarray(pc[:3], dtype=np.float32)
    n = pc.shape[1]
    keys = np.empty((4, n, 4), np.int64)
    bary = np.empty((4, n), np.float32)
    elmgr = np.empty((4, n), np.float32)
    lib(variant).efgh_oracle_keys(_fp(pc), n, n, _ip(keys), _fp(bary), _fp(elmgr))
    return keys, bary, elmgr


def generate(pc, scales_filter_map, variant="port"):
    """GenerateData.__call__ (reference nets/generate_data.py:117-193) as numpy.

    Returns a list of per-level dicts with the reference's keys (leading batch dim 1) plus
    'points' (the (3,N) cloud the level saw after its `*= scale`) for debugging.
    """
    L = lib(variant)
    pts = np.ascontiguousarray(pc[:3], dtype=np.float32)
    out = []
    nlev = len(scales_filter_map)
    for li, (scale, radius) in enumerate(scales_filter_map):
        n = pts.shape[1]
        has_next = li != nlev - 1
        if radius != -1:
            offs = np.ascontiguousarray(blur_offsets(radius))
            F = offs.shape[0]
        else:
            offs = np.zeros((1, 4), np.int64)
            F = -1
        h = L.efgh_oracle_level_build(_fp(pts), n, n, float(scale), _ip(offs), F, int(has_next))
        try:
            H = L.efgh_oracle_level_hash_cnt(h)
            scaled = np.empty((3, n), np.float32)
            bary = np.empty((4, n), np.float32)
            elmgr = np.empty((4, n), np.float32)
            loff = np.empty((4, n), np.int64)
            nbr = np.empty((F, H), np.int64) if F > 0 else np.zeros((1,), np.int64)
            nxt = np.empty((3, H), np.float32)
            L.efgh_oracle_level_export(h, _fp(scaled), _fp(bary), _fp(elmgr), _ip(loff),
                                       _ip(nbr) if F > 0 else None, _fp(nxt) if has_next else None)
        finally:
            L.efgh_oracle_level_free(h)
        out.append({"pc1_barycentric": bary[None], "pc1_el_minus_gr": elmgr[None],
                    "pc1_lattice_offset": loff[None], "pc1_blur_neighbors": nbr[None],
                    "pc1_hash_cnt": int(H), "points": scaled})
        if has_next:
            pts = nxt
    return out
