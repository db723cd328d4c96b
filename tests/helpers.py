"""Shared helpers for the parity tests."""
import glob
import hashlib
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def lattice_golden_names():
    return sorted(os.path.basename(p)[len("lattice_"):-4] for p in glob.glob(os.path.join(GOLDEN, "lattice_*.npz")))


def bcl_golden_names():
    return sorted(os.path.basename(p)[len("bcl_"):-4] for p in glob.glob(os.path.join(GOLDEN, "bcl_*.npz")))


def load_lattice_golden(name):
    z = np.load(os.path.join(GOLDEN, "lattice_%s.npz" % name))
    smap = [[float(s), int(r)] for s, r in z["scale_map"]]
    levels = []
    li = 0
    while "L%d_cnt" % li in z:
        levels.append({"pc1_barycentric": z["L%d_bary" % li], "pc1_el_minus_gr": z["L%d_elmgr" % li],
                       "pc1_lattice_offset": z["L%d_off" % li].astype(np.int64),
                       "pc1_blur_neighbors": z["L%d_nbr" % li].astype(np.int64),
                       "pc1_hash_cnt": int(z["L%d_cnt" % li])})
        li += 1
    return z["pc"], smap, z["keys0"].astype(np.int64), levels


def load_digests():
    out = {}
    with open(os.path.join(GOLDEN, "lattice_digests.txt")) as f:
        for line in f:
            k, v = line.split()
            out[k] = v
    return out


def digest_cases():
    return sorted({k.rsplit("/cnt", 1)[0] for k in load_digests() if k.endswith("/cnt")})


def bits(a):
    a = np.ascontiguousarray(a)
    return a.view(np.int32) if a.dtype == np.float32 else a


def assert_level_equal(got, want, ctx=""):
    """Bit-exact comparison of one level dict (numpy arrays) - ints AND float bit patterns."""
    assert int(got["pc1_hash_cnt"]) == int(want["pc1_hash_cnt"]), ctx + " hash_cnt"
    for k in ("pc1_lattice_offset", "pc1_blur_neighbors", "pc1_barycentric", "pc1_el_minus_gr"):
        g, w = np.asarray(got[k]), np.asarray(want[k])
        assert g.shape == w.shape, "%s %s shape %s vs %s" % (ctx, k, g.shape, w.shape)
        assert g.dtype == w.dtype, "%s %s dtype %s vs %s" % (ctx, k, g.dtype, w.dtype)
        if not np.array_equal(bits(g), bits(w)):
            bad = np.argwhere(bits(g) != bits(w))
            raise AssertionError("%s %s: %d mismatches, first at %s" % (ctx, k, len(bad), bad[0]))


def rel_err(got, want):
    got = np.asarray(got, np.float64)
    want = np.asarray(want, np.float64)
    return float(np.abs(got - want).max() / max(np.abs(want).max(), 1e-30))
