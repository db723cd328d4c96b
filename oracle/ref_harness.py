"""Live-reference harness (TEST INFRASTRUCTURE, this container only).

Imports the UNMODIFIED reference hot-path files from /root/reference by file path:
nets/transforms.py, nets/generate_data.py, nets/bilateralNN.py, with lib/khash*.h compiled by the
reference's own cffi build script into oracle/_ref/.  Nothing is copied into the repo; /root/reference
does not exist on the GPU box, so this module is only used (a) by oracle/make_golden.py to produce
tests/golden/*.npz and (b) by `-m "not gpu"` tests that are skipped when the reference is absent.

Shims: numba.cffi_support moved to numba.core.typing.cffi_utils; the `nets` package
__init__ drags in matplotlib/open3d, so a stub package is registered instead.
"""
import importlib.util
import os
import shutil
import subprocess
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_OUT = os.path.join(HERE, "_ref")
# Where the unmodified reference files are: the build container's /root/reference, else the git-ignored install
# baseline/_ref/ that baseline/install_ref.py copied from it (that copy travels to the GPU box; bench.py's
# `--impl reference` arm is its only user there).
INSTALLED = os.path.join(os.path.dirname(HERE), "baseline", "_ref")


def _find_root():
    for cand in (os.environ.get("EFGH_REFERENCE"), "/root/reference", INSTALLED):
        if cand and os.path.isfile(os.path.join(cand, "nets", "generate_data.py")):
            return cand
    return os.environ.get("EFGH_REFERENCE", "/root/reference")


REF_ROOT = _find_root()


def available():
    return os.path.isfile(os.path.join(REF_ROOT, "nets", "generate_data.py"))


def is_live():
    """True when the files come from the read-only reference checkout itself (build container)."""
    return available() and os.path.realpath(REF_ROOT) != os.path.realpath(INSTALLED)


def build_khash_ffi():
    """Run the reference's lib/build_khash_cffi.py in oracle/_ref/ (needs a writable cwd)."""
    os.makedirs(REF_OUT, exist_ok=True)
    built = [f for f in os.listdir(REF_OUT) if f.startswith("_khash_ffi") and f.endswith(".so")]
    if built:
        return
    inst = os.path.join(INSTALLED, "lib")
    if os.path.isdir(inst):                                  # built by baseline/install_ref.py
        for f in os.listdir(inst):
            if f.startswith("_khash_ffi") and f.endswith(".so"):
                shutil.copy(os.path.join(inst, f), REF_OUT)
                return
    tmp = os.path.join(REF_OUT, "_cffi_build")
    os.makedirs(tmp, exist_ok=True)
    for f in ("khash.h", "khash_int2int.h", "build_khash_cffi.py"):
        shutil.copy(os.path.join(REF_ROOT, "lib", f), tmp)
    subprocess.check_call([sys.executable, "build_khash_cffi.py"], cwd=tmp,
                          stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    for f in os.listdir(tmp):
        if f.startswith("_khash_ffi") and f.endswith(".so"):
            shutil.copy(os.path.join(tmp, f), REF_OUT)
    shutil.rmtree(tmp)  # reference sources never stay in the repo tree


_mods = {}


def load():
    """Returns (transforms, generate_data, bilateralNN) modules of the live reference."""
    if _mods:
        return _mods["t"], _mods["g"], _mods["b"]
    if not available():
        raise RuntimeError("reference not present at %s" % REF_ROOT)
    build_khash_ffi()
    if REF_OUT not in sys.path:
        sys.path.insert(0, REF_OUT)
    import numba
    import numba.core.typing.cffi_utils as cffi_utils
    if not hasattr(numba, "cffi_support"):
        numba.cffi_support = cffi_utils
        sys.modules["numba.cffi_support"] = cffi_utils
    pkg = types.ModuleType("refnets")
    pkg.__path__ = [os.path.join(REF_ROOT, "nets")]
    sys.modules["refnets"] = pkg

    def _imp(name):
        spec = importlib.util.spec_from_file_location(
            "refnets." + name, os.path.join(REF_ROOT, "nets", name + ".py"))
        m = importlib.util.module_from_spec(spec)
        sys.modules["refnets." + name] = m
        spec.loader.exec_module(m)
        return m

    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        _mods["t"] = _imp("transforms")
        _mods["g"] = _imp("generate_data")
        _mods["b"] = _imp("bilateralNN")
    return _mods["t"], _mods["g"], _mods["b"]


def load_net_utils():
    """The live reference's nets/net_utils.py (conv_1x1: the E-Net stem's building block, nets/enet.py:24-28)."""
    if not available():
        raise RuntimeError("reference not present at %s" % REF_ROOT)
    if "n" not in _mods:
        spec = importlib.util.spec_from_file_location("refnets_net_utils", os.path.join(REF_ROOT, "nets", "net_utils.py"))
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
        _mods["n"] = m
    return _mods["n"]


def load_common():
    """The live reference's common/torch_utils.py (range / depth image projections, :11-103) and
    data_loader/loader_utils.py (pcd_read :59-61, preproc_pcd :163-202), imported by path.  common/numpy_utils.py
    needs matplotlib and open3d (absent here; only its drawing helpers use them): stub modules stand in.  The two
    projection functions build their index tensors with torch.cuda.LongTensor / FloatTensor regardless of `device`
    (torch_utils.py:50-51); on this GPU-less container those two constructors are pointed at their CPU twins - the
    arithmetic that runs is the reference's own, on torch's CPU kernels.  Build container only."""
    if "c" in _mods:
        return _mods["c"]
    if not (available() and os.path.isfile(os.path.join(REF_ROOT, "common", "torch_utils.py"))):
        raise RuntimeError("reference common/ not present at %s" % REF_ROOT)
    from unittest import mock
    import torch
    for name in ("matplotlib", "matplotlib.pyplot", "open3d"):
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:
                sys.modules[name] = mock.MagicMock(name=name)
    pkg = types.ModuleType("common")
    pkg.__path__ = [os.path.join(REF_ROOT, "common")]
    saved = sys.modules.get("common")
    sys.modules["common"] = pkg
    try:
        out = {}
        for key, rel in (("numpy_utils", "common/numpy_utils.py"), ("torch_utils", "common/torch_utils.py"),
                         ("loader_utils", "data_loader/loader_utils.py")):
            modname = "common." + key if rel.startswith("common/") else "ref_" + key
            spec = importlib.util.spec_from_file_location(modname, os.path.join(REF_ROOT, rel))
            m = importlib.util.module_from_spec(spec)
            sys.modules[modname] = m
            spec.loader.exec_module(m)
            out[key] = m
    finally:
        if saved is not None:
            sys.modules["common"] = saved
    if not torch.cuda.is_available():
        torch.cuda.LongTensor = lambda t: t.to(torch.int64)
        torch.cuda.FloatTensor = lambda t: t.to(torch.float32)
    _mods["c"] = out
    return out
