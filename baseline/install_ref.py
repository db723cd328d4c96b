"""Installs the UNMODIFIED reference hot-path files into baseline/_ref/ (git-ignored, NOT gpurun-ignored, so the
directory travels to the GPU box with the snapshot) - the "offline install" of the reference arm.

The reference is not a pip package (no setup.py / pyproject; `pip install /root/reference` has nothing to build), so the
install is a verbatim copy of the files its lattice / BCL path consists of:

    nets/transforms.py  nets/generate_data.py  nets/bilateralNN.py  nets/net_utils.py
    lib/khash.h  lib/khash_int2int.h  lib/build_khash_cffi.py        (+ the cffi module they build)

`bench.py --impl reference` imports them from there (oracle/ref_harness.py, which only adds the numba.cffi_support
shim the reference's pinned numba 0.47 needs on numba 0.65 and a stub `nets` package - nets/__init__.py drags in
matplotlib / open3d through the rest of EFGHNet).  Nothing under baseline/_ref/ is ever committed.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
FILES = ["nets/transforms.py", "nets/generate_data.py", "nets/bilateralNN.py", "nets/net_utils.py",
         "lib/khash.h", "lib/khash_int2int.h", "lib/build_khash_cffi.py"]


def install(ref_root="/root/reference"):
    """Returns True when baseline/_ref holds the reference files (copied now or earlier), False when the
    reference is not available here and was never installed."""
    if not os.path.isfile(os.path.join(ref_root, "nets", "generate_data.py")):
        return os.path.isfile(os.path.join(DEST, "nets", "generate_data.py"))
    for f in FILES:
        dst = os.path.join(DEST, f)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(ref_root, f), dst)
    lib = os.path.join(DEST, "lib")
    if not any(f.startswith("_khash_ffi") and f.endswith(".so") for f in os.listdir(lib)):
        subprocess.check_call([sys.executable, "build_khash_cffi.py"], cwd=lib, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return True


if __name__ == "__main__":
    print("baseline/_ref installed:", install())
