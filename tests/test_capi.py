"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol include/efgh_b200.h declares."""
import ctypes
import os
import re

import pytest

from efgh_b200 import _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def library():
    _capi.build()
    return ctypes.CDLL(_capi.LIB_PATH)


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "efgh_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(efgh_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_expected_surface():
    syms = declared_symbols()
    for must in ("efgh_lattice_points", "efgh_lattice_vertices", "efgh_bcl_scatter", "efgh_bcl_gather",
                 "efgh_bcl_conv", "efgh_bcl_conv_dgrad", "efgh_bcl_conv_wgrad", "efgh_last_error"):
        assert must in syms


def test_library_exports_every_declared_symbol(library):
    for name in declared_symbols():
        assert hasattr(library, name), "libefgh_b200.so does not export %s" % name


def test_binding_covers_header():
    assert sorted(_capi.SIGNATURES) == declared_symbols()


def test_binding_arity_matches_header():
    text = open(os.path.join(ROOT, "include", "efgh_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    for name, (_, args) in _capi.SIGNATURES.items():
        m = re.search(r"\b%s\s*\(([^;]*?)\)\s*;" % name, text, flags=re.S)
        assert m, name
        params = m.group(1).strip()
        n = 0 if params in ("", "void") else params.count(",") + 1
        assert n == len(args), "%s: header has %d parameters, binding %d" % (name, n, len(args))


def test_pure_host_entry_points(library):
    # no compute calls without a GPU: only the size query and the error string
    library.efgh_lattice_workspace_bytes.restype = ctypes.c_size_t
    library.efgh_lattice_workspace_bytes.argtypes = [ctypes.c_int64]
    small = library.efgh_lattice_workspace_bytes(1000)
    big = library.efgh_lattice_workspace_bytes(131072)
    assert 0 < small < big < 64 * 1024 * 1024
    library.efgh_version.restype = ctypes.c_int
    assert library.efgh_version() >= 100


def test_host_side_size_queries():
    """The batch / gather-splat / stem size queries are pure host arithmetic (no GPU needed)."""
    L = _capi.lib()
    assert L.efgh_lattice_table_entries(131072, 4 * 131072) == 1 << 20          # 8 entries per point
    assert L.efgh_lattice_table_entries(131072, 2 * 131072) == 1 << 19          # sized by the vertex capacity
    assert L.efgh_lattice_table_entries(1, 1) == 1024                           # floor
    for B in (1, 2, 16, 64):
        n = L.efgh_lattice_batch_info_ints(B)
        assert n >= (B + 1) + B + 8 * B and n % 8 == 0
    assert L.efgh_lattice_vertex_offsets_ints(1000) > 1001
    one = L.efgh_lattice_batch_workspace_bytes(1, 1 << 20, 131072)
    four = L.efgh_lattice_batch_workspace_bytes(4, 1 << 20, 4 * 131072)
    assert 0 < one < four
    assert L.efgh_bcl_stem_weight_floats(3, 32, 32, 32) == 3 * 32 + 32 + 2 * (32 * 32 + 32)
    # tensor-core convolution: shapes of the five E-Net levels are supported, odd shapes are not
    for cin, (cmid, cout) in ((36, (32, 32)), (36, (64, 64)), (68, (128, 128)), (132, (256, 256)), (260, (256, 256))):
        assert L.efgh_bcl_conv_tc_supported(cin, 15, cmid, 3) == 1
        assert L.efgh_bcl_conv_tc_supported(cmid, 1, cout, 3) == 1
    assert L.efgh_bcl_conv_tc_supported(35, 15, 32, 3) == 0      # C not a multiple of 4
    assert L.efgh_bcl_conv_tc_supported(36, 15, 48, 3) == 0      # M not a multiple of 32
    assert L.efgh_bcl_conv_tc_groups(15 * 36, 32) == 1 and L.efgh_bcl_conv_tc_groups(15 * 132, 128) == 1     # cuts summed on chip
    assert L.efgh_bcl_conv_tc_groups(15 * 132, 256) == 8 and L.efgh_bcl_conv_tc_groups(256, 256) == 1        # K groups added in L2


def test_no_cpu_fallback():
    """The product modules refuse CPU devices instead of silently computing elsewhere."""
    import torch
    from efgh_b200.generate_data import GenerateData
    from efgh_b200.bilateralNN import BilateralConvFlex
    with pytest.raises(Exception):
        GenerateData(3, [[1.0, 1]], "cpu")
    m = BilateralConvFlex(3, 1, 4, [4, 4], "cpu", True, True, True, True, False, False)
    with pytest.raises(Exception):
        m(torch.zeros(1, 4, 8), torch.zeros(1, 4, 8), torch.zeros(1, 4, 8, dtype=torch.long),
          torch.zeros(1, 15, 3, dtype=torch.long), None, None)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "efgh_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("oracle/", "").lower() or f == "__init__.py" or \
                    "import oracle" not in src and "from oracle" not in src, f


def test_state_dict_keys_match_reference_layout():
    from efgh_b200.bilateralNN import BilateralConvFlex
    m = BilateralConvFlex(3, 1, 36, [32, 32], "cuda", True, True, True, True, False, False, chunk_size=-1)
    sd = m.state_dict()
    assert list(sd) == ["feat_indices", "blur_conv.0.weight", "blur_conv.0.bias", "blur_conv.2.weight", "blur_conv.2.bias"]
    assert tuple(sd["blur_conv.0.weight"].shape) == (32, 36, 15, 1) and tuple(sd["blur_conv.2.weight"].shape) == (32, 32, 1, 1)
    m2 = BilateralConvFlex(3, 1, 8, [16, 9], "cuda", True, True, True, True, True, True)
    assert "bias" in m2.state_dict() and "out_indices" in m2.state_dict()
