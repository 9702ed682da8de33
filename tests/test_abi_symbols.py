"""CPU: the C-ABI library builds/loads without a GPU and exports every symbol
include/sgg_b200.h declares (no compute calls here)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, 'include', 'sgg_b200.h')).read()
    return sorted(set(re.findall(r'SGG_API\s+[\w\s\*]+?\b(sgg_\w+)\s*\(', src)))


def test_header_declares_expected_entry_points():
    syms = declared_symbols()
    for s in ('sgg_mp_forward', 'sgg_l1_forward', 'sgg_graph_build', 'sgg_draw_union_boxes',
              'sgg_union_geom_forward', 'sgg_node_edge_features', 'sgg_linear_forward'):
        assert s in syms


def test_library_exports_every_declared_symbol(built_lib):
    lib = ctypes.CDLL(built_lib)
    for s in declared_symbols():
        assert hasattr(lib, s), 'missing export ' + s


def test_ctypes_signatures_cover_header(built_lib):
    from sgg_b200 import _lib
    assert sorted(_lib.SIGNATURES) == declared_symbols()
    lib = _lib.load()
    assert lib.sgg_abi_version() == 2
    # size queries are host-only and must work without a device
    assert lib.sgg_graph_workspace_bytes(240, 2400) > 0
    assert lib.sgg_mp_workspace_bytes(240, 2400, 512, 3) > (2 * 2640 * 512 * 4)
    assert lib.sgg_l1_workspace_bytes(0, 0, 512, 3) > 0


def test_ops_refuse_cpu_tensors(built_lib):
    import pytest, torch
    from sgg_b200 import ops
    from sgg_b200._lib import SggError
    with pytest.raises(SggError):
        ops.linear(torch.zeros(2, 16), torch.zeros(3, 16))
    with pytest.raises(SggError):
        ops.build_graph(torch.zeros((4, 2), dtype=torch.int64), 3)
