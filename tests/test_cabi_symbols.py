"""CPU-side checks of the drop-in boundary: the C-ABI library loads (no GPU needed: cudart is linked statically and
touches the driver lazily), exports every symbol include/tarl_b200.h declares, and the ctypes table matches it."""
import ctypes
import os
import re

import pytest

from tarl_simulator_b200 import _cabi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "tarl_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tarl_[a-z0-9_]+)\s*\(", text)))


def header_abi_version():
    text = open(os.path.join(ROOT, "include", "tarl_b200.h")).read()
    return int(re.search(r"#define\s+TARL_ABI_VERSION\s+(\d+)", text).group(1))


@pytest.fixture(scope="module")
def built_lib():
    from tarl_simulator_b200.build import build
    build()
    return ctypes.CDLL(_cabi.LIB_PATH)


def test_header_declares_entry_points():
    syms = declared_symbols()
    assert "tarl_core_step" in syms and "tarl_direction_forward" in syms and "tarl_response_forward" in syms


def test_library_exports_every_declared_symbol(built_lib):
    for name in declared_symbols():
        assert hasattr(built_lib, name), f"{name} declared in include/tarl_b200.h but not exported"


def test_ctypes_table_matches_header(built_lib):
    assert sorted(_cabi.SIGNATURES) == declared_symbols()
    assert _cabi.lib().tarl_abi_version() == _cabi.ABI_VERSION == header_abi_version()
    assert b"workspace" in _cabi.lib().tarl_error_string(-2)


def test_workspace_size_is_monotone():
    lib = _cabi.lib()
    assert lib.tarl_core_workspace_bytes(0) == 0
    assert lib.tarl_core_workspace_bytes(1000) >= 52 * 1000
    assert lib.tarl_core_workspace_bytes(2000) > lib.tarl_core_workspace_bytes(1000)


def test_argument_errors_are_reported_without_a_gpu():
    lib = _cabi.lib()
    csr = _cabi.DualCSR(4, 0, None, None, None, None, None, None)
    rc = lib.tarl_core_step(ctypes.byref(csr), None, 52, 15, None, None, None, None, 0.0, None, None, None, None, 0, None)
    assert rc == -1          # flags == NULL -> TARL_E_BADARG before any CUDA call


def test_compute_classes_refuse_cpu_tensors():
    import torch
    from tarl_simulator_b200.core import DirectionMPNN, SimulationCoreModel
    from tarl_simulator_b200.data import Data
    x = torch.zeros(3, 3 * 5 + 7)
    ei = torch.tensor([[0, 1], [1, 2]])
    with pytest.raises(RuntimeError, match="CUDA"):
        DirectionMPNN(Nmax=5)(x, ei, torch.ones(2, 1))
    g = Data(x=x, edge_index_routes=ei, edge_attr_routes=torch.ones(2, 1), num_roads=3)
    with pytest.raises(RuntimeError, match="CUDA"):
        SimulationCoreModel(Nmax=5, device="cpu", time=0)(g)


def test_uniform_in_edge_weight_hint():
    """topology.uniform_in_weights (what LinkStore parks in stat_a.w under TARL_STORE_UNIFORM_WEIGHTS): the weight when
    all in-edges of a link carry bitwise the same one, NaN when they differ or the link walks its CSR segment, 0 for a
    link without in-edges; -1 pads the unused ELL cells and must not take part in the comparison."""
    import torch
    from tarl_simulator_b200.topology import uniform_in_weights
    cols = torch.tensor([[0.25, 0.5, -1.0, 0.25, 0.3, 1.0 / 3.0],
                         [0.25, 0.25, -1.0, -1.0, 0.3, float(torch.tensor(1.0) / torch.tensor(3.0))],
                         [0.25, -1.0, -1.0, -1.0, 0.3, 0.33333334],
                         [0.25, -1.0, -1.0, -1.0, -2.0, -1.0]])
    deg = torch.tensor([4, 2, 0, 1, 7, 3])
    general = torch.tensor([False, False, False, False, True, False])
    out = uniform_in_weights(cols, deg, general)
    assert out[0] == 0.25 and torch.isnan(out[1]) and out[2] == 0.0 and out[3] == 0.25 and torch.isnan(out[4])
    assert out[5] == cols[0, 5]                               # the same float written three ways
    assert _cabi.STORE_UNIFORM_WEIGHTS == 1
    text = open(os.path.join(ROOT, "include", "tarl_b200.h")).read()
    assert re.search(r"#define\s+TARL_STORE_UNIFORM_WEIGHTS\s+1\b", text)
