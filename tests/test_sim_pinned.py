"""CPU pins for the simulator loop around the core step: (a) the oracle port of insert / withdraw / choice and of the
two step orders (oracle/agents_port.py) against golden trajectories produced by the UNMODIFIED reference
(tests/golden/sim_*.npz, oracle/gen_golden_sim.py), (b) the host-side MATSim readers of the product against the
graph tensors / agent_features the reference built from the same XML."""
import glob
import os

import numpy as np
import pytest
import torch

import agents_port
from core_port import Cols

HERE = os.path.dirname(os.path.abspath(__file__))
SIM_CASES = sorted(os.path.basename(f)[:-4] for f in glob.glob(os.path.join(HERE, "golden", "sim_*.npz"))
                   if "population" not in f)


def load_case(name):
    d = np.load(os.path.join(HERE, "golden", name + ".npz"))
    graph = {
        "edge_index": torch.from_numpy(d["g_edge_index"]), "edge_index_routes": torch.from_numpy(d["g_edge_index_routes"]),
        "edge_attr_routes": torch.from_numpy(d["g_edge_attr_routes"]), "num_roads": int(d["g_num_roads"]),
        "adj_matrix": torch.from_numpy(d["g_adj_matrix"]), "congestion_constant": torch.from_numpy(d["g_congestion_constant"]),
    }
    return d, graph


@pytest.mark.parametrize("name", SIM_CASES)
def test_port_reproduces_reference_trajectory(name):
    d, graph = load_case(name)
    c = Cols(int(d["Nmax"]))
    af = torch.from_numpy(d["af0"]).clone()
    rl = str(d["mode"]) == "rl"
    x = torch.from_numpy(d["x_reset"] if rl else d["g_x"]).clone()
    for s in range(len(d["t"])):
        t = float(d["t"][s])
        t = int(t) if t == int(t) else t
        u = torch.from_numpy(d["u_core"][s])
        if rl:
            out = agents_port.env_step(x, af, t, c, graph, torch.from_numpy(d["action"][s]), u)
            assert float(out["reward"]) == float(d["reward"][s])
            assert bool(out["done"]) == bool(d["done"][s])
            assert float(out["t_next"]) == float(d["obs_time"][s])
        else:
            out = agents_port.run_step(x, af, t, c, graph, torch.from_numpy(d["u_choice"][s]), u)
        assert torch.equal(x, torch.from_numpy(d["x"][s])), f"x differs after step {s}"
        assert torch.equal(af, torch.from_numpy(d["af"][s])), f"agent_features differ after step {s}"
        assert torch.equal(out["withdrawn"], torch.from_numpy(d["withdrawn"][s]))
        assert torch.equal(out["delta_tt"], torch.from_numpy(d["delta_tt"][s]))
        assert (out["pop"] is not None) == bool(d["has_pop"][s])
        if out["pop"] is not None:
            assert torch.equal(out["pop"], torch.from_numpy(d["pop"][s]))


def test_chain3_known_answers():
    """SURVEY.md §8c vector 3: hand-off at t=5,6,7 with exit times 10,11,12, then 15,16,17; DONE at 15,16,17."""
    d, _ = load_case("sim_chain3")
    af = d["af"][-1]
    assert af[1:, 8].tolist() == [1.0, 1.0, 1.0]
    assert af[1:, 3].tolist() == [15.0, 16.0, 17.0]
    assert int(np.sum(d["has_pop"])) == 6


@pytest.mark.parametrize("name", SIM_CASES)
def test_product_network_reader_matches_reference_graph(name, tmp_path):
    from tarl_simulator_b200.matsim_io import network_from_xml
    d, _ = load_case(name)
    (tmp_path / "network.xml").write_text(str(d["xml"]))
    g, Nmax = network_from_xml(str(tmp_path / "network"))
    assert Nmax == int(d["Nmax"])
    assert int(g.num_roads) == int(d["g_num_roads"])
    for key in ("x", "edge_index", "edge_attr", "edge_index_routes", "edge_attr_routes", "critical_number",
                "congestion_constant", "adj_matrix", "src_adj"):
        assert torch.equal(getattr(g, key), torch.from_numpy(d["g_" + key])), key


def test_product_population_reader_matches_reference(tmp_path):
    from tarl_simulator_b200.matsim_io import population_from_xml
    d = np.load(os.path.join(HERE, "golden", "sim_population_xml.npz"))
    (tmp_path / "network.xml").write_text(str(d["network_xml"]))
    (tmp_path / "population.xml").write_text(str(d["population_xml"]))
    rows = population_from_xml(str(tmp_path), verbose=False)
    assert torch.equal(torch.tensor(rows, dtype=torch.float32), torch.from_numpy(d["agent_features"]))


def test_gz_network_and_missing_file(tmp_path):
    import gzip
    from tarl_simulator_b200.matsim_io import network_from_xml
    d, _ = load_case("sim_twolink")
    with gzip.open(tmp_path / "network.xml.gz", "wt") as f:
        f.write(str(d["xml"]))
    g, Nmax = network_from_xml(str(tmp_path / "network"))
    assert g.x.shape == (6, 52) and Nmax == 15          # the reference's own test expectations (tests/conftest.py)
    assert g.edge_index.size(1) == 6 and g.edge_index_routes.size(1) == 2
    with pytest.raises(FileNotFoundError):
        network_from_xml(str(tmp_path / "nothing"))


def test_choice_rule_matches_injected_multinomial():
    """The port's sparse `choice` and the multinomial stand-in installed into the reference agree by construction on
    dense rows; checked on a random adjacency."""
    g = torch.Generator().manual_seed(0)
    adj = (torch.rand(12, 12, generator=g) < 0.3).float()
    adj = adj[adj.sum(1) > 0]
    u = torch.rand(adj.size(0), generator=g)
    got = agents_port.multinomial_rule(u)(adj / adj.sum(1, keepdim=True)).flatten()
    for i in range(adj.size(0)):
        cols = torch.nonzero(adj[i]).flatten()
        k = min(int(float(u[i]) * cols.numel()), cols.numel() - 1)
        assert int(got[i]) == int(cols[k])
