"""GPU parity of the in-place core step (C ABI: tarl_core_step / tarl_direction_forward / tarl_response_forward)
against the golden vectors produced by the unmodified reference and against the CPU oracle on seeded inputs.
Bit-exact: whole x, delta_travel_time, pop masks, history cadence."""
import glob
import os

import numpy as np
import pytest
import torch

import cases
import core_port

pytestmark = pytest.mark.gpu

CORE_CASES = sorted(os.path.basename(f)[:-4] for f in glob.glob(os.path.join(os.path.dirname(__file__), "golden", "core_*.npz")))


def make_graph(x, ei, w, use_static, Nmax):
    from tarl_simulator_b200.data import Data
    g = Data(x=x.cuda(), edge_index_routes=ei.cuda(), edge_attr_routes=w.cuda(), num_roads=x.size(0))
    if use_static:
        c = core_port.Cols(Nmax)
        crit, cc = core_port.static_factors(x, c)
        g.critical_number, g.congestion_constant = crit.cuda(), cc.cuda()
    return g


@pytest.mark.parametrize("name", CORE_CASES)
def test_core_step_matches_reference_goldens(name, golden_dir):
    from tarl_simulator_b200.core import SimulationCoreModel
    d = np.load(os.path.join(golden_dir, name + ".npz"))
    Nmax = int(d["Nmax"])
    c = core_port.Cols(Nmax)
    g = make_graph(torch.from_numpy(d["x0"]), torch.from_numpy(d["edge_index"]), torch.from_numpy(d["edge_attr"]),
                   bool(d["use_static"]), Nmax)
    model = SimulationCoreModel(Nmax=Nmax, device="cuda", time=0)
    n_hist = 0
    for s in range(len(d["t"])):
        g.x[:, c.SEL] = torch.from_numpy(d["sel"][s]).cuda()
        model.set_time(float(d["t"][s]))
        out = model(g, noise=torch.from_numpy(d["u"][s]).cuda())
        assert out is g
        assert torch.equal(g.x.cpu(), torch.from_numpy(d["x"][s])), f"x differs after step {s}"
        assert torch.equal(model.direction_mpnn.road_optimality_data["delta_travel_time"].cpu(), torch.from_numpy(d["delta_tt"][s]))
        hist = model.response_mpnn.update_history
        n_hist += int(d["has_pop"][s])
        assert len(hist) == n_hist
        if d["has_pop"][s]:
            assert hist[-1][0] == float(d["t"][s])
            assert torch.equal(hist[-1][1].cpu(), torch.from_numpy(d["pop"][s]))


@pytest.mark.parametrize("name", ["core_rand_a", "core_rand_b", "core_braess"])
def test_direction_then_response_standalone(name, golden_dir):
    """The two MPNNs called one after the other on a bare x (as the reference's tests do) equal the fused step."""
    from tarl_simulator_b200.core import DirectionMPNN, ResponseMPNN
    d = np.load(os.path.join(golden_dir, name + ".npz"))
    Nmax = int(d["Nmax"])
    c = core_port.Cols(Nmax)
    x = torch.from_numpy(d["x0"]).cuda()
    ei, w = torch.from_numpy(d["edge_index"]).cuda(), torch.from_numpy(d["edge_attr"]).cuda()
    cc = crit = None
    if bool(d["use_static"]):
        crit, cc = (t.cuda() for t in core_port.static_factors(torch.from_numpy(d["x0"]), c))
    dm, rm = DirectionMPNN(Nmax=Nmax), ResponseMPNN(Nmax=Nmax)
    for s in range(len(d["t"])):
        x[:, c.SEL] = torch.from_numpy(d["sel"][s]).cuda()
        dm.set_time(float(d["t"][s])); rm.set_time(float(d["t"][s]))
        out = dm(x, ei, w, critical_number=crit, congestion_constant=cc, noise=torch.from_numpy(d["u"][s]).cuda())
        assert out.shape == x.shape
        out = rm(out, ei, w)
        assert torch.equal(out.cpu(), torch.from_numpy(d["x"][s]))
        assert torch.equal(dm.road_optimality_data["delta_travel_time"].cpu(), torch.from_numpy(d["delta_tt"][s]))
    assert len(rm.update_history) == int(d["has_pop"].sum())


@pytest.mark.parametrize("seed,N,Nmax,sorted_src", [(11, 5000, 15, True), (12, 3000, 33, False), (13, 20000, 15, True)])
def test_core_step_matches_oracle_on_seeded_inputs(seed, N, Nmax, sorted_src):
    from tarl_simulator_b200.core import SimulationCoreModel
    g = torch.Generator().manual_seed(seed)
    ei, w = cases.random_dual_graph(g, N, 4, sort_by_source=sorted_src)
    x0, _ = cases.random_road_state(g, N, Nmax, 300.0, ei)
    c = core_port.Cols(Nmax)
    graph = make_graph(x0, ei, w, True, Nmax)
    cc = core_port.static_factors(x0, c)[1]
    model = SimulationCoreModel(Nmax=Nmax, device="cuda", time=300)
    x = x0.clone()
    total_pops = 0
    for s in range(8):
        sel = cases.random_selection(g, N, ei)
        u = cases.uniforms(g, ei.size(1))
        x[:, c.SEL] = sel
        t = 300.0 + s
        ref = core_port.core_step(x, ei, w, t, Nmax, u, cc)
        model.set_time(t)
        if s % 2:       # decisions handed to the kernel (fused "apply action, then core")
            model(graph, noise=u.cuda(), selected_road=sel.cuda())
        else:           # decisions written into graph.x by the caller, as the reference's loop does
            graph.x[:, c.SEL] = sel.cuda()
            model(graph, noise=u.cuda())
        assert torch.equal(graph.x.cpu(), x), f"x differs after step {s}"
        assert torch.equal(model.direction_mpnn.road_optimality_data["delta_travel_time"].cpu(), ref["delta_tt"])
        total_pops += 0 if ref["pop"] is None else int(ref["pop"].sum())
    hist = model.response_mpnn.update_history
    assert sum(int(m.sum()) for _, m in hist) == total_pops and total_pops > 0
    model.check_errors()


def test_overflow_sets_error_flag():
    from tarl_simulator_b200.core import SimulationCoreModel
    Nmax = 6
    c = core_port.Cols(Nmax)
    x = torch.zeros(2, c.F)
    x[:, c.MAXN], x[:, c.FFTT], x[:, c.RIDX] = 5, 3.0, torch.arange(2.0)
    x[0, c.NUM] = Nmax          # tail write would alias the arrival-time segment in the reference
    graph = make_graph(x, torch.tensor([[0], [1]]), torch.ones(1, 1), False, Nmax)
    model = SimulationCoreModel(Nmax=Nmax, device="cuda", time=0)
    model(graph, noise=torch.full((1,), 0.5).cuda())
    with pytest.raises(RuntimeError, match="NUMBER_OF_AGENT"):
        model.check_errors()


def test_empty_and_edgeless_graphs():
    from tarl_simulator_b200.core import SimulationCoreModel
    Nmax = 5
    c = core_port.Cols(Nmax)
    x = torch.zeros(3, c.F)
    x[:, c.MAXN], x[:, c.FFTT] = 4, 2.0
    ref = x.clone()
    core_port.core_step(ref, torch.zeros(2, 0, dtype=torch.long), torch.zeros(0, 1), 7.0, Nmax, torch.zeros(0))
    graph = make_graph(x, torch.zeros(2, 0, dtype=torch.long), torch.zeros(0, 1), False, Nmax)
    model = SimulationCoreModel(Nmax=Nmax, device="cuda", time=7)
    model(graph)
    assert torch.equal(graph.x.cpu(), ref)
    assert len(model.response_mpnn.update_history) == 0
