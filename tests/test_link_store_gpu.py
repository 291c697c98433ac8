"""GPU parity of the resident link store (C ABI: tarl_store_import / tarl_store_step / tarl_store_export): after every
step the exported rows must equal, bit for bit, the x the unmodified reference produced (golden vectors) or the CPU
oracle produces on seeded inputs — every cell, including what the reference leaves past the queue tails."""
import glob
import os

import numpy as np
import pytest
import torch

import cases
import core_port

pytestmark = pytest.mark.gpu

CORE_CASES = sorted(os.path.basename(f)[:-4] for f in glob.glob(os.path.join(os.path.dirname(__file__), "golden", "core_*.npz")))


def make_store(x0, ei, w, Nmax, use_static, replicas=1, seed=0, cluster=None):
    from tarl_simulator_b200.data import Data
    from tarl_simulator_b200.engine import LinkStore
    g = Data(x=x0.cuda(), edge_index_routes=ei.cuda(), edge_attr_routes=w.cuda(), num_roads=x0.size(0))
    if use_static:
        crit, cc = core_port.static_factors(x0, core_port.Cols(Nmax))
        g.critical_number, g.congestion_constant = crit.cuda(), cc.cuda()
    return LinkStore.from_graph(g, Nmax, replicas=replicas, seed=seed, cluster=cluster), g


VARIANTS = [0, 1]     # ELL (default), CSR


@pytest.mark.parametrize("cluster", [False, True])       # True: links kept in the store's own locality order
@pytest.mark.parametrize("variant", VARIANTS)
@pytest.mark.parametrize("name", CORE_CASES)
def test_store_matches_reference_goldens(name, golden_dir, variant, cluster):
    d = np.load(os.path.join(golden_dir, name + ".npz"))
    Nmax = int(d["Nmax"])
    x0 = torch.from_numpy(d["x0"])
    ei, w = torch.from_numpy(d["edge_index"]), torch.from_numpy(d["edge_attr"])
    store, _ = make_store(x0, ei, w, Nmax, bool(d["use_static"]), cluster=cluster)
    assert torch.equal(store.export_x()[0].cpu(), x0), "import -> export must be the identity"
    E = ei.size(1)
    dtt = torch.empty(1, E, device="cuda")
    for s in range(len(d["t"])):
        store.set_selected_road(torch.from_numpy(d["sel"][s]).cuda())
        pop = store.step(float(d["t"][s]), noise=torch.from_numpy(d["u"][s]).cuda(), delta_tt=dtt, variant=variant)
        assert torch.equal(store.export_x()[0].cpu(), torch.from_numpy(d["x"][s])), f"x differs after step {s}"
        assert torch.equal(dtt[0].cpu(), torch.from_numpy(d["delta_tt"][s]))
        assert torch.equal(pop[0].bool().cpu(), torch.from_numpy(d["pop"][s]))
        assert bool(store.flags[0].item() != 0) >= bool(d["has_pop"][s])
    store.check_errors()


@pytest.mark.parametrize("variant", VARIANTS)
@pytest.mark.parametrize("seed,N,Nmax,R,max_out", [(21, 4000, 15, 1, 4), (22, 1500, 7, 3, 4), (23, 2500, 40, 2, 4),
                                                    (24, 1200, 15, 2, 24)])   # 24: tiles above the staging capacity
def test_store_replicas_match_oracle(seed, N, Nmax, R, max_out, variant):
    g = torch.Generator().manual_seed(seed)
    ei, w = cases.random_dual_graph(g, N, max_out, sort_by_source=bool(seed % 2))
    x0, _ = cases.random_road_state(g, N, Nmax, 500.0, ei)
    c = core_port.Cols(Nmax)
    cc = core_port.static_factors(x0, c)[1]
    store, _ = make_store(x0, ei, w, Nmax, True, replicas=R, cluster=bool(seed % 2 == 0))
    xs = [x0.clone() for _ in range(R)]
    E = ei.size(1)
    dtt = torch.empty(R, E, device="cuda")
    pops = 0
    for s in range(10):
        t = 500.0 + s
        sel = torch.stack([cases.random_selection(g, N, ei) for _ in range(R)])
        u = torch.stack([cases.uniforms(g, E) for _ in range(R)])
        store.set_selected_road(sel.cuda())
        pop = store.step(t, noise=u.cuda(), delta_tt=dtt, variant=variant)
        out = store.export_x().cpu()
        for r in range(R):
            xs[r][:, c.SEL] = sel[r]
            ref = core_port.core_step(xs[r], ei, w, t, Nmax, u[r], cc)
            assert torch.equal(out[r], xs[r]), f"replica {r} differs after step {s}"
            assert torch.equal(dtt[r].cpu(), ref["delta_tt"])
            rp = ref["pop"] if ref["pop"] is not None else torch.zeros(N, dtype=torch.bool)
            assert torch.equal(pop[r].bool().cpu(), rp)
            pops += int(rp.sum())
    assert pops > 0
    store.check_errors()


def test_store_matches_inplace_path_at_scale():
    """Size-independent property at a size the CPU oracle would need minutes for: the store and the in-place kernels
    (already pinned to the oracle) produce identical rows after many steps on a 250k-link ring-radial network."""
    from tarl_simulator_b200 import synthetic
    from tarl_simulator_b200.core import SimulationCoreModel
    from tarl_simulator_b200.engine import LinkStore
    frm, to, n_nodes = synthetic.ring_radial_links(250, 250, device="cuda")
    g, Nmax = synthetic.build_graph(frm, to, n_nodes, with_full_edges=False)
    synthetic.warm_state(g, Nmax, 500_000, 1000.0, seed=3)
    N, E = int(g.num_roads), g.edge_index_routes.size(1)
    store = LinkStore.from_graph(g, Nmax)
    model = SimulationCoreModel(Nmax=Nmax, device="cuda", time=1000)
    gen = torch.Generator(device="cuda").manual_seed(5)
    for s in range(25):
        t = 1000.0 + s
        sel = synthetic.random_out_neighbour(g, 100 + s)
        u = torch.rand(E, device="cuda", generator=gen).clamp_(min=1e-7)
        store.set_selected_road(sel)
        store.step(t, noise=u)
        model.set_time(t)
        model(g, noise=u, selected_road=sel)
    assert torch.equal(store.export_x()[0], g.x[:N])
    assert float(store.num_agents().sum()) == float(g.x[:N, 3 * Nmax + 1].sum())
    store.check_errors(); model.check_errors()


def test_store_philox_stream_is_deterministic_and_seeded():
    g = torch.Generator().manual_seed(31)
    N, Nmax = 3000, 15
    ei, w = cases.random_dual_graph(g, N, 4)
    x0, _ = cases.random_road_state(g, N, Nmax, 50.0, ei, garbage=False)
    outs = []
    for seed in (7, 7, 8):
        store, _ = make_store(x0, ei, w, Nmax, True, seed=seed)
        for s in range(5):
            store.step(50.0 + s)
        outs.append(store.export_x()[0].cpu())
        store.check_errors()
    assert torch.equal(outs[0], outs[1])
    assert not torch.equal(outs[0], outs[2])
    c = core_port.Cols(Nmax)
    assert float((outs[0][:, c.NUM] - x0[:, c.NUM]).abs().sum()) > 0      # agents did move


@pytest.mark.parametrize("variant", VARIANTS)
@pytest.mark.parametrize("mode", ["tiny_attr", "extreme_noise"])
def test_store_literal_scan_outside_the_safe_bounds(mode, variant):
    """Edge weights below 1e-3 or uniforms outside [2^-24, 1-2^-24] switch the direction phase from the
    eligible-edges-only arg-max to the literal scan over every in-edge; both must agree with the oracle (a tiny
    weight lets an INELIGIBLE edge win the Gumbel arg-max, exactly as in src/direction_mpnn.py:136-144)."""
    g = torch.Generator().manual_seed(77)
    N, Nmax = 3000, 15
    ei, w = cases.random_dual_graph(g, N, 4)
    if mode == "tiny_attr":
        w = w * torch.where(torch.rand(w.shape, generator=g) < 0.5, 1e-9, 1.0)
    x0, _ = cases.random_road_state(g, N, Nmax, 500.0, ei)
    c = core_port.Cols(Nmax)
    cc = core_port.static_factors(x0, c)[1]
    store, _ = make_store(x0, ei, w, Nmax, True)
    x = x0.clone()
    E = ei.size(1)
    dtt = torch.empty(1, E, device="cuda")
    chosen_ineligible = 0
    for s in range(8):
        t = 500.0 + s
        sel = cases.random_selection(g, N, ei)
        u = cases.uniforms(g, E)
        if mode == "extreme_noise":
            k = torch.rand(E, generator=g)
            u = torch.where(k < 0.2, torch.full_like(u, 1e-30), u)           # Gumbel noise -4.2: still finite
            u = torch.where(k > 0.9, torch.full_like(u, 1.0 - 2.0 ** -24), u)
        store.set_selected_road(sel.cuda())
        pop = store.step(t, noise=u.cuda(), delta_tt=dtt, variant=variant)
        x[:, c.SEL] = sel
        ref = core_port.core_step(x, ei, w, t, Nmax, u, cc)
        assert torch.equal(store.export_x()[0].cpu(), x), f"x differs after step {s}"
        assert torch.equal(dtt[0].cpu(), ref["delta_tt"])
        chosen_ineligible += int((ref["chosen"] != 0).sum())
    assert chosen_ineligible > 0
    store.check_errors()


@pytest.mark.parametrize("variant", VARIANTS)
def test_store_run_equals_repeated_steps(variant):
    """tarl_store_run (n steps enqueued by one call, SELECTED_ROAD bank cycled) == n calls of tarl_store_step."""
    g = torch.Generator().manual_seed(5)
    N, Nmax, R = 5000, 15, 2
    ei, w = cases.random_dual_graph(g, N, 4)
    x0, _ = cases.random_road_state(g, N, Nmax, 100.0, ei, garbage=False)
    bank = [torch.stack([cases.random_selection(g, N, ei) for _ in range(R)]).reshape(-1).cuda() for _ in range(3)]
    a, _ = make_store(x0, ei, w, Nmax, True, replicas=R, seed=11)
    b, _ = make_store(x0, ei, w, Nmax, True, replicas=R, seed=11)
    E = ei.size(1)
    da, db = torch.empty(R, E, device="cuda"), torch.empty(R, E, device="cuda")
    n = 7
    for s in range(n):
        a.set_selected_road(bank[s % 3].view(R, N))
        pa = a.step(100.0 + s, delta_tt=da, variant=variant).clone()
    pb = b.run(100.0, n, dt=1.0, sel_bank=bank, delta_tt=db, variant=variant)
    assert torch.equal(a.export_x(), b.export_x())
    assert torch.equal(da, db) and torch.equal(pa, pb)
    assert a.step_id == b.step_id == n and a.t_last == b.t_last
    more = b.run(107.0, 2, sel_bank=None, variant=variant)          # even count, no bank: SELECTED_ROAD stays
    a.step(107.0, variant=variant); pa = a.step(108.0, variant=variant)
    assert torch.equal(a.export_x(), b.export_x()) and torch.equal(pa, more)
    a.check_errors(); b.check_errors()


def test_uniform_weight_hint_changes_nothing_but_the_bytes_read(monkeypatch):
    """stat_a.w = the weight shared by all in-edges of a link (TARL_STORE_UNIFORM_WEIGHTS): the ELL direction kernel
    reads it instead of the link's edge-weight column. A store with the hint and one without (every link reads its
    column, as before ABI 28) must stay bit-identical — in-kernel noise, so contested links draw Gumbel scores from the
    weights — on a graph that has links of both kinds."""
    g = torch.Generator().manual_seed(77)
    N, Nmax, R = 6000, 15, 2
    ei, w = cases.random_dual_graph(g, N, 4)
    x0, _ = cases.random_road_state(g, N, Nmax, 100.0, ei)
    bank = [torch.stack([cases.random_selection(g, N, ei) for _ in range(R)]).reshape(-1).cuda() for _ in range(3)]
    a, _ = make_store(x0, ei, w, Nmax, True, replicas=R, seed=5)
    monkeypatch.setenv("TARL_NO_UNIFORM_WEIGHTS", "1")
    b, _ = make_store(x0, ei, w, Nmax, True, replicas=R, seed=5)
    assert a.uniform_weights and not b.uniform_weights
    hint = a.stat_a[:N, 3]
    assert int(torch.isnan(hint).sum()) > N // 10 and int((hint > 0).sum()) > N // 10      # both kinds of links
    assert bool(torch.isnan(b.stat_a[:N, 3]).all())                                       # import leaves NaN
    E = ei.size(1)
    da, db = torch.empty(R, E, device="cuda"), torch.empty(R, E, device="cuda")
    for s in range(8):
        for st in (a, b):
            st.set_selected_road(bank[s % 3].view(R, N))
        pa = a.step(100.0 + s, delta_tt=da).clone()
        pb = b.step(100.0 + s, delta_tt=db)
        assert torch.equal(pa, pb) and torch.equal(da, db), f"step {s}"
    assert torch.equal(a.export_x(), b.export_x())
    a.import_x(a.export_x())                                    # a re-import must not lose the hint
    assert torch.equal(torch.nan_to_num(a.stat_a[:N, 3], nan=-1.0), torch.nan_to_num(hint, nan=-1.0))
    a.check_errors(); b.check_errors()
