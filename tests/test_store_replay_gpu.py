"""Parity of the BENCHMARKED configuration against the CPU oracle.

bench.py times `LinkStore.run(...)` with noise=None — the kernel instantiation that draws its uniforms in-kernel
(Philox4x32-10) — on grid100 (BASELINE configs[2], 39 600 links) and ring_radial_1m (configs[3], 999 000 links).
`tarl_store_noise` writes the uniforms of that stream out in the reference's [E] form, so the oracle
(oracle/core_port.core_step, pinned to the unmodified reference) can replay every step with exactly the numbers the
kernels consumed: whole x, delta_travel_time[E] and the pop mask must be bit-identical after every step, at full size.
The same two sizes also go through the drop-in `SimulationCoreModel.forward` with injected noise, and the in-kernel
picks are checked to be proportional to edge_attr on contested fan-ins (src/direction_mpnn.py:133-144).
"""
import pytest
import torch

import core_port

pytestmark = pytest.mark.gpu

T0 = 21600.0


def _workload(name):
    from tarl_simulator_b200 import synthetic
    g, Nmax, placed = synthetic.make_workload(name, device="cuda", t=T0, seed=0)
    return g, Nmax, placed


def _oracle_inputs(g, Nmax):
    N = int(g.num_roads)
    return (g.x[:N].cpu().clone(), g.edge_index_routes.cpu(), g.edge_attr_routes.cpu(),
            g.congestion_constant[:N].cpu(), core_port.Cols(Nmax))


def _unpack_bits(words, N):
    w = words.view(-1).cpu().to(torch.int64) & 0xFFFFFFFF
    bits = (w.unsqueeze(1) >> torch.arange(32)) & 1
    return bits.reshape(-1)[:N].bool()


@pytest.mark.parametrize("name,steps", [("grid100", 12), ("ring_radial_1m", 10)])
def test_store_run_with_inkernel_noise_replays_on_the_oracle(name, steps):
    from tarl_simulator_b200 import synthetic
    from tarl_simulator_b200.engine import LinkStore
    g, Nmax, _ = _workload(name)
    N, E = int(g.num_roads), g.edge_index_routes.size(1)
    x, ei, w, cc, c = _oracle_inputs(g, Nmax)
    store = LinkStore.from_graph(g, Nmax, seed=1234)
    twin = LinkStore.from_graph(g, Nmax, seed=1234)
    bank = [synthetic.random_out_neighbour(g, 1000 + i) for i in range(4)]
    contested = pops = 0
    for s in range(steps):
        t = T0 + s
        u = store.noise_of_step()                      # what the in-kernel stream yields for this step id
        assert float(u.min()) > 0.0 and float(u.max()) < 1.0
        store.run(t, 1, sel_bank=[bank[s % 4]], pop_bits=True)          # noise=None: tarl_store_run, Philox in-kernel
        x[:, c.SEL] = bank[s % 4].cpu()
        ref = core_port.core_step(x, ei, w, t, Nmax, u[0].cpu(), cc)
        assert torch.equal(store.export_x()[0].cpu(), x), f"{name}: x differs from the oracle after step {s}"
        assert torch.equal(store.expand_delta_tt()[0].cpu(), ref["delta_tt"]), f"delta_tt differs at step {s}"
        rp = ref["pop"] if ref["pop"] is not None else torch.zeros(N, dtype=torch.bool)
        assert torch.equal(store.pop[:N].bool().cpu(), rp), f"pop mask differs at step {s}"
        assert torch.equal(_unpack_bits(store.pop_bits[: store.words], N), rp), f"pop bits differ at step {s}"
        pops += int(rp.sum())
        # links where the Gumbel draw decided (two or more eligible in-edges): the oracle's prob > 0 count per target
        prob = ref["prob"] if "prob" in ref else None
        if prob is not None:
            contested += int((torch.zeros(N).scatter_add_(0, ei[1], (prob.flatten() > 0).float()) > 1).sum())
    store.check_errors()
    assert pops > steps * N // 50
    if "prob" in ref:
        assert contested > 0
    # the same steps enqueued by ONE call (what bench.py does) end in the same state
    twin.run(T0, steps, sel_bank=bank)
    assert torch.equal(twin.export_x(), store.export_x())
    assert torch.equal(twin.dtt_link, store.dtt_link) and torch.equal(twin.pop, store.pop)
    twin.check_errors()


@pytest.mark.parametrize("name,steps", [("grid100", 12), ("ring_radial_1m", 10)])
def test_dropin_forward_with_injected_noise_at_benchmark_sizes(name, steps):
    from tarl_simulator_b200 import synthetic
    from tarl_simulator_b200.core import SimulationCoreModel
    g, Nmax, _ = _workload(name)
    N, E = int(g.num_roads), g.edge_index_routes.size(1)
    x, ei, w, cc, c = _oracle_inputs(g, Nmax)
    model = SimulationCoreModel(Nmax=Nmax, device="cuda", time=T0)
    gen = torch.Generator().manual_seed(7)
    pops = 0
    for s in range(steps):
        t = T0 + s
        sel = synthetic.random_out_neighbour(g, 2000 + s)
        u = torch.rand(E, generator=gen).clamp_(min=1e-7)
        model.set_time(t)
        model(g, noise=u.cuda(), selected_road=sel)
        x[:, c.SEL] = sel.cpu()
        ref = core_port.core_step(x, ei, w, t, Nmax, u, cc)
        assert torch.equal(g.x[:N].cpu(), x), f"{name}: graph.x differs from the oracle after step {s}"
        dtt = model.direction_mpnn.road_optimality_data["delta_travel_time"]
        assert dtt.shape == (E,) and torch.equal(dtt.cpu(), ref["delta_tt"])
        rp = ref["pop"] if ref["pop"] is not None else torch.zeros(N, dtype=torch.bool)
        assert torch.equal(model.last_pop.bool().cpu(), rp)
        pops += int(rp.sum())
    model.check_errors()
    assert pops > 0
    assert len(model.response_mpnn.update_history) == steps


@pytest.mark.parametrize("name,steps", [("grid100", 9), ("ring_radial_1m", 6)])
def test_dropin_forward_with_host_buffers_matches_the_oracle(name, steps):
    """SimulationCoreModel.forward(graph, selected_road=<pinned host tensor>, host_out=...) — the call bench.py times
    end to end (tarl_store_step_host: upload, kernels and downloads enqueued by one library call on its own copy
    streams, two alternating slots) — against the oracle, step by step: graph.x, and the HOST copies of
    delta_travel_time per link (expanded to the reference's [E] form) and of the pop bits."""
    from tarl_simulator_b200 import synthetic
    from tarl_simulator_b200.core import SimulationCoreModel
    g, Nmax, _ = _workload(name)
    N, E = int(g.num_roads), g.edge_index_routes.size(1)
    x, ei, w, cc, c = _oracle_inputs(g, Nmax)
    model = SimulationCoreModel(Nmax=Nmax, device="cuda", time=T0, resident="always")
    words = (N + 31) // 32
    sels = [torch.empty(N, dtype=torch.float32).pin_memory() for _ in range(3)]
    outs = [{"delta_tt_link": torch.empty(N, dtype=torch.float32).pin_memory(),
             "pop_bits": torch.empty(words, dtype=torch.int32).pin_memory()} for _ in range(steps)]
    gen = torch.Generator().manual_seed(11)
    us, sel_log = [], []
    for s in range(steps):                       # every step is enqueued before anything is read back
        sel = synthetic.random_out_neighbour(g, 3000 + s).cpu()
        if s >= 3:
            model.host_sync(g)                   # the pinned input buffer about to be reused has been uploaded
        sels[s % 3].copy_(sel)
        u = torch.rand(E, generator=gen).clamp_(min=1e-7)
        us.append(u); sel_log.append(sel)
        model.set_time(T0 + s)
        model(g, noise=u.cuda(), selected_road=sels[s % 3], host_out=outs[s])
        assert model.last_path == "resident"
    model.host_sync(g)
    src = ei[0]
    for s in range(steps):
        x[:, c.SEL] = sel_log[s]
        ref = core_port.core_step(x, ei, w, T0 + s, Nmax, us[s], cc)
        assert torch.equal(outs[s]["delta_tt_link"][src], ref["delta_tt"]), f"{name}: host delta_tt differs at step {s}"
        rp = ref["pop"] if ref["pop"] is not None else torch.zeros(N, dtype=torch.bool)
        assert torch.equal(_unpack_bits(outs[s]["pop_bits"], N), rp), f"{name}: host pop bits differ at step {s}"
    assert torch.equal(g.x[:N].cpu(), x), f"{name}: graph.x differs from the oracle after {steps} host steps"
    model.check_errors()
    with pytest.raises(ValueError):              # host buffers must be pinned
        model(g, selected_road=torch.zeros(N), host_out=outs[0])


def _fan_in_case(weights, copies, Nmax=15):
    """`copies` independent motifs: k upstream links (each holding one due agent that selects d) -> one empty link d.
    Link ids: motif m owns [m*(k+1), (m+1)*(k+1)), d last. Returns x0, edge_index, edge_attr."""
    k = len(weights)
    c = core_port.Cols(Nmax)
    n = copies * (k + 1)
    x = torch.zeros(n, c.F)
    x[:, c.MAXN] = 14.0
    x[:, c.FFTT] = 7.2
    x[:, c.LENGTH] = 100.0
    x[:, c.MAX_FLOW] = 1800.0
    x[:, c.RIDX] = torch.arange(n, dtype=torch.float32)
    base = torch.arange(copies) * (k + 1)
    d = base + k
    src, dst, attr = [], [], []
    for j in range(k):
        u = base + j
        x[u, c.ID0] = (u + 1).float()                 # agent id = link id + 1
        x[u, c.ARR0] = 90.0
        x[u, c.DEP0] = 99.0                           # due at t = 100
        x[u, c.NUM] = 1.0
        x[u, c.SEL] = d.float()
        src.append(u); dst.append(d); attr.append(torch.full((copies,), float(weights[j])))
    ei = torch.stack([torch.stack(src, 1).reshape(-1), torch.stack(dst, 1).reshape(-1)])
    return x, ei, torch.stack(attr, 1).reshape(-1, 1), k


@pytest.mark.parametrize("weights", [(0.5, 0.3, 0.2), (0.05, 0.4, 0.1, 0.15, 0.2, 0.1), (0.1,) * 10])
def test_inkernel_picks_are_proportional_to_edge_attr(weights):
    """Gumbel-max over log(p) + g picks in-edge j with probability w_j / sum(w) (src/direction_mpnn.py:133-144). One
    step of many independent fan-ins with the in-kernel stream: frequencies within 4.5 sigma of the weights, a second
    seed draws differently, and the dumped noise replays the very same picks on the oracle. 3 in-edges: ELL width 4;
    6: width 8; 10: more than the ELL width (CSR segment)."""
    from tarl_simulator_b200.data import Data
    from tarl_simulator_b200.engine import LinkStore
    copies, Nmax = 30000, 15
    x0, ei, attr, k = _fan_in_case(weights, copies, Nmax)
    c = core_port.Cols(Nmax)
    g = Data(x=x0.cuda(), edge_index_routes=ei.cuda(), edge_attr_routes=attr.cuda(), num_roads=x0.size(0))
    crit, cc = core_port.static_factors(x0, c)
    g.critical_number, g.congestion_constant = crit.cuda(), cc.cuda()
    d = torch.arange(copies) * (k + 1) + k
    winners = {}
    for seed in (5, 6):
        store = LinkStore.from_graph(g, Nmax, seed=seed)
        u = store.noise_of_step()
        store.step(100.0)                              # noise=None
        out = store.export_x()[0].cpu()
        store.check_errors()
        assert bool((out[d, c.NUM] == 1).all()), "every fan-in admits exactly one agent"
        win = (out[d, c.ID0] - 1).long() - (d - k)     # which in-edge's head arrived
        assert int(win.min()) >= 0 and int(win.max()) < k
        winners[seed] = win
        freq = torch.bincount(win, minlength=k).double() / copies
        p = torch.tensor(weights, dtype=torch.float64) / sum(weights)
        sigma = torch.sqrt(p * (1 - p) / copies)
        assert bool(((freq - p).abs() < 4.5 * sigma).all()), f"pick frequencies {freq.tolist()} vs weights {p.tolist()}"
        x = x0.clone()
        core_port.core_step(x, ei, attr, 100.0, Nmax, u[0].cpu(), cc)
        assert torch.equal(out, x), "the dumped stream replays the same picks on the oracle"
    assert not torch.equal(winners[5], winners[6])
