"""Generate tests/golden/*.npz by executing the UNMODIFIED reference (/root/reference/src) behind oracle/shims.

Run in the authoring container only:  python oracle/gen_golden.py
TEST INFRASTRUCTURE ONLY. The vectors are committed; the GPU box never runs this (no /root/reference there).
"""
from __future__ import annotations

import contextlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import cases  # noqa: E402
import ref_loader  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


@contextlib.contextmanager
def injected_rand_like(u: torch.Tensor):
    """Make the reference's `torch.rand_like` (src/direction_mpnn.py:137) return the injected uniforms."""
    orig = torch.rand_like

    def fake(t, *a, **k):
        assert t.shape == u.shape, (t.shape, u.shape)
        return u.clone()

    torch.rand_like = fake
    try:
        yield
    finally:
        torch.rand_like = orig


def reference_core_trajectory(x0, ei, w, Nmax, t0, steps, g, use_static):
    """Drive the reference's SimulationCoreModel for `steps` steps; SELECTED_ROAD re-drawn before every step."""
    core_mod = ref_loader.load("src.simulation_core_model")
    from torch_geometric.data import Data
    x = x0.clone()
    N = x.size(0)
    graph = Data(x=x, edge_index_routes=ei, edge_attr_routes=w, num_roads=N)
    c = cases.Cols(Nmax)
    if use_static:
        crit = x[:, c.MAX_FLOW] * x[:, c.FFTT] / 3600
        graph.critical_number = crit
        graph.congestion_constant = x[:, c.FFTT] * (x[:, c.MAXN] + 10 - crit)
    model = core_mod.SimulationCoreModel(Nmax=Nmax, device="cpu", time=t0)
    rec = {k: [] for k in ("t", "sel", "u", "x", "delta_tt", "pop", "has_pop")}
    t = t0
    for s in range(steps):
        sel = cases.random_selection(g, N, ei) if s else x[:, c.SEL].clone()
        graph.x[:, c.SEL] = sel
        u = cases.uniforms(g, ei.size(1))
        model.set_time(t)
        n_hist = len(model.response_mpnn.update_history)
        with injected_rand_like(u):
            graph = model(graph)
        hist = model.response_mpnn.update_history
        has_pop = len(hist) > n_hist
        pop = hist[-1][1].clone() if has_pop else torch.zeros(N, dtype=torch.bool)
        rec["t"].append(float(t)); rec["sel"].append(sel.clone()); rec["u"].append(u)
        rec["x"].append(graph.x.clone())
        rec["delta_tt"].append(model.direction_mpnn.road_optimality_data["delta_travel_time"].clone())
        rec["pop"].append(pop); rec["has_pop"].append(has_pop)
        t += 1
    return rec


def save_core_case(name, x0, ei, w, Nmax, rec, use_static):
    np.savez_compressed(
        os.path.join(OUT, name + ".npz"),
        x0=x0.numpy(), edge_index=ei.numpy(), edge_attr=w.numpy(), Nmax=np.int64(Nmax),
        use_static=np.bool_(use_static), t=np.array(rec["t"], dtype=np.float64),
        sel=torch.stack(rec["sel"]).numpy(), u=torch.stack(rec["u"]).numpy() if ei.size(1) else np.zeros((len(rec["t"]), 0), np.float32),
        x=torch.stack(rec["x"]).numpy(), delta_tt=torch.stack(rec["delta_tt"]).numpy(),
        pop=torch.stack(rec["pop"]).numpy(), has_pop=np.array(rec["has_pop"]))


def gen_braess():
    """The reference's own `braess_graph` fixture (tests/conftest.py:45-91), re-built value by value."""
    Nmax = 100
    c = cases.Cols(Nmax)
    x = torch.zeros(3, c.F)
    for i, (maxn, num, fftt, sel) in enumerate([(2, 1, 3.0, 1), (2, 1, 1.0, 2), (2, 2, 1.0, 0)]):
        x[i, c.MAXN], x[i, c.NUM], x[i, c.FFTT] = maxn, num, fftt
        x[i, c.LENGTH], x[i, c.MAX_FLOW], x[i, c.SEL], x[i, c.RIDX] = 100.0, 10.0, sel, i
    x[0, 0], x[1, 0], x[2, 0], x[2, 1], x[2, 2 * Nmax + 1] = 1.0, 2.0, 3.0, 4.0, 1.0
    ei = torch.tensor([[0, 1, 2], [1, 2, 0]])
    g = torch.Generator().manual_seed(7)
    w = torch.rand(3, 1, generator=g)
    rec = reference_core_trajectory(x, ei, w, Nmax, 0, 3, g, use_static=False)
    save_core_case("core_braess", x, ei, w, Nmax, rec, False)


def gen_random(name, seed, N, Nmax, steps, max_out=4, sort_by_source=True, use_static=True, integral=False,
               garbage=True):
    g = torch.Generator().manual_seed(seed)
    ei, w = cases.random_dual_graph(g, N, max_out, sort_by_source)
    t0 = 100.0
    x0, _ = cases.random_road_state(g, N, Nmax, t0, ei, garbage=garbage, integral_times=integral)
    rec = reference_core_trajectory(x0, ei, w, Nmax, t0, steps, g, use_static)
    save_core_case(name, x0, ei, w, Nmax, rec, use_static)
    pops = int(torch.stack(rec["pop"]).sum())
    print(f"{name}: N={N} E={ei.size(1)} Nmax={Nmax} steps={steps} pops={pops}")


def gen_grid(name, seed, n, Nmax, steps):
    g = torch.Generator().manual_seed(seed)
    ei, w, _, _ = cases.grid_dual_graph(n)
    N = int(ei.max()) + 1
    x0, _ = cases.random_road_state(g, N, Nmax, 100.0, ei, garbage=False, integral_times=True, jam_fraction=0.05)
    rec = reference_core_trajectory(x0, ei, w, Nmax, 100.0, steps, g, True)
    save_core_case(name, x0, ei, w, Nmax, rec, True)
    print(f"{name}: N={N} E={ei.size(1)} pops={int(torch.stack(rec['pop']).sum())}")


def main():
    os.makedirs(OUT, exist_ok=True)
    gen_braess()
    gen_random("core_rand_a", 1, N=64, Nmax=15, steps=12)
    gen_random("core_rand_b", 2, N=97, Nmax=9, steps=10, sort_by_source=False, use_static=False)
    gen_random("core_rand_c", 3, N=33, Nmax=40, steps=8, max_out=6, integral=True)
    gen_random("core_rand_d", 4, N=1, Nmax=6, steps=3, max_out=0)              # single link, no dual edges
    gen_random("core_rand_e", 5, N=40, Nmax=15, steps=6, garbage=False, integral=True)
    gen_grid("core_grid6", 6, n=6, Nmax=15, steps=15)
    import gen_golden_mpnn
    gen_golden_mpnn.main(OUT)


if __name__ == "__main__":
    main()
