"""Generate tests/golden/sim_*.npz by executing the UNMODIFIED reference simulator loop behind oracle/shims.

Run in the authoring container only:  python oracle/gen_golden_sim.py
TEST INFRASTRUCTURE ONLY. The vectors are committed; the GPU box never runs this (no /root/reference there).

What is recorded, per scenario: the graph tensors config_network built from the MATSim XML (also committed, as text,
inside the npz), the initial agent_features, and after every step the whole x and agent_features plus the step's
side outputs. Random draws are injected (D4): torch.rand_like -> the recorded core uniforms, torch.multinomial ->
agents_port.multinomial_rule(recorded choice uniforms); torch.argsort is made stable (D3).
"""
from __future__ import annotations

import contextlib
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import agents_port  # noqa: E402
import cases  # noqa: E402
import ref_loader  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


@contextlib.contextmanager
def injected(u_core=None, u_choice=None):
    orig = (torch.rand_like, torch.multinomial, torch.argsort)

    def rand_like(t, *a, **k):
        assert u_core is not None and t.shape == u_core.shape, (t.shape, None if u_core is None else u_core.shape)
        return u_core.clone()

    def argsort(t, *a, **k):
        k.setdefault("stable", True)
        return orig[2](t, *a, **k)

    torch.rand_like = rand_like
    if u_choice is not None:
        torch.multinomial = agents_port.multinomial_rule(u_choice)
    torch.argsort = argsort
    try:
        yield
    finally:
        torch.rand_like, torch.multinomial, torch.argsort = orig


def network_xml(links, nodes=None, cell=7.5):
    """links: list of (id, from, to, length, capacity, freespeed, permlanes)."""
    out = ["<network>"]
    if nodes:
        out.append("  <nodes>")
        out += [f'    <node id="{i}" x="{x}" y="{y}"/>' for i, x, y in nodes]
        out.append("  </nodes>")
    out.append(f'  <links effectivecellsize="{cell}">')
    out += [f'    <link id="{i}" from="{a}" to="{b}" length="{ln}" capacity="{cap}" freespeed="{v}" permlanes="{pl}"/>'
            for i, a, b, ln, cap, v, pl in links]
    out += ["  </links>", "</network>"]
    return "\n".join(out)


def grid_links(n, g, length=(60.0, 140.0)):
    """n x n grid, both directions, per-link random length / capacity / lanes; node ids 'r_c' (string-sorted)."""
    links, nodes, k = [], [], 0
    for r in range(n):
        for col in range(n):
            nodes.append((f"{r}_{col}", 100.0 * col, 100.0 * r))
    for r in range(n):
        for col in range(n):
            for dr, dc in ((0, 1), (1, 0), (0, -1), (-1, 0)):
                rr, cc = r + dr, col + dc
                if 0 <= rr < n and 0 <= cc < n:
                    ln = float(torch.empty(1).uniform_(*length, generator=g))
                    cap = float(torch.randint(600, 2400, (1,), generator=g))
                    v = float(torch.randint(8, 20, (1,), generator=g))
                    pl = int(torch.randint(1, 3, (1,), generator=g))
                    links.append((k, f"{r}_{col}", f"{rr}_{cc}", round(ln, 2), cap, v, pl))
                    k += 1
    return links, nodes


def random_population(g, n_links, n_inter, A, t0, spread):
    af = torch.zeros(A + 1, 9)
    af[0, 2] = 48 * 3600.0
    o = torch.randint(0, n_inter, (A,), generator=g)
    d = (o + torch.randint(1, n_inter, (A,), generator=g)) % n_inter
    af[1:, 0] = (n_links + 2 * o).float()
    af[1:, 1] = (n_links + 2 * d + 1).float()
    af[1:, 2] = (t0 + torch.randint(0, spread, (A,), generator=g)).float()
    af[1:, 4] = torch.randint(18, 80, (A,), generator=g).float()
    return af


def graph_arrays(graph):
    return {
        "g_x": graph.x.numpy().copy(), "g_edge_index": graph.edge_index.numpy(), "g_edge_attr": graph.edge_attr.numpy(),
        "g_edge_index_routes": graph.edge_index_routes.numpy(), "g_edge_attr_routes": graph.edge_attr_routes.numpy(),
        "g_num_roads": np.int64(graph.num_roads), "g_critical_number": graph.critical_number.numpy(),
        "g_congestion_constant": graph.congestion_constant.numpy(), "g_adj_matrix": graph.adj_matrix.numpy(),
        "g_src_adj": graph.src_adj.numpy(),
    }


def classical(name, xml, af, t0, steps, seed, metrics=False):
    """TransportationSimulator.run loop (src/transportation_simulator.py:294-351). metrics=True also records what the
    reference's own compute_node_metrics (:563-669) returns at the end, and the per-link road-optimality series that
    plot_road_optimality (:482-488) reduces from road_optimality_values."""
    ts = ref_loader.load("src.transportation_simulator")
    g = torch.Generator().manual_seed(seed)
    with tempfile.TemporaryDirectory() as tmp:
        with open(os.path.join(tmp, "network.xml"), "w") as f:
            f.write(xml)
        sim = ts.TransportationSimulator("cpu")
        sim.config_network(os.path.join(tmp, "network"))
    rec = {"xml": np.array(xml), "Nmax": np.int64(sim.Nmax), "t0": np.float64(t0), "af0": af.numpy().copy(), **graph_arrays(sim.graph)}
    sim.agent.agent_features = af.clone()
    sim.config_parameters(start_time=t0)
    sim.agent.set_time(sim.time)
    N, n_nodes = sim.graph.num_roads, sim.graph.x.size(0)
    E = sim.graph.edge_index_routes.size(1)
    nodes, _, _ = agents_port.choosers_and_neighbours(sim.graph.edge_index, N, n_nodes)
    keys = ("t", "u_core", "u_choice", "x", "af", "withdrawn", "pop", "has_pop", "delta_tt")
    steps_rec = {k: [] for k in keys}
    for s in range(steps):
        u_core, u_choice = cases.uniforms(g, E), torch.rand(nodes.numel(), generator=g)
        n_hist = len(sim.model_core.response_mpnn.update_history)
        t = sim.time
        with injected(u_core, u_choice):
            sim.run()
        hist = sim.model_core.response_mpnn.update_history
        has_pop = len(hist) > n_hist
        steps_rec["t"].append(float(t)); steps_rec["u_core"].append(u_core.numpy()); steps_rec["u_choice"].append(u_choice.numpy())
        steps_rec["x"].append(sim.graph.x.numpy().copy()); steps_rec["af"].append(sim.agent.agent_features.numpy().copy())
        steps_rec["withdrawn"].append(sim.agent.withdraw_history[-1][1].numpy().copy())
        steps_rec["pop"].append(hist[-1][1].numpy().copy() if has_pop else np.zeros(N, dtype=bool))
        steps_rec["has_pop"].append(has_pop)
        steps_rec["delta_tt"].append(sim.model_core.direction_mpnn.road_optimality_data["delta_travel_time"].numpy().copy())
    rec.update({k: np.array(v) for k, v in steps_rec.items()})
    rec["mode"] = np.array("classical")
    if metrics:
        nm = sim.compute_node_metrics(output_dir=None)          # the unmodified reference method
        rec["nm_counts"] = np.array([nm[n]["hourly_counts"] for n in range(N)], dtype=np.int64)
        rec["nm_avg_vc"] = np.array([nm[n]["avg_vc"] for n in range(N)], dtype=np.float32)
        rec["nm_std_vc"] = np.array([nm[n]["std_vc"] for n in range(N)], dtype=np.float32)
        v_mat = torch.stack([v for _, v in sim.road_optimality_values], dim=0)                  # :483
        agg = torch.zeros(v_mat.size(0), N, dtype=v_mat.dtype)
        agg.scatter_add_(1, sim.graph.edge_index_routes[0].unsqueeze(0).expand(v_mat.size(0), -1), v_mat)   # :487-488
        rec["ro_agg"] = agg.numpy()
        rec["ro_times"] = np.array([t for t, _ in sim.road_optimality_values], dtype=np.float64)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **rec)
    done = int(sim.agent.agent_features[:, 8].sum())
    print(f"{name}: N={N} E={E} Nmax={sim.Nmax} agents={af.size(0) - 1} steps={steps} done={done} "
          f"handoff steps={int(np.sum(rec['has_pop']))} withdraw steps={int(rec['withdrawn'].any(axis=1).sum())}")


def rl_env(name, xml, af, steps, seed):
    """SimulatorEnv._reset + _step loop (src/reinforcement_learning.py:186-309) with random one-hot actions."""
    g = torch.Generator().manual_seed(seed)
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        os.makedirs(os.path.join(tmp, "data", "scn"))
        with open(os.path.join(tmp, "data", "scn", "network.xml"), "w") as f:
            f.write(xml)
        os.chdir(tmp)
        try:
            rl = ref_loader.load("src.reinforcement_learning")
            env = rl.SimulatorEnv(device="cpu", timestep_size=1, start_time=0, scenario="scn")
        finally:
            os.chdir(cwd)
    sim = env.simulator
    sim.agent.agent_features = af.clone()
    rec = {"xml": np.array(xml), "Nmax": np.int64(sim.Nmax), "af0": af.numpy().copy(), **graph_arrays(sim.graph)}
    td = env._reset()
    rec["t0"] = np.float64(sim.time)
    rec["x_reset"] = sim.graph.x.numpy().copy()
    ei = sim.graph.edge_index
    E, E_full = sim.graph.edge_index_routes.size(1), ei.size(1)
    srcs = torch.unique(ei[0])
    keys = ("t", "u_core", "action", "x", "af", "withdrawn", "pop", "has_pop", "delta_tt", "reward", "done", "obs_time")
    steps_rec = {k: [] for k in keys}
    for s in range(steps):
        u_core = cases.uniforms(g, E)
        action = torch.zeros(E_full, dtype=torch.bool)
        for v in srcs.tolist():                      # one out-edge per source node, uniformly
            out = torch.nonzero(ei[0] == v).flatten()
            action[out[int(torch.randint(0, out.numel(), (1,), generator=g))]] = True
        n_hist = len(sim.model_core.response_mpnn.update_history)
        t = sim.time
        with injected(u_core, None):
            out = env._step({"action": action})
        hist = sim.model_core.response_mpnn.update_history
        has_pop = len(hist) > n_hist
        N = sim.graph.num_roads
        steps_rec["t"].append(float(t)); steps_rec["u_core"].append(u_core.numpy()); steps_rec["action"].append(action.numpy())
        steps_rec["x"].append(sim.graph.x.numpy().copy()); steps_rec["af"].append(sim.agent.agent_features.numpy().copy())
        steps_rec["withdrawn"].append(sim.agent.withdraw_history[-1][1].numpy().copy())
        steps_rec["pop"].append(hist[-1][1].numpy().copy() if has_pop else np.zeros(N, dtype=bool))
        steps_rec["has_pop"].append(has_pop)
        steps_rec["delta_tt"].append(sim.model_core.direction_mpnn.road_optimality_data["delta_travel_time"].numpy().copy())
        steps_rec["reward"].append(float(out["reward"])); steps_rec["done"].append(bool(out["done"]))
        steps_rec["obs_time"].append(float(out["time"]))
    rec.update({k: np.array(v) for k, v in steps_rec.items()})
    rec["mode"] = np.array("rl")
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **rec)
    print(f"{name}: N={sim.graph.num_roads} E={E} E_full={E_full} steps={steps} done agents={int(sim.agent.agent_features[:, 8].sum())} "
          f"handoff steps={int(np.sum(rec['has_pop']))} last reward={rec['reward'][-1]}")


def population_case(name):
    """Agents.config_agents_from_xml (src/agents/base.py:38-242) on a small scenario with <nodes> and legacy x/y acts."""
    base = ref_loader.load("src.agents.base")
    g = torch.Generator().manual_seed(5)
    links, nodes = grid_links(3, g)
    net = network_xml(links, nodes)
    persons = []
    ids = [n[0] for n in nodes]
    for p in range(40):
        o, d, e = (ids[int(torch.randint(0, len(ids), (1,), generator=g))] for _ in range(3))
        car = "always" if p % 7 else "never"
        h1, m1 = 6 + p % 3, (7 * p) % 60
        acts = [f'<act type="h" link="{o}" end_time="{h1:02d}:{m1:02d}:00"/>', f'<act type="w" link="{d}" end_time="{h1 + 8:02d}:{m1:02d}"/>']
        if p % 4 == 0:
            acts.append(f'<act type="h" x="{100.0 * (p % 3) + 3}" y="{100.0 * ((p // 3) % 3) - 4}"/>')
        elif p % 4 == 1:
            acts.append(f'<act type="s" link="{e}"/>')
        sex = "f" if p % 2 else "m"
        persons.append(f'<person id="{p}" sex="{sex}" age="{20 + p}" car_avail="{car}" employed="{"yes" if p % 3 else "no"}">'
                       f'<plan>{"".join(acts)}</plan></person>')
    pop = "<population>" + "".join(persons) + "</population>"
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        os.makedirs(os.path.join(tmp, "data", "scn"))
        open(os.path.join(tmp, "data", "scn", "network.xml"), "w").write(net)
        open(os.path.join(tmp, "data", "scn", "population.xml"), "w").write(pop)
        os.chdir(tmp)
        try:
            a = base.Agents("cpu")
            a.config_agents_from_xml("scn", verbose=False)
        finally:
            os.chdir(cwd)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), network_xml=np.array(net), population_xml=np.array(pop),
                        agent_features=a.agent_features.numpy())
    print(f"{name}: agent_features {tuple(a.agent_features.shape)}")


def main():
    os.makedirs(OUT, exist_ok=True)
    only = set(sys.argv[1:])
    if only:                                  # e.g. `python oracle/gen_golden_sim.py sim_metrics_grid4`
        global classical, rl_env, population_case
        wrap = lambda f: (lambda name, *a, **k: f(name, *a, **k) if name in only else None)
        classical, rl_env, population_case = wrap(classical), wrap(rl_env), wrap(population_case)
    # the reference's own test network (tests/conftest.py:94-120): 2 links A<->B, one agent SRC(A) -> DEST(B)
    two = network_xml([(0, "A", "B", 100, 10, 10, 1), (1, "B", "A", 100, 10, 10, 1)])
    af = torch.zeros(2, 9); af[0, 2] = 25 * 3600.0; af[1, 0] = 2; af[1, 1] = 5
    classical("sim_twolink", two, af, t0=1, steps=14, seed=1)
    # 3-link chain A->B->C->D, 3 agents SRC(A) -> DEST(D) at t=0 (SURVEY.md §8c vector 3)
    chain = network_xml([(0, "A", "B", 100, 1800, 20, 1), (1, "B", "C", 100, 1800, 20, 1), (2, "C", "D", 100, 1800, 20, 1)])
    af = torch.zeros(4, 9); af[0, 2] = 25 * 3600.0; af[1:, 0] = 3; af[1:, 1] = 10
    classical("sim_chain3", chain, af, t0=0, steps=20, seed=2)
    # 4x4 grid, heterogeneous links, 400 agents over the first 60 s, random routing (classical loop)
    g = torch.Generator().manual_seed(3)
    links, nodes = grid_links(4, g)
    xml = network_xml(links, nodes)
    classical("sim_grid4", xml, random_population(g, len(links), 16, 400, 0, 60), t0=0, steps=160, seed=3)
    # short links -> small Nmax: queues fill, capacity clamps and the gridlock branch engage
    g = torch.Generator().manual_seed(4)
    links, nodes = grid_links(3, g, length=(20.0, 45.0))
    classical("sim_grid3_jam", network_xml(links, nodes), random_population(g, len(links), 9, 600, 0, 30), t0=0, steps=120, seed=4)
    # RL order (action, core, withdraw, insert) from _reset at 06:00-60s
    g = torch.Generator().manual_seed(6)
    links, nodes = grid_links(3, g)
    rl_env("sim_rl_grid3", network_xml(links, nodes), random_population(g, len(links), 9, 250, 21540, 40), steps=120, seed=6)
    population_case("sim_population_xml")
    # metrics side channels: the classical loop across an hour boundary (t = 3540 .. 3699), then the reference's own
    # compute_node_metrics and the road-optimality reduction
    g = torch.Generator().manual_seed(8)
    links, nodes = grid_links(4, g)
    classical("sim_metrics_grid4", network_xml(links, nodes), random_population(g, len(links), 16, 500, 3540, 80), t0=3540,
              steps=160, seed=8, metrics=True)


if __name__ == "__main__":
    main()
