"""GPU parity of the simulator loop around the core step, through the C ABI (tarl_agents_insert / _withdraw / _choice /
_apply_action, tarl_core_step, tarl_store_step, tarl_store_observe): after EVERY step the whole node table x and the
whole agent_features must equal, bit for bit, what the unmodified reference produced (tests/golden/sim_*.npz) or what
the CPU oracle port produces on seeded inputs — for the in-place drop-in classes (TransportationSimulator,
SimulatorEnv) and for the batched link-store environment (BatchedSimulatorEnv)."""
import glob
import os

import numpy as np
import pytest
import torch

import agents_port
import cases
from core_port import Cols

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
CLASSICAL = ["sim_twolink", "sim_chain3", "sim_grid4", "sim_grid3_jam"]
RL = ["sim_rl_grid3"]


def golden(name):
    return np.load(os.path.join(HERE, "golden", name + ".npz"))


def product_simulator(d, tmp_path):
    from tarl_simulator_b200.transportation_simulator import TransportationSimulator
    (tmp_path / "network.xml").write_text(str(d["xml"]))
    sim = TransportationSimulator("cuda")
    sim.config_network(str(tmp_path / "network"))
    sim.agent.agent_features = torch.from_numpy(d["af0"]).cuda()
    return sim


def as_time(v):
    v = float(v)
    return int(v) if v == int(v) else v


@pytest.mark.parametrize("name", CLASSICAL)
def test_transportation_simulator_run_matches_reference(name, tmp_path):
    d = golden(name)
    sim = product_simulator(d, tmp_path)
    assert torch.equal(sim.graph.x.cpu(), torch.from_numpy(d["g_x"]))
    sim.config_parameters(start_time=as_time(d["t0"]))
    sim.agent.set_time(sim.time)
    hist = sim.model_core.response_mpnn.update_history
    for s in range(len(d["t"])):
        n_hist = len(hist)
        sim.run(noise=torch.from_numpy(d["u_core"][s]).cuda(), choice_uniforms=torch.from_numpy(d["u_choice"][s]).cuda())
        assert torch.equal(sim.graph.x.cpu(), torch.from_numpy(d["x"][s])), f"x differs after step {s}"
        assert torch.equal(sim.agent.agent_features.cpu(), torch.from_numpy(d["af"][s])), f"agents differ after step {s}"
        assert torch.equal(sim.agent.withdraw_history[-1][1].cpu(), torch.from_numpy(d["withdrawn"][s]))
        assert sim.agent.withdraw_history[-1][0] == as_time(d["t"][s])
        assert torch.equal(sim.road_optimality_values[-1][1], torch.from_numpy(d["delta_tt"][s]))
        assert (len(hist) > n_hist) == bool(d["has_pop"][s])
        if d["has_pop"][s]:
            assert torch.equal(hist[-1][1].cpu(), torch.from_numpy(d["pop"][s]))
    assert sim.time == as_time(d["t"][-1]) + 1
    sim.agent.check_errors()
    sim.model_core.check_errors()


@pytest.mark.parametrize("name", RL)
def test_simulator_env_matches_reference(name, tmp_path):
    from tarl_simulator_b200.reinforcement_learning import SimulatorEnv
    d = golden(name)
    sim = product_simulator(d, tmp_path)
    env = SimulatorEnv(device="cuda", simulator=sim)
    obs = env.reset()
    assert float(obs["time"]) == float(d["t0"]) == 21540.0
    assert torch.equal(sim.graph.x.cpu(), torch.from_numpy(d["x_reset"]))
    assert obs["node_features"].shape == (sim.graph.x.size(0), 7) and obs["agent_index"].dtype == torch.int64
    for s in range(len(d["t"])):
        env.noise = torch.from_numpy(d["u_core"][s]).cuda()
        out = env._step({"action": torch.from_numpy(d["action"][s]).cuda()})
        assert torch.equal(sim.graph.x.cpu(), torch.from_numpy(d["x"][s])), f"x differs after step {s}"
        assert torch.equal(sim.agent.agent_features.cpu(), torch.from_numpy(d["af"][s])), f"agents differ after step {s}"
        assert float(out["reward"]) == float(d["reward"][s])
        assert bool(out["done"]) == bool(d["done"][s])
        assert float(out["time"]) == float(d["obs_time"][s])
        assert torch.equal(out["agent_index"].cpu(), torch.from_numpy(d["x"][s][:, 0]).long())
    sim.agent.check_errors()


def batched_env(d, sim, R, cluster=None):
    from tarl_simulator_b200.reinforcement_learning import BatchedSimulatorEnv
    return BatchedSimulatorEnv(sim.graph, int(d["Nmax"]), torch.from_numpy(d["af0"]), replicas=R, cluster=cluster)


@pytest.mark.parametrize("cluster", [False, True])
@pytest.mark.parametrize("name", RL)
def test_batched_env_matches_reference(name, tmp_path, cluster):
    d = golden(name)
    sim = product_simulator(d, tmp_path)
    R = 3
    env = batched_env(d, sim, R, cluster)
    env.reset()
    assert torch.equal(env.export_x()[1].cpu(), torch.from_numpy(d["x_reset"]))
    E = sim.graph.edge_index_routes.size(1)
    env.delta_tt = torch.empty(R, E, device="cuda")
    for s in range(len(d["t"])):
        u = torch.from_numpy(d["u_core"][s]).cuda().repeat(R, 1)
        a = torch.from_numpy(d["action"][s]).cuda().repeat(R, 1)
        # (cluster=False also takes the one-kernel insertion: every SRC node's SELECTED_ROAD comes from the action)
        out = env.step(a, noise=u, observe=True, direct_insert=not cluster)
        x = env.export_x().cpu()
        for r in range(R):
            assert torch.equal(x[r], torch.from_numpy(d["x"][s])), f"replica {r}: x differs after step {s}"
            assert torch.equal(env.agent_features[r].cpu(), torch.from_numpy(d["af"][s])), f"replica {r}, step {s}"
        assert out["reward"].tolist() == [float(d["reward"][s])] * R
        assert torch.equal(env.withdrawn[0].cpu(), torch.from_numpy(d["withdrawn"][s]))
        assert torch.equal(env.delta_tt[2].cpu(), torch.from_numpy(d["delta_tt"][s]))
        assert torch.equal(out["node_features"][0].cpu(), torch.from_numpy(d["x"][s][:, -7:]))
        assert torch.equal(out["agent_index"][1].cpu(), torch.from_numpy(d["x"][s][:, 0]).long())
        assert out["time"] == float(d["obs_time"][s])
    env.check_errors()
    done = int(d["af"][-1][:, 8].sum())
    assert env.counters[:, 1].tolist() == [done] * R


@pytest.mark.parametrize("cluster", [False, True])
@pytest.mark.parametrize("name", CLASSICAL)
def test_batched_store_classical_order_matches_reference(name, tmp_path, cluster):
    """insert -> withdraw -> choice -> core on the link store (the order of TransportationSimulator.run)."""
    d = golden(name)
    sim = product_simulator(d, tmp_path)
    R = 2
    env = batched_env(d, sim, R, cluster)
    for s in range(len(d["t"])):
        env.set_time(float(d["t"][s]))
        env.insert()
        env.withdraw()
        env.choice(uniforms=torch.from_numpy(d["u_choice"][s]).cuda().repeat(R, 1))
        env.store.step(env.time, noise=torch.from_numpy(d["u_core"][s]).cuda().repeat(R, 1))
        x = env.export_x().cpu()
        for r in range(R):
            assert torch.equal(x[r], torch.from_numpy(d["x"][s])), f"replica {r}: x differs after step {s}"
            assert torch.equal(env.agent_features[r].cpu(), torch.from_numpy(d["af"][s]))
        assert torch.equal(env.withdrawn[1].cpu(), torch.from_numpy(d["withdrawn"][s]))
    env.check_errors()


def synthetic_case(seed, n, A, spread, length):
    """n x n grid via the product's graph builder + a random population; dense adjacency for the oracle."""
    from tarl_simulator_b200.matsim_io import graph_from_links
    g = torch.Generator().manual_seed(seed)
    _, _, frm, to = cases.grid_dual_graph(n)
    L = frm.numel()
    ln = torch.empty(L).uniform_(*length, generator=g)
    graph, Nmax = graph_from_links([f"{int(v):04d}" for v in frm], [f"{int(v):04d}" for v in to], ln.tolist(),
                                   torch.randint(600, 2400, (L,), generator=g).float().tolist(),
                                   torch.randint(8, 20, (L,), generator=g).float().tolist(),
                                   torch.randint(1, 3, (L,), generator=g).float().tolist(), dense=True)
    af = torch.zeros(A + 1, 9)
    af[0, 2] = 48 * 3600.0
    o = torch.randint(0, n * n, (A,), generator=g)
    dd = (o + torch.randint(1, n * n, (A,), generator=g)) % (n * n)
    af[1:, 0], af[1:, 1] = (L + 2 * o).float(), (L + 2 * dd + 1).float()
    af[1:, 2] = torch.randint(0, spread, (A,), generator=g).float()
    return graph, Nmax, af, g


@pytest.mark.parametrize("seed,n,A,spread,length,steps", [(31, 10, 4000, 40, (60.0, 140.0), 90),
                                                          (32, 8, 6000, 10, (20.0, 40.0), 70)])
def test_inplace_and_store_match_oracle_on_random_grids(seed, n, A, spread, length, steps):
    from tarl_simulator_b200.reinforcement_learning import BatchedSimulatorEnv
    from tarl_simulator_b200.transportation_simulator import TransportationSimulator
    graph, Nmax, af, g = synthetic_case(seed, n, A, spread, length)
    c = Cols(Nmax)
    N = int(graph.num_roads)
    E = graph.edge_index_routes.size(1)
    ref_graph = {k: getattr(graph, k) for k in ("edge_index", "edge_index_routes", "edge_attr_routes", "adj_matrix",
                                                "congestion_constant")}
    ref_graph["num_roads"] = N
    x_ref, af_ref = graph.x.clone(), af.clone()
    sim = TransportationSimulator("cuda")
    sim.graph, sim.Nmax = graph.to("cuda"), Nmax
    from tarl_simulator_b200.feature_helpers import FeatureHelpers
    sim.h = FeatureHelpers(Nmax)
    sim.agent.agent_features = af.cuda()
    sim.config_parameters(start_time=0)
    sim.agent.set_time(0)
    sim.record_road_optimality = False
    env = BatchedSimulatorEnv(sim.graph, Nmax, af, replicas=2, cluster=bool(seed % 2))
    nodes, _, _ = agents_port.choosers_and_neighbours(graph.edge_index.cpu(), N, x_ref.size(0))
    moved = 0
    for s in range(steps):
        u_core, u_choice = cases.uniforms(g, E), torch.rand(nodes.numel(), generator=g)
        out = agents_port.run_step(x_ref, af_ref, s, c, ref_graph, u_choice, u_core)
        moved += 0 if out["pop"] is None else int(out["pop"].sum())
        sim.run(noise=u_core.cuda(), choice_uniforms=u_choice.cuda())
        assert torch.equal(sim.graph.x.cpu(), x_ref), f"in-place x differs after step {s}"
        assert torch.equal(sim.agent.agent_features.cpu(), af_ref), f"in-place agents differ after step {s}"
        env.set_time(float(s))
        env.insert(); env.withdraw(); env.choice(uniforms=u_choice.cuda().repeat(2, 1))
        env.store.step(env.time, noise=u_core.cuda().repeat(2, 1))
        if s % 10 == 9 or s == steps - 1:
            assert torch.equal(env.export_x()[1].cpu(), x_ref), f"store x differs after step {s}"
            assert torch.equal(env.agent_features[0].cpu(), af_ref), f"store agents differ after step {s}"
    assert moved > 100 and int(af_ref[:, 8].sum()) > 0
    sim.agent.check_errors(); env.check_errors()


def test_reference_agent_tests_on_gpu():
    """The reference's own tests/agents_test.py (insert_and_withdraw, insert_capacity_limit), on the device."""
    from tarl_simulator_b200.agents import Agents
    from tarl_simulator_b200.data import Data
    from tarl_simulator_b200.feature_helpers import FeatureHelpers
    h = FeatureHelpers(Nmax=5)

    def graph():
        x = torch.zeros((2, 3 * h.Nmax + 7))
        x[0, h.MAX_NUMBER_OF_AGENT] = 5
        x[0, h.ROAD_INDEX] = 0
        x[0, h.FREE_FLOW_TIME_TRAVEL] = 10
        edge_index = torch.tensor([[1, 0], [0, 0]])
        return Data(x=x.cuda(), edge_index=edge_index.cuda(), edge_index_routes=torch.empty((2, 0), dtype=torch.long).cuda(),
                    edge_attr_routes=torch.empty((0, 1)).cuda(), num_roads=1)

    agents = Agents("cuda")
    agents.agent_features = torch.tensor([[1.0, 0.0, 0.0, 0.0, 30.0, 0.0, 1.0, 0.0, 0.0],
                                          [1.0, 0.0, 0.0, 0.0, 25.0, 1.0, 0.0, 0.0, 0.0]]).cuda()
    g = graph()
    agents.time = 0
    g.x = agents.insert_agent_into_network(g, h)
    assert g.x[0, h.NUMBER_OF_AGENT] == 2
    assert torch.all(agents.agent_features[:2, agents.ON_WAY] == 1)
    g.x = agents.withdraw_agent_from_network(g, h)
    assert g.x[0, h.NUMBER_OF_AGENT] == 2
    agents.time = 10
    g.x = agents.withdraw_agent_from_network(g, h)
    assert g.x[0, h.NUMBER_OF_AGENT] == 0
    assert torch.all(agents.agent_features[:2, agents.DONE] == 1)
    assert len(agents.withdraw_history) == 2 and bool(agents.withdraw_history[1][1][0])

    agent = Agents("cuda")
    agent.agent_features = torch.tensor([[1.0, 0, 0, 0, 0, 0, 0, 0, 0]] * 4).cuda()
    g = graph()
    agent.time = 0
    g.x = agent.insert_agent_into_network(g, h)
    assert g.x[0, h.NUMBER_OF_AGENT] == 2
    assert torch.all(agent.agent_features[:2, agent.ON_WAY] == 1)
    assert torch.all(agent.agent_features[2:, agent.ON_WAY] == 0)


def test_insert_merges_origins_that_select_the_same_road():
    """Two SRC nodes point at one road: the admitted agents are the smallest ready ids of the union (base.py:275-291
    with a stable sort), whatever order the origin lists were pushed in."""
    from tarl_simulator_b200.agents import Agents
    from tarl_simulator_b200.data import Data
    from tarl_simulator_b200.feature_helpers import FeatureHelpers
    h = FeatureHelpers(Nmax=12)
    c = Cols(12)
    x = torch.zeros(5, c.F)
    x[0, c.MAXN], x[0, c.FFTT] = 11, 4.0            # road 0: room 8
    x[1, c.MAXN], x[1, c.FFTT], x[1, c.RIDX] = 11, 6.0, 1
    x[2:, c.RIDX] = -1
    x[2, c.SEL], x[3, c.SEL], x[4, c.SEL] = 0, 0, 1  # nodes 2 and 3 both select road 0
    g = torch.Generator().manual_seed(3)
    A = 40
    af = torch.zeros(A + 1, 9)
    af[0, 2] = 1e6
    af[1:, 0] = torch.randint(2, 5, (A,), generator=g).float()
    af[1:, 2] = torch.randint(0, 3, (A,), generator=g).float()
    af[1:, 1] = 1.0
    x_ref, af_ref = x.clone(), af.clone()
    graph = Data(x=x.cuda(), edge_index=torch.tensor([[2, 3, 4], [0, 0, 1]]).cuda(),
                 edge_index_routes=torch.empty((2, 0), dtype=torch.long).cuda(), edge_attr_routes=torch.empty((0, 1)).cuda(),
                 num_roads=2)
    agents = Agents("cuda")
    agents.agent_features = af.cuda()
    for t in range(4):
        agents_port.insert(x_ref, af_ref, t, c, None)
        agents.time = t
        agents.insert_agent_into_network(graph, h)
        assert torch.equal(graph.x.cpu(), x_ref) and torch.equal(agents.agent_features.cpu(), af_ref)
        x_ref[:2, c.NUM] = 0
        graph.x[:2, c.NUM] = 0                      # drain so that later steps admit again
    assert int(af_ref[:, 7].sum()) >= 24
    agents.check_errors()


def test_cpu_tensors_are_refused():
    from tarl_simulator_b200.agents import Agents
    from tarl_simulator_b200.data import Data
    from tarl_simulator_b200.feature_helpers import FeatureHelpers
    h = FeatureHelpers(Nmax=5)
    a = Agents("cpu")
    a.agent_features = torch.zeros(2, 9)
    g = Data(x=torch.zeros(2, 22), edge_index=torch.zeros(2, 1, dtype=torch.long), num_roads=1,
             edge_index_routes=torch.zeros(2, 0, dtype=torch.long))
    with pytest.raises(RuntimeError):
        a.insert_agent_into_network(g, h)


def _tiny_graph(Nmax=2, n_links=3):
    """A ring of `n_links` single-slot links (Nmax = 2: one ring slot per link), SRC/DEST nodes per intersection."""
    from tarl_simulator_b200.matsim_io import graph_from_links
    ids = [f"{i}" for i in range(n_links)]
    g, nm = graph_from_links(ids, ids[1:] + ids[:1], [5.0] * n_links, [1800.0] * n_links, [10.0] * n_links, [1.0] * n_links,
                             dense=True)
    assert nm == Nmax
    return g, nm


def test_minimal_queue_capacity_and_empty_population():
    """Nmax = 2 (MAXN = 1, so MAXN-3-NUM < 0: nothing is ever admitted) and a population with only the dummy agent:
    every kernel must be a clean no-op on both layouts, and the exported state must stay equal to the oracle's."""
    from tarl_simulator_b200.reinforcement_learning import BatchedSimulatorEnv
    from tarl_simulator_b200.transportation_simulator import TransportationSimulator
    from tarl_simulator_b200.feature_helpers import FeatureHelpers
    graph, Nmax = _tiny_graph()
    c = Cols(Nmax)
    N = int(graph.num_roads)
    af = torch.zeros(1, 9); af[0, 2] = 48 * 3600.0
    ref_graph = {k: getattr(graph, k) for k in ("edge_index", "edge_index_routes", "edge_attr_routes", "adj_matrix",
                                                "congestion_constant")}
    ref_graph["num_roads"] = N
    x_ref, af_ref = graph.x.clone(), af.clone()
    sim = TransportationSimulator("cuda")
    sim.graph, sim.Nmax, sim.h = graph.to("cuda"), Nmax, FeatureHelpers(Nmax)
    sim.agent.agent_features = af.cuda()
    sim.config_parameters(start_time=0); sim.agent.set_time(0)
    env = BatchedSimulatorEnv(sim.graph, Nmax, af, replicas=2)
    g = torch.Generator().manual_seed(0)
    nodes, _, _ = agents_port.choosers_and_neighbours(ref_graph["edge_index"], N, x_ref.size(0))
    E = ref_graph["edge_index_routes"].size(1)
    for s in range(5):
        u_core, u_choice = cases.uniforms(g, E), torch.rand(nodes.numel(), generator=g)
        agents_port.run_step(x_ref, af_ref, s, c, ref_graph, u_choice, u_core)
        sim.run(noise=u_core.cuda(), choice_uniforms=u_choice.cuda())
        env.set_time(float(s)); env.insert(); env.withdraw(); env.choice(uniforms=u_choice.cuda().repeat(2, 1))
        env.store.step(env.time, noise=u_core.cuda().repeat(2, 1))
        assert torch.equal(sim.graph.x.cpu(), x_ref)
        assert torch.equal(env.export_x()[1].cpu(), x_ref)
    assert float(x_ref[:, c.NUM].sum()) == 0.0
    sim.agent.check_errors(); env.check_errors()


def test_data_dependent_faults_are_flagged():
    """An origin whose SELECTED_ROAD is not a road while one of its agents is ready, and a queued agent id outside
    agent_features: the reference raises IndexError; here the sticky error word does."""
    from tarl_simulator_b200.agents import Agents
    from tarl_simulator_b200.data import Data
    from tarl_simulator_b200.feature_helpers import FeatureHelpers
    h = FeatureHelpers(Nmax=5)
    c = Cols(5)
    x = torch.zeros(3, c.F)
    x[0, c.MAXN], x[0, c.FFTT] = 5, 3.0
    x[1:, c.RIDX] = -1
    x[1, c.SEL] = 7                                   # the SRC node points outside the roads
    g = Data(x=x.cuda(), edge_index=torch.tensor([[1, 0], [0, 2]]).cuda(),
             edge_index_routes=torch.empty((2, 0), dtype=torch.long).cuda(), edge_attr_routes=torch.empty((0, 1)).cuda(),
             num_roads=1)
    a = Agents("cuda")
    a.agent_features = torch.tensor([[0, 0, 1e6, 0, 0, 0, 0, 0, 0], [1.0, 2, 0, 0, 0, 0, 0, 0, 0]]).cuda()
    a.time = 0
    a.insert_agent_into_network(g, h)
    with pytest.raises(IndexError):
        a.check_errors()
    g.x[0, 0], g.x[0, c.NUM], g.x[0, c.DEP0] = 99.0, 1.0, 0.0      # agent id 99 does not exist
    a.withdraw_agent_from_network(g, h)
    with pytest.raises(IndexError):
        a.check_errors()
    assert float(g.x[0, c.NUM]) == 1.0                              # nothing was removed


def test_full_size_environment_invariants():
    """BASELINE.json configs[3] at full size (999 000 links, 2 000 000 agents), where no oracle finishes: size-independent
    properties of the whole environment step on the link store — every agent is in at most one queue slot and every
    queued agent is flagged ON_WAY, the reward is minus the sum of the queue counters, the insertion / withdrawal
    counters agree with the agent table (inserted = on the way + arrived), NUM never leaves [0, MAXN], arrivals only
    ever increase, and two replicas fed the same noise stay bit-identical.

    NOT an invariant, by the reference's own behaviour: "queued agents == agents on the way". The response phase
    (src/response_mpnn.py:66-83) is evaluated on the post-direction state of ALL links at once, so an agent handed
    into an EMPTY link d is, for that one evaluation, the head of d and still the tail of the link u it came from; if
    d has a turn back into u (config_network creates U-turn edges, src/transportation_simulator.py:160) d pops it as
    well and the agent is in no queue any more while its ON_WAY flag stays set. The reference's golden trajectories
    show it (tests/golden/sim_rl_grid3: agent 54 at t = 21546) and the kernels reproduce it bit for bit; here it only
    means `queued <= on the way`."""
    from tarl_simulator_b200 import synthetic
    from tarl_simulator_b200.reinforcement_learning import BatchedSimulatorEnv
    frm, to, n_nodes = synthetic.ring_radial_links(500, 500, device="cuda")
    frm, to = synthetic.reorder_links(frm, to, "node")
    g, Nmax = synthetic.build_graph(frm, to, n_nodes)
    N = int(g.num_roads)
    assert N == 999_000 and Nmax == 15
    af = synthetic.population(g, 2_000_000, 21540, 60, seed=11)
    R = 2
    env = BatchedSimulatorEnv(g, Nmax, af, replicas=R, seed=3)
    env.reset()
    E = g.edge_index_routes.size(1)
    gen = torch.Generator(device="cuda").manual_seed(9)
    done_before = torch.zeros(R, device="cuda")
    for s in range(40):
        if s % 2 == 0:                                            # the same routing draw in both replicas
            env.choice(uniforms=torch.rand(env.side.n_choosers, device="cuda", generator=gen).repeat(R, 1))
        u = torch.rand(E, device="cuda", generator=gen).clamp_(min=1e-7).repeat(R, 1)
        out = env.step(None, noise=u)
        if s % 8 != 7:
            continue
        x = env.export_x()
        num = x[:, :N, 3 * Nmax + 1]
        maxn = x[:, :N, 3 * Nmax]
        assert bool((num >= 0).all()) and bool((num <= maxn).all())
        on_way = env.agent_features[..., 7].sum(1)
        done = env.agent_features[..., 8].sum(1)
        assert bool((num.sum(1) <= on_way).all())                                # see the docstring
        assert torch.equal(-out["reward"], num.sum(1))                           # reward = -occupancy
        assert torch.equal(env.counters[:, 0].float(), on_way + done)            # inserted = on the way + arrived
        assert torch.equal(env.counters[:, 1].float(), done)
        assert bool((done >= done_before).all())
        done_before = done
        ids = x[0, :N, :Nmax]
        live = torch.arange(Nmax, device="cuda").unsqueeze(0) < num[0].unsqueeze(1)
        q = ids[live].long()
        assert q.numel() == int(num[0].sum()) and q.numel() == torch.unique(q).numel() and int(q.min()) >= 1
        assert bool((env.agent_features[0, q, 7] == 1).all())                    # every queued agent is flagged ON_WAY
        assert torch.equal(x[0], x[1]) and torch.equal(env.agent_features[0], env.agent_features[1])
    assert float(on_way.min()) > 100_000 and float(num.sum(1).min()) > 100_000      # the network did fill up
    env.check_errors()


@pytest.mark.parametrize("inject_noise", [False, True])
def test_fused_step_withdraw_observe_equals_the_separate_passes(inject_noise):
    """The occupancy-only environment step (core step whose response phase also withdraws and leaves NUM / the reward
    behind, insertion patching them: tarl_store_step_withdraw) against the step with separate passes on a twin
    environment: whole exported state, agent table, withdrawn masks, counters, occupancy and the NUM frame after every
    one of 260 steps of short trips on a 6 x 6 grid (hundreds of withdrawals, queues that fill up)."""
    from tarl_simulator_b200 import synthetic
    from tarl_simulator_b200.reinforcement_learning import BatchedSimulatorEnv
    dev = torch.device("cuda")
    frm, to, n_nodes = synthetic.grid_links(6, device=dev)
    g, Nmax = synthetic.build_graph(frm, to, n_nodes)
    af = synthetic.population(g, 500, 21540, 90, seed=2)
    R = 3
    a = BatchedSimulatorEnv(g, Nmax, af, replicas=R, seed=9)
    b = BatchedSimulatorEnv(g, Nmax, af, replicas=R, seed=9)
    assert a.store.can_fuse_withdraw() and a.n_nodes - a.N <= a.N
    a.reset(); b.reset()
    gen = torch.Generator(device="cuda").manual_seed(4)
    E = a.store.E
    frame = torch.empty(R, a.n_nodes, device=dev)
    F = 3 * Nmax + 7
    for s in range(260):
        a.choice(seed=100 + s); b.choice(seed=100 + s)            # the same routing decisions in both
        noise = torch.rand(R, E, device=dev, generator=gen).clamp_(min=1e-6) if inject_noise else None
        a.step(None, noise=noise, compact_out=(frame, None, None), lean=True)      # fused
        out = b.step(None, noise=noise, observe=True)                               # separate passes
        xa, xb = a.export_x(), b.export_x()
        assert torch.equal(xa, xb), s
        assert torch.equal(a.agent_features, b.agent_features), s
        assert torch.equal(a.withdrawn, b.withdrawn) and torch.equal(a.counters, b.counters), s
        assert torch.equal(a.occupancy, b.occupancy) and torch.equal(out["reward"], -a.occupancy.float()), s
        assert torch.equal(frame[:, : a.N], xb[:, : a.N, F - 6]) and float(frame[:, a.N:].abs().sum()) == 0.0, s
        assert torch.equal(frame, out["node_features"][..., 1]), s
    assert int(a.counters[:, 1].min()) > 10                        # withdrawals did happen in every replica
    a.check_errors(); b.check_errors()


def test_direct_insertion_equals_the_list_insertion(monkeypatch):
    """tarl_agents_insert with road_origin (offer + admit as ONE kernel, for networks in which a road can be selected by
    one origin only) against the two-kernel form with per-road lists on a twin environment: whole state, agent table,
    counters and occupancy after every one of 200 steps of short trips on a 6 x 6 grid (origins that insert several
    agents in one step, full roads that admit only some of them). A SELECTED_ROAD naming another origin's road raises."""
    from tarl_simulator_b200 import synthetic
    from tarl_simulator_b200.reinforcement_learning import BatchedSimulatorEnv
    dev = torch.device("cuda")
    frm, to, n_nodes = synthetic.grid_links(6, device=dev)
    g, Nmax = synthetic.build_graph(frm, to, n_nodes)
    af = synthetic.population(g, 900, 21540, 60, seed=5)          # 15 agents per second over 36 origins
    R = 2
    a = BatchedSimulatorEnv(g, Nmax, af, replicas=R, seed=3)
    monkeypatch.setenv("TARL_NO_DIRECT_INSERT", "1")
    b = BatchedSimulatorEnv(g, Nmax, af, replicas=R, seed=3)
    monkeypatch.delenv("TARL_NO_DIRECT_INSERT")
    assert a._road_origin is not None and b._road_origin is None
    a.reset(); b.reset()
    for s in range(200):
        a.choice(seed=50 + s); b.choice(seed=50 + s)
        oa = a.step(None, observe=False, direct_insert=True); ob = b.step(None, observe=False, direct_insert=True)
        assert torch.equal(a.export_x(), b.export_x()), s
        assert torch.equal(a.agent_features, b.agent_features), s
        assert torch.equal(a.counters, b.counters) and torch.equal(oa["reward"], ob["reward"]), s
    assert int(a.counters[:, 0].min()) > 500                      # insertions did happen
    a.check_errors(); b.check_errors()
    # every SRC node selects road 0: all but one origin name somebody else's road while they have ready agents
    c = BatchedSimulatorEnv(g, Nmax, synthetic.population(g, 900, 21540, 1, seed=6), replicas=1, seed=3)
    c.reset()
    c.src_sel.zero_()
    c.set_time(21541.0)
    c.insert(direct=True)
    with pytest.raises(RuntimeError):
        c.check_errors()
