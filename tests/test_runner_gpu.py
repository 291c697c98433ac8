"""End-to-end runs of the drop-in entry points on the device: main.py --algo random / mpnn / mpnn+ppo on a synthetic
MATSim scenario written to a temp directory (the configurations BASELINE.json lists as configs[0] and configs[1], on
a small network), and PPO over a BatchedSimulatorEnv."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture
def scenario(tmp_path, monkeypatch):
    from tarl_simulator_b200 import synthetic
    synthetic.write_scenario(str(tmp_path), "grid5", "grid", (5,), n_agents=300, t0=21540, spread=60, seed=1)
    monkeypatch.chdir(tmp_path)
    return "grid5"


def test_main_random_eval(scenario):
    from main import main
    r = main(["--algo", "random", "--scenario", scenario, "--mode", "eval", "--start-end-time", "21540", "86400",
              "--steps", "400"])
    assert r.simulator.time == 21540 + 400
    assert r.summary["arrived"] > 0 and r.summary["average_travel_time"] > 0
    assert os.path.exists(os.path.join("save", scenario, "network.pt"))
    assert os.path.exists(os.path.join("save", scenario, "population.pt"))


def test_main_mpnn_eval(scenario):
    from main import main
    r = main(["--algo", "mpnn", "--scenario", scenario, "--mode", "eval", "--steps", "10"])
    assert r.env.simulator.time == 21540 + 10           # _reset starts every rollout at 06:00 - 60 s
    assert len(r.env.simulator.leg_histogram_values) == 10


def test_main_mpnn_ppo_train(scenario, tmp_path):
    from main import main
    r = main(["--algo", "mpnn+ppo", "--scenario", scenario, "--mode", "train", "--epochs", "3", "--rollout-steps", "24",
              "--steps", "8", "--output-dir", str(tmp_path / "runs")])
    h = r.history
    assert len(h) == 1 and h[0]["frames"] == 24
    assert all(torch.isfinite(torch.tensor(h[0][k])) for k in ("loss_objective", "loss_critic", "loss_entropy",
                                                                "grad_global_norm", "eval_return"))
    assert h[0]["grad_global_norm"] > 0
    assert os.path.exists(tmp_path / "runs" / "policy.pt") and os.path.exists(tmp_path / "runs" / "ppo_log.jsonl")


def test_ppo_on_batched_env_updates_parameters(scenario):
    from tarl_simulator_b200.mpnn_agent import MPNNPolicyNet, MPNNValueNetSimple
    from tarl_simulator_b200.reinforcement_learning import BatchedSimulatorEnv
    from tarl_simulator_b200.rl.ppo_trainer import PolicyModule, ValueModule, ppo_train
    from tarl_simulator_b200.transportation_simulator import TransportationSimulator
    sim = TransportationSimulator("cuda")
    sim.load_network(scenario)
    sim.agent.load(scenario)
    g = sim.graph
    env = BatchedSimulatorEnv(g, sim.Nmax, sim.agent.agent_features, replicas=8, seed=3)
    policy = MPNNPolicyNet(g.edge_index, g.x.size(0), torch.ones(g.edge_index.size(1)), "cuda")
    value = MPNNValueNetSimple(g.edge_index, g.x.size(0), "cuda")
    before = policy.nodes_embedding.weight.detach().clone()
    hist = ppo_train(env, PolicyModule(policy, g.edge_index), ValueModule(value), total_frames=40, frames_per_batch=20,
                     num_epochs=2, sub_batch_size=64, history=[])
    assert len(hist) == 2 and hist[0]["frames"] == 160
    assert not torch.equal(before, policy.nodes_embedding.weight.detach())
    assert int(env.counters[:, 0].min()) > 0                   # every replica inserted agents
    env.check_errors()


def test_occupancy_only_rollout_equals_the_full_rollout(scenario):
    """collect(occupancy_only=True) — the trajectory of nets that read NUMBER_OF_AGENT only — must hold the same
    occupancies, actions, log-probabilities and rewards as the full trajectory (the sampling kernel applies the action
    and draws its own uniforms in both); a rollout without the action sink is a valid episode too."""
    from tarl_simulator_b200.mpnn_agent import MPNNPolicyNet, MPNNValueNetSimple
    from tarl_simulator_b200.reinforcement_learning import BatchedSimulatorEnv
    from tarl_simulator_b200.rl.ppo_trainer import PolicyModule, ValueModule, _EnvAdapter, collect, occupancy_only
    from tarl_simulator_b200.transportation_simulator import TransportationSimulator
    sim = TransportationSimulator("cuda")
    sim.load_network(scenario)
    sim.agent.load(scenario)
    g = sim.graph
    torch.manual_seed(5)
    policy = MPNNPolicyNet(g.edge_index, g.x.size(0), torch.ones(g.edge_index.size(1)), "cuda")
    pm = PolicyModule(policy, g.edge_index)
    assert occupancy_only(pm, ValueModule(MPNNValueNetSimple(g.edge_index, g.x.size(0), "cuda")))
    runs = []
    for slim, use_sink in ((False, True), (True, True), (False, False)):
        env = BatchedSimulatorEnv(g, sim.Nmax, sim.agent.agent_features, replicas=8, seed=3)
        ad = _EnvAdapter(env)
        if not use_sink:
            ad.action_sink = lambda: None
        torch.manual_seed(11)
        runs.append(collect(ad, pm, 25, occupancy_only=slim))
        env.check_errors()
    full, slim, plain = runs
    assert slim["sel"] is None and slim["agent_index"] is None and full["sel"] is not None
    for k in ("num", "next_num", "action", "sample_log_prob", "reward", "done", "time"):
        assert torch.equal(full[k], slim[k]), k
    assert float(full["num"].sum()) > 0
    # Without the sink the uniforms come from torch.rand instead of the kernel's own stream: another episode of the same
    # process (the equivalence of the two write paths under the SAME uniforms is test_mpnn_gpu's sink test). Every
    # replica still selects exactly one edge per source group, and the observation is consistent with its reward.
    groups = int(torch.unique(g.edge_index[0]).numel())
    for run in (full, plain):
        assert bool((run["action"].sum(-1) == groups).all())
        assert torch.equal(run["next_num"].sum(-1), -run["reward"])
        sel_of_links = run["next_sel"][..., : env.N]
        assert bool(((sel_of_links >= 0) & (sel_of_links < g.x.size(0))).all())


def test_gae_values_once_over_the_frames_equal_both_shifted_views(scenario):
    """ppo_train evaluates the value net once over the T+1 frames of a rollout (frame t+1 is step t's next
    observation); that must equal evaluating it on the observations and on the next observations separately."""
    from tarl_simulator_b200.mpnn_agent import MPNNPolicyNet, MPNNValueNet, MPNNValueNetSimple
    from tarl_simulator_b200.reinforcement_learning import BatchedSimulatorEnv
    from tarl_simulator_b200.rl.ppo_trainer import PolicyModule, ValueModule, _EnvAdapter, _values, collect
    from tarl_simulator_b200.transportation_simulator import TransportationSimulator
    sim = TransportationSimulator("cuda")
    sim.load_network(scenario)
    sim.agent.load(scenario)
    g = sim.graph
    env = BatchedSimulatorEnv(g, sim.Nmax, sim.agent.agent_features, replicas=4, seed=3)
    ad = _EnvAdapter(env)
    pm = PolicyModule(MPNNPolicyNet(g.edge_index, g.x.size(0), torch.ones(g.edge_index.size(1)), "cuda"), g.edge_index)
    simple = ValueModule(MPNNValueNetSimple(g.edge_index, g.x.size(0), "cuda"))
    full = MPNNValueNet(g.edge_index, g.x.size(0), "cuda")
    full.agent_features = sim.agent.agent_features.cuda()
    full.eval()
    for vm, slim in ((simple, True), (simple, False), (ValueModule(full), False)):
        batch = collect(ad, pm, 12, occupancy_only=slim)
        with torch.no_grad():
            v_all = _values(ad, vm, batch["_frames"])
            v, nv = _values(ad, vm, batch), _values(ad, vm, batch, "next_")
        assert v_all.shape == (13, 4)
        assert torch.equal(v_all[:-1], v) and torch.equal(v_all[1:], nv)
