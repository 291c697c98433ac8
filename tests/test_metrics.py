"""Metrics side channels: hourly hand-off / withdrawal counters and the road-optimality aggregate.

Golden: tests/golden/sim_metrics_grid4.npz — the UNMODIFIED reference's classical loop across an hour boundary, its own
compute_node_metrics() result (src/transportation_simulator.py:563-669) and the per-link road-optimality series of
plot_road_optimality (:482-488), written by oracle/gen_golden_sim.py.
CPU: the history reduction and the V/C statistics of the product against that golden. GPU: the same trajectory
through TransportationSimulator with on-device counters (tarl_metrics_accumulate), bit-exact counts; the batched
link-store environment's counters against the masks of the RL golden."""
import os

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))


def golden(name):
    return np.load(os.path.join(HERE, "golden", name + ".npz"))


def as_time(v):
    v = float(v)
    return int(v) if v == int(v) else v


def test_history_reduction_and_vc_statistics_match_reference(tmp_path):
    from tarl_simulator_b200.metrics import counts_from_histories, node_metrics_from_counts
    d = golden("sim_metrics_grid4")
    T = len(d["t"])
    upd = [(as_time(d["t"][s]), torch.from_numpy(d["pop"][s])) for s in range(T) if d["has_pop"][s]]
    wdr = [(as_time(d["t"][s]), torch.from_numpy(d["withdrawn"][s])) for s in range(T)]
    counts = counts_from_histories(upd, wdr)
    assert counts.dtype == torch.int64 and counts.shape == d["nm_counts"].shape and counts.size(1) == 2   # hours 0 and 1
    assert torch.equal(counts, torch.from_numpy(d["nm_counts"]))
    nm = node_metrics_from_counts(counts, torch.from_numpy(d["g_x"])[:, 3 * int(d["Nmax"]) + 4], str(tmp_path))
    assert sorted(nm) == list(range(counts.size(0)))
    np.testing.assert_allclose([nm[n]["avg_vc"] for n in nm], d["nm_avg_vc"], rtol=1e-6)
    np.testing.assert_allclose([nm[n]["std_vc"] for n in nm], d["nm_std_vc"], rtol=1e-6, atol=1e-9)
    assert nm[5]["hourly_counts"] == d["nm_counts"][5].tolist()
    rows = (tmp_path / "node_metrics.csv").read_text().splitlines()
    assert rows[0] == "node_id,avg_vc,std_vc,count_0h,count_1h" and len(rows) == counts.size(0) + 1
    assert counts_from_histories([], []) is None


def test_hour_rule():
    from tarl_simulator_b200.metrics import LinkMetrics
    assert [LinkMetrics.hour_of(t) for t in (0, 3599, 3599.9, 3600, 7200.5, -5)] == [0, 0, 0, 1, 2, 0]


@pytest.mark.gpu
def test_on_device_counters_match_reference(tmp_path):
    from tarl_simulator_b200.transportation_simulator import TransportationSimulator
    d = golden("sim_metrics_grid4")
    (tmp_path / "network.xml").write_text(str(d["xml"]))

    def replay(node_metrics, keep):
        sim = TransportationSimulator("cuda")
        sim.config_network(str(tmp_path / "network"))
        sim.agent.agent_features = torch.from_numpy(d["af0"]).cuda()
        sim.config_parameters(start_time=as_time(d["t0"]))
        sim.agent.set_time(sim.time)
        sim.node_metrics = node_metrics
        sim.agent.keep_history = keep
        sim.model_core.response_mpnn.update_history.keep = keep
        for s in range(len(d["t"])):
            sim.run(noise=torch.from_numpy(d["u_core"][s]).cuda(), choice_uniforms=torch.from_numpy(d["u_choice"][s]).cuda())
        assert torch.equal(sim.graph.x.cpu(), torch.from_numpy(d["x"][-1]))
        sim.model_core.check_errors()
        return sim

    sim = replay(node_metrics=True, keep=False)          # counters only: no per-step masks are retained
    assert len(sim.agent.withdraw_history) == 0 and len(sim.model_core.response_mpnn.update_history) == 0
    m = sim.metrics
    assert m.steps == len(d["t"]) and m.max_hour == 1
    assert torch.equal(sim.hourly_counts().cpu(), torch.from_numpy(d["nm_counts"]))
    nm = sim.compute_node_metrics(output_dir=str(tmp_path / "out"))
    np.testing.assert_allclose([nm[n]["avg_vc"] for n in nm], d["nm_avg_vc"], rtol=1e-6)
    np.testing.assert_allclose([nm[n]["std_vc"] for n in nm], d["nm_std_vc"], rtol=1e-6, atol=1e-9)
    assert [nm[n]["hourly_counts"] for n in nm] == d["nm_counts"].tolist()
    assert (tmp_path / "out" / "node_metrics.csv").exists()
    # road optimality: the series the reference's plot reduces, the latest per-link aggregate and the hourly sums
    times, agg = sim.road_optimality_series()
    assert torch.equal(agg.cpu(), torch.from_numpy(d["ro_agg"]))
    np.testing.assert_allclose(times.numpy() * 3600.0, d["ro_times"], rtol=1e-6)
    assert torch.equal(m.optimality_now[0].cpu(), torch.from_numpy(d["ro_agg"][-1]))
    hours = torch.tensor([int(t) // 3600 for t in d["t"]])
    for h in (0, 1):
        ref = torch.from_numpy(d["ro_agg"])[hours == h].double().sum(0)
        np.testing.assert_allclose(m.optimality_sum[0, h].cpu().double().numpy(), ref.numpy(), rtol=1e-5, atol=1e-5)
    rows = sim.plot_daily_counts({0: 10.0, 7: 3.5}, output_dir=str(tmp_path / "out"))
    assert rows["simulated"] == [int(d["nm_counts"][0].sum()), int(d["nm_counts"][7].sum())]
    assert rows["difference"][1] == rows["simulated"][1] - 3.5

    sim2 = replay(node_metrics=False, keep=True)         # the reference's way: reduce the retained histories
    assert sim2.metrics is None
    assert torch.equal(sim2.hourly_counts().cpu(), torch.from_numpy(d["nm_counts"]))
    assert sim2.compute_node_metrics(output_dir=None)[3]["hourly_counts"] == d["nm_counts"][3].tolist()


@pytest.mark.gpu
def test_batched_env_counters_match_reference_masks(tmp_path):
    from tarl_simulator_b200.reinforcement_learning import BatchedSimulatorEnv, SimulatorEnv
    from tarl_simulator_b200.transportation_simulator import TransportationSimulator
    d = golden("sim_rl_grid3")
    (tmp_path / "network.xml").write_text(str(d["xml"]))
    sim = TransportationSimulator("cuda")
    sim.config_network(str(tmp_path / "network"))
    R, T = 2, len(d["t"])
    env = BatchedSimulatorEnv(sim.graph, int(d["Nmax"]), torch.from_numpy(d["af0"]), replicas=R)
    m = env.enable_metrics(optimality=True)
    env.reset()
    for s in range(T):
        env.step(torch.from_numpy(d["action"][s]).cuda().repeat(R, 1), noise=torch.from_numpy(d["u_core"][s]).cuda().repeat(R, 1))
    env.check_errors()
    hours = np.array([int(t) // 3600 for t in d["t"]])
    assert sorted(set(hours.tolist())) == [5, 6] and m.max_hour == 6          # 05:59:00 .. 06:00:59
    masks = d["pop"].astype(np.int64) + d["withdrawn"].astype(np.int64)
    want = {h: torch.from_numpy(masks[hours == h].sum(0)) for h in (5, 6)}
    hour = 6
    for r in range(R):
        for h in (5, 6):
            assert torch.equal(m.counts[r, h].cpu().long(), want[h])
        assert int(m.counts[r, :5].sum()) == 0
    N = int(d["g_num_roads"])
    agg = torch.zeros(N).scatter_add_(0, torch.from_numpy(d["g_edge_index_routes"][0]), torch.from_numpy(d["delta_tt"][-1]))
    assert torch.equal(m.optimality_now[1].cpu(), agg)
    # the same episode through the drop-in environment with node_metrics on
    sim.agent.agent_features = torch.from_numpy(d["af0"]).cuda()
    env1 = SimulatorEnv(device="cuda", simulator=sim)
    sim.node_metrics = True
    env1.reset()
    for s in range(T):
        env1.noise = torch.from_numpy(d["u_core"][s]).cuda()
        env1._step({"action": torch.from_numpy(d["action"][s]).cuda()})
    assert torch.equal(sim.metrics.counts[0, 5].cpu().long(), want[5])
    assert torch.equal(sim.hourly_counts()[:, hour].cpu(), want[6])
    env1.reset()
    assert sim.metrics.steps == 0 and int(sim.metrics.counts.sum()) == 0
