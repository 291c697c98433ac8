"""The device side of the PPO iteration (csrc/optim.cu, the captured rollout, the per-rank noise streams) against its
plain-torch specification: tarl_gae / tarl_standardise vs gae_host / standardise_host, tarl_ppo_clip_loss vs
clip_ppo_loss under torch autograd, tarl_adam_step vs torch.optim.Adam on the same gradients, a rollout replayed from its CUDA graph vs the same rollout launched eagerly,
and two shards of one job drawing different actions from one seed (every rank sampled the same actions)."""
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


def _env_and_nets(scenario, replicas, seed=3, first_replica=0):
    from tarl_simulator_b200.mpnn_agent import MPNNPolicyNet, MPNNValueNetSimple
    from tarl_simulator_b200.reinforcement_learning import BatchedSimulatorEnv
    from tarl_simulator_b200.rl.ppo_trainer import PolicyModule, ValueModule
    from tarl_simulator_b200.transportation_simulator import TransportationSimulator
    sim = TransportationSimulator("cuda")
    sim.load_network(scenario)
    sim.agent.load(scenario)
    g = sim.graph
    env = BatchedSimulatorEnv(g, sim.Nmax, sim.agent.agent_features, replicas=replicas, seed=seed,
                              first_replica=first_replica)
    torch.manual_seed(5)
    policy = MPNNPolicyNet(g.edge_index, g.x.size(0), torch.ones(g.edge_index.size(1)), "cuda")
    value = MPNNValueNetSimple(g.edge_index, g.x.size(0), "cuda")
    return env, PolicyModule(policy, g.edge_index), ValueModule(value)


@pytest.mark.parametrize("T,R", [(17, 3), (32, 1024), (1, 130)])
def test_gae_kernel_matches_the_torch_formula(T, R):
    from tarl_simulator_b200.rl import ppo_trainer as P
    g = torch.Generator().manual_seed(T * 1000 + R)
    v_frames = torch.randn(T + 1, R, generator=g) * 50
    reward = -torch.rand(T, R, generator=g) * 300
    done = torch.rand(T, R, generator=g) < 0.1
    adv_ref, target_ref = P.gae_host(v_frames[:-1], v_frames[1:], reward, done, done)
    std_ref = P.standardise_host(adv_ref)
    adv, target = P.gae_device(v_frames.cuda(), reward.cuda(), done.cuda(), done.cuda())
    scale = float(target_ref.abs().max())
    assert torch.allclose(target.cpu(), target_ref, rtol=1e-5, atol=1e-5 * scale)
    assert torch.allclose(adv.cpu(), std_ref, rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("n", [1, 32, 1000])
def test_clip_loss_kernel_matches_the_torch_formula_and_its_autograd(n):
    """Forward scalars and the gradient of loss_objective + loss_critic + loss_entropy w.r.t. log_prob, entropy and
    value: ratios on both sides of the clip interval with advantages of both signs (the four min / clamp regimes),
    value errors on both sides of smooth-L1's |d| = 1, and a non-unit upstream gradient. Tolerance: 1e-5 relative,
    1e-7 absolute (both sum n products of the same fp32 terms, in different orders)."""
    from tarl_simulator_b200.rl import ppo_trainer as P
    g = torch.Generator().manual_seed(n)
    slp = -torch.rand(n, generator=g) * 3
    lp0 = slp + torch.randn(n, generator=g) * 0.3              # ratios from ~0.4 to ~2.5: inside and outside [0.8, 1.2]
    adv = torch.randn(n, generator=g)
    ent0 = torch.rand(n, generator=g) * 2
    tgt = torch.randn(n, generator=g) * 3
    v0 = tgt + torch.randn(n, generator=g) * 1.5               # |v - target| below and above 1
    if n >= 4:
        lp0[0] = slp[0]; adv[1] = 0.0; v0[2] = tgt[2]; lp0[3] = slp[3] + 5.0     # ratio == 1, A == 0, d == 0, ratio >> 1
    outs = []
    for fn in (P.clip_ppo_loss, P.clip_ppo_loss_device):
        lp, ent, v = (x.clone().cuda().requires_grad_(True) for x in (lp0, ent0, v0))
        res = fn(lp, slp.cuda(), adv.cuda(), ent, v, tgt.cuda())
        total = res["loss"] if "loss" in res else res["loss_objective"] + res["loss_critic"] + res["loss_entropy"]
        (total * 1.7).backward()
        if "impossible_frames" in res:
            assert float(res["impossible_frames"]) == 0.0
        outs.append(({k: res[k].detach().cpu() for k in ("loss_objective", "loss_entropy", "loss_critic", "approx_kl",
                                                          "clip_fraction", "entropy")},
                     total.detach().cpu(), [x.grad.cpu() for x in (lp, ent, v)]))
    (ref, ref_total, ref_g), (our, our_total, our_g) = outs
    for k in ref:
        assert torch.allclose(our[k], ref[k], rtol=1e-5, atol=1e-7), k
    assert torch.allclose(our_total, ref_total, rtol=1e-5, atol=1e-7)
    for a, b, name in zip(our_g, ref_g, ("log_prob", "entropy", "value")):
        assert torch.allclose(a, b, rtol=1e-5, atol=1e-9), name
    with pytest.raises(RuntimeError):
        P.clip_ppo_loss_device(lp0, slp, adv, ent0, v0, tgt)                     # CPU tensors: no fallback


def test_clip_loss_kernel_leaves_impossible_frames_out_of_the_objective():
    """A frame whose action selects no edge in some group carries -inf in BOTH log-probabilities (GraphDistribution's
    marker, src/reinforcement_learning.py:82-93; about one draw in 10^8 — a few frames of a 128-replica grid100 rollout).
    The torch formula makes its ratio NaN (and the next Adam step every parameter); the kernel treats it as a frame
    with ratio 1 and advantage 0 — same scalars and gradients as the torch formula on such a frame — and counts it
    (declared divergence D8). A -inf or NaN on ONE side only still propagates."""
    from tarl_simulator_b200.rl import ppo_trainer as P
    n, bad = 32, [5, 17]
    g = torch.Generator().manual_seed(7)
    slp = -torch.rand(n, generator=g) * 3
    lp0 = slp + torch.randn(n, generator=g) * 0.3
    adv, ent0 = torch.randn(n, generator=g), torch.rand(n, generator=g) * 2
    tgt = torch.randn(n, generator=g) * 3
    v0 = tgt + torch.randn(n, generator=g) * 1.5
    lp_ref, slp_ref, adv_ref = lp0.clone(), slp.clone(), adv.clone()
    lp_ref[bad] = 0.0; slp_ref[bad] = 0.0; adv_ref[bad] = 0.0
    lp_dev, slp_dev = lp0.clone(), slp.clone()
    lp_dev[bad] = float("-inf"); slp_dev[bad] = float("-inf")
    assert not torch.isfinite(P.clip_ppo_loss(lp_dev, slp_dev, adv, ent0, v0, tgt)["loss_objective"])     # the formula: NaN
    lp, ent, v = (x.clone().cuda().requires_grad_(True) for x in (lp_ref, ent0, v0))
    ref = P.clip_ppo_loss(lp, slp_ref.cuda(), adv_ref.cuda(), ent, v, tgt.cuda())
    (ref["loss_objective"] + ref["loss_critic"] + ref["loss_entropy"]).backward()
    lp2, ent2, v2 = (x.clone().cuda().requires_grad_(True) for x in (lp_dev, ent0, v0))
    our = P.clip_ppo_loss_device(lp2, slp_dev.cuda(), adv.cuda(), ent2, v2, tgt.cuda())
    our["loss"].backward()
    assert float(our["impossible_frames"]) == len(bad)
    for k in ("loss_objective", "loss_entropy", "loss_critic", "approx_kl", "clip_fraction", "entropy"):
        assert torch.allclose(our[k], ref[k].detach(), rtol=1e-5, atol=1e-7), k
    for a, b in ((lp2, lp), (ent2, ent), (v2, v)):
        assert torch.isfinite(a.grad).all() and torch.allclose(a.grad, b.grad, rtol=1e-5, atol=1e-9)
    assert float(lp2.grad[bad].abs().max()) == 0.0
    one_sided = lp0.clone().cuda(); one_sided[3] = float("-inf")
    res = P.clip_ppo_loss_device(one_sided, slp.cuda(), adv.cuda(), ent0.cuda(), v0.cuda(), tgt.cuda())
    assert float(res["impossible_frames"]) == 0 and torch.isfinite(res["loss_objective"])   # exp(-inf) = 0: a legal ratio
    nan_in = lp0.clone().cuda(); nan_in[3] = float("nan")
    assert torch.isnan(P.clip_ppo_loss_device(nan_in, slp.cuda(), adv.cuda(), ent0.cuda(), v0.cuda(), tgt.cuda())["loss_objective"])


def test_flat_adam_matches_torch_adam():
    from tarl_simulator_b200.rl.ppo_trainer import FlatAdam
    torch.manual_seed(0)
    make = lambda: torch.nn.Sequential(torch.nn.Linear(37, 64), torch.nn.ReLU(), torch.nn.Linear(64, 5)).cuda()
    a, b = make(), make()
    b.load_state_dict(a.state_dict())
    ours, ref = FlatAdam(list(a.parameters()), lr=1e-3), torch.optim.Adam(b.parameters(), lr=1e-3)
    gen = torch.Generator(device="cuda").manual_seed(1)
    for k in range(25):
        x = torch.randn(16, 37, device="cuda", generator=gen)
        for net in (a, b):
            net(x).pow(2).mean().backward()
        ours.check_views()
        norm = ours.step()
        ref_norm = torch.norm(torch.stack([p.grad.norm() for p in b.parameters()]))
        assert torch.allclose(norm.reshape(()), ref_norm, rtol=1e-5)
        ref.step()
        ours.zero_grad(); ref.zero_grad()
        for p, q in zip(a.parameters(), b.parameters()):
            assert torch.allclose(p, q, rtol=1e-5, atol=1e-7), f"parameters drift apart at step {k}"
    assert all(p.data_ptr() >= ours.flat.data_ptr() for p in a.parameters())       # still views of the bucket


def test_captured_rollout_equals_the_eager_rollout(scenario, monkeypatch):
    """Three rollouts of one shape: eager, captured + replayed, replayed. With the graph switched off the same three
    seeds must give the same trajectories (same launches, noise keys in device words), and different seeds differ."""
    from tarl_simulator_b200.rl.ppo_trainer import _EnvAdapter, collect
    runs = {}
    for mode in ("graph", "eager"):
        if mode == "eager":
            monkeypatch.setenv("TARL_NO_ROLLOUT_GRAPH", "1")
        env, pm, _ = _env_and_nets(scenario, replicas=8)
        ad = _EnvAdapter(env)
        torch.manual_seed(21)
        out = []
        for it in range(3):
            b = collect(ad, pm, 20, occupancy_only=True)
            out.append({k: b[k].clone() for k in ("num", "next_num", "action", "sample_log_prob", "reward", "done", "time")})
            if mode == "graph":
                key = next(iter(ad._graphs))
                assert (ad._graphs[key]["graph"] is not None) == (it >= 1)
        env.check_errors()
        runs[mode] = (out, env.export_x().clone(), env.agent_features.clone(), env.time, env.store.step_id)
    g, e = runs["graph"], runs["eager"]
    for it in range(3):
        for k in g[0][it]:
            assert torch.equal(g[0][it][k], e[0][it][k]), f"rollout {it}: {k} differs between replay and eager launch"
    assert torch.equal(g[1], e[1]) and torch.equal(g[2], e[2]) and g[3:] == e[3:]
    assert not torch.equal(g[0][0]["action"], g[0][1]["action"])          # a new seed per rollout: new draws
    assert float(g[0][2]["num"].sum()) > 0


def test_shards_of_one_job_draw_different_actions(scenario):
    """Two ranks of a data-parallel job seed alike (identical initial parameters) and the policy's logits do not depend
    on the replica: without the global replica offset in the noise streams rank k's replica r would repeat rank 0's
    replica r. Same offset -> same rollout; different offset -> different actions and different hand-off noise."""
    from tarl_simulator_b200.rl.ppo_trainer import _EnvAdapter, collect
    outs = []
    for first in (0, 0, 8):
        env, pm, _ = _env_and_nets(scenario, replicas=8, seed=3, first_replica=first)
        torch.manual_seed(21)
        b = collect(_EnvAdapter(env), pm, 12, occupancy_only=True)
        outs.append((b["action"].clone(), b["num"].clone()))
        env.check_errors()
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    assert not torch.equal(outs[0][0], outs[2][0])
    same_rows = (outs[0][0] == outs[2][0]).all(-1).float().mean()
    assert float(same_rows) < 0.05, "replicas of different shards must not repeat each other's actions"


def test_ppo_iteration_trains_and_checkpoint_uses_reference_keys(scenario, tmp_path):
    from tarl_simulator_b200.rl import ppo_trainer as P
    env, pm, vm = _env_and_nets(scenario, replicas=8)
    before = pm.net.nodes_embedding.weight.detach().clone()
    w1_before = vm.net.final_mlp[0].weight.detach().clone()
    ck = tmp_path / "policy.pt"
    hist = P.ppo_train(env, pm, vm, total_frames=60, frames_per_batch=20, num_epochs=2, sub_batch_size=64, history=[],
                       checkpoint_path=ck)
    assert len(hist) == 3 and hist[0]["frames"] == 160
    assert all(torch.isfinite(torch.tensor(hist[-1][k])) for k in ("loss_total", "grad_global_norm", "loss_critic"))
    assert not torch.equal(before, pm.net.nodes_embedding.weight.detach())
    assert not torch.equal(w1_before, vm.net.final_mlp[0].weight.detach())
    state = torch.load(ck)
    assert all(k.startswith("module.0.module.") for k in state) and "module.0.module.nodes_embedding.weight" in state
    env2, pm2, _ = _env_and_nets(scenario, replicas=4)
    P.load_policy_state(pm2, state)
    assert torch.equal(pm2.net.nodes_embedding.weight.detach(), pm.net.nodes_embedding.weight.detach())
    P.load_policy_state(pm2, pm.net.state_dict())                          # bare keys load too
    env.check_errors()
