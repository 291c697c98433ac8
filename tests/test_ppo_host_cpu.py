"""CPU tests of the host-side PPO logic (no kernels involved): the plain-torch GAE / ClipPPOLoss formulas (the
specification the device kernels are tested against in tests/test_ppo_device_gpu.py) against hand-written loops, and
the multi-GPU plumbing ppo_train runs on every device — the flat parameter / gradient bucket with its in-place
all-reduce and broadcast, the global advantage statistics, replica sharding — under a world_size-2 gloo group."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tarl_simulator_b200.rl import ppo_trainer as P


def test_gae_matches_naive_recursion():
    g = torch.Generator().manual_seed(0)
    T, R = 17, 3
    v, nv, r = (torch.randn(T, R, generator=g) for _ in range(3))
    done = torch.rand(T, R, generator=g) < 0.15
    adv, target = P.gae_host(v, nv, r, done, done)
    for j in range(R):
        run = 0.0
        for t in reversed(range(T)):
            nd = 0.0 if done[t, j] else 1.0
            delta = float(r[t, j]) + 0.99 * float(nv[t, j]) * nd - float(v[t, j])
            run = delta + 0.99 * 0.95 * nd * run
            assert abs(run - float(adv[t, j])) < 1e-4
    assert torch.allclose(target, adv + v)


def test_standardise_single_process():
    a = torch.randn(50, 4)
    s = P.standardise_host(a)
    assert torch.allclose(s, (a - a.mean()) / a.std().clamp_min(1e-4), atol=1e-5)
    assert torch.equal(P.standardise_host(torch.zeros(8)), torch.zeros(8))       # std clamp: 0 / 1e-4


def test_clip_ppo_loss_formulas():
    g = torch.Generator().manual_seed(1)
    n = 64
    lp, slp, adv, ent, v, vt = (torch.randn(n, generator=g) for _ in range(6))
    out = P.clip_ppo_loss(lp, slp, adv, ent, v, vt)
    ratio = (lp - slp).exp()
    obj = -torch.minimum(ratio * adv, ratio.clamp(0.8, 1.2) * adv).mean()
    assert torch.allclose(out["loss_objective"], obj)
    assert torch.allclose(out["loss_entropy"], -0.01 * ent.mean())
    d = (v - vt).abs()
    assert torch.allclose(out["loss_critic"], torch.where(d < 1, 0.5 * d * d, d - 0.5).mean())


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(rank)                       # DIFFERENT initial parameters: the broadcast must level them
    net = torch.nn.Sequential(torch.nn.Linear(5, 4), torch.nn.Tanh(), torch.nn.Linear(4, 1))
    bucket = P.GradBucket(list(net.parameters()))
    init = bucket.flat.clone()
    bucket.broadcast(0)
    x = torch.randn(16, 5, generator=torch.Generator().manual_seed(100 + rank))
    net(x).pow(2).mean().backward()
    bucket.check_views()                          # autograd accumulated into the bucket, not into fresh tensors
    local = [p.grad.clone() for p in net.parameters()]
    world_seen = bucket.allreduce()
    adv = torch.randn(30, generator=torch.Generator().manual_seed(200 + rank)) * (1 + rank)
    torch.save({"local": local, "summed": [p.grad.clone() for p in net.parameters()], "adv": adv, "world": world_seen,
                "init": init, "start": bucket.flat.clone(), "std": P.standardise_host(adv)},
               os.path.join(out_dir, f"r{rank}.pt"))
    dist.destroy_process_group()


def test_two_rank_gradient_allreduce_and_global_statistics(tmp_path):
    world, port = 2, 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r = [torch.load(tmp_path / f"r{k}.pt") for k in range(world)]
    assert r[0]["world"] == r[1]["world"] == 2
    assert not torch.equal(r[0]["init"], r[1]["init"]) and torch.equal(r[0]["start"], r[1]["start"])
    assert torch.equal(r[0]["start"], r[0]["init"])                        # rank 0's parameters won
    for k, g in enumerate(r[0]["summed"]):
        assert torch.allclose(g, r[0]["local"][k] + r[1]["local"][k], atol=1e-7)
        assert torch.equal(g, r[1]["summed"][k])                           # every rank applies the same update
    both = torch.cat([r[0]["adv"], r[1]["adv"]])
    ref = (both - both.mean()) / both.std().clamp_min(1e-4)
    assert torch.allclose(torch.cat([r[0]["std"], r[1]["std"]]), ref, atol=1e-5)


def test_replica_sharding():
    from tarl_simulator_b200.parallel import shard_replicas
    assert [shard_replicas(1024, 8, r) for r in range(8)] == [(128 * r, 128) for r in range(8)]
    assert [shard_replicas(10, 4, r) for r in range(4)] == [(0, 3), (3, 3), (6, 2), (8, 2)]
    with pytest.raises(ValueError):
        shard_replicas(3, 4, 0)


def test_occupancy_only_predicate_and_trajectory_views():
    """Host logic of the slim rollouts: which pairs of nets qualify, and that the trajectory dict keeps the shifted
    views (next_* of step t is frame t+1, same storage) with None for the frames an occupancy-only rollout drops."""
    from types import SimpleNamespace
    from tarl_simulator_b200.rl.ppo_trainer import _trajectory, occupancy_only
    static_policy = SimpleNamespace(net=SimpleNamespace(reads_dynamic_features=False))
    dynamic_policy = SimpleNamespace(net=SimpleNamespace())
    simple_value = SimpleNamespace(net=SimpleNamespace(forward_occupancy=lambda n, t: n))
    full_value = SimpleNamespace(net=SimpleNamespace())
    assert occupancy_only(static_policy, simple_value)
    assert not occupancy_only(dynamic_policy, simple_value)
    assert not occupancy_only(static_policy, full_value)
    assert not occupancy_only(static_policy, None)
    T, R, M, E = 5, 2, 7, 11
    num = torch.arange((T + 1) * R * M, dtype=torch.float32).view(T + 1, R, M)
    times = torch.arange(T + 1, dtype=torch.float32).unsqueeze(1).expand(T + 1, R)
    action = torch.zeros(T, R, E, dtype=torch.bool)
    out = _trajectory(num, None, None, times, action, 3)
    assert out["sel"] is None and out["next_agent_index"] is None and out["_frames"]["sel"] is None
    assert out["num"].shape == (3, R, M) and out["_frames"]["num"].shape == (4, R, M)
    assert torch.equal(out["next_num"], num[1:4]) and out["next_num"].data_ptr() == num[1].data_ptr()
    assert torch.equal(out["next_time"][:, 0], torch.tensor([1.0, 2.0, 3.0]))
    sel = torch.zeros(T + 1, R, M)
    ai = torch.zeros(T + 1, R, M, dtype=torch.int64)
    out = _trajectory(num, sel, ai, times, action, T)
    assert out["sel"].shape == (T, R, M) and out["next_sel"].data_ptr() == sel[1].data_ptr()
    assert out["_frames"]["agent_index"].shape == (T + 1, R, M)
