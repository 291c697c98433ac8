"""GPU parity of the learned-MPNN path (C ABI: tarl_policy_embed_*, tarl_graphdist_*) against golden vectors from
the unmodified reference and against the CPU oracle port. Tolerance: 1e-5 relative (fp32), stated per assert;
integer outputs (mode, sampled actions, -inf rows) exact."""
import os

import numpy as np
import pytest
import torch

import mpnn_port

pytestmark = pytest.mark.gpu
RTOL, ATOL = 1e-5, 1e-6


def close(a, b, rtol=RTOL, atol=ATOL):
    torch.testing.assert_close(a.detach().cpu(), b.detach().cpu(), rtol=rtol, atol=atol)


@pytest.mark.parametrize("name", ["ring3", "rand1d", "rand2d", "single"])
def test_graph_distribution_goldens(name, golden_dir):
    from tarl_simulator_b200.distribution import GraphDistribution
    z = np.load(os.path.join(golden_dir, "mpnn_graphdist.npz"))
    g = lambda k: torch.from_numpy(z[f"{name}.{k}"])
    lg = g("logits").cuda().requires_grad_(True)
    ei = g("edge_index").cuda()
    d = GraphDistribution(lg, ei, temperature=float(z[f"{name}.temperature"]))
    close(d.proba, g("proba"))
    assert torch.equal(d.mode.cpu(), g("mode"))
    lp = d.log_prob(g("action").cuda())
    ent = d.entropy()
    assert lp.shape == g("log_prob").shape and ent.shape == g("entropy").shape
    close(lp, g("log_prob"))
    close(ent, g("entropy"))
    ((lp * g("w_lp").cuda()).sum() + (ent * g("w_ent").cuda()).sum()).backward()
    close(lg.grad, g("grad_logits"), rtol=1e-5, atol=2e-6)
    lp_bad = d.log_prob(g("bad_action").cuda()).detach().cpu()
    assert torch.equal(torch.isinf(lp_bad), torch.isinf(g("log_prob_bad"))) and torch.isinf(lp_bad).any()
    close(lp_bad, g("log_prob_bad"))
    if bool(z[f"{name}.stable_sort"]) and lg.dim() == 1:     # the reference's own (unstable) sort happened to be stable
        assert torch.equal(d.sample(uniforms=g("u").cuda()).cpu(), g("action"))


@pytest.mark.parametrize("seed,B", [(1, None), (2, 4)])
def test_graph_distribution_vs_oracle_with_sinks(seed, B):
    """Random graph with sink nodes (sources are NOT 0..K-1: the literal reference raises here, D1)."""
    from tarl_simulator_b200.distribution import GraphDistribution
    g = torch.Generator().manual_seed(seed)
    N, E = 500, 2100
    src = torch.randint(0, N // 2, (E,), generator=g) * 2          # odd nodes never a source
    dst = torch.randint(0, N, (E,), generator=g)
    ei = torch.stack([src, dst])
    shape = (E,) if B is None else (B, E)
    logits = torch.randn(*shape, generator=g) * 3
    ref_l = logits.clone().requires_grad_(True)
    ref = mpnn_port.GraphDistributionPort(ref_l, ei, 1.3)
    K = ref.K
    u = torch.rand(*(() if B is None else (B,)), K, generator=g)
    if B is None:
        act = ref.sample(u)
    else:
        act = torch.stack([mpnn_port.GraphDistributionPort(logits[b], ei, 1.3).sample(u[b]) for b in range(B)])
    lg = logits.cuda().requires_grad_(True)
    d = GraphDistribution(lg, ei.cuda(), temperature=1.3)
    assert d.nb_nodes == K
    assert torch.equal(d.sample(uniforms=u.cuda()).cpu(), act)
    close(d.proba, ref.proba)
    assert torch.equal(d.mode.cpu(), ref.mode)
    lp, ent = d.log_prob(act.cuda()), d.entropy()
    rlp, rent = ref.log_prob(act), ref.entropy()
    close(lp, rlp); close(ent, rent)
    wl, we = torch.randn(rlp.shape, generator=g), torch.randn(rent.shape, generator=g)
    ((rlp * wl).sum() + (rent * we).sum()).backward()
    ((lp * wl.cuda()).sum() + (ent * we.cuda()).sum()).backward()
    close(lg.grad, ref_l.grad, rtol=1e-5, atol=2e-6)
    # entropy alone (no action) and its gradient
    lg2 = logits.cuda().requires_grad_(True)
    e2 = GraphDistribution(lg2, ei.cuda(), temperature=1.3).entropy()
    e2.sum().backward()
    ref_l.grad = None
    mpnn_port.GraphDistributionPort(ref_l, ei, 1.3).entropy().sum().backward()
    close(lg2.grad, ref_l.grad, rtol=1e-5, atol=2e-6)


@pytest.mark.parametrize("tag", ["u", "b"])
def test_policy_and_value_simple_goldens(tag, golden_dir):
    from tarl_simulator_b200.mpnn_agent import MPNNPolicyNet, MPNNValueNetSimple
    z = np.load(os.path.join(golden_dir, "mpnn_nets.npz"))
    g = lambda k: torch.from_numpy(z[k])
    ei = g(f"value.{tag}.edge_index")
    nf, ef, ai, tm = (g(f"value.{tag}.{k}").cuda() for k in ("node_features", "edge_features", "agent_index", "time"))
    N = nf.size(-2)
    net = MPNNPolicyNet(ei.cuda(), N, torch.ones(ei.size(1)), "cuda")
    assert sorted(k for k, _ in net.named_parameters()) == list(z[f"policy.{tag}.param_names"])
    with torch.no_grad():
        net.nodes_embedding.weight.copy_(g(f"policy.{tag}.emb"))
    lg = net(nf, ef, ai)
    assert torch.equal(lg.cpu(), g(f"policy.{tag}.out"))            # a pure gather: exact
    (lg * g(f"policy.{tag}.w_out").cuda()).sum().backward()
    close(net.nodes_embedding.weight.grad, g(f"policy.{tag}.grad_emb"))
    net.check_errors()
    v = MPNNValueNetSimple(ei.cuda(), N, "cuda")
    with torch.no_grad():
        for k, p in v.named_parameters():
            p.copy_(g(f"simple.{tag}.param.{k}"))
    torch.backends.cuda.matmul.allow_tf32 = False
    out = v(nf, ef, ai, tm)
    close(out, g(f"simple.{tag}.out"))
    (out * g(f"simple.{tag}.w_out").cuda()).sum().backward()
    for k, p in v.named_parameters():
        close(p.grad, g(f"simple.{tag}.grad.{k}"), rtol=1e-4, atol=1e-6)


def test_policy_embed_negative_road_index_and_range_error():
    """D2: rows with ROAD_INDEX = -1 (SRC/DEST nodes) embed their own node id; an index beyond the table raises."""
    from tarl_simulator_b200.mpnn_agent import MPNNPolicyNet
    g = torch.Generator().manual_seed(3)
    N, E, B = 300, 1500, 3
    ei = torch.stack([torch.randint(0, N, (E,), generator=g), torch.randint(0, N, (E,), generator=g)])
    nf = torch.rand(B, N, 7, generator=g)
    ridx = torch.arange(N).float()
    ridx[200:] = -1.0
    nf[..., 6] = ridx
    net = MPNNPolicyNet(ei.cuda(), N, torch.ones(E), "cuda")
    w = net.nodes_embedding.weight.detach().cpu().clone().requires_grad_(True)
    ref = mpnn_port.policy_logits(w, nf, ei)
    out = net(nf.cuda(), None, None)
    assert torch.equal(out.cpu(), ref.detach())
    wo = torch.randn(B, E, generator=g)
    (ref * wo).sum().backward()
    (out * wo.cuda()).sum().backward()
    close(net.nodes_embedding.weight.grad, w.grad)
    nf[0, 5, 6] = N + 10
    net(nf.cuda(), None, None)
    with pytest.raises(IndexError):
        net.check_errors()


@pytest.mark.parametrize("tag", ["u", "b"])
def test_value_net_goldens(tag, golden_dir):
    """MPNNValueNet (eval mode) against outputs and parameter gradients of the unmodified reference."""
    from tarl_simulator_b200.mpnn_agent import MPNNValueNet
    z = np.load(os.path.join(golden_dir, "mpnn_nets.npz"))
    g = lambda k: torch.from_numpy(z[k])
    ei = g(f"value.{tag}.edge_index")
    nf, ef, ai, tm = (g(f"value.{tag}.{k}").cuda() for k in ("node_features", "edge_features", "agent_index", "time"))
    N = nf.size(-2)
    net = MPNNValueNet(ei.cuda(), N, "cuda")
    names = sorted(k for k, _ in net.named_parameters())
    assert names == sorted(k[len(f"value.{tag}.param."):] for k in z.files if k.startswith(f"value.{tag}.param."))
    net.agent_features = g(f"value.{tag}.agent_features").cuda()
    with torch.no_grad():
        for k, p in net.named_parameters():
            p.copy_(g(f"value.{tag}.param.{k}"))
    net.eval()
    torch.backends.cuda.matmul.allow_tf32 = False
    out = net(nf, ef, ai, tm)
    close(out, g(f"value.{tag}.out"))
    (out * g(f"value.{tag}.w_out").cuda()).sum().backward()
    for k, p in net.named_parameters():
        close(p.grad, g(f"value.{tag}.grad.{k}"), rtol=1e-5, atol=1e-6)
    net.check_errors()


@pytest.mark.parametrize("tag", ["u", "b"])
def test_value_net_train_mode_goldens(tag, golden_dir):
    """MPNNValueNet in TRAIN mode with the message-dropout mask the unmodified reference drew injected: outputs and
    parameter gradients of the reference (tests/golden/mpnn_value_train.npz)."""
    from tarl_simulator_b200.mpnn_agent import MPNNValueNet
    z = np.load(os.path.join(golden_dir, "mpnn_value_train.npz"))
    g = lambda k: torch.from_numpy(z[k])
    ei = g(f"{tag}.edge_index")
    nf, ef, ai, tm = (g(f"{tag}.{k}").cuda() for k in ("node_features", "edge_features", "agent_index", "time"))
    net = MPNNValueNet(ei.cuda(), nf.size(-2), "cuda")
    net.agent_features = g(f"{tag}.agent_features").cuda()
    with torch.no_grad():
        for k, p in net.named_parameters():
            p.copy_(g(f"{tag}.param.{k}"))
    net.train()
    net.time_net[1].p = 0.0                  # as in the generator: the output depends on the message mask alone
    net.time_net[4].p = 0.0
    net.keep_bits = g(f"{tag}.keep_bits").cuda()
    torch.backends.cuda.matmul.allow_tf32 = False
    out = net(nf, ef, ai, tm)
    close(out, g(f"{tag}.out"))
    (out * g(f"{tag}.w_out").cuda()).sum().backward()
    for k, p in net.named_parameters():
        close(p.grad, g(f"{tag}.grad.{k}"), rtol=1e-5, atol=1e-6)
    net.check_errors()
    assert torch.equal(net.dropout_words().cpu().reshape(-1), g(f"{tag}.keep_bits").reshape(-1))


@pytest.mark.parametrize("B,injected", [(None, True), (5, True), (4, False), (32, False)])
def test_value_net_train_mode_vs_oracle_large(B, injected):
    """Train mode on thousands of nodes: injected random masks, and the in-kernel Philox stream read back through
    dropout_words() and handed to the oracle — forward and every gradient must agree with the SAME mask."""
    from tarl_simulator_b200.mpnn_agent import MPNNValueNet
    g = torch.Generator().manual_seed(23 + (B or 0))
    N, E, A = 3000, 11000, 500
    src = torch.randint(0, N - 300, (E,), generator=g)           # the last 300 nodes have no out-edge
    ei = torch.stack([src, torch.randint(0, N - 100, (E,), generator=g)])    # ... and 100 nodes no in-edge
    lead = () if B is None else (B,)
    nf = torch.rand(*lead, N, 7, generator=g) * 3
    ef = torch.rand(*lead, E, 1, generator=g)
    ai = torch.randint(0, A + 1, (*lead, N), generator=g)
    tm = torch.rand(*lead, 1, generator=g) * 10
    af = torch.rand(A + 1, 9, generator=g) * 2
    net = MPNNValueNet(ei.cuda(), N, "cuda")
    net.agent_features = af.cuda()
    net.train()
    net.time_net[1].p = 0.0
    net.time_net[4].p = 0.0
    net.message_mlp[0].p = 0.2 if injected else 0.05
    if injected:
        keep = torch.rand(*lead, E, 17, generator=g) >= 0.2
        net.keep_bits = mpnn_port.pack_keep_bits(keep).cuda()
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(99)
    out = net(nf.cuda(), ef.cuda(), ai.cuda(), tm.cuda())
    words = net.dropout_words().cpu().reshape(*lead, E)
    keep_used = mpnn_port.unpack_keep_bits(words)
    if injected:
        assert torch.equal(keep_used, keep)
    else:
        assert int(words.max()) < (1 << 17)
        frac = keep_used.float().mean().item()                  # 17 * B * E Bernoulli(0.95) draws
        assert abs(frac - 0.95) < 4 * (0.95 * 0.05 / keep_used.numel()) ** 0.5 + 1e-4, frac
        per_input = keep_used.float().reshape(-1, 17).mean(0)
        assert (per_input - 0.95).abs().max() < 0.01
    p = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in net.named_parameters()}
    ref = mpnn_port.value_net_forward(p, nf, ef, af, ai, tm, ei, keep=keep_used, drop_p=net.message_mlp[0].p)
    close(out, ref.detach(), rtol=1e-5, atol=1e-6)
    w = torch.randn(ref.shape, generator=g)
    (ref * w).sum().backward()
    (out * w.cuda()).sum().backward()
    for k, v in net.named_parameters():
        close(v.grad, p[k].grad, rtol=2e-5, atol=2e-6)
    net.check_errors()
    if not injected:
        # the stream follows torch's default generator: same seed -> same words and value, next draw -> different
        torch.manual_seed(99)
        out2 = net(nf.cuda(), ef.cuda(), ai.cuda(), tm.cuda())
        assert torch.equal(out2, out) and torch.equal(net.dropout_words().cpu().reshape(*lead, E), words)
        net(nf.cuda(), ef.cuda(), ai.cuda(), tm.cuda())
        assert not torch.equal(net.dropout_words().cpu().reshape(*lead, E), words)
        # eval mode is untouched by all this
        net.eval()
        ev = net(nf.cuda(), ef.cuda(), ai.cuda(), tm.cuda())
        close(ev, mpnn_port.value_net_forward(p, nf, ef, af, ai, tm, ei).detach(), rtol=1e-5, atol=1e-6)


def test_value_net_train_mode_drop_everything():
    """p = 1: nn.Dropout zeroes every message input, the message is tanh(bias) on every edge."""
    from tarl_simulator_b200.mpnn_agent import MPNNValueNet
    g = torch.Generator().manual_seed(3)
    N, E = 50, 200
    ei = torch.stack([torch.randint(0, N, (E,), generator=g), torch.randint(0, N, (E,), generator=g)])
    net = MPNNValueNet(ei.cuda(), N, "cuda")
    net.agent_features = torch.rand(8, 9, generator=g).cuda()
    net.train()
    net.time_net[1].p = 0.0
    net.time_net[4].p = 0.0
    net.message_mlp[0].p = 1.0
    nf, ef = torch.rand(N, 7, generator=g), torch.rand(E, 1, generator=g)
    ai, tm = torch.randint(0, 8, (N,), generator=g), torch.rand(1, generator=g)
    out = net(nf.cuda(), ef.cuda(), ai.cuda(), tm.cuda())
    assert int(net.dropout_words().abs().max()) == 0
    p = {k: v.detach().cpu() for k, v in net.named_parameters()}
    has_out = torch.zeros(N).scatter_(0, ei[0], 1.0)
    v = torch.tanh(p["node_mlp.0.weight"][0, 0] * torch.tanh(p["message_mlp.1.bias"][0]) * has_out + p["node_mlp.0.bias"][0])
    te = net.time_net(tm.cuda()).detach().cpu()
    ref = torch.cat((v, te)) @ p["final_mlp.0.weight"][0] + p["final_mlp.0.bias"]
    close(out, ref, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("B", [None, 4])
def test_value_net_vs_oracle_large(B):
    """Thousands of nodes, nodes without out-edges, repeated edges: outputs and all gradients vs the oracle port."""
    from tarl_simulator_b200.mpnn_agent import MPNNValueNet
    g = torch.Generator().manual_seed(11 + (B or 0))
    N, E, A = 3000, 11000, 500
    src = torch.randint(0, N - 300, (E,), generator=g)           # the last 300 nodes have no out-edge
    ei = torch.stack([src, torch.randint(0, N, (E,), generator=g)])
    lead = () if B is None else (B,)
    nf = torch.rand(*lead, N, 7, generator=g) * 3
    ef = torch.rand(*lead, E, 1, generator=g)
    ai = torch.randint(0, A + 1, (*lead, N), generator=g)
    tm = torch.rand(*lead, 1, generator=g) * 10
    af = torch.rand(A + 1, 9, generator=g) * 2
    net = MPNNValueNet(ei.cuda(), N, "cuda")
    net.agent_features = af.cuda()
    net.eval()
    p = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in net.named_parameters()}
    ref = mpnn_port.value_net_forward(p, nf, ef, af, ai, tm, ei)
    torch.backends.cuda.matmul.allow_tf32 = False
    out = net(nf.cuda(), ef.cuda(), ai.cuda(), tm.cuda())
    close(out, ref.detach(), rtol=1e-5, atol=1e-6)
    w = torch.randn(ref.shape, generator=g)
    (ref * w).sum().backward()
    (out * w.cuda()).sum().backward()
    for k, v in net.named_parameters():
        close(v.grad, p[k].grad, rtol=2e-5, atol=2e-6)


def _value_mlp_case(M, N, pad, seed):
    from tarl_simulator_b200.mpnn_agent import MPNNValueNetSimple
    g = torch.Generator(device="cuda").manual_seed(seed)
    ei = torch.zeros(2, 1, dtype=torch.long, device="cuda")
    net = MPNNValueNetSimple(ei, N, "cuda")
    with torch.no_grad():
        for p in net.parameters():
            p.copy_(torch.randn(p.shape, device="cuda", generator=g) * (0.05 if p.dim() == 2 and p.size(1) > 64 else 0.3))
    buf = torch.zeros(M, N + pad, device="cuda")
    num = buf[:, :N]
    num.copy_(torch.randint(0, 15, (M, N), device="cuda", generator=g).float() * (torch.rand(M, N, device="cuda", generator=g) < 0.7))
    num[:, ::7] += torch.rand(M, num[:, ::7].size(1), device="cuda", generator=g)          # non-integers too
    time = (torch.arange(M, device="cuda", dtype=torch.float32).reshape(M, 1) % 7.0) * (3600.0 if M == 300 else 1.0)
    return net, num, time


def _within(got, ref, what):
    """The contract's bar, per element: |got - ref| <= 1e-5 |ref| + 1e-5 mean|ref| (the second term is the stated
    absolute tolerance: an output that is a near-cancellation of O(mean) terms cannot be held to its own size)."""
    got, ref = got.double(), ref.double()
    bound = 1e-5 * ref.abs() + 1e-5 * ref.abs().mean()
    worst = float(((got - ref).abs() / bound.clamp_min(1e-300)).max())
    assert worst <= 1.0, f"{what}: {worst:.2f} x the 1e-5 tolerance"


# pad = extra floats in the row pitch: 0 / 4 keep it TMA-addressable, 1 does not (and N = 1003 never is)
@pytest.mark.parametrize("M,N,pad", [(1, 40, 0), (100, 1000, 0), (128, 1003, 1), (128, 1000, 1), (96, 1000, 4),
                                     (300, 4096, 0), (1024, 59600, 0), (130, 24, 0)])
def test_value_mlp_tensor_core_forward_matches_fp64(M, N, pad):
    """MPNNValueNetSimple.forward_occupancy (tcgen05, 3xTF32) against the same MLP in float64, 1e-5 relative per element
    (the fp32 library GEMM is measured against the same bar beside it). Covers the K tail (N not a multiple of 32), the
    M tail, split-K, and row pitches TMA cannot address — which must reach the SAME kernel through a padded copy, never
    a library GEMM (`last_path` says which way the call went)."""
    net, num, time = _value_mlp_case(M, N, pad, M * 7 + N)
    addressable = M == 1 or num.stride(0) % 4 == 0
    with torch.no_grad():
        got = net.forward_occupancy(num, time)
        assert net.last_path == ("tcgen05" if addressable else "tcgen05+pad")
        assert torch.equal(net(torch.stack([num] * 7, dim=-1), None, None, time), got)     # the reference's signature
        lib = net.final_mlp(torch.cat((num, time), dim=-1))
        ref = net.final_mlp.double()(torch.cat((num, time), dim=-1).double())
        net.final_mlp.float()
    assert got.shape == (M, 1)
    _within(got, ref, "tcgen05 path")
    lib_err = float((lib.double() - ref).abs().max())
    assert float((got.double() - ref).abs().max()) <= 4 * lib_err + 1e-6 * float(ref.abs().mean()), "worse than fp32 SGEMM"


@pytest.mark.parametrize("M,N,pad", [(32, 59600, 0), (32, 1003, 0), (7, 40, 0), (200, 4096, 4)])
def test_value_mlp_backward_matches_fp64_autograd(M, N, pad):
    """The PPO update's forward + backward of MPNNValueNetSimple (src/rl/ppo_trainer.py:132-145 evaluates it with
    gradients on a 32-frame minibatch) on the hand-written kernels: output and all six parameter gradients against
    float64 autograd through the same nn.Sequential."""
    net, num, time = _value_mlp_case(M, N, pad, 1000 + M + N)
    w = torch.randn(M, 1, device="cuda", generator=torch.Generator(device="cuda").manual_seed(3))
    out = net.forward_occupancy(num, time)
    assert out.requires_grad and net.last_path.startswith("tcgen05")
    (out * w).sum().backward()
    got = {k: v.grad.clone() for k, v in net.final_mlp.named_parameters()}
    for v in net.final_mlp.parameters():
        v.grad = None
    net.final_mlp.double()
    ref_out = net.final_mlp(torch.cat((num, time), dim=-1).double())
    (ref_out * w.double()).sum().backward()
    _within(out.detach(), ref_out.detach(), "forward")
    for k, v in net.final_mlp.named_parameters():
        assert got[k].shape == v.grad.shape
        _within(got[k], v.grad, f"grad of final_mlp.{k}")
    net.final_mlp.float()


@pytest.mark.parametrize("B", [4, 32, 64])
def test_fast_sampling_path_matches_generic_kernel_and_its_log_prob(B):
    """Edge-major logits with B in {4, 8, 16, 32k} and a byte one-hot take k_gd_sample_em4 (4 rows per thread, fused
    log-probability): the drawn edges must equal those of the generic kernel on a row-major copy of the same logits
    with the same uniforms, and the fused log_prob must equal log_prob(action) (1e-5 relative)."""
    from tarl_simulator_b200.distribution import GraphDistribution
    g = torch.Generator(device="cuda").manual_seed(B)
    N, E = 3000, 14000
    ei = torch.stack([torch.randint(0, N - 300, (E,), device="cuda", generator=g),
                      torch.randint(0, N, (E,), device="cuda", generator=g)])
    ei[0, :40] = 7                                        # one group longer than the register cache
    base = torch.randn(E, B, device="cuda", generator=g) * 2.0
    lg_em = base.t()                                      # [B, E], strides (1, B)
    lg_rm = lg_em.contiguous()
    d_em, d_rm = GraphDistribution(lg_em, ei), GraphDistribution(lg_rm, ei)
    u = torch.rand(B, d_em.nb_nodes, device="cuda", generator=g)
    a_em, lp = d_em.sample(uniforms=u, dtype=torch.bool, return_log_prob=True)
    a_rm = d_rm.sample(uniforms=u, dtype=torch.bool)
    assert a_em.stride(0) == 1                            # stayed edge-major: the fast path ran
    assert torch.equal(a_em, a_rm)
    assert bool((a_em.sum(1) == d_em.nb_nodes).all())     # exactly one edge per source group
    ref = d_rm.log_prob(a_rm)
    assert torch.allclose(lp, ref, rtol=1e-5, atol=1e-5 * float(ref.abs().max()))
    # uniforms in group-major memory (what sample() draws itself) give the same result
    a2 = d_em.sample(uniforms=u.t().contiguous().t(), dtype=torch.bool)
    assert torch.equal(a2, a_em)


def test_value_mlp_split_cache_follows_the_weights():
    """The TF32 hi/lo split of W1 is cached in the module's workspace: an optimiser step (in-place update), a
    load_state_dict and a change of the number of rows must all be noticed."""
    from tarl_simulator_b200.mpnn_agent import MPNNValueNetSimple
    torch.manual_seed(0)
    N = 2048
    net = MPNNValueNetSimple(torch.zeros(2, 1, dtype=torch.long, device="cuda"), N, "cuda")
    num = torch.randint(0, 9, (64, N), device="cuda").float()
    tm = torch.rand(64, 1, device="cuda")

    def both(n=64):
        with torch.no_grad():
            return net.forward_occupancy(num[:n], tm[:n]), net.final_mlp(torch.cat((num[:n], tm[:n]), dim=-1))

    a, b = both()
    assert torch.allclose(a, b, rtol=1e-5, atol=1e-5)
    opt = torch.optim.SGD(net.parameters(), lr=0.5)
    net.final_mlp(torch.cat((num, tm), dim=-1)).sum().backward()
    opt.step()                                           # in place: same storage, new version
    a2, b2 = both()
    assert not torch.allclose(b2, b, rtol=1e-3, atol=1e-3) and torch.allclose(a2, b2, rtol=1e-5, atol=1e-5)
    sd = {k: v * 0.5 for k, v in net.state_dict().items()}
    net.load_state_dict(sd)
    a3, b3 = both()
    assert torch.allclose(a3, b3, rtol=1e-5, atol=1e-5) and not torch.allclose(b3, b2, rtol=1e-3, atol=1e-3)
    a4, b4 = both(32)                                     # another problem shape on the same workspace
    assert torch.allclose(a4, b4, rtol=1e-5, atol=1e-5)
    a5, b5 = both(64)
    assert torch.allclose(a5, b5, rtol=1e-5, atol=1e-5)


def test_policy_on_an_expanded_observation_and_sampling_from_one_logits_row():
    """A rollout hands MPNNPolicyNet the static observation expanded over the replicas (batch stride 0): the net
    computes one row and returns it expanded; sampling / log_prob / entropy from such logits must equal what the
    materialised [B, N, 7] observation gives, and gradients must flow through the expansion."""
    from tarl_simulator_b200.distribution import GraphDistribution
    from tarl_simulator_b200.mpnn_agent import MPNNPolicyNet
    g = torch.Generator(device="cuda").manual_seed(3)
    N, E, B = 500, 2100, 32
    ei = torch.stack([torch.randint(0, N - 20, (E,), device="cuda", generator=g), torch.randint(0, N, (E,), device="cuda", generator=g)])
    one = torch.rand(1, N, 7, device="cuda", generator=g)
    one[..., 6] = torch.arange(N, device="cuda").float()
    net = MPNNPolicyNet(ei, N, torch.ones(E, device="cuda"), "cuda")
    lg_x = net(one.expand(B, -1, -1), None, None)                 # stride-0 batch
    lg_m = net(one.repeat(B, 1, 1), None, None)                   # materialised
    assert lg_x.shape == lg_m.shape == (B, E) and lg_x.stride(0) == 0
    assert torch.equal(lg_x, lg_m)
    d_x, d_m = GraphDistribution(lg_x, ei), GraphDistribution(lg_m, ei)
    u = torch.rand(B, d_x.nb_nodes, device="cuda", generator=g)
    out = torch.empty(E, B, dtype=torch.bool, device="cuda").t()
    a_x, lp_x = d_x.sample(uniforms=u, dtype=torch.bool, out=out, return_log_prob=True)
    a_m, lp_m = d_m.sample(uniforms=u, dtype=torch.bool, return_log_prob=True)
    assert a_x.data_ptr() == out.data_ptr() and torch.equal(a_x, a_m)
    assert torch.allclose(lp_x, lp_m, rtol=1e-6, atol=1e-4)
    assert torch.allclose(d_x.log_prob(a_x), d_m.log_prob(a_m), rtol=1e-6, atol=1e-4)
    assert torch.allclose(d_x.entropy(), d_m.entropy(), rtol=1e-6, atol=1e-4)
    (d_x.log_prob(a_x).sum() + d_x.entropy().sum()).backward()
    gx = net.nodes_embedding.weight.grad.clone()
    net.nodes_embedding.weight.grad = None
    (d_m.log_prob(a_m).sum() + d_m.entropy().sum()).backward()
    assert torch.allclose(gx, net.nodes_embedding.weight.grad, rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("R", [4, 36, 128, 1024])
def test_sampling_kernel_applies_the_action_to_the_environment(R):
    """Rollout fast path: with one logits row for all replicas, GraphDistribution.sample(sink=env.action_sink()) writes
    SELECTED_ROAD in the pass that draws the edges. One-hot, log-probability and every SELECTED_ROAD cell must equal
    sample() followed by env.apply_action(action); a uniform of 1.0 (no hit) leaves the node's decision untouched."""
    from tarl_simulator_b200 import synthetic
    from tarl_simulator_b200.distribution import GraphDistribution
    from tarl_simulator_b200.reinforcement_learning import BatchedSimulatorEnv
    dev = torch.device("cuda")
    gen = torch.Generator(device="cuda").manual_seed(R)
    frm, to, n_nodes = synthetic.grid_links(7, device=dev)
    g, Nmax = synthetic.build_graph(frm, to, n_nodes)
    af = synthetic.population(g, 300, 21540, 60, seed=1)
    env = BatchedSimulatorEnv(g, Nmax, af, replicas=R, seed=5)
    E = g.edge_index.size(1)
    row = torch.randn(1, E, device=dev, generator=gen) * 2
    d = GraphDistribution(row.expand(R, -1), g.edge_index)
    u = torch.rand(d.nb_nodes, R, device=dev, generator=gen).t()
    u[R // 2, 3] = 1.0                                          # no hit for (replica R/2, group 3)
    sel0 = torch.rand(R, env.N, device=dev, generator=gen).mul(50).floor()
    src0 = torch.rand(env.src_sel.shape, device=dev, generator=gen).mul(50).floor()

    def start():
        env.store.sel[: R * env.N].view(R, env.N).copy_(sel0)
        env.src_sel.copy_(src0)

    start()
    out_a = torch.empty(E, R, dtype=torch.bool, device=dev).t()
    a, lp_a = d.sample(uniforms=u, dtype=torch.bool, out=out_a, return_log_prob=True)
    env.apply_action(a)
    sel_a, src_a = env.store.sel[: R * env.N].clone(), env.src_sel.clone()
    start()
    sink = env.action_sink()
    out_b = torch.empty(E, R, dtype=torch.bool, device=dev).t()
    b, lp_b = d.sample(uniforms=u, dtype=torch.bool, out=out_b, return_log_prob=True, sink=sink)
    assert sink.applied
    assert torch.equal(a, b)
    assert torch.allclose(lp_a, lp_b, rtol=1e-5, atol=1e-4) and bool(torch.isinf(lp_a[R // 2])) and int(torch.isinf(lp_a).sum()) == 1
    assert torch.equal(env.store.sel[: R * env.N], sel_a) and torch.equal(env.src_sel, src_a)
    assert not torch.equal(sel_a.view(R, env.N), sel0)          # ... and the action did change the decisions
    node = int(d.nodes[3])
    untouched = sel0[R // 2, node] if node < env.N else src0[R // 2, node - env.N]
    now = sel_a.view(R, env.N)[R // 2, node] if node < env.N else src_a[R // 2, node - env.N]
    assert float(now) == float(untouched)
    # without the log-probability, and a sink of another environment size is refused (falls back, not applied)
    start()
    c = d.sample(uniforms=u, dtype=torch.bool, sink=sink)
    assert sink.applied and torch.equal(c, a) and torch.equal(env.store.sel[: R * env.N], sel_a)
    if R > 4:
        d4 = GraphDistribution(row.expand(4, -1), g.edge_index)
        d4.sample(uniforms=u[:4], dtype=torch.bool, sink=sink)
        assert not sink.applied


def test_one_logits_row_sampling_matches_the_generic_kernel_on_long_groups():
    """k_gd_sample_bcast keeps 8 edges of a group in shared memory; longer groups walk the logits row. Same edges as
    the generic kernel on the materialised logits, with and without the action sink."""
    from tarl_simulator_b200.distribution import ActionSink, GraphDistribution
    from tarl_simulator_b200.topology import group_csr_for
    gen = torch.Generator(device="cuda").manual_seed(77)
    N, E, B = 400, 2600, 64
    ei = torch.stack([torch.randint(0, N - 30, (E,), device="cuda", generator=gen),
                      torch.randint(0, N, (E,), device="cuda", generator=gen)])
    ei[0, :40] = 11                                               # a 40+-edge group
    ei[0, 40:49] = 12                                             # a 9+-edge group (one past the cache)
    row = torch.randn(1, E, device="cuda", generator=gen) * 2
    d_x, d_m = GraphDistribution(row.expand(B, -1), ei), GraphDistribution(row.repeat(B, 1), ei)
    u = torch.rand(B, d_x.nb_nodes, device="cuda", generator=gen)
    a_m = d_m.sample(uniforms=u, dtype=torch.bool)
    a_x, lp = d_x.sample(uniforms=u, dtype=torch.bool, return_log_prob=True)
    assert torch.equal(a_x, a_m)
    ref = d_m.log_prob(a_m)
    assert torch.allclose(lp, ref, rtol=1e-5, atol=1e-5 * float(ref.abs().max()))
    n_links = 250                                                 # nodes >= 250 play the part of the SRC nodes
    sel_l = torch.full((B, n_links), -1.0, device="cuda")
    sel_s = torch.full((B, N - n_links), -1.0, device="cuda")
    sink = ActionSink(ei, group_csr_for(ei, "source_rank"), sel_l, sel_s, n_links, N)
    out = torch.empty(E, B, dtype=torch.bool, device="cuda").t()
    a_s = d_x.sample(uniforms=u, dtype=torch.bool, out=out, sink=sink)
    assert sink.applied and torch.equal(a_s, a_m)
    want = torch.full((B, N), -1.0, device="cuda")
    b_idx, e_idx = torch.nonzero(a_m, as_tuple=True)
    want[b_idx, ei[0][e_idx]] = ei[1][e_idx].float()
    assert torch.equal(torch.cat((sel_l, sel_s), dim=1), want)


def test_sampling_kernel_own_uniform_stream():
    """Without injected uniforms the rollout sampling kernel draws its own (Philox keyed by a seed taken from torch's
    default generator): empirical frequencies over 8192 replicas must follow the softmax of the shared logits row,
    torch.manual_seed must reproduce a draw, consecutive draws must differ, and replicas must not repeat each other."""
    from tarl_simulator_b200.distribution import ActionSink, GraphDistribution
    from tarl_simulator_b200.topology import group_csr_for
    gen = torch.Generator(device="cuda").manual_seed(5)
    N, E, R = 60, 260, 8192
    ei = torch.stack([torch.randint(0, N, (E,), device="cuda", generator=gen), torch.randint(0, N, (E,), device="cuda", generator=gen)])
    row = torch.randn(1, E, device="cuda", generator=gen)
    d = GraphDistribution(row.expand(R, -1), ei)
    sel = torch.zeros(R, N, device="cuda")
    sink = ActionSink(ei, group_csr_for(ei, "source_rank"), sel, None, N, N)

    def draw():
        out = torch.empty(E, R, dtype=torch.bool, device="cuda").t()
        a, lp = d.sample(dtype=torch.bool, out=out, return_log_prob=True, sink=sink)
        assert sink.applied
        return a.clone(), lp.clone()

    torch.manual_seed(123)
    a1, lp1 = draw()
    a2, _ = draw()
    torch.manual_seed(123)
    a3, lp3 = draw()
    assert torch.equal(a1, a3) and torch.equal(lp1, lp3) and not torch.equal(a1, a2)
    assert bool((a1.sum(1) == d.nb_nodes).all())                       # one edge per source group in every replica
    assert len({bytes(r.cpu().numpy().tobytes()) for r in a1[:64]}) > 60        # replicas draw independently
    p = d.proba[0] if d.proba.dim() == 2 else d.proba                  # softmax of the shared row, per edge
    freq = a1.float().mean(0)
    sigma = (p * (1 - p) / R).sqrt()
    assert bool(((freq - p).abs() <= 5 * sigma + 1e-3).all()), float(((freq - p).abs() / (sigma + 1e-6)).max())
    assert torch.allclose(lp1, d.log_prob(a1), rtol=1e-5, atol=1e-4)


# ---------------------------------------------------------------------------------------------------- per-edge MLPs
def _edge_mlp_net(d, tag, which):
    from tarl_simulator_b200.mpnn_agent import MPNNPolicyNet
    ei = torch.from_numpy(d[f"{tag}.edge_index"]).cuda()
    nf = torch.from_numpy(d[f"{tag}.node_features"]).cuda()
    net = MPNNPolicyNet(ei, nf.size(-2), None if nf.size(-2) > 4096 else torch.ones(ei.size(1), device="cuda"), "cuda")
    net.agent_features = torch.from_numpy(d[f"{tag}.agent_features"]).cuda()
    with torch.no_grad():
        for k, p in getattr(net, which).named_parameters():
            p.copy_(torch.from_numpy(d[f"{tag}.param.{which}.{k}"]))
    return net, nf, torch.from_numpy(d[f"{tag}.edge_features"]).cuda(), torch.from_numpy(d[f"{tag}.agent_index"]).cuda()


@pytest.mark.parametrize("tensor_cores", [False, True])
@pytest.mark.parametrize("tag", ["u", "b"])
@pytest.mark.parametrize("which", ["edge_mlp", "edge_mlp_test"])
def test_edge_mlp_matches_reference_modules(which, tag, tensor_cores, golden_dir):
    """MPNNPolicyNet.edge_logits (csrc/edge_mlp.cu; edge_mlp's forward also on tcgen05, csrc/edge_mlp_tc.cu) against the
    reference's own modules on the formula of the commented-out update_edges bodies: logits and every parameter
    gradient within 1e-5 relative (atol stated)."""
    d = np.load(os.path.join(golden_dir, "mpnn_edge_mlp.npz"))
    net, nf, ef, ai = _edge_mlp_net(d, tag, which)
    out = net.edge_logits(nf, ef, ai, which=which, tensor_cores=tensor_cores)
    ref = torch.from_numpy(d[f"{tag}.{which}.out"])
    assert out.shape == ref.shape
    close(out, ref, rtol=1e-5, atol=2e-5)
    (out * torch.from_numpy(d[f"{tag}.{which}.w_out"]).cuda()).sum().backward()
    for k, p in getattr(net, which).named_parameters():
        gref = torch.from_numpy(d[f"{tag}.grad.{which}.{k}"])
        close(p.grad, gref, rtol=1e-5, atol=1e-5 * float(gref.abs().max()) + 1e-6)
    if which == "edge_mlp" and tensor_cores:
        from tarl_simulator_b200 import _cabi
        assert net.last_edge_path == ("tcgen05" if _cabi.lib().tarl_edge_mlp_tc_available() else "fp32")
    net.check_errors()


@pytest.mark.parametrize("grid", [False, True])
@pytest.mark.parametrize("B,N,E", [(1, 50, 1), (2, 300, 1000), (32, 2000, 9000), (3, 40, 127)])
def test_edge_mlp_matches_port_on_seeded_inputs(B, N, E, grid):
    """Both MLPs on random graphs / batches against oracle/mpnn_port.edge_mlp_logits evaluated in float64: tile tails, a
    single edge, an edge_attr row shared by the batch (stride 0), fp32 pipe and tensor cores.
    grid=False: random floats, LOGITS within 1e-5 relative. A ReLU net's parameter gradients are discontinuous where a
    pre-activation crosses zero, and among 10^7 units some sit within fp32 rounding of it — no two summation orders agree
    on those masks — so the GRADIENTS are checked on grid=True inputs: everything a multiple of 1/8 in a small range,
    which makes every pre-activation exact in fp32 whatever the order (and exact under the 3xTF32 split), the masks
    identical, and leaves only the rounding of the final sums over the pairs."""
    from tarl_simulator_b200.mpnn_agent import MPNNPolicyNet
    g = torch.Generator().manual_seed(B * 1000 + E + int(grid))
    q = (lambda t, lo, hi: torch.randint(lo, hi + 1, t.shape, generator=g).float() / 8.0) if grid else None
    ei = torch.stack([torch.randint(0, N, (E,), generator=g), torch.randint(0, N, (E,), generator=g)])
    nf = torch.rand(B, N, 7, generator=g) * 4
    ai = torch.randint(0, 30, (B, N), generator=g)
    af = torch.rand(30, 9, generator=g) * 2
    ef = torch.rand(1, E, 1, generator=g)
    if grid:
        nf, af, ef = q(nf, 0, 16), q(af, 0, 8), q(ef, 0, 8)
    ef = ef.expand(B, -1, -1)
    net = MPNNPolicyNet(ei.cuda(), N, torch.ones(E, device="cuda"), "cuda")
    net.agent_features = af.cuda()
    with torch.no_grad():
        for p in list(net.edge_mlp.parameters()) + list(net.edge_mlp_test.parameters()):
            p.copy_((q(p, -4, 4) if grid else torch.randn(p.shape, generator=g) * 0.25).cuda())
    sd = {k: v.detach().cpu().double() for k, v in net.state_dict().items() if k.startswith("edge_mlp")}
    w = q(torch.empty(B, E), -8, 8) if grid else torch.randn(B, E, generator=g)
    for which in ("edge_mlp", "edge_mlp_test"):
        pd = {k: v.clone().requires_grad_(True) for k, v in sd.items() if k.startswith(which + ".")}
        ref = mpnn_port.edge_mlp_logits(pd, nf.double(), ef.double(), af.double(), ai, ei, which)
        (ref * w.double()).sum().backward()
        for tc in ((False, True) if which == "edge_mlp" else (False,)):
            for p in getattr(net, which).parameters():
                p.grad = None
            out = net.edge_logits(nf.cuda(), ef.cuda(), ai.cuda(), which=which, tensor_cores=tc)
            _within(out.detach().cpu(), ref.detach().float(), f"{which} tc={tc} logits")
            if not grid:
                continue
            (out * w.cuda()).sum().backward()
            for k, p in getattr(net, which).named_parameters():
                _within(p.grad.cpu(), pd[f"{which}.{k}"].grad.float(), f"{which} tc={tc} grad {k}")
    net.check_errors()
