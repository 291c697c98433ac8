"""Pins the CPU oracle (oracle/core_port.py, oracle/mpnn_port.py) against golden vectors that were produced by the
UNMODIFIED reference (oracle/gen_golden.py) and — when /root/reference is present — against the live reference."""
import glob
import os

import numpy as np
import pytest
import torch

import cases
import core_port
import mpnn_port
import ref_loader

CORE_CASES = sorted(os.path.basename(f)[:-4] for f in glob.glob(os.path.join(os.path.dirname(__file__), "golden", "core_*.npz")))


def replay_core_case(d, step_fn):
    """Replay a golden trajectory with `step_fn(x, ei, w, t, Nmax, u, cc) -> dict(delta_tt, pop)`; bit-exact asserts."""
    x = torch.from_numpy(d["x0"]).clone()
    ei, w = torch.from_numpy(d["edge_index"]), torch.from_numpy(d["edge_attr"])
    Nmax = int(d["Nmax"])
    c = core_port.Cols(Nmax)
    cc = core_port.static_factors(x, c)[1] if bool(d["use_static"]) else None
    for s in range(len(d["t"])):
        x[:, c.SEL] = torch.from_numpy(d["sel"][s])
        out = step_fn(x, ei, w, float(d["t"][s]), Nmax, torch.from_numpy(d["u"][s]), cc)
        assert torch.equal(x, torch.from_numpy(d["x"][s])), f"x differs after step {s}"
        assert torch.equal(out["delta_tt"], torch.from_numpy(d["delta_tt"][s]))
        assert (out["pop"] is not None) == bool(d["has_pop"][s])
        if out["pop"] is not None:
            assert torch.equal(out["pop"], torch.from_numpy(d["pop"][s]))


@pytest.mark.parametrize("name", CORE_CASES)
def test_core_port_matches_reference_goldens(name, golden_dir):
    assert CORE_CASES, "golden vectors missing"
    replay_core_case(np.load(os.path.join(golden_dir, name + ".npz")), core_port.core_step)


def test_braess_known_answer(golden_dir):
    """on the reference's braess fixture one core step changes exactly three cells."""
    d = np.load(os.path.join(golden_dir, "core_braess.npz"))
    x0, x1 = d["x0"], d["x"][0]
    changed = np.argwhere(x0 != x1)
    assert changed.tolist() == [[0, 201], [1, 201], [2, 202]]
    assert x1[0, 201] == np.float32(3.2704544067382812)
    assert x1[1, 201] == np.float32(1.0906565189361572)
    assert x1[2, 202] == np.float32(1.199722170829773)
    assert not d["has_pop"].any() and (d["delta_tt"][0] == 0).all()


@pytest.mark.skipif(not ref_loader.available(), reason="needs /root/reference (authoring container only)")
@pytest.mark.parametrize("seed", [101, 102, 103])
def test_core_port_matches_live_reference(seed):
    import gen_golden
    g = torch.Generator().manual_seed(seed)
    N, Nmax = 150, 12
    ei, w = cases.random_dual_graph(g, N, 5, sort_by_source=bool(seed % 2))
    x0, _ = cases.random_road_state(g, N, Nmax, 50.0, ei)
    g2 = torch.Generator().manual_seed(seed + 1)
    rec = gen_golden.reference_core_trajectory(x0, ei, w, Nmax, 50.0, 6, g2, use_static=True)
    x = x0.clone()
    c = core_port.Cols(Nmax)
    cc = core_port.static_factors(x, c)[1]
    for s in range(6):
        x[:, c.SEL] = rec["sel"][s]
        out = core_port.core_step(x, ei, w, rec["t"][s], Nmax, rec["u"][s], cc)
        assert torch.equal(x, rec["x"][s])
        assert torch.equal(out["delta_tt"], rec["delta_tt"][s])


# ----------------------------------------------------------------------------------------------------------------
GD_CASES = ["ring3", "rand1d", "rand2d", "single"]


@pytest.mark.parametrize("name", GD_CASES)
def test_graph_distribution_port(name, golden_dir):
    z = np.load(os.path.join(golden_dir, "mpnn_graphdist.npz"))
    g = lambda k: torch.from_numpy(z[f"{name}.{k}"])
    lg = g("logits").clone().requires_grad_(True)
    d = mpnn_port.GraphDistributionPort(lg, g("edge_index"), float(z[f"{name}.temperature"]), sort_index=g("sort_index"))
    torch.testing.assert_close(d.proba.detach(), g("proba"), rtol=1e-6, atol=1e-7)
    assert torch.equal(d.mode, g("mode"))
    if lg.dim() == 1:
        assert torch.equal(d.sample(g("u")), g("action"))
    lp, ent = d.log_prob(g("action")), d.entropy()
    torch.testing.assert_close(lp.detach(), g("log_prob"), rtol=1e-6, atol=1e-6)
    torch.testing.assert_close(ent.detach(), g("entropy"), rtol=1e-6, atol=1e-6)
    ((lp * g("w_lp")).sum() + (ent * g("w_ent")).sum()).backward()
    torch.testing.assert_close(lg.grad, g("grad_logits"), rtol=1e-5, atol=1e-6)
    assert torch.equal(d.log_prob(g("bad_action")).detach(), g("log_prob_bad"))
    # stable order (the contract) gives the same group-wise quantities as the reference's unstable order
    d2 = mpnn_port.GraphDistributionPort(g("logits"), g("edge_index"), float(z[f"{name}.temperature"]))
    torch.testing.assert_close(d2.entropy(), g("entropy"), rtol=1e-6, atol=1e-6)
    torch.testing.assert_close(d2.log_prob(g("action")), g("log_prob"), rtol=1e-6, atol=1e-6)


def test_graph_distribution_known_answer(golden_dir):
    """SURVEY.md §8c(4)."""
    z = np.load(os.path.join(golden_dir, "mpnn_graphdist.npz"))
    assert abs(float(z["ring3.entropy"][0]) - 1.3129) < 1e-4
    assert z["ring3.mode"].tolist() == [0, 1, 1, 0, 0, 1]


@pytest.mark.parametrize("tag", ["u", "b"])
def test_value_net_train_mode_port(tag, golden_dir):
    """MPNNValueNet in train mode: the port with the message-dropout mask the unmodified reference drew (recovered by
    oracle/gen_golden_mpnn.py) reproduces the reference's output and parameter gradients."""
    z = np.load(os.path.join(golden_dir, "mpnn_value_train.npz"))
    g = lambda k: torch.from_numpy(z[k])
    ei = g(f"{tag}.edge_index")
    nf, ef, ai, tm, af = (g(f"{tag}.{k}") for k in ("node_features", "edge_features", "agent_index", "time", "agent_features"))
    keep = mpnn_port.unpack_keep_bits(g(f"{tag}.keep_bits"))
    assert 0 < int((~keep).sum()) < keep.numel() // 8            # the golden really drops something
    assert torch.equal(mpnn_port.pack_keep_bits(keep), g(f"{tag}.keep_bits"))
    names = [k[len(f"{tag}.param."):] for k in z.files if k.startswith(f"{tag}.param.")]
    p = {k: g(f"{tag}.param.{k}").clone().requires_grad_(True) for k in names}
    v = mpnn_port.value_net_forward(p, nf, ef, af, ai, tm, ei, keep=keep)
    torch.testing.assert_close(v.detach(), g(f"{tag}.out"), rtol=1e-5, atol=1e-6)
    (v * g(f"{tag}.w_out")).sum().backward()
    for k in names:
        torch.testing.assert_close(p[k].grad, g(f"{tag}.grad.{k}"), rtol=1e-4, atol=1e-6)
    # and the mask matters: eval mode gives a different value
    v_eval = mpnn_port.value_net_forward(p, nf, ef, af, ai, tm, ei)
    assert not torch.allclose(v_eval.detach(), g(f"{tag}.out"), rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("tag", ["u", "b"])
def test_nets_port(tag, golden_dir):
    z = np.load(os.path.join(golden_dir, "mpnn_nets.npz"))
    g = lambda k: torch.from_numpy(z[k])
    ei = g(f"value.{tag}.edge_index")
    nf, ef, ai, tm, af = (g(f"value.{tag}.{k}") for k in ("node_features", "edge_features", "agent_index", "time", "agent_features"))
    # MPNNValueNet
    names = [k[len(f"value.{tag}.param."):] for k in z.files if k.startswith(f"value.{tag}.param.")]
    p = {k: g(f"value.{tag}.param.{k}").clone().requires_grad_(True) for k in names}
    v = mpnn_port.value_net_forward(p, nf, ef, af, ai, tm, ei)
    torch.testing.assert_close(v.detach(), g(f"value.{tag}.out"), rtol=1e-5, atol=1e-6)
    (v * g(f"value.{tag}.w_out")).sum().backward()
    for k in names:
        torch.testing.assert_close(p[k].grad, g(f"value.{tag}.grad.{k}"), rtol=1e-4, atol=1e-6)
    # MPNNValueNetSimple
    names = [k[len(f"simple.{tag}.param."):] for k in z.files if k.startswith(f"simple.{tag}.param.")]
    p = {k: g(f"simple.{tag}.param.{k}").clone().requires_grad_(True) for k in names}
    v = mpnn_port.value_simple_forward(p, nf, tm)
    torch.testing.assert_close(v.detach(), g(f"simple.{tag}.out"), rtol=1e-5, atol=1e-6)
    (v * g(f"simple.{tag}.w_out")).sum().backward()
    for k in names:
        torch.testing.assert_close(p[k].grad, g(f"simple.{tag}.grad.{k}"), rtol=1e-4, atol=1e-6)
    # MPNNPolicyNet active path
    emb = g(f"policy.{tag}.emb").clone().requires_grad_(True)
    lg = mpnn_port.policy_logits(emb, nf, ei)
    assert torch.equal(lg.detach(), g(f"policy.{tag}.out"))
    (lg * g(f"policy.{tag}.w_out")).sum().backward()
    torch.testing.assert_close(emb.grad, g(f"policy.{tag}.grad_emb"), rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("tag", ["u", "b"])
@pytest.mark.parametrize("which", ["edge_mlp", "edge_mlp_test"])
def test_edge_mlp_port_matches_reference_modules(tag, which, golden_dir):
    """oracle/mpnn_port.edge_mlp_logits against what the reference's own edge_mlp / edge_mlp_test modules return for
    the formula of the commented-out update_edges bodies (tests/golden/mpnn_edge_mlp.npz), outputs and gradients."""
    d = np.load(os.path.join(golden_dir, "mpnn_edge_mlp.npz"))
    g = lambda k: torch.from_numpy(d[f"{tag}.{k}"])
    p = {k[len(f"{tag}.param."):]: torch.from_numpy(d[k]).clone().requires_grad_(True)
         for k in d.files if k.startswith(f"{tag}.param.{which}.")}
    out = mpnn_port.edge_mlp_logits(p, g("node_features"), g("edge_features"), g("agent_features"), g("agent_index"),
                                    g("edge_index"), which)
    torch.testing.assert_close(out, g(f"{which}.out"), rtol=1e-6, atol=1e-6)
    (out * g(f"{which}.w_out")).sum().backward()
    for k, v in p.items():
        gref = torch.from_numpy(d[f"{tag}.grad.{k}"])            # sums over ~1e3 pairs with cancellation: atol scales with the tensor
        torch.testing.assert_close(v.grad, gref, rtol=1e-5, atol=1e-5 * max(1.0, float(gref.abs().max())))


def test_largest_uniform_wins_the_gumbel_race_when_the_gap_is_clear():
    """The premise of the direction kernel's logarithm-free pick (csrc/engine.cu, kClearGap): among edges that carry ONE
    weight w, the oracle's fp32 score log(w + 1e-12) + (-log(-log u)) (oracle/core_port.py:79-80, src/direction_mpnn.py:
    137-138) orders two edges exactly as their uniforms do whenever those are at least 1e-4 apart — checked on the
    worst case, the nearest pair of the kernel's 2^-23 uniform grid that is still 1e-4 apart, for 4 M random positions
    per weight. The smallest score gap must stay far above the evaluation error of a score (< 5e-6)."""
    g = torch.Generator().manual_seed(3)
    worst = 1.0
    for w in (0.25, 1.0 / 3.0, 0.5, 1.0, 1e-3):
        lpa = torch.log(torch.tensor(w, dtype=torch.float32) + 1e-12)
        k = torch.randint(0, 2 ** 23, (4_000_000,), generator=g)
        u1 = ((k.double() + 0.5) / 2 ** 23).float()
        k2 = torch.floor((u1.double() - 1e-4) * 2 ** 23 - 0.5).long()
        u2 = ((k2.clamp_min(0).double() + 0.5) / 2 ** 23).float()
        ok = (k2 >= 0) & ((u1 - u2) >= 1e-4)                      # the kernel's own test, in fp32
        s1 = lpa + (-torch.log(-torch.log(u1[ok])))
        s2 = lpa + (-torch.log(-torch.log(u2[ok])))
        assert int(ok.sum()) > 3_900_000 and bool((s1 > s2).all())
        worst = min(worst, float((s1 - s2).min()))
    assert worst > 2.5e-4                                        # e * 1e-4 = 2.7e-4 in exact arithmetic


def test_logarithm_free_pick_rule_equals_the_literal_arg_max():
    """The decision rule of k_ell_select_append on links whose in-edges share one weight (csrc/engine.cu), restated in
    torch: winner = eligible edge with the largest uniform when the two largest are >= 1e-4 apart, else the literal
    fp32 scores with strict '>' in ascending edge id — against the oracle's literal arg-max over the eligible edges
    (oracle/core_port.py:79-84) on 2 M random fan-ins of 2..4 eligible edges, with near-ties and exact ties of the
    uniforms planted (the kernel's 2^-23 grid makes exact ties a one-in-10^7 event per pair)."""
    g = torch.Generator().manual_seed(11)
    n, W = 2_000_000, 4
    k = torch.randint(0, 2 ** 23, (n, W), generator=g)
    k[: n // 20, 1] = k[: n // 20, 0]                                         # exact ties: lowest edge id must win
    k[n // 20: n // 10, 2] = (k[n // 20: n // 10, 0] + torch.randint(-400, 401, (n // 10 - n // 20,), generator=g)).clamp(0, 2 ** 23 - 1)
    u = ((k.double() + 0.5) / 2 ** 23).float()                                # near ties: inside the 1e-4 gap
    elig = torch.rand(n, W, generator=g) < 0.7
    elig[:, 0] = True; elig[:, 1] |= ~elig[:, 1:].any(1)                       # at least two eligible edges
    w = torch.tensor([0.25, 1.0 / 3.0, 0.5, 1.0])[torch.randint(0, 4, (n,), generator=g)].unsqueeze(1)
    score = torch.log(w + 1e-12) + (-torch.log(-torch.log(u)))                # the oracle's fp32 scores
    literal = torch.where(elig, score, torch.full_like(score, float("-inf"))).argmax(1)   # first maximum = lowest id
    assert bool((torch.where(elig, score, torch.full_like(score, float("-inf"))).max(1).values > float("-inf")).all())
    # torch.argmax returns the first of equal maxima on CPU; make that explicit
    best = torch.full((n,), float("-inf")); lit = torch.zeros(n, dtype=torch.long)
    for j in range(W):
        better = elig[:, j] & (score[:, j] > best)
        best = torch.where(better, score[:, j], best); lit = torch.where(better, torch.full_like(lit, j), lit)
    assert torch.equal(lit, literal)
    m1 = torch.full((n,), -1.0); m2 = torch.full((n,), -1.0); j1 = torch.zeros(n, dtype=torch.long)
    for j in range(W):                                                        # the kernel's running two largest
        uj = u[:, j]
        top = elig[:, j] & (uj > m1)
        second = elig[:, j] & ~top & (uj > m2)
        m2 = torch.where(top, m1, torch.where(second, uj, m2))
        j1 = torch.where(top, torch.full_like(j1, j), j1)
        m1 = torch.where(top, uj, m1)
    clear = (m1 - m2) >= 1e-4
    rule = torch.where(clear, j1, lit)
    assert torch.equal(rule, lit)
    assert 0.02 < float((~clear).float().mean()) < 0.2                        # the planted ties took the literal path
