"""Golden vectors for GraphDistribution and the MPNN nets, from the UNMODIFIED reference behind oracle/shims.
Called by oracle/gen_golden.py. TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import contextlib
import os

import numpy as np
import torch

import ref_loader


@contextlib.contextmanager
def injected_rand(u):
    orig = torch.rand

    def fake(*size, **k):
        shape = tuple(size[0]) if len(size) == 1 and not isinstance(size[0], int) else tuple(size)
        assert shape == tuple(u.shape), (shape, u.shape)
        return u.clone()

    torch.rand = fake
    try:
        yield
    finally:
        torch.rand = orig


def dense_source_graph(g, K, extra):
    """Every node 0..K-1 has >=1 out-edge (ring) plus `extra` random edges, shuffled: the only kind of graph the
    literal GraphDistribution accepts (sources must be exactly 0..K-1)."""
    src = torch.cat([torch.arange(K), torch.randint(0, K, (extra,), generator=g)])
    dst = torch.cat([(torch.arange(K) + 1) % K, torch.randint(0, K, (extra,), generator=g)])
    perm = torch.randperm(src.numel(), generator=g)
    return torch.stack([src[perm], dst[perm]])


def gen_graph_distribution(out):
    rl = ref_loader.load("src.reinforcement_learning")
    cases = {}
    # known-answer case from SURVEY.md §8c(4): 3-node ring
    ei = torch.tensor([[0, 0, 1, 1, 2, 2], [1, 2, 2, 0, 0, 1]])
    logits = torch.tensor([0.0, 1.0, 2.0, 0.0, -1.0, 1.0])
    cases["ring3"] = (ei, logits, 1.0)
    g = torch.Generator().manual_seed(11)
    ei = dense_source_graph(g, 23, 60)
    cases["rand1d"] = (ei, torch.randn(ei.size(1), generator=g) * 2, 1.0)
    ei = dense_source_graph(g, 17, 40)
    cases["rand2d"] = (ei, torch.randn(5, ei.size(1), generator=g) * 3, 0.7)
    ei = dense_source_graph(g, 9, 0)                      # every group has exactly one edge
    cases["single"] = (ei, torch.randn(ei.size(1), generator=g), 1.0)
    blob = {}
    for name, (ei, logits, temp) in cases.items():
        lg = logits.clone().requires_grad_(True)
        d = rl.GraphDistribution(lg, ei, temperature=temp)
        K = d.nb_nodes
        stable = torch.equal(d.index, torch.sort(ei[0], stable=True)[1])
        mode = d.mode.clone()
        if logits.dim() == 1:
            u = torch.rand(K, generator=g)
            with injected_rand(u):
                action = d.sample()
        else:
            # Batched sampling AND batched `mode` are broken in the reference (r_bsum slices the batch axis,
            # src/reinforcement_learning.py:73; `n = arange(B).repeat(K).view(B,K)` scrambles rows, :51-53).
            # Declared divergence D7: the batched contract is "row b behaves like the 1-D distribution of logits[b]",
            # so the golden action/mode rows come from the reference's own 1-D code path, row by row.
            u = torch.rand(logits.size(0), K, generator=g)
            rows_a, rows_m = [], []
            for b in range(logits.size(0)):
                db = rl.GraphDistribution(logits[b], ei, temperature=temp)
                assert torch.equal(db.index, d.index)
                with injected_rand(u[b]):
                    rows_a.append(db.sample())
                rows_m.append(db.mode.clone())
            action, mode = torch.stack(rows_a), torch.stack(rows_m)
        lp = d.log_prob(action)
        ent = d.entropy()
        wl, we = torch.randn(lp.shape, generator=g), torch.randn(ent.shape, generator=g)
        ((lp * wl).sum() + (ent * we).sum()).backward()
        bad = action.clone()
        first = int(torch.nonzero(bad.reshape(-1, bad.size(-1))[0])[0])
        bad.reshape(-1, bad.size(-1))[0, first] = 0        # drop one selection -> impossible action
        lp_bad = d.log_prob(bad)
        blob.update({f"{name}.{k}": v for k, v in dict(
            edge_index=ei.numpy(), logits=logits.numpy(), temperature=np.float64(temp), proba=d.proba.detach().numpy(),
            mode=mode.numpy(), u=u.numpy(), sort_index=d.index.numpy(), action=action.numpy(), stable_sort=np.bool_(stable),
            log_prob=lp.detach().numpy(), entropy=ent.detach().numpy(), w_lp=wl.numpy(), w_ent=we.numpy(),
            grad_logits=lg.grad.numpy(), bad_action=bad.numpy(), log_prob_bad=lp_bad.detach().numpy()).items()})
        print(f"graphdist {name}: K={K} E={ei.size(1)} entropy={ent.detach().flatten()[:2].tolist()} stable_sort={stable}")
    np.savez_compressed(os.path.join(out, "mpnn_graphdist.npz"), **blob)


def _small_net_inputs(g, N, E, A, B=None):
    ei = torch.stack([torch.randint(0, N, (E,), generator=g), torch.randint(0, N, (E,), generator=g)])
    lead = () if B is None else (B,)
    nf = torch.rand(*lead, N, 7, generator=g) * 5
    nf[..., 6] = torch.arange(N).float()                       # ROAD_INDEX >= 0 everywhere (literal reference runs)
    nf[..., 1] = torch.randint(0, 9, (*lead, N), generator=g).float()
    ef = torch.rand(*lead, E, 1, generator=g)
    ai = torch.randint(0, A + 1, (*lead, N), generator=g)
    tm = torch.rand(*lead, 1, generator=g) * 10
    af = torch.rand(A + 1, 9, generator=g) * 3
    return ei, nf, ef, ai, tm, af


def gen_nets(out):
    mp = ref_loader.load("src.agents.mpnn_agent")
    g = torch.Generator().manual_seed(21)
    blob = {}
    for tag, B in (("u", None), ("b", 3)):
        N, E, A = 14, 37, 20
        ei, nf, ef, ai, tm, af = _small_net_inputs(g, N, E, A, B)
        torch.manual_seed(5)
        # --- MPNNValueNet (eval mode: dropout off)
        net = mp.MPNNValueNet(ei, N, "cpu")
        net.agent_features = af
        net.eval()
        v = net(nf, ef, ai, tm)
        wv = torch.randn(v.shape, generator=g)
        (v * wv).sum().backward()
        sd = {k: p.detach().clone() for k, p in net.named_parameters()}
        gr = {k: p.grad.clone() for k, p in net.named_parameters()}
        blob.update({f"value.{tag}.{k}": t.numpy() for k, t in dict(
            edge_index=ei, node_features=nf, edge_features=ef, agent_index=ai, time=tm, agent_features=af,
            out=v.detach(), w_out=wv).items()})
        blob.update({f"value.{tag}.param.{k}": t.numpy() for k, t in sd.items()})
        blob.update({f"value.{tag}.grad.{k}": t.numpy() for k, t in gr.items()})
        # --- MPNNValueNetSimple
        net = mp.MPNNValueNetSimple(ei, N, "cpu")
        v = net(nf, ef, ai, tm)
        wv = torch.randn(v.shape, generator=g)
        (v * wv).sum().backward()
        blob.update({f"simple.{tag}.out": v.detach().numpy(), f"simple.{tag}.w_out": wv.numpy()})
        blob.update({f"simple.{tag}.param.{k}": p.detach().numpy() for k, p in net.named_parameters()})
        blob.update({f"simple.{tag}.grad.{k}": p.grad.numpy() for k, p in net.named_parameters()})
        # --- MPNNPolicyNet (active path = embedding of ROAD_INDEX gathered at edge targets)
        ff = torch.rand(E, generator=g) + 0.5
        net = mp.MPNNPolicyNet(ei, N, ff, "cpu")
        net.agent_features = af
        lg = net(nf, ef, ai)
        wl = torch.randn(lg.shape, generator=g)
        (lg * wl).sum().backward()
        blob.update({f"policy.{tag}.out": lg.detach().numpy(), f"policy.{tag}.w_out": wl.numpy(),
                     f"policy.{tag}.emb": net.nodes_embedding.weight.detach().numpy(),
                     f"policy.{tag}.grad_emb": net.nodes_embedding.weight.grad.numpy(),
                     f"policy.{tag}.param_names": np.array(sorted(k for k, _ in net.named_parameters()))})
        print(f"nets {tag}: value={v.detach().flatten().tolist()[:2]} logits[:3]={lg.detach().flatten()[:3].tolist()}")
    np.savez_compressed(os.path.join(out, "mpnn_nets.npz"), **blob)


def gen_value_train(out):
    """MPNNValueNet in TRAIN mode: the message dropout (nn.Dropout(0.05) on the [B*E, 17] message input,
    src/agents/mpnn_agent.py:278) draws from torch's global generator. The mask the unmodified reference used is
    recovered by replaying the same draw (same seed, same shape) and stored with the outputs; the two dropouts of
    time_net are set to p = 0 on the instance so that the output is a function of that one mask."""
    mp = ref_loader.load("src.agents.mpnn_agent")
    import mpnn_port
    g = torch.Generator().manual_seed(33)
    blob = {}
    for tag, B in (("u", None), ("b", 4)):
        N, E, A = 14, 37, 20
        ei, nf, ef, ai, tm, af = _small_net_inputs(g, N, E, A, B)
        torch.manual_seed(7)
        net = mp.MPNNValueNet(ei, N, "cpu")
        net.agent_features = af
        net.train()
        net.time_net[1].p = 0.0
        net.time_net[4].p = 0.0
        rows = (B or 1) * E
        torch.manual_seed(1234)
        keep = torch.nn.functional.dropout(torch.ones(rows, 17), net.message_mlp[0].p, True) != 0
        keep = keep.view(*(() if B is None else (B,)), E, 17)
        torch.manual_seed(1234)
        v = net(nf, ef, ai, tm)
        wv = torch.randn(v.shape, generator=g)
        (v * wv).sum().backward()
        sd = {k: p.detach().clone() for k, p in net.named_parameters()}
        chk = mpnn_port.value_net_forward(sd, nf, ef, af, ai, tm, ei, keep=keep)
        assert torch.allclose(chk, v.detach(), rtol=1e-6, atol=1e-7), "recovered mask is not the one the reference drew"
        blob.update({f"{tag}.{k}": t.numpy() for k, t in dict(
            edge_index=ei, node_features=nf, edge_features=ef, agent_index=ai, time=tm, agent_features=af,
            keep_bits=mpnn_port.pack_keep_bits(keep), out=v.detach(), w_out=wv).items()})
        blob.update({f"{tag}.param.{k}": t.numpy() for k, t in sd.items()})
        blob.update({f"{tag}.grad.{k}": p.grad.numpy() for k, p in net.named_parameters()})
        print(f"value train {tag}: dropped {int((~keep).sum())} of {keep.numel()} inputs, value={v.detach().flatten().tolist()[:2]}")
    np.savez_compressed(os.path.join(out, "mpnn_value_train.npz"), **blob)


def gen_edge_mlp(out):
    """MPNNPolicyNet.edge_mlp / edge_mlp_test evaluated by the UNMODIFIED reference modules on the formula of the two
    commented-out bodies of update_edges (src/agents/mpnn_agent.py:220-231), with the x forward() assembles (:163-167):
    outputs and parameter gradients, unbatched and batched (the reference batches by offsetting node ids: row b of the
    batched result is the unbatched result of sample b)."""
    mp = ref_loader.load("src.agents.mpnn_agent")
    g = torch.Generator().manual_seed(44)
    blob = {}
    for tag, B in (("u", None), ("b", 3)):
        N, E, A = 19, 301, 25                                  # E spans three 128-pair tiles with a ragged tail
        ei, nf, ef, ai, tm, af = _small_net_inputs(g, N, E, A, B)
        torch.manual_seed(9)
        net = mp.MPNNPolicyNet(ei, N, torch.rand(E, generator=g) + 0.5, "cpu")
        net.agent_features = af
        with torch.no_grad():                                   # the reference's init is +-0.1 with zero biases: widen it
            for p in list(net.edge_mlp.parameters()) + list(net.edge_mlp_test.parameters()):
                p.copy_(torch.randn(p.shape, generator=g) * (0.3 if p.dim() == 2 else 0.2))
        rows = [None] if B is None else list(range(B))
        outs = {"edge_mlp": [], "edge_mlp_test": []}
        for b in rows:
            sl = (lambda t: t) if b is None else (lambda t: t[b])
            x = torch.cat((sl(nf), af[sl(ai)]), dim=-1)                                    # :181-184
            x_i, x_j = x[ei[0]], x[ei[1]]                                                  # :228-229
            outs["edge_mlp"].append(net.edge_mlp(torch.cat([x_i, x_j, sl(ef)], dim=1)).squeeze(-1))   # :230-231
            outs["edge_mlp_test"].append(net.edge_mlp_test(torch.cat([x_i, x_j], dim=1)).squeeze(-1))  # :224-225 on x
        for which, seq in (("edge_mlp", net.edge_mlp), ("edge_mlp_test", net.edge_mlp_test)):
            o = outs[which][0] if B is None else torch.stack(outs[which])
            w = torch.randn(o.shape, generator=g)
            for p in seq.parameters():
                p.grad = None
            (o * w).sum().backward()
            blob[f"{tag}.{which}.out"] = o.detach().numpy()
            blob[f"{tag}.{which}.w_out"] = w.numpy()
            for k, p in seq.named_parameters():
                blob[f"{tag}.param.{which}.{k}"] = p.detach().numpy()
                blob[f"{tag}.grad.{which}.{k}"] = p.grad.numpy()
        blob.update({f"{tag}.{k}": t.numpy() for k, t in dict(edge_index=ei, node_features=nf, edge_features=ef,
                                                             agent_index=ai, agent_features=af).items()})
        print(f"edge mlp {tag}: logits[:3]={blob[f'{tag}.edge_mlp.out'].reshape(-1)[:3].tolist()}")
    np.savez_compressed(os.path.join(out, "mpnn_edge_mlp.npz"), **blob)


def main(out, only_new=False):
    if not only_new:
        gen_graph_distribution(out)
        gen_nets(out)
    gen_value_train(out)
    gen_edge_mlp(out)
