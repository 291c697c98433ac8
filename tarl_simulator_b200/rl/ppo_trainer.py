"""PPO training loop over SimulatorEnv / BatchedSimulatorEnv.

Drop-in for the reference's src/rl/ppo_trainer.py:12-160 (`ppo_train` keeps its keyword set). torchrl is not part of
this stack; the pieces the reference takes from torchrl 0.5.0 are restated on plain torch with the same formulas and
the reference's settings:

  collector   frames_per_batch steps per iteration, environment reset at each iteration, actions sampled
  GAE         gamma 0.99, lambda 0.95; delta_t = r_t + gamma V(s_{t+1}) (1 - terminated) - V(s_t);
              A_t = delta_t + gamma lambda (1 - done_t) A_{t+1}; value_target = A + V; then A standardised over the
              whole batch with std clamped at 1e-4 (average_gae=True)
  ClipPPOLoss clip 0.2, entropy bonus 0.01, critic coefficient 1.0 with smooth-L1, no advantage normalisation
  optimiser   Adam(lr 1e-3); per epoch: advantages recomputed, min(sub_batch_size, frames) frames drawn without
              replacement, one step; gradient norm measured, not clipped

Multi-GPU (one process per GPU, torch.distributed initialised by the launcher): every rank rolls out its own
replicas with no communication; per optimiser step ONE all-reduce of the flat fp32 gradient bucket (averaged), plus
a 3-scalar all-reduce for the advantage statistics, so that all ranks apply identical updates.

On a CUDA device nothing of an iteration runs on the host but the launches themselves: the rollout of a
BatchedSimulatorEnv is captured in ONE CUDA graph (all T steps; noise keys live in device words, so every replay draws
anew), the value net sees all (T+1) R frames in one call, GAE is one kernel (csrc/optim.cu), the minibatch is drawn with
a device permutation, parameters and gradients live in one flat bucket each — the all-reduce and the single-launch
Adam step work on them in place — and the logged scalars cross the bus in one copy per iteration. The *_host functions
are the same formulas in plain torch: the specification the kernels are tested against, and what the world-size-2 gloo
tests of the plumbing run on CPU tensors; ppo_train itself never calls them on a CUDA environment.
"""
from __future__ import annotations

import ctypes as C
import json
import os
import time

import torch
import torch.distributed as dist
import torch.nn as nn
import torch.nn.functional as F

from .. import _cabi
from ..distribution import GraphDistribution
from ..feature_helpers import ObservationFeatureHelpers as OBS


def _stream(device) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class PolicyModule(nn.Module):
    """Stands where the reference wires TensorDictModule(policy_net) → ProbabilisticActor(GraphDistribution)
    (src/runner.py:83-95): observation dict → logits → distribution → one-hot action (+ its log-probability)."""

    def __init__(self, policy_net, edge_index, return_log_prob: bool = True):
        super().__init__()
        self.net = policy_net
        self.edge_index = edge_index
        self.return_log_prob = return_log_prob

    def dist(self, obs) -> GraphDistribution:
        logits = self.net(obs["node_features"], obs.get("edge_features"), obs.get("agent_index"))
        return GraphDistribution(logits, self.edge_index)

    def forward(self, obs, mode: bool = False, out: torch.Tensor | None = None, sink=None):
        """out (optional): a bool [.., E] buffer (e.g. one frame of a preallocated trajectory) for the action.
        sink (optional): the environment's ActionSink — see GraphDistribution.sample; the result carries
        "applied": True when the sampling kernel already wrote SELECTED_ROAD."""
        d = self.dist(obs)
        if mode:
            action = d.mode.to(torch.bool)
            if out is not None:
                out.copy_(action.reshape(out.shape))
                action = out
            lp = d.log_prob(action).detach() if self.return_log_prob else None
        elif self.return_log_prob:
            action, lp = d.sample(dtype=torch.bool, out=out, return_log_prob=True, sink=sink)
        else:
            action, lp = d.sample(dtype=torch.bool, out=out, sink=sink), None
        out = {"action": action, "applied": bool(sink is not None and not mode and sink.applied)}
        if lp is not None:
            out["sample_log_prob"] = lp
        return out


class ValueModule(nn.Module):
    """ValueOperator stand-in (src/runner.py:97-105)."""

    def __init__(self, value_net):
        super().__init__()
        self.net = value_net

    def forward(self, obs):
        return self.net(obs["node_features"], obs.get("edge_features"), obs.get("agent_index"), obs["time"])


# ---------------------------------------------------------------------------------------------------------------
def gae_host(value, next_value, reward, done, terminated, gamma=0.99, lmbda=0.95):
    """torchrl 0.5.0 generalized_advantage_estimate over the leading time dimension, in plain torch. All inputs
    [T, ...]. The specification of tarl_gae (tests) — ppo_train uses gae_device."""
    not_term = 1.0 - terminated.to(value.dtype)
    not_done = 1.0 - done.to(value.dtype)
    delta = reward + gamma * next_value * not_term - value
    adv = torch.empty_like(value)
    running = torch.zeros_like(value[0])
    for t in range(value.size(0) - 1, -1, -1):
        running = delta[t] + gamma * lmbda * not_done[t] * running
        adv[t] = running
    return adv, adv + value


def standardise_host(adv, group=None):
    """average_gae=True: (A - mean) / std.clamp_min(1e-4), unbiased std; statistics over every rank's frames. Plain
    torch (specification / gloo tests)."""
    n = torch.tensor(float(adv.numel()), device=adv.device, dtype=torch.float64)
    stats = reduce_stats(torch.stack([n, adv.double().sum(), (adv.double() ** 2).sum()]), group)
    n, s, ss = stats[0], stats[1], stats[2]
    mean = s / n
    var = (ss - n * mean * mean) / torch.clamp(n - 1, min=1.0)
    std = torch.sqrt(torch.clamp(var, min=0.0)).clamp_min(1e-4)
    return ((adv.double() - mean) / std).to(adv.dtype)


def gae_device(v_frames, reward, done, terminated=None, gamma=0.99, lmbda=0.95, group=None):
    """GAE + standardisation on the device (tarl_gae, tarl_standardise). v_frames: V over the T+1 frames of a rollout,
    [T+1, R] contiguous (value = v_frames[:-1], next_value = v_frames[1:]); reward [T, R] fp32; done / terminated [T, R]
    bool. Returns (standardised advantage [T, R], value_target [T, R]). The batch statistics are summed over the ranks
    by one 3-double all-reduce; nothing comes back to the host."""
    if not v_frames.is_cuda:
        raise RuntimeError("gae_device computes on CUDA devices only (gae_host is the plain-torch formula)")
    T, R = reward.shape
    dev = reward.device
    terminated = done if terminated is None else terminated
    v_frames = v_frames.to(torch.float32).contiguous()
    reward = reward.to(torch.float32).contiguous()
    done8 = done.contiguous().view(torch.uint8)
    term8 = terminated.contiguous().view(torch.uint8)
    adv = torch.empty(T, R, dtype=torch.float32, device=dev)
    target = torch.empty(T, R, dtype=torch.float32, device=dev)
    lib = _cabi.lib()
    partials = torch.empty(max(lib.tarl_gae_partial_count(R), 1), 2, dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        rc = lib.tarl_gae(v_frames.data_ptr(), v_frames[1:].data_ptr() if T > 0 else v_frames.data_ptr(), R,
                          reward.data_ptr(), done8.data_ptr(), term8.data_ptr(), T, R, gamma, lmbda, adv.data_ptr(),
                          target.data_ptr(), partials.data_ptr(), _stream(dev))
    _cabi.check(rc, "tarl_gae")
    # (torch.full is a fill kernel; `stats[0] = python_float` would be a synchronising host-to-device copy)
    stats = torch.cat([torch.full((1,), float(T * R), dtype=torch.float64, device=dev), partials.sum(0)])
    reduce_stats(stats, group)
    with torch.cuda.device(dev):
        rc = lib.tarl_standardise(adv.data_ptr(), T * R, stats.data_ptr(), _stream(dev))
    _cabi.check(rc, "tarl_standardise")
    return adv, target


def clip_ppo_loss(log_prob, sample_log_prob, advantage, entropy, value, value_target, clip_epsilon=0.2,
                  entropy_coef=0.01, critic_coef=1.0):
    """torchrl 0.5.0 ClipPPOLoss.forward with the reference's settings, in plain torch: the specification of
    tarl_ppo_clip_loss (tests) — ppo_train uses clip_ppo_loss_device."""
    log_weight = log_prob - sample_log_prob
    ratio = log_weight.exp()
    gain1 = ratio * advantage
    gain2 = ratio.clamp(1.0 - clip_epsilon, 1.0 + clip_epsilon) * advantage
    out = {
        "loss_objective": -torch.minimum(gain1, gain2).mean(),
        "loss_entropy": -entropy_coef * entropy.mean(),
        "loss_critic": critic_coef * F.smooth_l1_loss(value, value_target, reduction="none").mean(),
    }
    with torch.no_grad():
        out["approx_kl"] = (-log_weight).mean()
        out["clip_fraction"] = ((ratio - 1.0).abs() > clip_epsilon).float().mean()
        out["entropy"] = entropy.mean()
    return out


class _ClipLoss(torch.autograd.Function):
    """tarl_ppo_clip_loss behind autograd: forward leaves the six scalars and the three gradient vectors of
    loss_objective + loss_critic + loss_entropy; backward scales them by the incoming gradient of the total."""

    @staticmethod
    def forward(ctx, log_prob, entropy, value, sample_log_prob, advantage, value_target, clip, ent_coef, critic_coef):
        dev, n = log_prob.device, log_prob.numel()
        f32 = lambda x: x.detach().reshape(-1).to(torch.float32).contiguous()
        lp, ent, val = f32(log_prob), f32(entropy), f32(value)
        slp, adv, tgt = f32(sample_log_prob), f32(advantage), f32(value_target)
        if not (ent.numel() == val.numel() == slp.numel() == adv.numel() == tgt.numel() == n):
            raise ValueError("clip_ppo_loss_device: the six inputs must hold one element per frame")
        out = torch.empty(7, dtype=torch.float32, device=dev)
        grads = torch.empty(3, n, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            rc = _cabi.lib().tarl_ppo_clip_loss(lp.data_ptr(), slp.data_ptr(), adv.data_ptr(), ent.data_ptr(), val.data_ptr(),
                                                tgt.data_ptr(), n, float(clip), float(ent_coef), float(critic_coef),
                                                out.data_ptr(), grads[0].data_ptr(), grads[1].data_ptr(),
                                                grads[2].data_ptr(), _stream(dev))
        _cabi.check(rc, "tarl_ppo_clip_loss")
        ctx.save_for_backward(grads)
        ctx.shapes = (log_prob.shape, entropy.shape, value.shape)
        total = out[0] + out[2] + out[1]          # objective + critic + entropy, the order of src/rl/ppo_trainer.py:136-138
        ctx.mark_non_differentiable(out)
        return total, out

    @staticmethod
    def backward(ctx, g_total, _g_out):
        (grads,) = ctx.saved_tensors
        s0, s1, s2 = ctx.shapes
        g = grads * g_total
        return g[0].reshape(s0), g[1].reshape(s1), g[2].reshape(s2), None, None, None, None, None, None


def clip_ppo_loss_device(log_prob, sample_log_prob, advantage, entropy, value, value_target, clip_epsilon=0.2,
                         entropy_coef=0.01, critic_coef=1.0):
    """clip_ppo_loss on the device in ONE launch (tarl_ppo_clip_loss, csrc/optim.cu) — forward and the gradient with
    respect to log_prob, entropy and value. Returns the same dict plus "loss" = loss_objective + loss_critic +
    loss_entropy, the only entry that carries a gradient (the reference differentiates exactly that sum,
    src/rl/ppo_trainer.py:136-139) — and "impossible_frames": how many frames of the minibatch carry GraphDistribution's
    -inf marker in both log-probabilities (an action without an edge in some group, about one draw in 10^8). The torch
    formula turns such a frame into a NaN loss and, one Adam step later, NaN parameters; the kernel leaves it out of
    the objective (declared divergence D8, include/tarl_b200.h)."""
    if not log_prob.is_cuda:
        raise RuntimeError("clip_ppo_loss_device computes on CUDA devices only (clip_ppo_loss is the plain-torch formula)")
    total, out = _ClipLoss.apply(log_prob, entropy, value, sample_log_prob, advantage, value_target, clip_epsilon,
                                 entropy_coef, critic_coef)
    names = ("loss_objective", "loss_entropy", "loss_critic", "approx_kl", "clip_fraction", "entropy", "impossible_frames")
    res = {k: out[i] for i, k in enumerate(names)}
    res["loss"] = total
    return res


def reduce_stats(stats: torch.Tensor, group=None) -> torch.Tensor:
    """{count, sum, sum of squares} of the advantages summed over the ranks, in place (one tiny all-reduce)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(stats, group=group)
    return stats


class GradBucket:
    """Parameters and gradients of a set of nn.Parameters as ONE flat fp32 buffer each (plain torch plumbing, any
    device): every parameter is re-pointed at a slice of `flat` and its .grad at a slice of `grad`, so that the
    backward pass accumulates straight into the buffer the gradient all-reduce — and, on CUDA, the single-launch
    optimiser step — works on: no concatenation, no copy back."""

    def __init__(self, params):
        self.params = list(params)
        if not self.params:
            raise ValueError("no parameters")
        dev = self.params[0].device
        if any(p.device != dev or p.dtype != torch.float32 for p in self.params):
            raise RuntimeError("GradBucket needs fp32 parameters on one device")
        n = sum(p.numel() for p in self.params)
        self.n, self.device = n, dev
        self.flat = torch.empty(n, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(n, dtype=torch.float32, device=dev)
        off = 0
        with torch.no_grad():
            for p in self.params:
                k = p.numel()
                self.flat[off:off + k].copy_(p.detach().reshape(-1))
                p.data = self.flat[off:off + k].view(p.shape)
                p.grad = self.grad[off:off + k].view(p.shape)
                off += k

    def zero_grad(self):
        self.grad.zero_()

    def holds(self, params) -> bool:
        """True when `params` are exactly the parameters of this bucket and still live in it."""
        params = list(params)
        if len(params) != len(self.params) or any(a is not b for a, b in zip(params, self.params)):
            return False
        off = 0
        for p in self.params:
            if p.data_ptr() != self.flat.data_ptr() + 4 * off:
                return False
            off += p.numel()
        return True

    def check_views(self):
        """autograd keeps accumulating into the bucket only while every .grad is still the view handed out here."""
        off = 0
        for p in self.params:
            if p.grad is None or p.grad.data_ptr() != self.grad.data_ptr() + 4 * off:
                raise RuntimeError("a parameter's .grad was replaced: the gradient bucket is out of sync")
            off += p.numel()

    def broadcast(self, src: int = 0, group=None):
        """Every rank starts from rank `src`'s parameters (only gradients are exchanged afterwards)."""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.broadcast(self.flat, src=src, group=group)

    def allreduce(self, group=None) -> int:
        """Sum of the gradient bucket over the ranks, in place (NCCL over NVLink on the GPU box, gloo in the CPU
        tests). Returns the world size: the consumer scales by its inverse."""
        if dist.is_available() and dist.is_initialized():
            world = dist.get_world_size(group)
            if world > 1:
                dist.all_reduce(self.grad, group=group)
            return world
        return 1


class FlatAdam(GradBucket):
    """Adam over the flat bucket in one launch (tarl_adam_step, csrc/optim.cu) — the reference's torch.optim.Adam(lr)
    at src/rl/ppo_trainer.py:39 steps ~8 tensors with ~6 launches each — with the averaging of the all-reduced gradient
    and the global gradient norm the reference logs (:141) folded into the same pass."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        super().__init__(params)
        if self.device.type != "cuda":
            raise RuntimeError("FlatAdam steps on a CUDA device (no CPU fallback)")
        n, dev = self.n, self.device
        self.exp_avg = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(n, dtype=torch.float32, device=dev)
        self.lr, self.betas, self.eps, self.steps = float(lr), (float(betas[0]), float(betas[1])), float(eps), 0
        self.norm_partials = torch.empty(max(_cabi.lib().tarl_adam_partial_count(n), 1), dtype=torch.float64, device=dev)
        self.grad_norm = torch.zeros(1, dtype=torch.float32, device=dev)

    def reset_state(self, lr=None):
        """A fresh optimiser on the same bucket (what constructing torch.optim.Adam anew does at the reference's
        src/rl/ppo_trainer.py:39): moments and step count start over, parameter storage — which captured rollouts
        read — stays where it is."""
        self.exp_avg.zero_(); self.exp_avg_sq.zero_(); self.grad.zero_()
        self.steps = 0
        if lr is not None:
            self.lr = float(lr)
        off = 0
        for p in self.params:
            p.grad = self.grad[off:off + p.numel()].view(p.shape)
            off += p.numel()

    def step(self, group=None):
        """All-reduce of the gradient bucket, then one launch: scale by 1 / world, global gradient norm (left in
        self.grad_norm, a device scalar), Adam update of every parameter."""
        world = self.allreduce(group)
        self.steps += 1
        with torch.cuda.device(self.device):
            rc = _cabi.lib().tarl_adam_step(self.flat.data_ptr(), self.grad.data_ptr(), self.exp_avg.data_ptr(),
                                            self.exp_avg_sq.data_ptr(), self.n, self.lr, self.betas[0], self.betas[1],
                                            self.eps, self.steps, 1.0 / world, self.norm_partials.data_ptr(),
                                            self.grad_norm.data_ptr(), _stream(self.device))
        _cabi.check(rc, "tarl_adam_step")
        for p in self.params:           # the kernel wrote through raw pointers: whoever caches by version (the TF32
            torch.autograd.graph.increment_version(p)      # split of the value net's W1, csrc/value_mlp.cu) must see it
        return self.grad_norm


# ---------------------------------------------------------------------------------------------------------------
class _EnvAdapter:
    """Uniform [R]-batched view of SimulatorEnv (R = 1, reference row layout) and BatchedSimulatorEnv (link store)."""

    @classmethod
    def of(cls, env):
        """The adapter of an environment, kept on it: trajectory buffers, captured rollouts and the parameter bucket
        survive from one ppo_train / collect call to the next."""
        hit = env.__dict__.get("_tarl_adapter")
        if hit is None:
            hit = env.__dict__["_tarl_adapter"] = cls(env)
        return hit

    def __init__(self, env):
        self.env = env
        self.batched = hasattr(env, "store")
        g = env.graph if self.batched else env.simulator.graph
        self.graph = g
        self.R = env.R if self.batched else 1
        self.n_nodes = g.x.size(-2)
        self.device = g.x.device
        Nmax = env.Nmax if self.batched else env.simulator.Nmax
        self.static = g.x[..., 3 * Nmax:].reshape(-1, 7)[: self.n_nodes].clone()      # MAXN .. ROAD_INDEX template
        self.edge_features = g.edge_attr
        self._traj = {}           # (T, slim) -> preallocated trajectory buffers, reused by every rollout of that shape
        self._graphs = {}         # (T, ...) -> captured rollout (torch.cuda.CUDAGraph + what it leaves on the host side)
        self._side = None         # the stream the draws of an overlapped rollout run on

    def side_stream(self):
        if self._side is None:
            self._side = torch.cuda.Stream(self.device)
        return self._side

    def trajectory_buffers(self, T: int, slim: bool):
        """[T+1, R, ..] frames + per-step outputs, allocated once per (T, slim): a rollout writes them in place (frame t+1
        of a step is the next step's frame t), the update reads them, the next rollout overwrites them — and a rollout
        captured in a CUDA graph needs its outputs at fixed addresses anyway."""
        key = (int(T), bool(slim))
        hit = self._traj.get(key)
        if hit is None:
            R, M, dev = self.R, self.n_nodes, self.device
            E = self.graph.edge_index.size(1)
            hit = {"num": torch.empty(T + 1, R, M, dtype=torch.float32, device=dev),
                   "sel": None if slim else torch.empty(T + 1, R, M, dtype=torch.float32, device=dev),
                   "ai": None if slim else torch.empty(T + 1, R, M, dtype=torch.int64, device=dev),
                   "times": torch.empty(T + 1, R, dtype=torch.float32, device=dev),
                   "action": (torch.empty(T, E, R, dtype=torch.bool, device=dev).permute(0, 2, 1) if R > 1 else
                              torch.empty(T, R, E, dtype=torch.bool, device=dev)),
                   "occ": torch.zeros(T, R, dtype=torch.int32, device=dev),
                   "lps": torch.zeros(T, R, dtype=torch.float32, device=dev)}
            if len(self._traj) >= 4:
                self._traj.clear(); self._graphs.clear()
            self._traj[key] = hit
        return hit

    def reset(self):
        if self.batched:
            self.env.reset()
        else:
            self.env._reset()

    def time(self):
        return float(self.env.time if self.batched else self.env.simulator.time)

    def dynamic(self, out=None):
        """(NUM [R, N_tot], SELECTED_ROAD [R, N_tot], head agent id [R, N_tot]) of the current state, optionally
        written into the three given buffers (frames of a preallocated trajectory)."""
        if self.batched:
            return self.env.compact_state(out)
        x, _, _, ai = self.env.simulator.state()
        vals = (x[:, OBS.NUMBER_OF_AGENT].unsqueeze(0), x[:, OBS.SELECTED_ROAD].unsqueeze(0), ai.unsqueeze(0))
        if out is None:
            return vals[0].clone(), vals[1].clone(), vals[2]
        for o, v in zip(out, vals):
            o.copy_(v)
        return out

    def action_sink(self):
        return self.env.action_sink() if self.batched else None

    def step(self, action, out=None, applied: bool = False):
        """action [R, E_full] bool. Returns (reward [R], done [R]); `out` (optional): the three buffers of dynamic()
        for the post-step state (on the link store they are filled by the pass that computes the reward).
        applied=True: the policy step already wrote this action into SELECTED_ROAD (ActionSink)."""
        if self.batched:
            res = self.env.step(None if applied else action, compact_out=out)
            return res["reward"], res["done"]
        res = self.env._step({"action": action[0]})
        if out is not None:
            self.dynamic(out=out)
        return res["reward"].reshape(1).to(torch.float32), res["done"].reshape(1).to(self.device)

    def observation(self, num, sel, agent_index, time, dynamic: bool = True):
        """Observation dict with a leading batch dimension from compact dynamic columns. time: [B] tensor.
        dynamic=False (for consumers that only read static columns, e.g. MPNNPolicyNet's ROAD_INDEX) hands out the
        static template expanded over the batch (stride 0) instead of materialising [B, N_tot, 7]."""
        B = num.size(0)
        if not dynamic:
            nf = self.static.unsqueeze(0).expand(B, -1, -1)
        else:
            nf = self.static.unsqueeze(0).repeat(B, 1, 1)
            nf[..., OBS.NUMBER_OF_AGENT] = num
            if sel is not None:           # None: an occupancy-only trajectory (its consumers do not read this column)
                nf[..., OBS.SELECTED_ROAD] = sel
        return {"node_features": nf, "edge_features": self.edge_features.unsqueeze(0).expand(B, -1, -1),
                "agent_index": agent_index, "time": time.reshape(B, 1).to(torch.float32)}


@torch.no_grad()
def occupancy_only(policy_module, value_module) -> bool:
    """True when neither net reads anything of the dynamic observation but NUMBER_OF_AGENT: the policy's active path
    embeds the static ROAD_INDEX (`reads_dynamic_features = False`) and the value net offers `forward_occupancy`
    (MPNNValueNetSimple, the pair src/runner.py wires). A rollout for them keeps the occupancy frames only."""
    return (not getattr(policy_module.net, "reads_dynamic_features", True)
            and value_module is not None and hasattr(value_module.net, "forward_occupancy"))


def collect(adapter: _EnvAdapter, policy_module: PolicyModule, frames: int, mode: bool = False,
            break_when_any_done: bool = False, occupancy_only: bool = False):
    """`frames` steps of every replica after a reset. Returns a dict of [T, R, ...] tensors (compact observations:
    NUM, SELECTED_ROAD and head ids per node; the static columns are re-attached when a minibatch is formed).
    The trajectory is written in place into preallocated [T+1, R, ..] buffers — frame t+1 of a step is the next
    step's frame t, so `next_*` are views shifted by one — and the one-hot actions go straight from the sampling
    kernel into their frame (edge-major inside a frame: [T, R, E] with the replica innermost).
    occupancy_only=True (see occupancy_only()): SELECTED_ROAD and head-id frames are neither written nor kept
    ("sel" / "agent_index" are None in the result) — 4 instead of 16 bytes per node and frame."""
    R, M, dev = adapter.R, adapter.n_nodes, adapter.device
    T = int(frames)
    slim = bool(occupancy_only) and adapter.batched
    buf = adapter.trajectory_buffers(T, slim)
    num, sel, ai, times, action = buf["num"], buf["sel"], buf["ai"], buf["times"], buf["action"]
    frame = (lambda t: (num[t], None, None)) if slim else (lambda t: (num[t], sel[t], ai[t]))
    dynamic = getattr(policy_module.net, "reads_dynamic_features", True)
    sink = adapter.action_sink() if isinstance(policy_module, PolicyModule) else None
    if adapter.batched and not dynamic and isinstance(policy_module, PolicyModule) and not mode:
        return _collect_static_policy(adapter, policy_module, T, frame, buf, sink, break_when_any_done)
    adapter.reset()
    small = {k: [] for k in ("sample_log_prob", "reward", "done")}
    adapter.dynamic(out=frame(0))
    times[0] = adapter.time()
    n = 0
    for t in range(T):
        obs = adapter.observation(*frame(t), times[t], dynamic=dynamic)
        act = policy_module(obs, mode=mode, out=action[t], sink=sink)
        reward, done = adapter.step(act["action"], out=frame(t + 1), applied=act.get("applied", False))
        times[t + 1] = adapter.time()
        small["sample_log_prob"].append(act.get("sample_log_prob", torch.zeros(R, device=dev)))
        small["reward"].append(reward)
        small["done"].append(done)
        n = t + 1
        if bool(done.any()):
            if break_when_any_done:
                break
            # the reference's SyncDataCollector resets a finished environment and carries on: frame t+1 becomes the
            # first observation of the new episode (step t is marked done, so nothing bootstraps across the seam)
            adapter.reset()
            adapter.dynamic(out=frame(t + 1))
            times[t + 1] = adapter.time()
    out = _trajectory(num, sel, ai, times, action, n)
    out.update({k: torch.stack(v) for k, v in small.items()})
    return out


def _trajectory(num, sel, ai, times, action, n):
    cut = lambda x, a, b: None if x is None else x[a:b]
    out = {"num": num[:n], "sel": cut(sel, 0, n), "agent_index": cut(ai, 0, n), "time": times[:n], "action": action[:n],
           "next_num": num[1:n + 1], "next_sel": cut(sel, 1, n + 1), "next_agent_index": cut(ai, 1, n + 1),
           "next_time": times[1:n + 1]}
    # frame t+1 IS the next observation of step t (no reset inside a rollout): consumers that evaluate something on
    # every observation (the value net in GAE) can do it once over the n+1 frames instead of on both shifted views
    out["_frames"] = {"num": num[:n + 1], "sel": cut(sel, 0, n + 1), "agent_index": cut(ai, 0, n + 1), "time": times[:n + 1]}
    return out


def _collect_static_policy(adapter, policy_module, T, frame, buf, sink, break_when_any_done):
    """The loop of collect() for a policy whose logits do not depend on the dynamic observation (MPNNPolicyNet's
    active path) on the link store: the distribution is built ONCE per rollout (the parameters do not change inside
    one) and a step is one sampling call + one environment step, with no per-step tensors on the host side — rewards
    accumulate in a [T, R] int32 buffer the environment's occupancy pointer walks through, times and done flags are
    host numbers.

    All of it — the policy's logits row, the T sampling launches and the T environment steps — is a fixed sequence of
    launches on fixed buffers: the second rollout of a shape is captured in a CUDA graph and every later one is a single
    replay (at 128 replicas per GPU a step's launches took the host longer, 0.36 ms, than the GPU, 0.27 ms). What varies
    between rollouts lives on the device: the parameters (read by the captured policy kernels) and the two noise keys
    (BatchedSimulatorEnv.begin_rollout). TARL_NO_ROLLOUT_GRAPH=1 keeps the eager loop (same results: same launches)."""
    env, R, dev = adapter.env, adapter.R, adapter.device
    from ..reinforcement_learning import EPISODE_END
    num, sel, ai, action, occ, lps = buf["num"], buf["sel"], buf["ai"], buf["action"], buf["occ"], buf["lps"]
    seed = int(torch.randint(0, 2 ** 62, (1,)))          # torch's default CPU generator: torch.manual_seed reproduces
    env.begin_rollout(seed)
    adapter.reset()
    adapter.dynamic(out=frame(0))
    t0, dt = adapter.time(), float(env.timestep)
    steps_to_done = T
    for k in range(T):                                    # host arithmetic: after which step does the episode end?
        if t0 + dt * (k + 1) > EPISODE_END:
            steps_to_done = k + 1
            break
    n = steps_to_done if (break_when_any_done and steps_to_done < T) else T
    want_lp = bool(policy_module.return_log_prob)
    if sink is not None:
        sink = adapter.action_sink()                      # begin_rollout may have rebuilt it

    # Two streams: the draw of step t+1 (sampling kernel + log-probability finish) depends on nothing the environment
    # computes — static logits, its own noise key — so it runs on a side stream while the main stream steps the
    # environment through step t. The routing decisions alternate between two pairs of SELECTED_ROAD buffers
    # (BatchedSimulatorEnv.sel_pair): draw t writes pair t & 1 (carrying over, from the other pair, the entries of
    # nodes it does not decide), released when the core step of step t-1 is through — i.e. next to the insertion of
    # step t-1, not next to its bandwidth-bound core step. At 128 replicas per GPU the
    # insertion kernels are chains of dependent loads that leave most of the device idle (43 + 20 us per step against
    # 42 + 4 us of sampling): measured 10.7 -> 9.6 ms per 32-step rollout. Captured as a fork / join inside the graph.
    # (Releasing draw t one whole step earlier — when step t-2, the last reader of its pair, is through — and / or a
    # high-priority side stream: 9.41 -> 9.43 ms, measured and dropped: the kernels' summed time, not the join, is
    # what the rollout costs.)
    overlap = sink is not None and not os.environ.get("TARL_NO_ROLLOUT_OVERLAP") and n >= 2

    def draw(t):
        if want_lp:
            _, lp = d_holder[0].sample(dtype=torch.bool, out=action[t], return_log_prob=True, sink=sink)
            lps[t].copy_(lp)
        else:
            d_holder[0].sample(dtype=torch.bool, out=action[t], sink=sink)

    d_holder = [None]

    def body():
        obs = adapter.observation(frame(0)[0], None, None, torch.zeros(R, device=dev), dynamic=False)
        d_holder[0] = policy_module.dist(obs)
        keep_occ = env.occupancy
        try:
            if not overlap:
                for t in range(n):
                    draw(t)
                    applied = sink is not None and sink.applied
                    env.occupancy = occ[t]
                    env.step(None if applied else action[t], compact_out=frame(t + 1), lean=True,
                             direct_insert=applied and R <= 256)
                    if env.time > EPISODE_END and t + 1 < n:  # auto-reset (see collect): not part of a captured rollout
                        adapter.reset()
                        adapter.dynamic(out=frame(t + 1))
                return
            main = torch.cuda.current_stream(dev)
            side = adapter.side_stream()
            N_links, M = env.N, env.n_nodes
            env.sync_sel_pairs()
            side.wait_stream(main)
            core_done = [None] * n

            def mark_core(t):
                def mark():
                    core_done[t] = torch.cuda.Event()
                    core_done[t].record(main)
                return mark

            for t in range(n):
                cur, other = env.sel_pair(t & 1)
                with torch.cuda.stream(side):
                    # not before the core step of step t-1 is through: that is when the device starts to idle (and
                    # step t-2, the last reader of this pair, finished long before)
                    if t >= 1:
                        side.wait_event(core_done[t - 1])
                    sink.retarget(cur[0][: R * N_links].view(R, N_links), cur[1] if M > N_links else None,
                                  other[0][: R * N_links].view(R, N_links), other[1] if M > N_links else None)
                    draw(t)
                    drawn = torch.cuda.Event()
                    drawn.record(side)
                if not sink.applied:
                    raise RuntimeError("the overlapped rollout needs the sampling kernel to write SELECTED_ROAD itself")
                main.wait_event(drawn)
                env.occupancy = occ[t]
                env.step(None, compact_out=frame(t + 1), lean=True, after_core=mark_core(t), direct_insert=R <= 256)
                if env.time > EPISODE_END and t + 1 < n:
                    adapter.reset()
                    adapter.dynamic(out=frame(t + 1))
            main.wait_stream(side)
            last = (n - 1) & 1
            cur, other = env.sel_pair(0)                      # leave the decisions of the last step in pair 0
            if last == 1:
                cur[0].copy_(other[0]); cur[1].copy_(other[1])
            sink.retarget(cur[0][: R * N_links].view(R, N_links), cur[1] if M > N_links else None)
        finally:
            env.occupancy = keep_occ

    crosses_end = steps_to_done < T
    use_graph = (not os.environ.get("TARL_NO_ROLLOUT_GRAPH") and not crosses_end and sink is not None
                 and env.metrics is None and n == T)
    key = (T, want_lp, num.data_ptr(), action.data_ptr(), id(policy_module), t0, dt,
           tuple(p.data_ptr() for p in policy_module.parameters()))
    if not want_lp:
        lps.zero_()
    entry = adapter._graphs.get(key) if use_graph else None
    if entry is not None and entry["graph"] is not None:
        entry["graph"].replay()
        env.time = entry["time"]
        st = env.store
        st.cur, st.step_id, st.t_last = entry["cur"], entry["step_id"], entry["t_last"]
        sink.draw_id, sink.applied = entry["draw_id"], True
    elif use_graph and entry is not None:                 # second rollout of this shape: capture it, then run it
        torch.cuda.synchronize(dev)
        graph = torch.cuda.CUDAGraph()
        st = env.store
        before = (env.time, st.cur, st.step_id, st.t_last, sink.draw_id)
        with torch.cuda.graph(graph):
            body()
        entry.update(graph=graph, time=env.time, cur=st.cur, step_id=st.step_id, t_last=st.t_last, draw_id=sink.draw_id)
        env.time, st.cur, st.step_id, st.t_last, sink.draw_id = before        # capture ran nothing
        graph.replay()
        env.time = entry["time"]
        st.cur, st.step_id, st.t_last = entry["cur"], entry["step_id"], entry["t_last"]
        sink.draw_id = entry["draw_id"]
    else:
        body()
        if use_graph:
            adapter._graphs[key] = {"graph": None}        # seen once (kernels loaded, caches warm): capture next time
    # times and done flags are host arithmetic and the same for every rollout of this shape: built (and copied to the
    # device) once, so that no rollout ends in a host-to-device copy the host has to wait for
    ck = ("clock", n, t0, dt, bool(break_when_any_done))
    clock = buf.get(ck)
    if clock is None:
        host_time, host_done, now = [t0], [], t0
        for k in range(n):
            now += dt
            host_done.append(now > EPISODE_END)
            if now > EPISODE_END and k + 1 < n:           # frames after a seam carry the new episode's clock
                from ..reinforcement_learning import EPISODE_START
                now = float(EPISODE_START)
            host_time.append(now)
        clock = buf[ck] = (torch.tensor(host_time, dtype=torch.float32, device=dev).unsqueeze(1).expand(n + 1, R),
                           torch.tensor(host_done, dtype=torch.bool, device=dev).unsqueeze(1).expand(n, R))
    times, done = clock
    out = _trajectory(num, sel, ai, times, action, n)
    out["sample_log_prob"] = lps[:n]
    out["reward"] = -occ[:n].to(torch.float32)
    out["done"] = done
    return out


def _values(adapter, value_module, batch, prefix=""):
    """V(s) for every frame of the batch: [T, R]. A value net that reads the occupancy only sees all T R frames in ONE
    call (the frames of a rollout are one contiguous [T, R, N_tot] array: a [T R, N_tot] occupancy matrix for the
    tcgen05 kernel, whose grid then has work for every SM even at 128 replicas per GPU); other nets go one time step
    (R frames) at a time to bound the transient memory."""
    T, R = batch["num"].shape[:2]
    fast = getattr(value_module.net, "forward_occupancy", None)
    num, tm = batch[prefix + "num"], batch[prefix + "time"]
    if fast is not None and num.is_contiguous() and T * R <= 65536:
        return fast(num.reshape(T * R, -1), tm.reshape(T * R, 1).to(torch.float32).contiguous()).reshape(T, R)
    out = []
    for t in range(T):
        time = tm[t].reshape(R, 1).to(torch.float32)
        if fast is not None:
            out.append(fast(num[t], time).reshape(R))
        else:
            obs = adapter.observation(num[t], batch[prefix + "sel"][t], batch[prefix + "agent_index"][t], tm[t])
            out.append(value_module(obs).reshape(R))
    return torch.stack(out)


def ppo_train(env, policy_module, value_module, *, total_frames=128, frames_per_batch=32, num_epochs=1,
              sub_batch_size=32, device=None, checkpoint_path=None, log_dir=None, eval_env=None, eval_interval=0,
              log_interval=1, stochastic_eval=False, lr=1e-3, seed=0, history=None):
    """See the module docstring. `total_frames` / `frames_per_batch` count environment steps (each step advances
    every replica of a BatchedSimulatorEnv); `history` (optional list) receives one dict of scalars per iteration."""
    adapter = _EnvAdapter.of(env)
    eval_adapter = _EnvAdapter.of(eval_env) if eval_env is not None else None
    params = [p for p in list(policy_module.parameters()) + list(value_module.parameters()) if p.requires_grad]
    seen, uniq = set(), []
    for p in params:
        if id(p) not in seen:
            seen.add(id(p)); uniq.append(p)
    params = uniq
    optim = adapter.__dict__.get("_optim")
    if optim is not None and optim.holds(params):
        optim.reset_state(lr)         # same nets as the previous call: their storage (and captured rollouts) stay valid
    else:
        optim = adapter._optim = FlatAdam(params, lr=lr)
    optim.broadcast(0)                # ranks seeded differently would otherwise train divergent replicas silently
    rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
    dev = adapter.device
    gen = torch.Generator(device=dev).manual_seed(seed + 7919 * rank)      # minibatch draws, on the device
    writer = None
    log_file = None
    if log_dir is not None and rank == 0:
        os.makedirs(log_dir, exist_ok=True)
        log_file = open(os.path.join(log_dir, "ppo_log.jsonl"), "a")
        try:
            from torch.utils.tensorboard import SummaryWriter
            writer = SummaryWriter(log_dir)
        except Exception:
            writer = None
    global_step = 0
    n_iters = max(-(-total_frames // frames_per_batch), 1)       # the collector yields until total_frames are covered
    for it in range(n_iters):
        t0 = time.perf_counter()
        batch = collect(adapter, policy_module, frames_per_batch, occupancy_only=occupancy_only(policy_module, value_module))
        T, R = batch["reward"].shape
        global_step += frames_per_batch
        rollout_s = time.perf_counter() - t0      # host time to enqueue the rollout (no synchronisation here)
        for _ in range(num_epochs):
            with torch.no_grad():
                if "_frames" in batch:          # V over the T+1 frames once: value = V[:-1], next_value = V[1:]
                    v_all = _values(adapter, value_module, batch["_frames"])
                else:
                    v_all = torch.cat([_values(adapter, value_module, batch),
                                       _values(adapter, value_module, batch, "next_")[-1:]])
                adv, target = gae_device(v_all, batch["reward"], batch["done"], batch["done"])
            n = min(sub_batch_size, T * R)
            pick = torch.randperm(T * R, generator=gen, device=dev)[:n]
            ti, ri = pick // R, pick % R

            def flat(x):          # n frames of a [T, R, ...] trajectory array, whatever its strides
                if x is None:
                    return None
                if x.is_contiguous():     # rows of the [T R, ...] view: one vectorised row gather
                    return x.reshape(T * R, *x.shape[2:]).index_select(0, pick)
                return x[ti, ri]

            obs = adapter.observation(flat(batch["num"]), flat(batch["sel"]), flat(batch["agent_index"]), flat(batch["time"]))
            d = policy_module.dist(obs)
            log_prob = d.log_prob(flat(batch["action"]))
            entropy = d.entropy()
            v = value_module(obs).reshape(n)
            # ClipPPOLoss forward + backward as one launch
            losses = clip_ppo_loss_device(log_prob, flat(batch["sample_log_prob"]), flat(adv), entropy, v, flat(target))
            loss = losses.pop("loss")
            loss.backward()
            optim.check_views()
            grad_norm = optim.step().clone()          # all-reduce + averaged-gradient norm + Adam: one launch
            optim.zero_grad()
        rec = {"iteration": it, "global_step": global_step, "frames": T * R, "rollout_s": round(rollout_s, 4)}
        # every logged scalar crosses the bus in ONE copy, and only when somebody will look at it
        if history is not None or (rank == 0 and it % max(log_interval, 1) == 0 and (log_file or writer)):
            names = ["avg_step_reward", "episode_return", "loss_total", "grad_global_norm"] + list(losses)
            vals = torch.stack([batch["reward"].mean(), batch["reward"].sum(0).mean(), loss.detach(),
                                grad_norm.reshape(())] + [losses[k].detach().to(torch.float32) for k in losses]).cpu()
            rec.update({k: float(x) for k, x in zip(names, vals.tolist())})
        if eval_adapter is not None and eval_interval and it % eval_interval == 0:
            e0 = time.perf_counter()
            ev = collect(eval_adapter, policy_module, frames_per_batch, mode=True, break_when_any_done=True)
            rec["eval_return"] = float(ev["reward"].sum(0).mean())
            rec["eval_episode_len"] = int(ev["reward"].size(0))
            rec["eval_ms"] = round((time.perf_counter() - e0) * 1e3, 2)
            sim = getattr(eval_env, "simulator", None)          # node metrics of the evaluation episode
            if sim is not None:                                 # (src/rl/ppo_trainer.py:117-127)
                counts = sim.hourly_counts()
                if counts is not None:
                    cap = sim.graph.x[: counts.size(0), sim.h.MAX_FLOW].to(torch.float32).clone()
                    cap[cap == 0] = float("nan")
                    vc = counts.float() / cap.unsqueeze(1)
                    rec["eval_avg_vc_mean"] = float(torch.nanmean(torch.nanmean(vc, dim=1)))
                    rec["eval_std_vc_mean"] = float(torch.nanmean(torch.std(vc, dim=1, unbiased=False)))
            if stochastic_eval:
                ev = collect(eval_adapter, policy_module, frames_per_batch, mode=False, break_when_any_done=True)
                rec["eval_stochastic_return"] = float(ev["reward"].sum(0).mean())
        if history is not None:
            history.append(rec)
        if rank == 0 and it % max(log_interval, 1) == 0:
            if log_file is not None:
                log_file.write(json.dumps(rec) + "\n"); log_file.flush()
            if writer is not None:
                for k, v in rec.items():
                    if isinstance(v, (int, float)) and k not in ("iteration", "global_step"):
                        writer.add_scalar(f"PPO/{k}", v, global_step)
    if writer is not None:
        writer.close()
    if log_file is not None:
        log_file.close()
    if checkpoint_path is not None and rank == 0:
        torch.save(reference_state_dict(policy_module), checkpoint_path)
    return history


REFERENCE_POLICY_PREFIX = "module.0.module."


def reference_state_dict(policy_module):
    """policy.pt as the reference writes it (src/rl/ppo_trainer.py:157-159): the state dict of the
    ProbabilisticActor(TensorDictModule(policy_net)) wrapper, i.e. the net's keys under "module.0.module.". A file
    written here therefore loads into the reference's wrapped module, and load_policy_state reads both forms."""
    return {REFERENCE_POLICY_PREFIX + k: v.detach().clone() for k, v in policy_module.net.state_dict().items()}


def load_policy_state(policy_module, state):
    """Loads a policy.pt written by this package or by the reference (wrapper-prefixed keys), or bare net keys."""
    if any(k.startswith(REFERENCE_POLICY_PREFIX) for k in state):
        state = {k[len(REFERENCE_POLICY_PREFIX):]: v for k, v in state.items() if k.startswith(REFERENCE_POLICY_PREFIX)}
    return policy_module.net.load_state_dict(state)
