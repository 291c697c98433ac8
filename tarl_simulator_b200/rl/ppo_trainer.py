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
"""
from __future__ import annotations

import json
import os
import time

import torch
import torch.distributed as dist
import torch.nn as nn
import torch.nn.functional as F

from ..distribution import GraphDistribution
from ..feature_helpers import ObservationFeatureHelpers as OBS


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
def gae(value, next_value, reward, done, terminated, gamma=0.99, lmbda=0.95):
    """torchrl 0.5.0 generalized_advantage_estimate over the leading time dimension. All inputs [T, ...]."""
    not_term = 1.0 - terminated.to(value.dtype)
    not_done = 1.0 - done.to(value.dtype)
    delta = reward + gamma * next_value * not_term - value
    adv = torch.empty_like(value)
    running = torch.zeros_like(value[0])
    for t in range(value.size(0) - 1, -1, -1):
        running = delta[t] + gamma * lmbda * not_done[t] * running
        adv[t] = running
    return adv, adv + value


def standardise(adv, group=None):
    """average_gae=True: (A - mean) / std.clamp_min(1e-4), unbiased std; statistics over every rank's frames."""
    n = torch.tensor(float(adv.numel()), device=adv.device, dtype=torch.float64)
    stats = torch.stack([n, adv.double().sum(), (adv.double() ** 2).sum()])
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(stats, group=group)
    n, s, ss = stats[0], stats[1], stats[2]
    mean = s / n
    var = (ss - n * mean * mean) / torch.clamp(n - 1, min=1.0)
    std = torch.sqrt(torch.clamp(var, min=0.0)).clamp_min(1e-4)
    return ((adv.double() - mean) / std).to(adv.dtype)


def clip_ppo_loss(log_prob, sample_log_prob, advantage, entropy, value, value_target, clip_epsilon=0.2,
                  entropy_coef=0.01, critic_coef=1.0):
    """torchrl 0.5.0 ClipPPOLoss.forward with the reference's settings."""
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


def allreduce_gradients(params, group=None):
    """One flat fp32 bucket, averaged over ranks (NCCL over NVLink on the GPU box, gloo in the CPU tests)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, group=group)
    flat /= dist.get_world_size(group)
    off = 0
    for g in grads:
        g.copy_(flat[off:off + g.numel()].view_as(g))
        off += g.numel()


# ---------------------------------------------------------------------------------------------------------------
class _EnvAdapter:
    """Uniform [R]-batched view of SimulatorEnv (R = 1, reference row layout) and BatchedSimulatorEnv (link store)."""

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
    adapter.reset()
    R, M, dev = adapter.R, adapter.n_nodes, adapter.device
    E = adapter.graph.edge_index.size(1)
    T = int(frames)
    num = torch.empty(T + 1, R, M, dtype=torch.float32, device=dev)
    slim = bool(occupancy_only) and adapter.batched
    sel = None if slim else torch.empty(T + 1, R, M, dtype=torch.float32, device=dev)
    ai = None if slim else torch.empty(T + 1, R, M, dtype=torch.int64, device=dev)
    frame = (lambda t: (num[t], None, None)) if slim else (lambda t: (num[t], sel[t], ai[t]))
    times = torch.empty(T + 1, R, dtype=torch.float32, device=dev)
    if R > 1:
        action = torch.empty(T, E, R, dtype=torch.bool, device=dev).permute(0, 2, 1)
    else:
        action = torch.empty(T, R, E, dtype=torch.bool, device=dev)
    small = {k: [] for k in ("sample_log_prob", "reward", "done")}
    adapter.dynamic(out=frame(0))
    times[0] = adapter.time()
    dynamic = getattr(policy_module.net, "reads_dynamic_features", True)
    sink = adapter.action_sink() if isinstance(policy_module, PolicyModule) else None
    n = 0
    if adapter.batched and not dynamic and isinstance(policy_module, PolicyModule) and not mode:
        return _collect_static_policy(adapter, policy_module, T, frame, num, sel, ai, action, sink, break_when_any_done)
    for t in range(T):
        obs = adapter.observation(*frame(t), times[t], dynamic=dynamic)
        act = policy_module(obs, mode=mode, out=action[t], sink=sink)
        reward, done = adapter.step(act["action"], out=frame(t + 1), applied=act.get("applied", False))
        times[t + 1] = adapter.time()
        small["sample_log_prob"].append(act.get("sample_log_prob", torch.zeros(R, device=dev)))
        small["reward"].append(reward)
        small["done"].append(done)
        n = t + 1
        if break_when_any_done and bool(done.any()):
            break
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


def _collect_static_policy(adapter, policy_module, T, frame, num, sel, ai, action, sink, break_when_any_done):
    """The loop of collect() for a policy whose logits do not depend on the dynamic observation (MPNNPolicyNet's
    active path) on the link store: the distribution is built ONCE per rollout (the parameters do not change inside
    one) and a step is one sampling call + one environment step, with no per-step tensors on the host side — rewards
    accumulate in a [T, R] int32 buffer the environment's occupancy pointer walks through, times and done flags are
    host numbers turned into tensors at the end. At 128 replicas per GPU the per-step host work (0.41 ms) was longer
    than the step's kernels (0.32 ms)."""
    env, R, dev = adapter.env, adapter.R, adapter.device
    from ..reinforcement_learning import EPISODE_END
    obs = adapter.observation(frame(0)[0], None, None, torch.zeros(R, device=dev), dynamic=False)
    d = policy_module.dist(obs)
    occ = torch.zeros(T, R, dtype=torch.int32, device=dev)
    lps = torch.empty(T, R, dtype=torch.float32, device=dev)
    host_time, host_done = [adapter.time()], []
    keep_occ = env.occupancy
    n = 0
    try:
        for t in range(T):
            if policy_module.return_log_prob:
                _, lp = d.sample(dtype=torch.bool, out=action[t], return_log_prob=True, sink=sink)
                lps[t] = lp
            else:
                d.sample(dtype=torch.bool, out=action[t], sink=sink)
                lps[t] = 0.0
            applied = sink is not None and sink.applied
            env.occupancy = occ[t]
            env.step(None if applied else action[t], compact_out=frame(t + 1), lean=True)
            host_time.append(adapter.time())
            host_done.append(env.time > EPISODE_END)
            n = t + 1
            if break_when_any_done and host_done[-1]:
                break
    finally:
        env.occupancy = keep_occ
    times = torch.tensor(host_time, dtype=torch.float32, device=dev).unsqueeze(1).expand(n + 1, R)
    out = _trajectory(num, sel, ai, times, action, n)
    out["sample_log_prob"] = lps[:n]
    out["reward"] = -occ[:n].to(torch.float32)
    out["done"] = torch.tensor(host_done, dtype=torch.bool, device=dev).unsqueeze(1).expand(n, R)
    return out


def _values(adapter, value_module, batch, prefix=""):
    """V(s) for every frame of the batch, one time step (R frames) at a time to bound the transient memory."""
    T, R = batch["num"].shape[:2]
    fast = getattr(value_module.net, "forward_occupancy", None)
    out = []
    for t in range(T):
        time = batch[prefix + "time"][t].reshape(R, 1).to(torch.float32)
        if fast is not None:
            out.append(fast(batch[prefix + "num"][t], time).reshape(R))
        else:
            obs = adapter.observation(batch[prefix + "num"][t], batch[prefix + "sel"][t], batch[prefix + "agent_index"][t],
                                      batch[prefix + "time"][t])
            out.append(value_module(obs).reshape(R))
    return torch.stack(out)


def ppo_train(env, policy_module, value_module, *, total_frames=128, frames_per_batch=32, num_epochs=1,
              sub_batch_size=32, device=None, checkpoint_path=None, log_dir=None, eval_env=None, eval_interval=0,
              log_interval=1, stochastic_eval=False, lr=1e-3, seed=0, history=None):
    """See the module docstring. `total_frames` / `frames_per_batch` count environment steps (each step advances
    every replica of a BatchedSimulatorEnv); `history` (optional list) receives one dict of scalars per iteration."""
    adapter = _EnvAdapter(env)
    eval_adapter = _EnvAdapter(eval_env) if eval_env is not None else None
    params = [p for p in list(policy_module.parameters()) + list(value_module.parameters()) if p.requires_grad]
    seen, uniq = set(), []
    for p in params:
        if id(p) not in seen:
            seen.add(id(p)); uniq.append(p)
    params = uniq
    optim = torch.optim.Adam(params, lr=lr)
    rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
    gen = torch.Generator(device="cpu").manual_seed(seed + 7919 * rank)
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
    n_iters = max(total_frames // frames_per_batch, 1)
    for it in range(n_iters):
        t0 = time.perf_counter()
        batch = collect(adapter, policy_module, frames_per_batch, occupancy_only=occupancy_only(policy_module, value_module))
        T, R = batch["reward"].shape
        global_step += frames_per_batch
        rollout_s = time.perf_counter() - t0
        for _ in range(num_epochs):
            with torch.no_grad():
                if "_frames" in batch:          # V over the T+1 frames once: value = V[:-1], next_value = V[1:]
                    v_all = _values(adapter, value_module, batch["_frames"])
                    value, next_value = v_all[:-1], v_all[1:]
                else:
                    value = _values(adapter, value_module, batch)
                    next_value = _values(adapter, value_module, batch, "next_")
                adv, target = gae(value, next_value, batch["reward"], batch["done"], batch["done"])
                adv = standardise(adv)
            n = min(sub_batch_size, T * R)
            pick = torch.randperm(T * R, generator=gen)[:n].to(adapter.device)
            ti, ri = pick // R, pick % R
            flat = lambda x: None if x is None else x[ti, ri]     # gathers n frames whatever the strides of the trajectory
            obs = adapter.observation(flat(batch["num"]), flat(batch["sel"]), flat(batch["agent_index"]), flat(batch["time"]))
            d = policy_module.dist(obs)
            log_prob = d.log_prob(flat(batch["action"]))
            entropy = d.entropy()
            v = value_module(obs).reshape(n)
            losses = clip_ppo_loss(log_prob, flat(batch["sample_log_prob"]), flat(adv), entropy, v, flat(target))
            loss = losses["loss_objective"] + losses["loss_critic"] + losses["loss_entropy"]
            loss.backward()
            allreduce_gradients(params)
            grads = [p.grad for p in params if p.grad is not None]
            grad_norm = torch.norm(torch.stack([g.norm() for g in grads])) if grads else torch.zeros(())
            optim.step()
            optim.zero_grad()
        rec = {"iteration": it, "global_step": global_step, "frames": T * R, "rollout_s": round(rollout_s, 4),
               "avg_step_reward": float(batch["reward"].mean()), "episode_return": float(batch["reward"].sum(0).mean()),
               "loss_total": float(loss.detach()), "grad_global_norm": float(grad_norm),
               **{k: float(v.detach()) for k, v in losses.items()}}
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
        torch.save(policy_module.net.state_dict(), checkpoint_path)
    return history
