#!/usr/bin/env python
"""bench.py — headline benchmark of the hot path on B200 (contract: see the task brief / DESIGN.md §Measurement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference] [--workload NAME]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Metric (BASELINE.json): sim link-steps/s of the per-timestep network step (SimulationCoreModel.forward semantics) on
the synthetic 1M-link ring-radial network with 2M agents; the MPNN fwd+bwd edges/s figure rides along under "mpnn".
A "step" = one pass of the core step over the whole network (E uniforms drawn on the device + the three kernels).
At N>1 every rank steps its own independent replica of the workload (weak scaling, no data-path collective).
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC, UNIT = "sim link-steps/s", "link-steps/s"
T0 = 21600.0


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=500)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=["native", "reference"], default="native")
    ap.add_argument("--workload", default="ring_radial_1m")
    ap.add_argument("--cpu-steps", type=int, default=8, help="oracle steps timed for cpu_baseline (N=1, rank 0)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-mpnn", action="store_true")
    ap.add_argument("--no-ppo", action="store_true")
    ap.add_argument("--ppo-replicas", type=int, default=1024, help="environment replicas in total (sharded over ranks)")
    ap.add_argument("--ppo-steps", type=int, default=32, help="rollout steps per PPO iteration in the ppo section")
    ap.add_argument("--mpnn-batch", type=int, default=32, help="batch rows (frames) of the MPNN fwd+bwd measurement")
    ap.add_argument("--replicas", type=int, default=1, help="independent network replicas stepped per GPU")
    ap.add_argument("--link-order", default="node", choices=["node", "direction", "shuffled"])
    ap.add_argument("--variant", type=int, default=0, help="tarl_store_step kernel family: 0 ELL (default), 1 CSR")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clock and throttle reasons during the timed region (NVML; nvidia-smi as fallback)."""

    BAD = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown")

    def __init__(self, index: int, period=0.02):
        self.index, self.period = index, period
        self.sm, self.reasons, self.sm_max = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        self._nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nvml = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self._nvml = None

    def _reasons_nvml(self):
        nv = self._nvml
        try:
            bits = nv.nvmlDeviceGetCurrentClocksEventReasons(self._h)
        except Exception:
            try:
                bits = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
            except Exception:
                return
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                 "hw_power_brake_slowdown": 0x80, "sync_boost": 0x10}
        for k, b in names.items():
            if bits & b:
                self.reasons.add(k)

    def _loop(self):
        while not self._stop.is_set():
            self.sample_now()
            self._stop.wait(self.period)

    def sample_now(self):
        if True:
            try:
                if self._nvml is not None:
                    self.sm.append(self._nvml.nvmlDeviceGetClockInfo(self._h, self._nvml.NVML_CLOCK_SM))
                    self._reasons_nvml()
                else:
                    out = subprocess.run(
                        ["nvidia-smi", f"--id={self.index}", "--query-gpu=clocks.sm,clocks.max.sm,"
                         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap",
                         "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout.strip()
                    f = [v.strip() for v in out.split(",")]
                    self.sm.append(int(f[0])); self.sm_max = int(f[1])
                    for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[2:]):
                        if v == "Active":
                            self.reasons.add(name)
            except Exception:
                pass

    def __enter__(self):
        self._thread = threading.Thread(target=self._loop, daemon=True)
        self._thread.start()
        return self

    def mark(self):
        """The timed region starts here: samples taken before it (warm-up) are dropped."""
        self.sm, self.reasons = [], set()

    def __exit__(self, *exc):
        self._stop.set()
        self._thread.join(timeout=5)

    def summary(self):
        return {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": self.sm_max,
                "reasons": sorted(self.reasons), "samples": len(self.sm)}


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


TRAFFIC_FILE = "traffic_r02.json"


def measured_traffic(workload, replicas, kernel):
    """DRAM bytes per launch of `kernel` from the committed ncu --set full capture (profiles/traffic_r02.json), valid
    only for the workload / replica count it was captured on; None otherwise."""
    try:
        with open(os.path.join(ROOT, "profiles", TRAFFIC_FILE)) as f:
            t = json.load(f)
        if t["workload"] == workload and t["replicas"] == replicas:
            return t["kernels"][kernel]["dram_bytes_per_launch"]
    except Exception:
        pass
    return None


def step_bytes(N, E, Nmax, p, noise_read=False):
    """algorithmic bytes of one core step, B_step = N (69 + p (24 (Nmax-1) + 4)) + 20 E. 4 of the
    20 B per dual edge are the uniform the direction phase reads; SURVEY sets them to 0 when the noise is generated
    in-kernel, which is what the timed run does (noise_read=False -> 16 E)."""
    return N * (69 + p * (24 * (Nmax - 1) + 4)) + (20 if noise_read else 16) * E


def phase_bytes(N, E, Nmax, p, noise_read=False):
    """The same budget split over the two phases of the store step: direction = per link read head
    triplet 12 + {MAXN, NUM, FFTT, SEL, RIDX} 20 + cc 4, write tail triplet 12 + NUM 4 (52 N); per dual edge col-idx
    4 + edge_attr 4 (+ noise 4) + delta_tt 4. Response = 17 N + 4 E + p N (24 (Nmax-1) + 4)."""
    return {"direction": N * 52 + (16 if noise_read else 12) * E,
            "response": N * 17 + 4 * E + p * N * (24 * (Nmax - 1) + 4)}


# ---------------------------------------------------------------------------------------------------------------
def run_native(args):
    import torch
    import torch.distributed as dist
    from tarl_simulator_b200 import _cabi, synthetic
    from tarl_simulator_b200.core import SimulationCoreModel
    from tarl_simulator_b200.engine import PHASE_RESPOND_POP, PHASE_SELECT_APPEND, LinkStore

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl native needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # NCCL_DEBUG is left as the environment has it: whatever NCCL prints goes to stderr (fd 1 is pointed at
        # stderr for the whole run, see __main__), the JSON line is written to the saved stdout
        dist.init_process_group("nccl", device_id=dev)

    g, Nmax, placed = synthetic.make_workload(args.workload, device=dev, t=T0, seed=rank, order=args.link_order)
    N, E = int(g.num_roads), g.edge_index_routes.size(1)
    stream = torch.cuda.current_stream(dev)
    # The resident link store holds the state between steps (same semantics as SimulationCoreModel.forward, exact
    # import/export: tests/test_link_store_gpu.py). Noise: drawn in-kernel (Philox), as the reference draws inside
    # aggregate (src/direction_mpnn.py:137). delta_travel_time[E] and the pop mask[N] are produced every step.
    store = LinkStore.from_graph(g, Nmax, replicas=args.replicas, seed=1234 + rank)
    R = args.replicas
    # Agents.choice re-draws SELECTED_ROAD for every link every step (src/agents/base.py:446-494); the draw itself is
    # not part of the core step, so a bank of pre-drawn decision vectors is cycled through as the step's input.
    sel_bank = [synthetic.random_out_neighbour(g, 1000 + 17 * rank + i).repeat(R) for i in range(8)]
    state = {"t": T0, "i": 0}
    ell = args.variant == 0

    def use_bank(i):      # this step's routing decisions (a pointer swap when the store keeps link-id order)
        if store.slot_link is None:
            store.sel = sel_bank[i % len(sel_bank)]
        else:
            store.set_selected_road(sel_bank[i % len(sel_bank)].view(R, N))

    def step(mask=PHASE_SELECT_APPEND | PHASE_RESPOND_POP):
        use_bank(state["i"])
        store.step(state["t"], noise=None, delta_tt_link=True, pop_bits=True, phase_mask=mask, variant=args.variant)
        if mask & PHASE_RESPOND_POP:
            state["t"] += 1.0
            state["i"] += 1

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    warm = max(args.warmup, 3)

    def run(n):     # n steps enqueued by one library call (tarl_store_run), routing decisions cycled from the bank
        bank = [sel_bank[(state["i"] + k) % len(sel_bank)] for k in range(len(sel_bank))]
        store.run(state["t"], n, dt=1.0, sel_bank=bank, variant=args.variant, delta_tt_link=True, pop_bits=True)
        state["t"] += float(n)
        state["i"] += n

    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # The sampler (NVML initialisation: tens of ms) is set up BEFORE the warm-up: the timed region is ~1 ms long, and a
    # device left idle between warm-up and the timed region spends a good part of it getting back to its clocks.
    with ClockSampler(local, period=0.002) as clk:
        run(warm)
        barrier()
        clk.mark()
        ev0.record(stream)
        run(args.steps)
        ev1.record(stream)
        clk.sample_now()          # the device is still inside the timed region (the enqueue runs ahead of it)
        torch.cuda.synchronize(dev)
    barrier()
    ms = ev0.elapsed_time(ev1)
    if world > 1:
        tms = torch.tensor([ms], device=dev)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms = float(tms.item())
    value = N * R * world * args.steps / (ms / 1e3)
    store.check_errors()

    # ---- per-kernel durations (CUDA events on the launching stream) and the pop fraction p
    names = ["k_ell_select_append", "k_ell_respond_pop"] if ell else ["k_csr_select_append", "k_csr_respond_pop"]
    masks = [PHASE_SELECT_APPEND, PHASE_RESPOND_POP]
    K = len(names)
    pairs = {k: 0.0 for k in names}
    reps = min(args.steps, 20)
    pops_dev = torch.zeros((), dtype=torch.int64, device=dev)
    evs = [torch.cuda.Event(enable_timing=True) for _ in range((K + 1) * reps)]
    for i in range(reps):          # (a) an event pair around every single launch: includes the launch gap of each
        for j in range(K):
            evs[(K + 1) * i + j].record(stream)
            step(masks[j])
        evs[(K + 1) * i + K].record(stream)
        pops_dev += store.pop[: N * R].sum()
    torch.cuda.synchronize(dev)
    for i in range(reps):
        for j in range(K):
            pairs[names[j]] += evs[(K + 1) * i + j].elapsed_time(evs[(K + 1) * i + j + 1])
    pops = int(pops_dev.item())
    p = pops / (reps * N * R)
    pairs = {k: v / reps for k, v in pairs.items()}
    # (b) the direction kernel alone, `reps` launches back to back between ONE event pair (it reads the current
    # records and writes the other buffer, so repeating it on a fixed state repeats exactly the same traffic; the
    # launches are captured in a CUDA graph and replayed, so that the host's per-call time — Python + ctypes, ~35 us,
    # the same order as the kernel — is not part of what the event pair brackets). The response kernel's share is what
    # remains of the pipelined step. This is the per-launch duration the roofline uses.
    use_bank(state["i"])
    for _ in range(3):
        store.step(state["t"], noise=None, delta_tt_link=True, pop_bits=True, phase_mask=PHASE_SELECT_APPEND, variant=args.variant)
    torch.cuda.synchronize(dev)
    side = torch.cuda.Stream(dev)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        with torch.cuda.graph(graph, stream=side):
            for _ in range(reps):
                store.step(state["t"], noise=None, delta_tt_link=True, pop_bits=True, phase_mask=PHASE_SELECT_APPEND,
                           variant=args.variant)
    for _ in range(3):
        graph.replay()
    torch.cuda.synchronize(dev)
    ev0.record(stream)
    graph.replay()
    ev1.record(stream)
    torch.cuda.synchronize(dev)
    sel_ms = ev0.elapsed_time(ev1) / reps
    del graph
    step(PHASE_RESPOND_POP)          # complete the step that the repeated direction phase left half done
    per = {names[0]: sel_ms, names[1]: max(ms / args.steps - sel_ms, 0.0)}
    store.check_errors()
    peak, peak_src = peaks()
    # algorithmic bytes per launch: SURVEY.md §8(d)'s per-unit figures, noise term dropped (drawn in-kernel)
    ph = phase_bytes(N, E, Nmax, p)
    pb = {names[0]: R * ph["direction"], names[1]: R * ph["response"]}
    dom = max(per, key=per.get)
    achieved = pb[dom] / (per[dom] / 1e3) / 1e9
    step_ms = ms / args.steps
    sb = R * step_bytes(N, E, Nmax, p)
    # what this formulation has to move at the least: delta_tt is emitted per upstream LINK (4 N
    # instead of 4 E; the [E] form is materialised by whoever reads it)
    # (and with the uniform-weight hint in stat_a.w the direction kernel does not read its 4 B/edge weight column)
    moved = {"direction": N * 56 + (4 if getattr(store, "uniform_weights", False) else 8) * E, "response": ph["response"]}
    traffic = measured_traffic(args.workload, R, dom)
    per_kernel = {}
    for k in names[:2]:
        tr = measured_traffic(args.workload, R, k)
        per_kernel[k] = {"ms": round(per[k], 4), "algorithmic_bytes": int(pb[k]),
                         "achieved": round(pb[k] / (per[k] / 1e3) / 1e9, 1), "frac": round(pb[k] / (per[k] / 1e3) / 1e9 / peak, 4),
                         "traffic": tr}
    roofline = {"bound": "hbm", "kernel": dom, "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 4), "traffic": traffic,
                "traffic_source": f"profiles/{TRAFFIC_FILE} (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum, per launch)",
                "peak_source": peak_src,
                "algorithmic_bytes_per_launch": int(pb[dom]),
                "algorithmic_bytes_how": "SURVEY.md 8(d): direction 52 N + 12 E (noise drawn in-kernel: its 4 B/edge are "
                                         "not counted), response 17 N + 4 E + p N (24 (Nmax-1) + 4)",
                "kernel_ms": round(per[dom], 4),
                "kernels_ms": {k: round(v, 4) for k, v in per.items()},
                "per_kernel": per_kernel,
                "kernels_ms_how": f"{names[0]}: {reps} launches captured in one CUDA graph, replayed between one CUDA-event "
                                  f"pair; {names[1]}: pipelined step time minus that",
                "kernels_ms_event_pair_per_launch": {k: round(v, 4) for k, v in pairs.items()},
                "pop_fraction": round(p, 4),
                "dram_frac_of_measured_traffic": (round(traffic / (per[dom] / 1e3) / 1e9 / peak, 4)
                                                  if traffic is not None else None),
                "step": {"algorithmic_bytes": int(sb), "achieved": round(sb / (step_ms / 1e3) / 1e9, 1),
                         "frac": round(sb / (step_ms / 1e3) / 1e9 / peak, 4),
                         "bytes_this_formulation_must_move": int(R * (moved["direction"] + moved["response"]))}}

    # ---- end to end through the public drop-in API, host buffers both ways
    # SimulationCoreModel.forward keeps the road state in its resident link store while graph.x is not edited between
    # calls (core.py); graph.x itself is brought up to date when somebody reads it (nobody does inside this loop, as
    # nobody does inside TransportationSimulator.run between two metrics reads).
    store.export_x(out=g.x[:N].unsqueeze(0)) if R == 1 else g.x[:N].copy_(store.export_x()[0])
    model = SimulationCoreModel(Nmax=Nmax, device=str(dev), time=state["t"], resident="always", seed=4321 + rank)
    model.pack_pop = True
    words = (N + 31) // 32
    sel_hosts = [b[:N].cpu().pin_memory() for b in sel_bank]
    dtt_hosts = [torch.empty(N, dtype=torch.float32).pin_memory() for _ in range(2)]
    pop_hosts = [torch.empty(words, dtype=torch.int32).pin_memory() for _ in range(2)]
    host_outs = [{"delta_tt_link": dtt_hosts[j], "pop_bits": pop_hosts[j]} for j in range(2)]
    e2e_steps = min(args.steps, 100)
    pending = {"k": 0}

    def e2e_step():
        # ONE public call per step, host buffers on both sides: this step's routing decisions (pinned host memory) in,
        # delta_travel_time per link + pop bits out into pinned host buffers. The library enqueues upload, kernels and
        # downloads on its own copy streams (tarl_store_step_host), alternating two slots, so the copies of neighbouring
        # steps overlap this step's kernels; every byte is moved inside the timed region.
        k = pending["k"]
        model.set_time(state["t"])
        model(g, selected_road=sel_hosts[state["i"] % len(sel_hosts)], host_out=host_outs[k % 2])
        pending["k"] = k + 1
        state["t"] += 1.0
        state["i"] += 1

    # Warm-up long enough for the caching allocator to reach its steady state, then several windows of e2e_steps; the
    # MEDIAN window is reported and every window is listed (a single window is at the mercy of one host hiccup).
    for _ in range(20):
        e2e_step()
    model.host_sync(g)
    model.response_mpnn.update_history.resolve()
    assert model.last_path == "resident"
    windows, enqueue = [], []
    for _ in range(9):
        barrier()
        w0 = time.perf_counter()
        ev0.record(stream)
        for _ in range(e2e_steps):
            e2e_step()
        enqueue.append((time.perf_counter() - w0) * 1e3)      # host time to enqueue the window (no synchronisation in it)
        model.host_join(g)                                    # the window ends when its last download has landed
        ev1.record(stream)
        torch.cuda.synchronize(dev)
        windows.append(max(ev0.elapsed_time(ev1), (time.perf_counter() - w0) * 1e3))
        model.response_mpnn.update_history.resolve()
    e2e_ms = statistics.median(windows)
    if os.environ.get("TARL_BENCH_DEBUG"):
        print(f"[rank {rank}] e2e windows (ms for {e2e_steps} steps): {[round(w, 2) for w in windows]}", file=sys.stderr)
    # the [E] contract of road_optimality_data["delta_travel_time"] on the host side: expanding the per-link vector
    # that crossed the bus with edge_index_routes[0] gives exactly what the device would have materialised
    model.host_sync(g)
    k_last = (pending["k"] - 1) % 2
    full = model.direction_mpnn.road_optimality_data["delta_travel_time"]
    contract_ok = bool(torch.equal(dtt_hosts[k_last][g.edge_index_routes[0].cpu()], full.cpu()))
    # What the box's host <-> device path allows for exactly these copies, with no kernel in between: the same bytes per
    # step on two copy streams, all ranks at once. e2e close to this floor = bound by the copies (PCIe / host memory),
    # not by anything this repository runs.
    h2d_s, d2h_s = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    stage_d = torch.empty(N, dtype=torch.float32, device=dev)
    dtt_d = torch.empty(N, dtype=torch.float32, device=dev)
    bits_d = torch.empty(words, dtype=torch.int32, device=dev)
    floors = []
    for _ in range(3):
        barrier()
        ev0.record(stream)
        h2d_s.wait_stream(stream); d2h_s.wait_stream(stream)
        for k in range(e2e_steps):
            with torch.cuda.stream(h2d_s):
                stage_d.copy_(sel_hosts[k % len(sel_hosts)], non_blocking=True)
            with torch.cuda.stream(d2h_s):
                dtt_hosts[k % 2].copy_(dtt_d, non_blocking=True)
                pop_hosts[k % 2].copy_(bits_d, non_blocking=True)
        stream.wait_stream(h2d_s); stream.wait_stream(d2h_s)
        ev1.record(stream)
        torch.cuda.synchronize(dev)
        floors.append(ev0.elapsed_time(ev1))
    floor_ms = statistics.median(floors)
    if world > 1:
        tms = torch.tensor([e2e_ms, floor_ms], device=dev)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        e2e_ms, floor_ms = float(tms[0].item()), float(tms[1].item())
    e2e = {"value": round(N * world * e2e_steps / (e2e_ms / 1e3), 1), "unit": UNIT,
           "h2d_bytes_per_step": int(N * 4), "d2h_bytes_per_step": int(N * 4 + words * 4),
           "steps": e2e_steps, "windows_ms": [round(w, 3) for w in windows], "window": "median of 9",
           "host_enqueue_ms": [round(w, 3) for w in enqueue],
           "copies_alone": {"value": round(N * world * e2e_steps / (floor_ms / 1e3), 1), "unit": UNIT,
                            "ms": round(floor_ms, 3),
                            "what": "the same H2D + D2H bytes per step on two copy streams with NO kernel, all ranks at "
                                    "once (max over ranks): the ceiling the box's PCIe / host memory puts on e2e"},
           "kernel_path": model.last_path, "delta_tt_edge_form_reproduced_on_host": contract_ok,
           "api": "SimulationCoreModel.forward(graph, selected_road=<pinned host tensor>, host_out={pinned host "
                  "buffers}): ONE public call per step; state resident on the device (link store behind graph.x, exported "
                  "when graph.x is read); per step H2D = SELECTED_ROAD decisions [N] fp32, D2H = delta_travel_time per "
                  "upstream link [N] fp32 (the [E] vector of road_optimality_data is that value repeated on each "
                  "out-edge: expanded on whichever side reads it) + pop mask as bits [N/8 bytes]; the library enqueues "
                  "upload, kernels and downloads itself (tarl_store_step_host: two copy streams, two slots), so the "
                  "copies of neighbouring steps overlap this step's kernels; noise drawn on the device"}

    out = {"metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": args.steps,
           "warmup": warm, "ms_per_step": round(step_ms, 5), "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": args.workload, "links": N, "dual_edges": E, "agents": placed, "Nmax": Nmax,
                      "link_order": args.link_order, "replicas_per_gpu": R, "kernel_variant": {0: "ell", 1: "csr"}[args.variant], "parallelism": f"independent replicas x{world}",
                      "state": "resident link store (tarl_store_run), noise drawn in-kernel; delta_travel_time per upstream "
                               "link, pop mask (bytes) and pop bits written every step",
                      "l2": "per-step working set larger than the 126 MB L2" if N * R * 150 > 130e6 else
                      "per-step working set fits in L2 (small workload)"},
           "e2e": e2e, "gpu_launches": 2 * args.steps, "roofline": roofline, "clocks": clk.summary()}

    # The secondary blocks must not take the headline line down with them: a failure is reported in place. (With
    # several ranks an exception on one rank would leave the others in a collective, so it is still fatal there.)
    def guarded(fn, *a):
        if world > 1:
            return fn(*a)
        try:
            return fn(*a)
        except Exception as exc:        # noqa: BLE001
            import traceback
            traceback.print_exc(file=sys.stderr)
            return {"error": f"{type(exc).__name__}: {exc}"}

    if not args.no_mpnn:
        out["mpnn"] = guarded(mpnn_bench, args, g, dev, world, rank, peak)
        out["gpu_launches"] += out["mpnn"].pop("_launches_in_headline", 0)
    if not args.no_ppo:
        out["ppo"] = guarded(ppo_bench, args, dev, world, rank)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(args, sample_steps=args.cpu_steps)
        if "mpnn" in out and "error" not in out["mpnn"]:
            out["mpnn"]["cpu_baseline"] = mpnn_cpu_baseline(args)
    if world > 1:
        dist.destroy_process_group()
    if rank == 0:
        emit(out)


# ---------------------------------------------------------------------------------------------------------------
def mpnn_bench(args, g, dev, world, rank, peak):
    """Second half of BASELINE.json's metric: MPNN fwd+bwd edges/s on the FULL graph of the same workload (links +
    SRC/DEST nodes): policy logits -> GraphDistribution log_prob + entropy -> PPO-style loss -> backward to the
    embedding gradient; and, reported separately, MPNNValueNet forward + backward. Timed with CUDA events on the
    current stream, max over ranks; every rank works on its own copy (weak scaling)."""
    import torch
    import torch.distributed as dist
    from tarl_simulator_b200.distribution import GraphDistribution
    from tarl_simulator_b200.mpnn_agent import MPNNPolicyNet, MPNNValueNet

    B = args.mpnn_batch
    ei = g.edge_index
    E_full, N_tot = ei.size(1), g.x.size(0)
    Nmax = (g.x.size(1) - 7) // 3
    nf = g.x[:, 3 * Nmax:].unsqueeze(0).repeat(B, 1, 1).contiguous()
    policy = MPNNPolicyNet(ei, N_tot, None, str(dev)) if N_tot > 4096 else MPNNPolicyNet(ei, N_tot, torch.ones(E_full, device=dev), str(dev))
    gen = torch.Generator(device=dev).manual_seed(5 + rank)
    with torch.no_grad():
        d0 = GraphDistribution(policy(nf, None, None), ei)
        action = d0.sample(uniforms=torch.rand(B, d0.nb_nodes, device=dev, generator=gen), dtype=torch.bool)
    adv = torch.randn(B, device=dev, generator=gen)
    stream = torch.cuda.current_stream(dev)

    def policy_iter():
        policy.nodes_embedding.weight.grad = None
        dd = GraphDistribution(policy(nf, None, None), ei)
        lp = dd.log_prob(action)
        ent = dd.entropy()
        (-(lp * adv).mean() - 0.01 * ent.mean()).backward()

    value = MPNNValueNet(ei, N_tot, str(dev))
    value.agent_features = torch.rand(1024, 9, device=dev, generator=gen)
    value.eval()
    ef = g.edge_attr.reshape(1, E_full, 1).expand(B, -1, -1)      # what the environment hands out: one row, batch stride 0
    ai = torch.randint(0, 1024, (B, N_tot), device=dev, generator=gen)
    tm = torch.full((B, 1), 21600.0, device=dev)
    wv = torch.randn(B, 1, device=dev, generator=gen)

    def value_iter():
        for p_ in value.parameters():
            p_.grad = None
        (value(nf, ef, ai, tm) * wv).sum().backward()

    def timed(fn, iters, warm=3):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(iters):
            fn()
        e1.record(stream)
        torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1)
        if world > 1:
            tms = torch.tensor([ms], device=dev)
            dist.all_reduce(tms, op=dist.ReduceOp.MAX)
            ms = float(tms.item())
        return ms / iters

    iters = max(min(args.steps // 10, 50), 5)
    pol_ms = timed(policy_iter, iters)
    val_ms = timed(value_iter, iters)
    # the same net in TRAIN mode: nn.Dropout(0.05) on the [B*E, 17] message input (src/agents/mpnn_agent.py:278); the
    # keep words are drawn in the kernel; time_net's own dropouts are torch modules and stay on
    value.train()
    val_train_ms = timed(value_iter, max(iters // 5, 3))
    value.eval()
    del policy, value, nf, ai, action
    torch.cuda.empty_cache()
    emlp = edge_mlp_bench(g, dev, timed, world)
    torch.cuda.empty_cache()
    vmlp = value_mlp_bench(dev, timed, world, peak)
    # algorithmic bytes per edge: policy 28 B + GraphDistribution 24 B = 52 B/edge (+4 B/node)
    pol_bytes = B * (52 * E_full + 4 * N_tot)
    val_bytes = B * (2 * (12 * E_full) + 2 * 68 * N_tot)          # fwd + bwd: 12 B/edge + 68 B/node each
    train_bytes = B * (36 * E_full + 92 * N_tot)                  # train mode: a keep word and a message per (row, edge)
    return {"metric": "MPNN fwd+bwd edges/s", "unit": "edges/s", "batch_rows": B, "edges_full_graph": E_full,
            "nodes_full_graph": N_tot, "iters": iters, "edge_mlp": emlp,
            "policy_distribution": {"value": round(world * B * E_full / (pol_ms / 1e3), 1), "ms_per_iter": round(pol_ms, 4),
                                    "what": "MPNNPolicyNet.forward -> GraphDistribution.log_prob + entropy -> backward (7 kernels)",
                                    "roofline": {"bound": "hbm", "algorithmic_bytes": int(pol_bytes),
                                                 "achieved": round(pol_bytes / (pol_ms / 1e3) / 1e9, 1), "peak": peak,
                                                 "unit": "GB/s", "frac": round(pol_bytes / (pol_ms / 1e3) / 1e9 / peak, 4)}},
            "value_net": {"value": round(world * B * E_full / (val_ms / 1e3), 1), "ms_per_iter": round(val_ms, 4),
                          "what": "MPNNValueNet.forward (eval) -> backward (project, aggregate, node_grad, edge_grad, finish)",
                          "roofline": {"bound": "hbm", "algorithmic_bytes": int(val_bytes),
                                       "achieved": round(val_bytes / (val_ms / 1e3) / 1e9, 1), "peak": peak, "unit": "GB/s",
                                       "frac": round(val_bytes / (val_ms / 1e3) / 1e9 / peak, 4)}},
            "value_net_train_mode": {"value": round(world * B * E_full / (val_train_ms / 1e3), 1),
                                     "ms_per_iter": round(val_train_ms, 4),
                                     "roofline": {"bound": "hbm", "algorithmic_bytes": int(train_bytes),
                                                  "achieved": round(train_bytes / (val_train_ms / 1e3) / 1e9, 1),
                                                  "peak": peak, "unit": "GB/s",
                                                  "frac": round(train_bytes / (val_train_ms / 1e3) / 1e9 / peak, 4),
                                                  "how": "per (row, edge): message 4 w + 4 r, keep word 4 w + 8 r, d z 4 w + "
                                                         "4 r, gm gather 4 = 36 B; per (row, node): the 16 message inputs "
                                                         "assembled twice (28 + 8 B each) + mean / v / gm 20 B = 92 B"},
                                     "what": "MPNNValueNet.forward (train: message dropout p = 0.05, keep words drawn in "
                                             "the kernel) -> backward (message_dropout, aggregate_msg, node_grad, "
                                             "edge_grad_dropout, finish): the per-node projection does not factor "
                                             "through a per-edge mask, every (row, edge) gathers its 16 inputs' products"},
            "value_mlp": vmlp}


def edge_mlp_bench(g, dev, timed, world):
    """MPNNPolicyNet.edge_mlp (33 -> 64 -> 32 -> 1 per edge; dormant in the reference, src/agents/mpnn_agent.py:38-44,
    227-231) on the full graph of the workload at 8 rows: forward on tcgen05 (csrc/edge_mlp_tc.cu) beside the fp32-pipe
    forward, and forward + backward (parameter gradients, fp32 pipe). Tensor-bound: 8.5 kflop per pair on 132 gathered
    bytes; `tensor` relates the TF32 flop the MMAs issue (3xTF32, K padded to 40) to half the measured dense bf16 peak."""
    import torch
    from tarl_simulator_b200.mpnn_agent import MPNNPolicyNet
    B = 8
    ei = g.edge_index
    E, N_tot = ei.size(1), g.x.size(0)
    Nmax = (g.x.size(1) - 7) // 3
    gen = torch.Generator(device=dev).manual_seed(17)
    nf = g.x[:, 3 * Nmax:].unsqueeze(0).repeat(B, 1, 1).contiguous()
    net = MPNNPolicyNet(ei, N_tot, None, str(dev)) if N_tot > 4096 else MPNNPolicyNet(ei, N_tot, torch.ones(E, device=dev), str(dev))
    net.agent_features = torch.rand(1024, 9, device=dev, generator=gen)
    ai = torch.randint(0, 1024, (B, N_tot), device=dev, generator=gen)
    ef = g.edge_attr.reshape(1, E, 1).expand(B, -1, -1)
    w = torch.randn(B, E, device=dev, generator=gen)
    with torch.no_grad():
        a = net.edge_logits(nf, ef, ai, tensor_cores=True)
        path = net.last_edge_path
        b = net.edge_logits(nf, ef, ai, tensor_cores=False)
        rel = float((a - b).abs().max() / b.abs().max().clamp_min(1e-6))
        tc_ms = timed(lambda: net.edge_logits(nf, ef, ai, tensor_cores=True), 5)
        fp_ms = timed(lambda: net.edge_logits(nf, ef, ai, tensor_cores=False), 3)

    def train():
        for p_ in net.edge_mlp.parameters():
            p_.grad = None
        (net.edge_logits(nf, ef, ai, tensor_cores=True) * w).sum().backward()

    tr_ms = timed(train, 3)
    pairs = B * E
    useful = 2 * (64 * 34 + 32 * 64 + 32) * pairs
    issued = 2 * 3 * (64 * 40 + 32 * 64) * pairs
    tf32_peak = None
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            tf32_peak = float(json.load(f)["bf16_tflops"]) / 2
    except Exception:
        pass
    return {"metric": "edge MLP (row, edge) pairs/s", "rows": B, "edges": E, "forward_path": path,
            "tcgen05": {"value": round(world * pairs / (tc_ms / 1e3), 1), "ms": round(tc_ms, 4),
                        "what": "x assembly + tarl_edge_mlp_forward: gather -> TMEM -> tcgen05.mma kind::tf32 (3xTF32) for both "
                                "hidden layers, 32 -> 1 in registers"},
            "fp32_pipe": {"value": round(world * pairs / (fp_ms / 1e3), 1), "ms": round(fp_ms, 4)},
            "forward_backward": {"ms": round(tr_ms, 4), "what": "tcgen05 forward + k_edge_mlp_bwd (recompute, per-tile outer-"
                                 "product sums on the fp32 pipe, deterministic finish)"},
            "max_rel_diff_tc_vs_fp32": rel,
            "tensor": {"bound": "tensor", "useful_tflops": round(useful / (tc_ms / 1e3) / 1e12, 2),
                       "issued_tf32_tflops": round(issued / (tc_ms / 1e3) / 1e12, 2), "peak": tf32_peak, "unit": "TFLOP/s",
                       "frac": round(issued / (tc_ms / 1e3) / 1e12 / tf32_peak, 4) if tf32_peak else None,
                       "peak_source": "MEASURED_PEAKS.json bf16_tflops / 2 (TF32 runs at half the bf16 rate)"}}


def value_mlp_bench(dev, timed, world, peak):
    """MPNNValueNetSimple (the value net Runner wires, src/runner.py:68) evaluated for one environment step of 1024
    grid100 replicas: [1024, 59600] occupancies -> values. tcgen05 path (csrc/value_mlp.cu: 3xTF32, TMA, TMEM) beside
    the library GEMM path of the same module. The layer reads the occupancy matrix once: HBM-bound on the tensor pipe
    (32 flop per byte), so the roofline is algorithmic bytes (the matrix) over the measured HBM peak; the tensor-pipe
    share is in the ncu summary under profiles/."""
    import torch
    from tarl_simulator_b200.mpnn_agent import MPNNValueNetSimple
    M, N_tot = 1024, 59600
    net = MPNNValueNetSimple(torch.zeros(2, 1, dtype=torch.long, device=dev), N_tot, str(dev))
    g = torch.Generator(device=dev).manual_seed(11)
    num = torch.randint(0, 12, (M, N_tot), device=dev, generator=g).float()
    tm = torch.full((M, 1), 21600.0, device=dev)
    with torch.no_grad():
        a = net.forward_occupancy(num, tm)
        assert net.last_path == "tcgen05"
        b = net.final_mlp(torch.cat((num, tm), dim=-1))
        rel = float((a - b).abs().max() / b.abs().max().clamp_min(1e-6))

        def tc():
            net.forward_occupancy(num, tm)

        def lib():
            net.final_mlp(torch.cat((num, tm), dim=-1))

        tc_ms = timed(tc, 20)
        lib_ms = timed(lib, 20)
        # the PPO update evaluates the net on all (T + 1) R frames of a rollout in ONE call: 8192 rows here (1.95 GB)
        M8 = 8192
        num8 = torch.randint(0, 12, (M8, N_tot), device=dev, generator=g).float()
        tm8 = torch.full((M8, 1), 21600.0, device=dev)
        big_ms = timed(lambda: net.forward_occupancy(num8, tm8), 10)
        del num8, tm8
    # the PPO update's use of the same net: forward + backward on a 32-frame minibatch (src/rl/ppo_trainer.py:132-145)
    B = 32
    numb, tmb = num[:B].contiguous(), tm[:B].contiguous()
    wv = torch.randn(B, 1, device=dev, generator=g)

    def train_ours():
        for p_ in net.parameters():
            p_.grad = None
        (net.forward_occupancy(numb, tmb) * wv).sum().backward()

    def train_lib():
        for p_ in net.parameters():
            p_.grad = None
        (net.final_mlp(torch.cat((numb, tmb), dim=-1)) * wv).sum().backward()

    tr_ms = timed(train_ours, 20)
    tr_lib_ms = timed(train_lib, 20)
    tr_bytes = 4 * (B * N_tot + 2 * 64 * (N_tot + 1) + 2 * 64 * N_tot)      # A twice, W1 hi+lo read, dW1 written
    a_bytes = M * N_tot * 4
    flops = 2 * M * (N_tot + 1) * 64
    return {"metric": "value MLP observation rows/s", "rows": M, "nodes": N_tot,
            "tcgen05": {"value": round(world * M / (tc_ms / 1e3), 1), "ms": round(tc_ms, 4),
                        "what": "tarl_value_mlp_forward, calls back to back: TMA -> TMEM split -> tcgen05.mma kind::tf32 "
                                "(3xTF32), split-K, fused 64x64 + 64x1 tail; W1 hi/lo split cached"},
            "library": {"value": round(world * M / (lib_ms / 1e3), 1), "ms": round(lib_ms, 4),
                        "what": "torch.cat + nn.Linear x3 (cuBLAS fp32 SIMT)"},
            "max_rel_diff_vs_library": rel,
            "train_32_frames": {"ms": round(tr_ms, 4), "library_ms": round(tr_lib_ms, 4),
                                "what": "forward (tcgen05, pre-activations kept) + backward (k_value_mlp_bwd_small, "
                                        "k_value_mlp_dw1: dW1 = g^T A on the fp32 pipe, HBM-bound) vs cat + nn.Linear x3 "
                                        "with autograd (cuBLAS)",
                                "roofline": {"bound": "hbm", "algorithmic_bytes": tr_bytes,
                                             "achieved": round(tr_bytes / (tr_ms / 1e3) / 1e9, 1), "peak": peak,
                                             "unit": "GB/s", "frac": round(tr_bytes / (tr_ms / 1e3) / 1e9 / peak, 4)}},
            "rows_8192": {"ms": round(big_ms, 4),
                          "roofline": {"bound": "hbm", "algorithmic_bytes": M8 * N_tot * 4,
                                       "achieved": round(M8 * N_tot * 4 / (big_ms / 1e3) / 1e9, 1), "peak": peak,
                                       "unit": "GB/s", "frac": round(M8 * N_tot * 4 / (big_ms / 1e3) / 1e9 / peak, 4)}},
            "roofline": {"bound": "hbm", "algorithmic_bytes": a_bytes, "achieved": round(a_bytes / (tc_ms / 1e3) / 1e9, 1),
                         "peak": peak, "unit": "GB/s", "frac": round(a_bytes / (tc_ms / 1e3) / 1e9 / peak, 4),
                         "useful_tflops": round(flops / (tc_ms / 1e3) / 1e12, 2),
                         "issued_tf32_tflops": round(3 * flops / (tc_ms / 1e3) / 1e12, 2)}}


def ppo_bench(args, dev, world, rank):
    """BASELINE.json configs[4]: PPO on the 100x100 grid with `--ppo-replicas` environment replicas sharded over the
    ranks (strong scaling: the total is fixed), one gradient all-reduce per optimiser step over NCCL. Reports rollout
    throughput (environment steps/s and link-steps/s through the whole _step: action, core, withdraw, insert, reward,
    plus policy forward + sampling) and the time of one update. Timed on the device, max over ranks."""
    import torch
    import torch.distributed as dist
    from tarl_simulator_b200 import synthetic
    from tarl_simulator_b200.mpnn_agent import MPNNPolicyNet, MPNNValueNetSimple
    from tarl_simulator_b200.parallel import shard_replicas
    from tarl_simulator_b200.reinforcement_learning import BatchedSimulatorEnv
    from tarl_simulator_b200.rl.ppo_trainer import PolicyModule, ValueModule, _EnvAdapter, collect, occupancy_only, ppo_train

    total = max(args.ppo_replicas, world)
    first, R = shard_replicas(total, world, rank)
    frm, to, n_nodes = synthetic.grid_links(100, device=dev)
    frm, to = synthetic.reorder_links(frm, to, "node")
    g, Nmax = synthetic.build_graph(frm, to, n_nodes)
    af = synthetic.population(g, 100_000, 21540, 600, seed=7)
    env = BatchedSimulatorEnv(g, Nmax, af, replicas=R, seed=100, first_replica=first)
    N, N_tot, E_full = int(g.num_roads), g.x.size(0), g.edge_index.size(1)
    torch.manual_seed(0)                                   # identical initial parameters on every rank
    policy = MPNNPolicyNet(g.edge_index, N_tot, None, str(dev))
    value = MPNNValueNetSimple(g.edge_index, N_tot, str(dev))
    pm, vm = PolicyModule(policy, g.edge_index), ValueModule(value)
    adapter = _EnvAdapter.of(env)
    T = args.ppo_steps
    stream = torch.cuda.current_stream(dev)

    def timed(fn):
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        fn()
        e1.record(stream)
        torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1)
        if world > 1:
            tms = torch.tensor([ms], device=dev)
            dist.all_reduce(tms, op=dist.ReduceOp.MAX)
            ms = float(tms.item())
        return ms

    # warm-up: three full iterations (optimiser and parameter bucket, CSR builds, allocator; the rollout is launched
    # eagerly once, captured in a CUDA graph on its second run and replayed from then on)
    ppo_train(env, pm, vm, total_frames=3 * T, frames_per_batch=T, num_epochs=1, sub_batch_size=32)
    slim = occupancy_only(pm, vm)                         # what ppo_train itself passes for this pair of nets
    # median of three each: an iteration starts with a few hundred microseconds of host work (optimiser state reset,
    # generator) during which the device waits — on a single call that jitter is as large as the update itself
    hist = []
    roll_all = sorted(timed(lambda: collect(adapter, pm, T, occupancy_only=slim)) for _ in range(3))
    train_all = sorted(timed(lambda: ppo_train(env, pm, vm, total_frames=T, frames_per_batch=T, num_epochs=1,
                                               sub_batch_size=32, history=hist)) for _ in range(3))
    roll_ms, train_ms = roll_all[1], train_all[1]
    env.check_errors()
    n_params = sum(p.numel() for p in list(policy.parameters()) + list(value.parameters()) if p.requires_grad)
    in_sync = None
    if world > 1:       # every rank must hold bit-identical parameters after the all-reduced update
        flat = torch.cat([p.detach().reshape(-1) for p in list(policy.parameters()) + list(value.parameters())])
        lo, hi = flat.clone(), flat.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        in_sync = bool(torch.equal(lo, hi))
    return {"workload": "grid100 (39 600 links, 100 000 agents per replica), PPO as wired by the reference",
            "replicas_total": total, "replicas_per_gpu": R, "rollout_steps": T, "scaling": "strong",
            "rollout_ms": round(roll_ms, 2), "env_steps_per_s": round(total * T / (roll_ms / 1e3), 1),
            "link_steps_per_s": round(total * T * N / (roll_ms / 1e3), 1),
            "iteration_ms": round(train_ms, 2), "update_ms": round(max(train_ms - roll_ms, 0.0), 2),
            "rollout_ms_all": [round(x, 2) for x in roll_all], "iteration_ms_all": [round(x, 2) for x in train_all],
            "impossible_frames_in_minibatches": float(sum(h.get("impossible_frames", 0.0) for h in hist)),
            "allreduce_bytes_per_update": 4 * n_params if world > 1 else 0, "parameters_identical_across_ranks": in_sync,
            "inserted_agents_per_replica": float(env.counters[:, 0].float().mean()),
            "rollout_from_cuda_graph": any(e.get("graph") is not None for e in adapter._graphs.values()),
            "what": "rollout = episode reset + policy forward + per step: sample, env step (action, core step, withdraw, "
                    "insert, reward) for every replica, replayed from one CUDA graph; iteration = rollout + value net over "
                    "all (T+1) R frames + GAE kernel + one clipped-PPO minibatch step (32 frames) + gradient all-reduce + "
                    "one-launch Adam on the flat bucket"}


def mpnn_cpu_baseline(args):
    """The oracle port of the same policy -> distribution -> backward iteration on the host cores, on a bounded
    sample: one batch row of the full graph of the workload."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import mpnn_port
    from tarl_simulator_b200 import synthetic
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    g, Nmax, _ = synthetic.make_workload(args.workload, device="cpu", t=T0, seed=0, order=args.link_order)
    ei = g.edge_index
    E_full, N_tot = ei.size(1), g.x.size(0)
    nf = g.x[:, 3 * Nmax:].contiguous()
    w = torch.randn(N_tot, 1, requires_grad=True)
    gen = torch.Generator().manual_seed(0)
    times = []
    action = None
    for it in range(3):
        a = time.perf_counter()
        w.grad = None
        d = mpnn_port.GraphDistributionPort(mpnn_port.policy_logits(w, nf, ei), ei, 1.0)
        if action is None:
            action = d.sample(torch.rand(d.K, generator=gen))
            a = time.perf_counter()
        lp, ent = d.log_prob(action), d.entropy()
        (-(lp.mean()) - 0.01 * ent.mean()).backward()
        times.append(time.perf_counter() - a)
    med = statistics.median(times[1:])
    return {"value": round(E_full / med, 1), "unit": "edges/s", "cores": cores, "kind": "port",
            "sample": f"2 iterations, 1 batch row of the full {args.workload} graph ({E_full} edges), oracle/mpnn_port.py",
            "ms_per_iter": round(med * 1e3, 1)}


# ---------------------------------------------------------------------------------------------------------------
def cpu_baseline(args, sample_steps, warmup=1, workload=None):
    """The CPU oracle port (op-for-op restatement of the reference's torch code, all host threads) on the same
    workload. The only place outside tests/smoke where oracle/ is executed — as the thing timed beside the GPU."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import core_port
    from tarl_simulator_b200 import synthetic
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    name = workload or args.workload
    g, Nmax, placed = synthetic.make_workload(name, device="cpu", t=T0, seed=0, order=getattr(args, "link_order", "node"))
    N = int(g.num_roads)
    x = g.x[:N].clone()
    ei, w, cc = g.edge_index_routes, g.edge_attr_routes, g.congestion_constant[:N]
    gen = torch.Generator().manual_seed(0)
    times = []
    t = T0
    for s in range(warmup + sample_steps):
        u = torch.rand(ei.size(1), generator=gen).clamp_(min=1e-7)
        a = time.perf_counter()
        core_port.core_step(x, ei, w, t, Nmax, u, cc)
        b = time.perf_counter()
        if s >= warmup:
            times.append(b - a)
        t += 1.0
    med = statistics.median(times)
    return {"value": round(N / med, 1), "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{sample_steps} core steps of the full {name} workload ({N} links), median; "
                      f"oracle/core_port.py (op-for-op torch restatement of the reference, {cores} ATen threads)",
            "ms_per_step": round(med * 1e3, 2), "links": N, "dual_edges": int(ei.size(1)), "agents": int(placed),
            "Nmax": int(Nmax)}


def run_reference(args):
    """Reference arm: the reference's own CPU implementation of the path. The reference is pure Python on wheels
    that are not installable here, and /root/reference does not exist on the GPU box, so this times the oracle port
    (kind "port"). Rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    name = args.workload
    budget_s = 150.0
    est = {"ring_radial_1m": 1.4, "grid100": 0.045}.get(name, 0.05)
    steps, warm = args.steps, args.warmup
    if (steps + warm) * est > budget_s:      # a bounded sample: the CPU path needs ~0.3-1.6 s per 1M-link step
        warm = min(warm, 2)
        steps = max(3, min(steps, int(budget_s / est) - warm))
    t0 = time.perf_counter()
    cb = cpu_baseline(args, sample_steps=steps, warmup=warm, workload=name)
    out = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
           "steps": steps, "warmup": warm, "ms_per_step": cb["ms_per_step"], "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": name, "links": cb["links"], "dual_edges": cb["dual_edges"], "agents": cb["agents"],
                      "Nmax": cb["Nmax"], "link_order": getattr(args, "link_order", "node"), "replicas_per_gpu": 1,
                      "kernel_variant": "cpu", "parallelism": "ONE CPU replica on rank 0 whatever --gpus is (the other "
                                                              "ranks exit): only the N = 1 ratio is like for like",
                      "state": "reference row layout on the host, noise injected per step (torch.rand outside the timed call)",
                      "l2": "n/a (host)"},
           "steps_requested": args.steps, "warmup_requested": args.warmup,
           "cpu_baseline": cb, "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "wall_s": round(time.perf_counter() - t0, 1)}
    emit(out)


def emit(out):
    """The ONE JSON line, on the process's original stdout."""
    os.write(_REAL_STDOUT, (json.dumps(out) + "\n").encode())


_REAL_STDOUT = 1

if __name__ == "__main__":
    a = parse()
    # stdout carries exactly one JSON line: anything a library prints there (NCCL's version banner under torchrun, for
    # one) goes to stderr instead — fd 1 is pointed at stderr for the whole run and the line is written to the saved fd
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    if a.impl == "reference":
        run_reference(a)
    else:
        run_native(a)
