#!/usr/bin/env python
"""Benchmark of EgoPack's temporal-graph hot path on B200 (BASELINE.json metric: graph-nodes/s, forward+backward).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference] [--videos V] [--nodes n]

Workload (BASELINE.json configs[1], "c2"): multi-task AR+LTA+PNR training step over ONE shared temporal GNN
(experiments/mtl.yaml: k=1, hidden 1024, depth 3, TRN hidden 1024, dropout 0.5), synthetic Ego4D-shaped features
``[N, 3, 1536]`` fp32.  A step = zero_grad + 3 Graph forwards + 3 task heads + losses + ONE backward + gradient
all-reduce (N>1) + Adam step -- i.e. the body of main_temporal.py:76-130; nothing is skipped.  Per GPU every
task batch holds ``--videos`` graphs of ``--nodes`` segments (weak scaling: per-GPU work is fixed as N grows).

One JSON line on rank 0.  ``value`` = nodes/s with inputs resident in HBM; ``e2e`` = the same step fed through the
package's public feed (``egopack_b200.feed.DeviceFeeder``) from pinned HOST memory every step (H2D of features /
labels / structure inside the timed region, device-side graph transforms, D2H of the loss);
``roofline`` = the dominant kernel family (tcgen05 GEMMs) traced per launch with CUDA events on the launching
stream, ``roofline_hbm`` = every memory-bound kernel of the step the same way; ``cpu_baseline`` = the oracle (CPU
restatement of the reference) on a bounded sample of the same workload; ``extra`` (N=1) = the other BASELINE
configurations (c3 EgoPack backpack, c4 long video) measured device-resident in the same run.
``--impl reference`` times the CPU oracle alone (the reference's own torch_geometric stack is not installable, and
/root/reference does not exist on the GPU box).

Features: the loader stores the Omnivore features as bf16 (``Batch.to_feature_dtype``; ``--feature-dtype fp32`` keeps
the reference's fp32 storage).  In the bf16 compute mode that is bit-identical to feeding fp32 features (the first
GEMM's operand is the same round-to-nearest-even either way, tests/test_gpu_models.py) and halves the PCIe bytes.
PNR features are one vector per node repeated over the three segments, as the reference's dataset builds them
(data/ego4d_oscc.py:291) -- for every arm, the CPU one included; the e2e feed ships that vector once and repeats it on the
device (``egopack_b200.data.replicated_base``; ``EGP_BENCH_COMPACT=0`` ships the materialised tensor instead).
``step_ms`` = per-step spread of the timed regions (median / min / max, cudaMalloc calls inside them).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

METRIC = "temporal_graph_nodes_per_sec_fwd_bwd"
UNIT = "nodes/s"
TASKS = ("ar", "lta", "pnr")
HIDDEN, DEPTH, TRN_HIDDEN, DROPOUT, K_RADIUS = 1024, 3, 1024, 0.5, 1
SEED = 1


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--videos", type=int, default=256, help="graphs per task batch per GPU")
    ap.add_argument("--nodes", type=int, default=128, help="segments (nodes) per graph")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--cpu-videos", type=int, default=64, help="graphs per task batch of the CPU sample (bounded sample)")
    ap.add_argument("--feature-dtype", default="auto", choices=["auto", "bf16", "fp32"],
                    help="host storage of the features: auto = bf16 in the bf16 compute mode, fp32 in the fp32 parity mode")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --videos graphs per task PER GPU; strong: --videos graphs per task in TOTAL, split over the ranks")
    ap.add_argument("--optimizer", default="flat", choices=["flat", "torch"],
                    help="flat = egopack_b200.optim.FlatAdam (one kernel over flat buffers, writes the bf16 weight copies); "
                         "torch = torch.optim.Adam(fused=True)")
    ap.add_argument("--no-extra", action="store_true", help="skip the c3 / c4 sub-lines of the N=1 run")
    ap.add_argument("--cpu-sweep", action="store_true", help="(--impl reference) also time 16 / 64 / 256 graphs per task, one step each")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="c2", choices=["c2", "c3", "c4"],
                    help="c2 = MTL AR+LTA+PNR (the BASELINE metric's config); c3 = EgoPack OSCC + AR/LTA/PNR prototype backpack; "
                         "c4 = long-video stress (AR, 256 graphs x 2048 segments, radius 16, 4 GNN layers)")
    ap.add_argument("--protos", type=int, default=4096, help="prototypes per bank (c3)")
    ap.add_argument("--quick", action="store_true", help="device-resident timing only (for profiler runs)")
    ap.add_argument("--trace-out", default=None, help="write the per-launch trace summary to this JSON file")
    ap.add_argument("--gap-profile", default=None,
                    help="(with --quick) profile two more steps with torch.profiler and write kernel busy/idle statistics here")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"hbm": p["hbm_gbs"], "tensor_burst": p["bf16_tflops"],
                "tensor_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]), "source": "measured"}
    return {"hbm": 6650.0, "tensor_burst": 1590.0, "tensor_sustained": 1400.0, "source": "fallback"}


# ------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle (reference restatement) -- used for cpu_baseline and for --impl reference
# ------------------------------------------------------------------------------------------------------------
def cpu_oracle_run(videos: int, nodes: int, steps: int, warmup: int):
    from oracle import egopack_oracle as eo
    from oracle import pyg_restated as pyg
    from egopack_b200 import synthetic as syn
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    torch.manual_seed(SEED)
    model = eo.GraphOracle(syn.FEATURE_DIM, HIDDEN, DEPTH, temporal_pooling={"hidden_size": TRN_HIDDEN, "dropout": DROPOUT},
                           num_segments=syn.NUM_SEGMENTS)
    heads = (syn.N_VERBS, syn.N_NOUNS)
    tasks = {"ar": eo.RecognitionTaskOracle(HIDDEN, HIDDEN, heads), "lta": eo.LTATaskOracle(HIDDEN, HIDDEN, heads),
             "pnr": eo.PNRTaskOracle(HIDDEN, HIDDEN)}
    params = list(model.parameters()) + [p for t in tasks.values() for p in t.parameters()]
    opt = torch.optim.Adam(params, lr=1e-5, weight_decay=1e-5)
    model.train()
    for t in tasks.values():
        t.train()
    gen = syn.generator(SEED, 2, 0)
    batches = {}
    for t in TASKS:
        b = syn.make_batch(t, videos, nodes, gen)
        d = pyg.Data(x=b.x, pos=b.pos, y=b.y)
        d.batch, d.ptr = b.batch, b.ptr
        if t == "lta":                                       # per-sample transform, then collate offsets
            eis = []
            for g in range(videos):
                s = pyg.Data(x=b.x[g * nodes:(g + 1) * nodes], pos=b.pos[g * nodes:(g + 1) * nodes], y=b.y[g * nodes:(g + 1) * nodes])
                eis.append(eo.lta_temporal_connectivity(s, K_RADIUS + 0.5).edge_index + g * nodes)
            d.edge_index = torch.cat(eis, 1)
        else:
            d.edge_index = pyg.radius_graph(b.pos, K_RADIUS + 0.5, b.batch)
        batches[t] = d
    n_nodes = videos * nodes * len(TASKS)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        opt.zero_grad()
        loss, _ = eo.mtl_step(model, tasks, batches)
        loss.backward()
        opt.step()
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    med = statistics.median(times)
    return {"value": n_nodes / med, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{len(TASKS)} task batches x {videos} graphs x {nodes} nodes ({n_nodes} nodes/step), fp32, "
                      f"{steps} timed steps (median {med:.3f} s/step), torch {torch.__version__} CPU, "
                      f"oracle = CPU restatement of the reference (torch_geometric is not installable here)",
            "ms_per_step": med * 1e3}


# ------------------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """`nvidia-smi -lms` in the background (started BEFORE the warm-up so it is already streaming); `window()` keeps
    the samples whose timestamps fall inside the timed region (padded by one sampling period)."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    PERIOD_MS = int(os.environ.get("EGP_BENCH_SAMPLER_MS", "50"))   # NVML queries take driver locks: see profiles/README.md

    def __init__(self, index: int):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        if self.PERIOD_MS <= 0:
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", str(self.PERIOD_MS)],
                                         stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, t0: float, t1: float):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(2 * self.PERIOD_MS / 1e3)
        self.proc.terminate()
        pad = 2 * self.PERIOD_MS / 1e3
        inside = [r for t, r in self.rows if t0 - pad <= t <= t1 + pad]
        scope = "timed region"
        if len(inside) < 3:                                  # very short timed regions: fall back to the whole run
            inside, scope = [r for _, r in self.rows], "warm-up + timed region"
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in inside:
            try:
                sm.append(float(r[1]))
                mx = float(r[2])
                for nme, v in zip(names, r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
            except (ValueError, IndexError):
                continue
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm), "scope": scope}


# ------------------------------------------------------------------------------------------------------------
# native arm
# ------------------------------------------------------------------------------------------------------------
WORKLOAD_TEXT = {
    "c2": "c2: MTL AR+LTA+PNR shared temporal GNN (experiments/mtl.yaml: k=1, hidden 1024, depth 3, TRN hidden 1024, "
          "dropout 0.5), full train step = zero_grad+fwd+loss+bwd+grad-allreduce+Adam",
    "c3": "c3: EgoPack OSCC primary + frozen AR/LTA/PNR prototype backpack (k=4, depth 3, residual, {protos} "
          "prototypes/bank, late fusion), Graph trainable, full train step",
    "c4": "c4: long-video stress (BASELINE configs[3]): AR over 2048-segment graphs, temporal radius 16 (33-wide band), "
          "4 GNN layers, hidden 1024, full train step",
}
# algorithmic HBM bytes are attached to the traced launches in egopack_b200/ops.py (2*C*b per node for the
# aggregation, 3*C*b / 5*C*b graph-LN fwd / bwd, 2*C*b / 3-4*C*b row-LN fwd / bwd, 2*C*b posenc, C*b pooling)
HBM_KERNELS = ("sage_mean_band", "sage_mean_band_star", "sage_mean_csr", "graph_layernorm_fwd", "graph_layernorm_bwd",
               "row_layernorm_fwd", "row_layernorm_bwd", "posenc_add", "act_bwd_colsum", "colsum", "cast",
               "segment_max_pool_fwd", "proto_max_gather", "max_combine_fwd")


def load_traffic():
    """DRAM bytes per launch measured under `ncu --set full` for the kernels that ship in THIS tree (profiles/r2_*)."""
    for name in ("r2_ncu_traffic.json",):
        path = os.path.join(ROOT, "profiles", name)
        if os.path.exists(path):
            return {k: v for k, v in json.load(open(path)).items() if isinstance(v, dict)}
    return {}


def native_run(args, rank: int, world: int, local_rank: int, workload: str = "c2", full: bool = True):
    """One workload on this rank's GPU.  `full` adds the e2e, small-batch and roofline sections (the main line);
    the c3 / c4 sub-lines of an N=1 run only take the device-resident timing plus the kernel trace."""
    import gc

    import egopack_b200
    from egopack_b200 import _lib, ops, steps
    from egopack_b200 import synthetic as syn
    from egopack_b200.data import replicated_base
    from egopack_b200.dp import GradientAllReduce
    from egopack_b200.feed import DeviceFeeder, bind_host_memory_to_gpu
    from egopack_b200.models.graph import Graph
    from egopack_b200.models.graphONE.graphONE import GraphONE
    from egopack_b200.models.tasks import LTATask, OSCCTask, PNRTask, RecognitionTask
    from egopack_b200.models.transforms import LTATemporalConnectivity, RadiusGraph

    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    numa = bind_host_memory_to_gpu(local_rank) if full else None   # before any pinned allocation
    egopack_b200.set_precision(args.precision)
    torch.manual_seed(SEED)                                   # identical replicas on every rank
    c3, c4 = workload == "c3", workload == "c4"
    k_radius, depth = (16, 4) if c4 else (K_RADIUS, DEPTH)
    videos, nodes = args.videos, args.nodes
    if c4:                                                    # BASELINE configs[3]: 256 videos x 2048 segments
        videos, nodes = (256, 2048) if (args.videos, args.nodes) == (256, 128) else (args.videos, args.nodes)
    if args.scaling == "strong":
        from egopack_b200.dp import shard_graphs
        videos = len(shard_graphs(videos, rank, world))
    feat_dtype = {"auto": torch.bfloat16 if args.precision == "bf16" else torch.float32,
                  "bf16": torch.bfloat16, "fp32": torch.float32}[args.feature_dtype]
    model = Graph(syn.FEATURE_DIM, HIDDEN, depth, temporal_pooling={"hidden_size": TRN_HIDDEN, "dropout": DROPOUT},
                  num_segments=syn.NUM_SEGMENTS).to(dev)
    heads = (syn.N_VERBS, syn.N_NOUNS)
    graphone = None
    if c3:
        # EgoPack novel task (experiments/egopack/oscc.yaml): OSCC primary, frozen AR/LTA/PNR banks, k=4, depth 3,
        # residual, late fusion with averaged logits, Graph trainable
        aux = ("ar", "lta", "pnr")
        tasks = {"oscc": OSCCTask(HIDDEN, HIDDEN, aux_tasks=aux, average_logits=True, head_dropout=0.5).to(dev),
                 "ar": RecognitionTask(HIDDEN, HIDDEN, heads).to(dev), "lta": LTATask(HIDDEN, HIDDEN, heads).to(dev),
                 "pnr": PNRTask(HIDDEN, HIDDEN).to(dev)}
        banks = syn.make_banks(aux, args.protos, HIDDEN, syn.generator(SEED, 3, 0))
        graphone = GraphONE(banks, features_size=HIDDEN, hidden_size=HIDDEN, k=4, depth=3, residual=True).to(dev)
        graphone.train()
        task_names = ("oscc",)
    elif c4:
        tasks = {"ar": RecognitionTask(HIDDEN, HIDDEN, heads).to(dev)}
        task_names = ("ar",)
    else:
        tasks = {"ar": RecognitionTask(HIDDEN, HIDDEN, heads).to(dev), "lta": LTATask(HIDDEN, HIDDEN, heads).to(dev),
                 "pnr": PNRTask(HIDDEN, HIDDEN).to(dev)}
        task_names = TASKS
    model.train()
    for t in tasks.values():
        t.train()
    params = list(model.parameters()) + [p for t in tasks.values() for p in t.parameters()]
    if graphone is not None:
        params += [p for p in graphone.parameters() if p.requires_grad]
    if args.optimizer == "flat":
        from egopack_b200.optim import FlatAdam
        opt = FlatAdam(params, lr=1e-5, weight_decay=1e-5)
    else:
        opt = torch.optim.Adam(params, lr=1e-5, weight_decay=1e-5, fused=True)
    sync = GradientAllReduce(params) if world > 1 else None
    feed_tf = {t: (LTATemporalConnectivity(r=k_radius + 0.5) if t == "lta" else RadiusGraph(r=k_radius + 0.5))
               for t in task_names}

    gen = syn.generator(SEED, 2, rank)                        # every rank draws its own shard of graphs
    n_nodes = videos * nodes * len(task_names)
    on_device = not full                                      # sub-lines: synthetic features drawn on the device
    if on_device:
        host = None
        dgen = torch.Generator(device=dev).manual_seed(SEED + 1000 * 2 + rank)
        resident = {}
        for t in task_names:
            hb = syn.make_batch(t, videos, nodes, gen, feature_dim=1, num_segments=1, band_k=k_radius)   # labels / structure
            d = egopack_b200.Batch()
            d.x = torch.randn(videos * nodes, syn.NUM_SEGMENTS, syn.FEATURE_DIM, generator=dgen, device=dev,
                              dtype=torch.float32).to(feat_dtype)
            for k in ("pos", "y", "batch", "ptr"):
                setattr(d, k, getattr(hb, k).to(dev))
            d.pos_unit_spaced = True
            resident[t] = feed_tf[t](d)
        h2d_bytes = 0
    else:
        # compact: PNR features are one vector per node repeated over the segments (the reference's dataset,
        # data/ego4d_oscc.py:291); the loader hands them over as a stride-0 view, so the base alone crosses PCIe
        host = {t: syn.make_batch(t, videos, nodes, gen, band_k=k_radius, pin=True, feature_dtype=feat_dtype,
                                   compact=os.environ.get("EGP_BENCH_COMPACT", "1") != "0")
                for t in task_names}

        def wire(x):                                          # the tensor that actually crosses the bus
            base = replicated_base(x)
            return x if base is None else base
        h2d_bytes = sum(v.numel() * v.element_size() for b in host.values() for v in (wire(b.x), b.pos, b.y, b.batch, b.ptr))

    def host_loader(n):
        for _ in range(n):
            yield host

    def step(batches):
        opt.zero_grad(set_to_none=True)
        if c3:
            loss, _ = steps.egopack_losses(model, tasks, batches, graphone)
        else:
            loss, _ = steps.mtl_losses(model, tasks, batches)
        loss.backward()
        if sync is not None:
            sync.finish()
        opt.step()
        return loss

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms: float) -> float:
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident timing ------------------------------------------------------------------------
    if not on_device:
        resident = next(iter(DeviceFeeder(host_loader(1), dev, feed_tf)))
    clocks = ClockSampler(local_rank)
    clocks.start()
    for _ in range(args.warmup):
        loss = step(resident)    # held across the next step exactly as in the timed loop: same allocation pattern, so the
    barrier()                    # timed region finds every block it needs in the allocator's cache (no cudaMalloc in it)
    _lib.CALL_COUNTS.clear()
    for k_ in ops.KNN_STATS:
        ops.KNN_STATS[k_] = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # keep full cyclic-GC sweeps cheap inside a timed region: everything alive now (modules, parameters, the loaders'
    # batches) moves to the permanent generation, so a generation-2 pass only walks what the steps themselves create.
    # (Switching the collector OFF is not an option: a step leaves reference cycles that hold its activations, and
    # without the collector the allocator grows by a step's worth of memory per step -- measured 11.8 -> 15.4 ms/step.)
    gc.collect()
    gc.freeze()
    marks, host_t = [], []                                   # one event per step: where a slow run lost its time
    dbg_steps = [] if os.environ.get("EGP_BENCH_DEBUG_STEPS") == "1" else None
    mallocs0 = torch.cuda.memory_stats(dev).get("num_device_alloc", 0)
    t_begin = time.time()
    e0.record()
    for _ in range(args.steps):
        loss = step(resident)
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        marks.append(ev)
        host_t.append(time.time())
        if dbg_steps is not None:
            st = torch.cuda.memory_stats(dev)
            dbg_steps.append((st.get("num_device_alloc", 0), st.get("reserved_bytes.all.current", 0), gc.get_count()))
    e1.record()
    barrier()
    t_end = time.time()
    gc.unfreeze()
    mallocs = torch.cuda.memory_stats(dev).get("num_device_alloc", 0) - mallocs0
    step_ms = [a.elapsed_time(b) for a, b in zip([e0] + marks[:-1], marks)]
    host_ms = [(b - a) * 1e3 for a, b in zip([t_begin] + host_t[:-1], host_t)]
    if dbg_steps is not None:
        for i, (h, g_, d_) in enumerate(zip(host_ms, step_ms, dbg_steps)):
            print(f"[steps] {i} host {h:.2f} ms gpu {g_:.2f} ms mallocs {d_[0]} reserved {d_[1] >> 20} MiB gc {d_[2]}", file=sys.stderr)
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    clock_info = clocks.stop(t_begin, t_end)
    launches = _lib.kernel_launches()
    knn_stats = dict(ops.KNN_STATS)
    ms_per_step = ms_total / args.steps
    value = world * n_nodes / (ms_per_step / 1e3)
    last = float(loss.item())

    if args.quick and args.gap_profile and rank == 0:
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(2):
                step(resident)
            torch.cuda.synchronize()
        ks = sorted(((e.time_range.start, e.time_range.end, e.name) for e in prof.events()
                     if e.device_type == torch.autograd.DeviceType.CUDA), key=lambda t: t[0])
        busy = sum(b - a for a, b, _ in ks)
        gaps, end = [], ks[0][1]
        for a, b, nme in ks[1:]:
            if a > end:
                gaps.append((a - end, nme))
            end = max(end, b)
        span = end - ks[0][0]
        gaps.sort(reverse=True)
        hist = {"<1us": 0, "1-2us": 0, "2-5us": 0, "5-20us": 0, ">20us": 0}
        for g, _ in gaps:
            hist["<1us" if g < 1 else "1-2us" if g < 2 else "2-5us" if g < 5 else "5-20us" if g < 20 else ">20us"] += 1
        by_name = {}
        for a, b, nme in ks:
            r = by_name.setdefault(nme[:90], [0, 0.0])
            r[0] += 1
            r[1] += b - a
        os.makedirs(os.path.dirname(os.path.abspath(args.gap_profile)), exist_ok=True)
        json.dump({"steps": 2, "kernels": len(ks), "span_us": span, "busy_us": busy, "idle_us": span - busy,
                   "idle_frac": (span - busy) / span, "gap_hist": hist,
                   "largest_gaps_us_before": [(round(g, 1), n[:80]) for g, n in gaps[:25]],
                   "kernel_time_us": {k: [v[0], round(v[1], 1)] for k, v in sorted(by_name.items(), key=lambda kv: -kv[1][1])[:60]}},
                  open(args.gap_profile, "w"), indent=1)
    if args.quick:
        return {"metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": round(ms_per_step, 3), "gpu_launches": int(launches), "quick": True,
                "step_ms": {"median": round(statistics.median(step_ms), 3), "max": round(max(step_ms), 3),
                            "argmax": int(step_ms.index(max(step_ms))), "host_max": round(max(host_ms), 3),
                            "cuda_mallocs": int(mallocs)}}

    # ---- end-to-end timing: the public feed (DeviceFeeder) from pinned host memory every step, loss read back ------
    # the feeder enqueues the copies of step i+1 (and its device-side transforms) on its copy stream before it hands
    # out step i, so they overlap the step; nothing in the hand-over waits on the GPU
    e2e = None
    if full:
        for b in DeviceFeeder(host_loader(max(4, args.warmup)), dev, feed_tf):   # the feed's staging buffers reach steady state
            float(step(b).item())
        barrier()
        # what the platform gives this rank for the feature copies ALONE (all ranks copying at once, GPU otherwise idle):
        # the ceiling of any feed, reported next to the e2e rate
        probe_dst = {t: torch.empty(wire(hb.x).shape, dtype=hb.x.dtype, device=dev) for t, hb in host.items()}
        pe0, pe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for rep in range(4):
            if rep == 1:
                barrier()
                pe0.record()
            for t, hb in host.items():
                probe_dst[t].copy_(wire(hb.x), non_blocking=True)
        pe1.record()
        barrier()
        probe_ms = max_over_ranks(pe0.elapsed_time(pe1)) / 3
        h2d_alone = sum(wire(hb.x).numel() * hb.x.element_size() for hb in host.values()) / (probe_ms / 1e3) / 1e9
        del probe_dst
        feeder = DeviceFeeder(host_loader(args.steps), dev, feed_tf)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        gc.collect()
        gc.freeze()
        e2e_marks = []
        e2e_mallocs0 = torch.cuda.memory_stats(dev).get("num_device_alloc", 0)
        e0.record()
        for b in feeder:
            loss = step(b)
            last = float(loss.item())                          # D2H of the loss every step
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            e2e_marks.append(ev)
        e1.record()
        barrier()
        gc.unfreeze()
        e2e_mallocs = torch.cuda.memory_stats(dev).get("num_device_alloc", 0) - e2e_mallocs0
        e2e_step_ms = [a.elapsed_time(b) for a, b in zip([e0] + e2e_marks[:-1], e2e_marks)]
        e2e_ms = max_over_ranks(e0.elapsed_time(e1)) / args.steps
        assert feeder.h2d_bytes == h2d_bytes * args.steps, (feeder.h2d_bytes, h2d_bytes)
        e2e = {"value": round(world * n_nodes / (e2e_ms / 1e3), 1), "unit": UNIT, "h2d_bytes_per_step": int(h2d_bytes),
               "d2h_bytes_per_step": 4, "ms_per_step": round(e2e_ms, 3),
               "h2d_gbps_per_gpu": round(h2d_bytes / (e2e_ms / 1e3) / 1e9, 1),
               "h2d_alone_gbps_per_gpu": round(h2d_alone, 1),
               "step_ms": {"median": round(statistics.median(e2e_step_ms), 3), "min": round(min(e2e_step_ms), 3),
                           "max": round(max(e2e_step_ms), 3), "argmax": int(e2e_step_ms.index(max(e2e_step_ms))),
                           "cuda_mallocs": int(e2e_mallocs)},
               "h2d_bytes_per_step_if_pnr_were_materialised": int(sum(
                   v.numel() * v.element_size() for b in host.values() for v in (b.x, b.pos, b.y, b.batch, b.ptr))),
               "note": "egopack_b200.feed.DeviceFeeder: pinned host -> device copy of every step's inputs on a copy stream "
                       "(enqueued before the previous step is handed out), device-side graph transforms (band_k / star hints, "
                       "lazy edge_index), loss.item() per step; PNR features (one vector per node repeated over the segments, "
                       "data/ego4d_oscc.py:291) cross the bus once and are repeated on the device; `pnr_materialised` = the same "
                       "loop with that tensor shipped as [N,3,1536]"}
        b = None
        # the same loop with the PNR features shipped as the reference's dataset materialises them ([N,3,1536], the vector
        # three times): the number to compare when the loader cannot be changed
        if any(replicated_base(hb.x) is not None for hb in host.values()):
            from egopack_b200.data import Batch as _Batch
            host_mat = {}
            for t, hb in host.items():
                if replicated_base(hb.x) is None:
                    host_mat[t] = hb
                    continue
                mb = _Batch()
                for k_, v_ in hb._fields.items():
                    setattr(mb, k_, v_)
                mb.x = hb.x.contiguous().pin_memory()
                host_mat[t] = mb
            mat_bytes = sum(v.numel() * v.element_size() for hb in host_mat.values() for v in (hb.x, hb.pos, hb.y, hb.batch, hb.ptr))

            def mat_loader(n):
                for _ in range(n):
                    yield host_mat
            for b in DeviceFeeder(mat_loader(3), dev, feed_tf):
                float(step(b).item())
            barrier()
            feeder = DeviceFeeder(mat_loader(args.steps), dev, feed_tf)
            m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            m0.record()
            for b in feeder:
                float(step(b).item())
            m1.record()
            barrier()
            mat_ms = max_over_ranks(m0.elapsed_time(m1)) / args.steps
            assert feeder.h2d_bytes == mat_bytes * args.steps
            e2e["pnr_materialised"] = {"value": round(world * n_nodes / (mat_ms / 1e3), 1), "ms_per_step": round(mat_ms, 3),
                                       "h2d_bytes_per_step": int(mat_bytes),
                                       "h2d_gbps_per_gpu": round(mat_bytes / (mat_ms / 1e3) / 1e9, 1)}
            e2e.pop("h2d_bytes_per_step_if_pnr_were_materialised", None)
            b = None
            del host_mat, feeder

    # ---- launch-bound regime: the reference's own batch size (16 graphs/task), eager vs one CUDA graph per step ------
    small = None
    if full and world == 1 and workload == "c2":
        try:
            from egopack_b200.graphs import GraphedStep
            # AccumulateGrad nodes created on the default stream by the eager steps above must not survive into the
            # capture stream: drop every reference to the old autograd graphs first
            loss = None
            gc.collect()
            sv, sn = 16, 16
            sgen = syn.generator(SEED, 9, rank)
            sdev = {}
            for t in task_names:
                d = syn.make_batch(t, sv, sn, sgen, band_k=k_radius, feature_dtype=feat_dtype).to(dev)
                d.pos_unit_spaced = True
                sdev[t] = feed_tf[t](d)
            # FlatAdam reads its step counter and learning rate from device memory: capturable as is
            opt2 = opt if args.optimizer == "flat" else torch.optim.Adam(params, lr=1e-5, weight_decay=1e-5, fused=True,
                                                                         capturable=True)

            def small_step(b):
                opt2.zero_grad(set_to_none=True)
                l, _ = steps.mtl_losses(model, tasks, b)
                l.backward()
                opt2.step()
                return l

            def time_it(fn, n=30):
                for _ in range(5):
                    fn()
                torch.cuda.synchronize()
                a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in range(n):
                    fn()
                b_.record()
                torch.cuda.synchronize()
                return a.elapsed_time(b_) / n

            eager_ms = time_it(lambda: small_step(sdev))
            gc.collect()
            runner = GraphedStep(small_step, sdev)
            graph_ms = time_it(lambda: runner())
            nn_small = sv * sn * len(task_names)
            small = {"graphs_per_task": sv, "nodes_per_graph": sn, "nodes_per_step": nn_small,
                     "eager_ms_per_step": round(eager_ms, 3), "cuda_graph_ms_per_step": round(graph_ms, 3),
                     "eager_nodes_per_s": round(nn_small / eager_ms * 1e3, 1),
                     "cuda_graph_nodes_per_s": round(nn_small / graph_ms * 1e3, 1),
                     "note": "reference-scale batch (launch-bound): whole step (fwd+bwd+Adam) replayed as ONE CUDA graph"}
        except Exception as ex:  # noqa: BLE001 -- a probe; never let it take the bench line down
            small = {"error": f"{type(ex).__name__}: {ex}"[:300]}

    # ---- per-launch trace for the rooflines (separate, untimed steps) ---------------------------------------
    pk = peaks()
    ops.TRACE = []
    for _ in range(2):
        step(resident)
    torch.cuda.synchronize()
    trace, ops.TRACE = ops.TRACE, None
    agg = {}
    shapes = {}
    for name, work, unit, a, b, detail in trace:
        ms = a.elapsed_time(b)
        if (name.startswith("sage_mean") or name == "colsum") and detail:   # aggregation kernels differ by radius, column
            name = f"{name} {detail}"                                         # sums by width (heads: a few hundred columns)
            detail = ""
        r = agg.setdefault(name, {"work": 0.0, "ms": 0.0, "n": 0, "unit": unit})
        r["work"] += work
        r["ms"] += ms
        r["n"] += 1
        if detail:
            d = shapes.setdefault(f"{name} {detail}", {"work": 0.0, "ms": 0.0, "n": 0})
            d["work"] += work
            d["ms"] += ms
            d["n"] += 1
    traffic = load_traffic()

    def traffic_of(kernel):
        t = traffic.get(kernel)
        return (int(t["bytes"]), f"ncu dram bytes of one launch: {t['launch']} ({t['source']})") if t else (None, None)

    roof, roof_hbm = None, []
    if "gemm_tcgen05" in agg or "gemm_fp32" in agg:
        g = agg.get("gemm_tcgen05") or agg["gemm_fp32"]
        ach = g["work"] / (g["ms"] / 1e3) / 1e12
        peak = pk["tensor_sustained"]
        roof = {"bound": "tensor", "kernel": "tc_gemm_kernel (tcgen05/TMEM/TMA)" if "gemm_tcgen05" in agg else "sgemm_kernel",
                "achieved": round(ach, 1), "peak": peak, "unit": "TFLOP/s", "frac": round(ach / peak, 4),
                "traffic": traffic_of("tc_gemm_kernel")[0], "traffic_note": traffic_of("tc_gemm_kernel")[1],
                "launches_per_step": g["n"] // 2, "ms_per_step": round(g["ms"] / 2, 3),
                "share_of_step": round(g["ms"] / 2 / ms_per_step, 3), "peak_source": f"{pk['source']} (sustained bf16 GEMM)"}
    for nme, a in agg.items():
        if nme.split(" ")[0] not in HBM_KERNELS:
            continue
        ach = a["work"] / (a["ms"] / 1e3) / 1e9
        roof_hbm.append({"bound": "hbm", "kernel": nme, "achieved": round(ach, 1), "peak": pk["hbm"], "unit": "GB/s",
                         "frac": round(ach / pk["hbm"], 4), "traffic": traffic_of(nme)[0], "traffic_note": traffic_of(nme)[1],
                         "launches_per_step": a["n"] // 2, "ms_per_step": round(a["ms"] / 2, 3),
                         "mb_per_launch": round(a["work"] / a["n"] / 1e6, 2), "peak_source": pk["source"]})
    roof_hbm.sort(key=lambda r: -r["ms_per_step"])
    if args.trace_out and rank == 0 and full:
        os.makedirs(os.path.dirname(os.path.abspath(args.trace_out)), exist_ok=True)
        summary = {k: {**v, "ms_per_step": v["ms"] / 2} for k, v in agg.items()}
        summary["gemm_shapes"] = {k: {"launches_per_step": v["n"] // 2, "ms_per_step": round(v["ms"] / 2, 4),
                                      "tflops": round(v["work"] / (v["ms"] / 1e3) / 1e12, 1)}
                                  for k, v in sorted(shapes.items(), key=lambda kv: -kv[1]["ms"])}
        json.dump(summary, open(args.trace_out, "w"), indent=1)

    feat_name = "bf16" if feat_dtype == torch.bfloat16 else "fp32"
    config = {"workload": WORKLOAD_TEXT[workload].format(protos=args.protos),
              "graphs_per_task_per_gpu": videos, "nodes_per_graph": nodes,
              "nodes_per_step_per_gpu": n_nodes,
              "features": f"[N,3,1536] {feat_name} N(0,1)" + (" (stored by the loader as bf16: Batch.to_feature_dtype; bit-identical "
                                                             "to fp32 features in the bf16 compute mode)" if feat_name == "bf16" else "")
                          + ("; PNR rows are one 1536-vector per node repeated over the 3 segments, as the reference's dataset builds "
                             "them (data/ego4d_oscc.py:291) -- the feed ships the vector once and repeats it on the device"
                             if "pnr" in task_names else ""),
              "parallelism": f"dp{world}", "optimizer": "FlatAdam (egp_adam_step)" if args.optimizer == "flat" else "torch fused Adam",
              "l2": (f"inputs larger than L2 ({n_nodes * 4608 * (2 if feat_name == 'bf16' else 4) / 2**20:.0f} MiB of features per step, "
                     f"{h2d_bytes / 2**20:.0f} MiB of them over PCIe)" if not on_device else
                     f"inputs larger than L2 ({n_nodes * 4608 * 2 / 2**20:.0f} MiB of features per step, drawn on the device)"),
              "final_loss": round(last, 4)}
    if numa is not None:
        config["host_memory_binding"] = numa
    if c3:
        config["knn_guard"] = {"rows_scored_on_tensor_cores": knn_stats["rows"], "rows_rerun_exactly": knn_stats["flagged"],
                               "note": "cosine k-NN miss detector: rows where bf16 rounding could hide a neighbour are re-ranked in fp32"}
    out = {
        "metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(ms_per_step, 3), "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": args.precision, "data": "synthetic", "config": config, "clocks": clock_info,
        "e2e": e2e, "gpu_launches": int(launches), "roofline": roof,
        # per-step spread of the timed region (one CUDA event after every step; `value` stays total time / K as the contract
        # says): a host-side stall longer than the launch queue hides shows up as ONE slow step with a matching host gap
        "step_ms": {"median": round(statistics.median(step_ms), 3), "min": round(min(step_ms), 3),
                    "max": round(max(step_ms), 3), "host_max": round(max(host_ms), 3), "cuda_mallocs": int(mallocs)},
    }
    if roof_hbm:
        out["roofline_hbm"] = roof_hbm
    if small is not None:
        out["small_batch"] = small
    # release this workload's memory before the next one
    del model, tasks, opt, params, resident
    gc.collect()
    torch.cuda.empty_cache()
    return out


def main():
    # stdout carries exactly one JSON line.  Libraries write to file descriptor 1 on their own (NCCL prints a
    # "NCCL version ..." banner there at any NCCL_DEBUG level >= VERSION), so fd 1 is pointed at stderr for the whole run
    # and the result line goes to a private duplicate of the original stdout.
    sys.stdout.flush()
    result_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if rank != 0:
            return
        r = cpu_oracle_run(args.cpu_videos, args.nodes, max(args.steps, 1), max(args.warmup, 0))
        line = {
            "impl": "reference", "metric": METRIC, "value": round(r["value"], 1), "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(r["ms_per_step"], 1),
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
            # the arm's own config; every CPU step is a bounded sample of it (cpu_baseline.sample says which)
            "config": {"workload": WORKLOAD_TEXT["c2"], "graphs_per_task_per_gpu": args.videos,
                       "nodes_per_graph": args.nodes, "nodes_per_step_per_gpu": args.videos * args.nodes * len(TASKS),
                       "features": "[N,3,1536] fp32 N(0,1)", "parallelism": "cpu",
                       "sample_graphs_per_task": args.cpu_videos},
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": round(r["value"], 1), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }
        if args.cpu_sweep:
            line["sweep_nodes_per_s"] = {str(v): round(cpu_oracle_run(v, args.nodes, 1, 1)["value"], 1) for v in (16, 64, 256)}
        print(json.dumps(line), file=result_out, flush=True)
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py (native arm) needs a CUDA device; there is no CPU fallback. Use --impl reference for the CPU arm.")
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    out = native_run(args, rank, world, local_rank, args.workload, full=True)
    if world == 1 and not args.quick and not args.no_extra and args.workload == "c2":
        # the other BASELINE configurations, device-resident, in the same run (configs[2] and configs[3])
        extra = {}
        sub = argparse.Namespace(**vars(args))
        sub.steps, sub.warmup, sub.trace_out = min(args.steps, 5), 3, None
        for wl in ("c3", "c4"):
            try:
                r = native_run(sub, rank, world, local_rank, wl, full=False)
                extra[wl] = {k: r[k] for k in ("value", "unit", "ms_per_step", "steps", "warmup", "config", "clocks",
                                               "gpu_launches", "roofline", "roofline_hbm") if k in r}
            except Exception as ex:  # noqa: BLE001 -- never let a sub-line take the main line down
                extra[wl] = {"error": f"{type(ex).__name__}: {ex}"[:300]}
        out["extra"] = extra
    if rank == 0:
        if not args.no_cpu_baseline and world == 1 and not args.quick:
            r = cpu_oracle_run(args.cpu_videos, args.nodes, 3, 1)
            out["cpu_baseline"] = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}
        elif not args.no_cpu_baseline:
            out["cpu_baseline"] = None
        print(json.dumps(out), file=result_out, flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
