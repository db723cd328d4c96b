#!/usr/bin/env python
"""bench.py - scans/s of the EFGHNet lattice hot path (BASELINE.json metric) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A STEP is one pass of the hot path over one batch of synthetic scans: for every scan of the rank's batch,
the 5-level permutohedral lattice build + the five E-Net BCL layers forward (reference
nets/enet.py:107-141 between generate_data(...) and bcn5(...)).  Workload at N=1 = BASELINE.json
configs[1]: full Ouster OS1-64 scans, 131 072 points each.  Scans are independent, so N GPUs each take
their own batch (weak scaling, no data-path collective).

One JSON line on stdout (rank 0).  `value` = scans/s with inputs resident in HBM; `e2e` = the same
through ScanPipeline.forward_host with pinned HOST buffers (H2D + D2H inside the timed region);
`roofline` = the dominant kernel against the measured peak; `cpu_baseline` = the CPU oracle port on this
box's host cores (rank 0, N=1 only).  --impl reference times that CPU path alone.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "scans/sec lattice build + BCL fwd (131k pts)"
UNIT = "scans/s"
SENSOR = "os1-64"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=32, help="scans per GPU per step")
    ap.add_argument("--streams", type=int, default=1,
                    help="concurrent COMPUTE streams per GPU (the end-to-end leg always double-buffers its copies on separate streams)")
    ap.add_argument("--scan-batch", type=int, default=16,
                    help="scans per launch sequence (ragged batched lattices, SURVEY §8 f2); 1 = one launch sequence per scan")
    ap.add_argument("--sensor", default=SENSOR)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--stages", action="store_true", help="print a per-stage timing table to stderr")
    ap.add_argument("--stem", action="store_true",
                    help="SURVEY §8 f1: compute E-Net's pointwise stem inside the level-0 splat; the cloud is then the only input")
    ap.add_argument("--int32-only", action="store_true",
                    help="do not write the reference-format int64 copies of lattice_offset / blur_neighbors (the BCL kernels read int32)")
    ap.add_argument("--atomic-splat", action="store_true", help="splat with vector atomics instead of the gather-form splat")
    ap.add_argument("--no-graph", action="store_true", help="enqueue kernel by kernel instead of replaying CUDA graphs")
    return ap.parse_args()


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port (C lattice build, 1 thread; torch-CPU BCL forward, all host threads)
# ------------------------------------------------------------------------------------------------
def cpu_scan_seconds(pc, feat0, weights, reps):
    import numpy as np
    import torch
    from oracle import lattice as ol, bcl as obcl
    from efgh_b200 import synth
    variant = "ref" if ol.has_ref() else "port"
    times, split = [], None
    for _ in range(reps):
        t0 = time.perf_counter()
        data = ol.generate(pc, synth.SCALE_MAP, variant)
        t1 = time.perf_counter()
        with torch.no_grad():
            prev = torch.from_numpy(feat0)[None]
            for li, d in enumerate(data):
                x = torch.cat((torch.from_numpy(d["pc1_el_minus_gr"]), prev), 1)
                prev = obcl.bcl_forward(x, torch.from_numpy(d["pc1_barycentric"]), torch.from_numpy(d["pc1_lattice_offset"]),
                                        torch.from_numpy(d["pc1_blur_neighbors"]), weights[li], dtype=torch.float32)
        t2 = time.perf_counter()
        times.append(t2 - t0)
        split = (t1 - t0, t2 - t1)
    return float(np.median(times)), split, variant


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np
    import torch
    from efgh_b200 import synth
    from efgh_b200.pipeline import make_enet_weights
    from oracle import lattice as ol
    ol.build()
    try:   # torchrun exports OMP_NUM_THREADS=1; the CPU arm may use every host core it is allowed to
        torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
    except Exception:
        pass
    weights = make_enet_weights(synth.ENET_BCL)
    rng = np.random.default_rng(0)
    pc = synth.synth_scan(0, args.sensor)
    feat0 = rng.standard_normal((32, pc.shape[1])).astype(np.float32)
    for _ in range(max(args.warmup, 1) if args.warmup < 2 else 1):
        cpu_scan_seconds(pc, feat0, weights, 1)
    t0 = time.perf_counter()
    splits = []
    for k in range(args.steps):
        _, sp, variant = cpu_scan_seconds(pc, feat0, weights, 1)
        splits.append(sp)
    dt = time.perf_counter() - t0
    v = args.steps / dt
    cores = torch.get_num_threads()
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "configs[1]: one %s scan (%d pts) per step, 5-level lattice build + 5 BCL fwd" % (args.sensor, pc.shape[1])},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": "1 scan per step; C oracle lattice build (1 thread, %s hash map) %.3f s + torch-CPU BCL fwd (%d threads) %.3f s"
                                       % ("reference khash" if variant == "ref" else "ported", float(np.median([s[0] for s in splits])),
                                          cores, float(np.median([s[1] for s in splits])))},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
class ClockSampler(object):
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for nm, val in zip(names, f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from efgh_b200 import synth, _capi
    from efgh_b200.pipeline import ScanPipeline, make_enet_weights

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the product path has no CPU fallback (use --impl reference for the CPU arm)")
    _capi.lib()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    B = args.batch
    G = max(1, min(args.scan_batch, B))                 # scans per launch sequence
    if B % G:
        raise SystemExit("bench.py: --batch must be a multiple of --scan-batch")
    NG = B // G                                         # launch sequences ("groups") per step
    P = max(1, min(args.streams, NG))                   # compute streams
    NP = max(P, min(NG, 2))                             # pipelines (buffer sets): >= 2 so that copies of one batch overlap kernels of another
    weights = make_enet_weights(synth.ENET_BCL)
    # this rank's scans: global scan index = rank + world * j  (scan-index sharding)
    seeds = [rank + world * j for j in range(B)]
    clouds = [synth.synth_scan(sd, args.sensor) for sd in seeds]
    N = clouds[0].shape[1]
    rng = np.random.default_rng(1000 + rank)
    feats = [rng.standard_normal((32, N)).astype(np.float32) for _ in range(B)]
    # resident inputs, one (3, G*N) / (32, G*N) pair per group: scan b of the group in columns [b*N, (b+1)*N)
    pc_dev = [torch.from_numpy(np.concatenate(clouds[g * G:(g + 1) * G], axis=1)).to(dev) for g in range(NG)]
    ft_dev = [torch.from_numpy(np.concatenate(feats[g * G:(g + 1) * G], axis=1)).to(dev) for g in range(NG)]
    stem = None
    if args.stem:   # random-init conv_in (reference nets/enet.py:24-28): 3 -> 32 -> 32 -> 32, LeakyReLU(0.1)
        gs_ = torch.Generator().manual_seed(5)
        stem = ([(torch.randn(co, ci, 1, generator=gs_) * 0.3, torch.randn(co, generator=gs_) * 0.1) for ci, co in ((3, 32), (32, 32), (32, 32))], True)
        ft_dev = [None] * NG
    pipes = [ScanPipeline(N, synth.SCALE_MAP, synth.ENET_BCL, weights, dev, vertex_cap_factor=1.0, batch=G,
                          gather_splat=not args.atomic_splat, stem=stem, emit_int64=not args.int32_only) for _ in range(NP)]
    pipe1 = pipes[0] if G == 1 else ScanPipeline(N, synth.SCALE_MAP, synth.ENET_BCL, weights, dev, vertex_cap_factor=1.0,
                                                 gather_splat=not args.atomic_splat, stem=stem)
    streams = [torch.cuda.Stream(dev) for _ in range(P)]
    copy_streams = [torch.cuda.Stream(dev) for _ in range(NP)]
    main = torch.cuda.current_stream(dev)

    use_graph = not args.no_graph
    graphs = None
    if use_graph:   # one CUDA graph per resident group (captures the whole 5-level launch sequence on that group's stream)
        graphs = [pipes[j % NP].graph_for(pc_dev[j], ft_dev[j], streams[j % P]) for j in range(NG)]

    def step(timers=None):
        for j in range(NG):
            if use_graph and timers is None:
                with torch.cuda.stream(streams[j % P]):
                    graphs[j].replay()
            else:
                pipes[j % NP].enqueue(pc_dev[j], ft_dev[j], stream=streams[j % P], timers=timers)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    def timed_region(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(main)
        for st in streams + copy_streams:
            st.wait_event(e0)
        for _ in range(steps):
            fn()
        for st in streams + copy_streams:
            ev = torch.cuda.Event()
            ev.record(st)
            main.wait_event(ev)
        e1.record(main)
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    # ---- warm-up (also validates capacities: raises if any level overflowed)
    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize(dev)
    counts = pipes[0].counts()                          # totals over the G scans of a group
    for p in pipes[1:]:
        p.counts()
    vs = pipes[0].vertex_starts()
    counts_scan0 = [v[1] - v[0] for v in vs]

    # ---- timed region, inputs resident in HBM; the dominant kernel carries CUDA events on its own stream
    DOM = "L0.conv1"
    sampler = ClockSampler(local) if rank == 0 else None
    ms = timed_region(step, args.steps)
    clocks = sampler.stop() if sampler else None
    scans = B * world * args.steps
    value = scans / (ms * 1e-3)
    # the dominant kernel, timed live with CUDA events on its own stream over one more pass of the same work
    # (eager launches: events cannot be read back from inside a replayed graph)
    timers = {DOM: []}
    step(timers)
    torch.cuda.synchronize(dev)
    dom_ms = float(np.mean([a.elapsed_time(b) for a, b in timers[DOM]]))

    # ---- end to end: pinned host buffers -> H2D -> scan -> D2H of the result rows + level records
    out_rows = 2048
    pc_pin = [torch.from_numpy(c).pin_memory() for c in clouds]
    ft_pin = [torch.from_numpy(f).pin_memory() for f in feats]
    out_pin = [torch.empty((out_rows, synth.ENET_BCL[-1][1][-1]), dtype=torch.float32).pin_memory() for _ in range(NG)]
    st_pin = [torch.empty((len(synth.SCALE_MAP), 24), dtype=torch.int32).pin_memory() for _ in range(NG)]
    vs_pin = [torch.empty((len(synth.SCALE_MAP), G + 1), dtype=torch.int32).pin_memory() for _ in range(NG)]

    def step_e2e():
        for j in range(NG):
            # copies on the pipeline's own stream, kernels on a shared compute stream: H2D of batch j+1 overlaps batch j
            if G == 1:
                pipes[j % NP].forward_host(pc_pin[j], ft_pin[j], out_pin[j], st_pin[j], stream=copy_streams[j % NP],
                                           use_graph=use_graph, compute_stream=streams[j % P])
            else:
                pipes[j % NP].forward_host(pc_pin[j * G:(j + 1) * G], ft_pin[j * G:(j + 1) * G], out_pin[j], st_pin[j],
                                           stream=copy_streams[j % NP], use_graph=use_graph, starts_host=vs_pin[j],
                                           compute_stream=streams[j % P])

    for _ in range(2):
        step_e2e()
    e2e_steps = max(2, args.steps // 2)
    ms_e2e = timed_region(step_e2e, e2e_steps)
    e2e_value = B * world * e2e_steps / (ms_e2e * 1e-3)
    h2d = B * (clouds[0].nbytes + (feats[0].nbytes if stem is None else 0))
    d2h = NG * (out_pin[0].numel() * 4 + st_pin[0].numel() * 4 + (vs_pin[0].numel() * 4 if G > 1 else 0))
    assert int(st_pin[0][0, 1]) == counts[0] or NG > NP  # the records really came back

    # ---- single-scan latency and per-stage table (outside the timed region)
    lat = []
    pc1_dev = pc_dev[0][:, :N].contiguous()
    ft1_dev = ft_dev[0][:, :N].contiguous() if stem is None else None
    g1 = pipe1.graph_for(pc1_dev, ft1_dev, streams[0]) if use_graph else None
    for _ in range(5):
        torch.cuda.synchronize(dev)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(streams[0])
        if use_graph:
            with torch.cuda.stream(streams[0]):
                g1.replay()
        else:
            pipe1.enqueue(pc1_dev, ft1_dev, stream=streams[0])
        b.record(streams[0])
        torch.cuda.synchronize(dev)
        lat.append(a.elapsed_time(b))
    stage_t = {"*": []}
    for _ in range(3):
        pipes[0].enqueue(pc_dev[0], ft_dev[0], stream=streams[0], timers=stage_t)
    torch.cuda.synchronize(dev)
    stages = {k: float(np.median([a.elapsed_time(b) for a, b in v])) * 1e3 for k, v in stage_t.items() if k != "*"}
    if args.stages and rank == 0:
        for k in sorted(stages, key=lambda k: (k.split(".")[0], -stages[k])):
            print("  %-14s %9.1f us" % (k, stages[k]), file=sys.stderr)
        print("  sum            %9.1f us; single-scan latency %.1f us" % (sum(stages.values()), 1e3 * float(np.median(lat))), file=sys.stderr)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (level-0 gather-convolution)
    peaks = measured_peaks()
    lv0 = pipes[0].levels[0]
    H0 = counts[0]
    K0 = lv0["F"] * lv0["cin"]
    dom_bytes = 4 * lv0["cin"] * (H0 + 1) + 4 * (H0 + 1) + 4 * lv0["F"] * H0 + 4 * lv0["cmid"] * H0 + 4 * K0 * lv0["cmid"]
    dom_flops = 2.0 * H0 * K0 * lv0["cmid"]
    total_bytes, _ = pipes[0].algorithmic_bytes(counts)
    total_bytes /= G                                    # per scan
    traffic = None
    tp = os.path.join(ROOT, "profiles", "r1_dominant_kernel.json")   # dram bytes per launch from one `ncu --set full` capture
    if os.path.exists(tp):
        try:
            tj = json.load(open(tp))
            traffic = tj.get("dram_bytes_per_launch") if tj.get("scans_per_launch", 1) == G else None
        except Exception:
            traffic = None
    ach = dom_bytes / (dom_ms * 1e-3) / 1e9
    roof = {"kernel": "k_conv_tc level 0 (%d scans per launch): neighbour gather + (15,1) convolution (%s)" % (G, pipes[0].precision), "bound": "hbm",
            "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": ach / peaks["hbm_gbs"], "traffic": traffic,
            "peak_source": peaks["source"] + " copy bandwidth",
            "kernel_ms": dom_ms, "algorithmic_bytes": dom_bytes, "tensor_tflops": dom_flops / (dom_ms * 1e-3) / 1e12,
            "share_of_scan": stages.get(DOM, 0.0) / max(sum(stages.values()), 1e-9),
            "note": "instruction-issue bound (profiles/r1_conv_tc_stalls_batch8_v2.txt): 62 % of issue slots, L2 hit rate 85 %, "
                    "DRAM traffic 1.3x algorithmic; FLOPs are 3x this figure on the tensor pipe (3xTF32)"}
    scan_ms = ms / (B * args.steps)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": "configs[1]: %s scans (%d pts), 5-level lattice build + 5 E-Net BCL fwd" % (args.sensor, N),
                   "scans_per_gpu_per_step": B, "scans_per_launch_sequence": G, "compute_streams": P, "pipelines": NP, "levels_H": counts_scan0,
                   "l2_policy": ("inputs larger than L2 (%d scans x %.1f MB resident, cycled)" % (B, (clouds[0].nbytes + feats[0].nbytes) / 1e6))
                                if stem is None else
                                ("working set larger than L2: every step streams %.0f MB of lattice / feature buffers (%d scans x %.0f MB algorithmic)"
                                 % (B * total_bytes / 1e6, B, total_bytes / 1e6)),
                   "splat": "levels 1-4 gather through vertex -> contributions lists, level 0 atomic scatter" if pipes[0].gather_splat else "atomic scatter",
                   "stem": "conv_in fused into the level-0 splat (input = cloud only)" if stem is not None else "stem features are an input (32 x N f32)",
                   "lattice_index_dtype": "int64 (reference format) + int32 copies for the BCL kernels" if pipes[0].emit_int64 else "int32 only",
                   "conv_precision": pipes[0].precision, "cuda_graphs": use_graph, "single_scan_latency_ms": float(np.median(lat)),
                   "algorithmic_MB_per_scan": total_bytes / 1e6,
                   "scan_roofline_frac": total_bytes / (scan_ms * 1e-3) / 1e9 / peaks["hbm_gbs"]},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": pipes[0].launches_per_scan * NG * args.steps,
        "roofline": roof,
        "stages_us": stages,          # per launch sequence (G scans), eager single-stream pass
    }
    if world == 1 and not args.no_cpu_baseline:
        try:
            try:
                torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
            except Exception:
                pass
            sec, split, variant = cpu_scan_seconds(clouds[0], feats[0], weights, 3)
            line["cpu_baseline"] = {"value": 1.0 / sec, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                                    "sample": "1 scan (seed 0), median of 3: C oracle lattice build (1 thread, %s hash map) %.3f s + torch-CPU BCL fwd (%d threads) %.3f s"
                                              % ("reference khash" if variant == "ref" else "ported", split[0], torch.get_num_threads(), split[1])}
        except Exception as e:  # the oracle is optional infrastructure; the product numbers stand without it
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": "unavailable: %r" % (e,)}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
