#!/usr/bin/env python
"""bench.py - scans/s of the EFGHNet lattice hot path (BASELINE.json metric) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--train]

A STEP is one pass of the hot path over one batch of synthetic scans: for every scan of the rank's batch, the 5-level
permutohedral lattice build + the five E-Net BCL layers forward (reference nets/enet.py:107-141 between
generate_data(...) and bcn5(...)), with E-Net's pointwise stem computed inside the level-0 splat so that the cloud is
the path's only input (reference nets/enet.py:103-111: the input is `pc`).  Workload at N=1 = BASELINE.json
configs[1]: full Ouster OS1-64 scans, 131 072 points each.  Scans are independent, so N GPUs each take their own batch
(weak scaling, no data-path collective).

One JSON line on stdout (rank 0):
  value        scans/s with inputs resident in HBM;
  e2e          the same through ScanPipeline.forward_host with pinned HOST clouds: H2D of every scan's cloud, the
               scan, D2H of the WHOLE last-level output + level records, all inside the timed region;
  roofline     the dominant kernel against the measured peak, `traffic` = its DRAM bytes measured IN THIS RUN by an
               `ncu --metrics dram__bytes_*` pass over one launch sequence (null + reason when ncu cannot run);
  roofline_levels / stages_us / stages_traffic   per-stage detail;
  module_path  the reference's own operator API (drop-in GenerateData + 5 BilateralConvFlex), forward, scans/s;
  train        BASELINE configs[3] (8 x 65 536-point scans per GPU: fwd + bwd + NCCL gradient all-reduce + Adam);
  cpu_baseline the reference's CPU path on this box's host cores (rank 0, N=1 only).
--impl reference times the CPU path alone: the UNMODIFIED reference files (baseline/_ref, kind "reference") when they
are installed, else the C/torch oracle port (kind "port").  --train makes configs[3] the headline of the line.
"""
import argparse
import csv
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "scans/sec lattice build + BCL fwd (131k pts)"
METRIC_TRAIN = "scans/sec lattice build + BCL fwd+bwd + grad allreduce + Adam (65k pts)"
UNIT = "scans/s"
SENSOR = "os1-64"
TRAIN_SENSOR = "os1-64-64k"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--train", action="store_true", help="headline = BASELINE configs[3]: training step on 8 x 65 536-point scans per GPU")
    ap.add_argument("--train-impl", default="auto", choices=["auto", "batched", "module"],
                    help="batched: one ScanPipeline launch sequence forward + backward; module: drop-in modules under torch.autograd, scan by scan")
    ap.add_argument("--train-scans", type=int, default=8, help="scans per GPU per training step")
    ap.add_argument("--batch", type=int, default=384, help="scans per GPU per step (cycled over --resident distinct scans)")
    ap.add_argument("--resident", type=int, default=32, help="distinct scans resident per GPU")
    ap.add_argument("--streams", type=int, default=1, help="concurrent COMPUTE streams per GPU")
    ap.add_argument("--scan-batch", type=int, default=16,
                    help="scans per launch sequence (ragged batched lattices, SURVEY §8 f2); 1 = one launch sequence per scan")
    ap.add_argument("--sensor", default=SENSOR)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the module-path, feature-input, train and ncu-traffic legs")
    ap.add_argument("--stages", action="store_true", help="print a per-stage timing table to stderr")
    ap.add_argument("--no-stem", action="store_true",
                    help="take E-Net's (32, N) stem features as a second input instead of computing them in the level-0 splat")
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


def workload_config(args, train):
    """The `config` object - identical in both arms (ours / reference) for the same command line."""
    from efgh_b200 import synth
    if train:
        n = synth.SENSORS[TRAIN_SENSOR][0] * synth.SENSORS[TRAIN_SENSOR][1]
        return {"workload": "configs[3]: training step on %d %s scans (%d pts) per GPU: 5-level lattice build + E-Net stem + 5 BCL "
                            "fwd+bwd, gradient all-reduce, Adam" % (args.train_scans, TRAIN_SENSOR, n),
                "sensor": TRAIN_SENSOR, "points_per_scan": n, "levels": len(synth.SCALE_MAP),
                "l2_policy": "working set larger than L2: one step streams > 1 GB of lattice / feature / gradient buffers"}
    n = synth.SENSORS[args.sensor][0] * synth.SENSORS[args.sensor][1]
    return {"workload": "configs[1]: %s scans (%d pts), 5-level lattice build + 5 E-Net BCL fwd (value: cloud + stem features resident in "
                        "HBM; e2e: %s)" % (args.sensor, n, "cloud + features from pinned host memory" if args.no_stem else
                                           "cloud only from pinned host memory, E-Net stem fused into the level-0 splat"),
            "sensor": args.sensor, "points_per_scan": n, "levels": len(synth.SCALE_MAP),
            "l2_policy": "working set larger than L2: every launch sequence streams > 1 GB of lattice / feature buffers; "
                         "%d distinct scans resident per GPU, cycled" % args.resident}


# ------------------------------------------------------------------------------------------------
# CPU arm.  kind "reference": the unmodified reference files (baseline/_ref or /root/reference) - GenerateData on
# CPU (torch-CPU + numpy + numba/khash) + 5 BilateralConvFlex(DEVICE="cpu") with E-Net's shapes and wiring
# (reference nets/enet.py:107-141).  kind "port": the C oracle lattice build (1 thread) + torch-CPU BCL.
# ------------------------------------------------------------------------------------------------
class CpuArm(object):
    def __init__(self, weights, want_reference=True, train=False):
        import torch
        from oracle import lattice as ol, ref_harness
        from efgh_b200 import synth
        self.torch, self.synth, self.ol = torch, synth, ol
        try:   # torchrun exports OMP_NUM_THREADS=1; the CPU arm may use every host core it is allowed to
            torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
        except Exception:
            pass
        self.cores = torch.get_num_threads()
        self.weights = weights
        self.train = train
        self.kind = "port"
        self.ref = None
        if want_reference and ref_harness.available():
            try:
                _, g, b = ref_harness.load()
                self.gd = g.GenerateData(3, synth.SCALE_MAP, "cpu")
                self.bcls = []
                for li, (cin, nout) in enumerate(synth.ENET_BCL):
                    m = b.BilateralConvFlex(3, 1, cin, nout, "cpu", True, True, True, True, False, False, chunk_size=-1)
                    (W0, b0), (W1, b1) = weights[li]
                    with torch.no_grad():
                        m.blur_conv[0].weight.copy_(W0); m.blur_conv[0].bias.copy_(b0)
                        m.blur_conv[2].weight.copy_(W1); m.blur_conv[2].bias.copy_(b1)
                    self.bcls.append(m)
                self.kind = "reference"
                self.ref = ref_harness
            except Exception as e:           # numba / cffi missing on this box: fall back to the port
                self.why_port = repr(e)
        ol.build()
        self.hash_variant = "ref" if ol.has_ref() else "port"

    def describe(self):
        if self.kind == "reference":
            return ("unmodified reference nets/generate_data.py + nets/transforms.py (numba + khash cffi) + nets/bilateralNN.py on CPU "
                    "(lattice build is single-threaded by construction, torch ops use %d threads)" % self.cores)
        return ("oracle port: C lattice build (1 thread, %s hash map) + torch-CPU BCL (%d threads)"
                % ("reference khash" if self.hash_variant == "ref" else "ported", self.cores))

    def scan_seconds(self, pc, feat0):
        """One scan: (lattice seconds, BCL seconds)."""
        torch = self.torch
        t0 = time.perf_counter()
        if self.kind == "reference":
            _, data = self.gd(torch.from_numpy(pc))
            t1 = time.perf_counter()
            prev = torch.from_numpy(feat0)[None]
            if self.train:
                prev = prev.clone().requires_grad_(True)
            with torch.enable_grad() if self.train else torch.no_grad():
                for d, m in zip(data, self.bcls):
                    prev = m(torch.cat((d["pc1_el_minus_gr"], prev), 1), d["pc1_barycentric"], d["pc1_lattice_offset"],
                             d["pc1_blur_neighbors"], None, None)
                if self.train:
                    (0.5 * prev.square().mean()).backward()
        else:
            from oracle import bcl as obcl
            data = self.ol.generate(pc, self.synth.SCALE_MAP, self.hash_variant)
            t1 = time.perf_counter()
            prev = torch.from_numpy(feat0)[None]
            ws = self.weights
            if self.train:
                prev = prev.clone().requires_grad_(True)
                ws = [[(W.clone().requires_grad_(True), b.clone().requires_grad_(True)) for W, b in lv] for lv in self.weights]
            with torch.enable_grad() if self.train else torch.no_grad():
                for li, d in enumerate(data):
                    x = torch.cat((torch.from_numpy(d["pc1_el_minus_gr"]), prev), 1)
                    prev = obcl.bcl_forward(x, torch.from_numpy(d["pc1_barycentric"]), torch.from_numpy(d["pc1_lattice_offset"]),
                                            torch.from_numpy(d["pc1_blur_neighbors"]), ws[li], dtype=torch.float32)
                if self.train:
                    (0.5 * prev.square().mean()).backward()
        t2 = time.perf_counter()
        return t1 - t0, t2 - t1


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np
    from efgh_b200 import synth
    from efgh_b200.pipeline import make_enet_weights
    weights = make_enet_weights(synth.ENET_BCL)
    sensor = TRAIN_SENSOR if args.train else args.sensor
    arm = CpuArm(weights, True, train=args.train)
    rng = np.random.default_rng(0)
    pc = synth.synth_scan(0, sensor)
    feat0 = rng.standard_normal((32, pc.shape[1])).astype(np.float32)
    for _ in range(1 if arm.kind == "reference" else max(1, min(args.warmup, 2))):     # (numba JIT / first-touch)
        arm.scan_seconds(pc, feat0)
    t0 = time.perf_counter()
    splits = [arm.scan_seconds(pc, feat0) for _ in range(args.steps)]
    dt = time.perf_counter() - t0
    v = args.steps / dt
    port = None
    if arm.kind == "reference":      # the faster C port beside it, for the record
        parm = CpuArm(weights, False, train=args.train)
        parm.scan_seconds(pc, feat0)
        ps = [parm.scan_seconds(pc, feat0) for _ in range(3)]
        port = 1.0 / float(np.median([a + b for a, b in ps]))
    line = {"impl": "reference", "metric": METRIC_TRAIN if args.train else METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, args.train),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": arm.cores, "kind": arm.kind,
                             "sample": "1 scan (seed 0, %d pts) per step, %s: %s; median lattice build %.3f s + BCL %s %.3f s"
                                       % (pc.shape[1], "fwd+bwd" if args.train else "fwd", arm.describe(),
                                          float(np.median([s[0] for s in splits])), "fwd+bwd" if args.train else "fwd",
                                          float(np.median([s[1] for s in splits]))),
                             "port_value": port},
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


def bind_near_gpu(local):
    """Multi-rank runs: pin this process (and so its pinned host buffers, first-touch) to the CPUs of the GPU's NUMA
    node - 8 ranks allocating their staging buffers on one socket would send half the H2D traffic across the
    inter-socket link.  Best effort; returns a description."""
    try:
        import torch
        p = torch.cuda.get_device_properties(local)
        addr = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % addr).read())
        if node < 0:
            return "numa node unknown"
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return "numa node %d: no allowed cpu" % node
        os.sched_setaffinity(0, cpus)
        return "numa node %d (%d cpus)" % (node, len(cpus))
    except Exception as e:
        return "unbound (%s)" % type(e).__name__


# ------------------------------------------------------------------------------------------------
# Algorithmic bytes per stage (SURVEY.md §8d's per-unit figures x the units one launch processes)
# ------------------------------------------------------------------------------------------------
def stage_model(pipe, counts, n_points):
    """{stage: (bytes, flops)} for one launch sequence; n = points entering the level, H = its vertices."""
    out, n = {}, n_points
    for li, lv in enumerate(pipe.levels):
        H, cin, cmid, cout, F = counts[li], lv["cin"], lv["cmid"], lv["cout"], lv["F"]
        idx = 12 if pipe.emit_int64 else 4
        out["L%d.points" % li] = (12 * n + 32 * n, 0)
        out["L%d.vertices" % li] = (4 * idx * n + F * idx * H + (12 * H if lv["next"] is not None else 0), 0)
        if li == 0 and getattr(pipe, "gs0", False):            # stem (or transpose) writes point-major rows, the splat gathers them
            out["L0.stem"] = (12 * n + 4 * (cin - 4) * n, 2.0 * n * (3 * 32 + 32 * 32 + 32 * 32) if pipe.stem is not None else 0)
            feat_in = 4 * (cin - 4) * n
        else:
            feat_in = (12 * n if li == 0 and pipe.stem is not None else 4 * (cin - 4) * n)
        out["L%d.splat" % li] = (feat_in + 48 * n + 4 * cin * (H + 1), 0)
        wbytes = 4 * F * cin * cmid * (2 if pipe.nsplit == 3 else 1)
        out["L%d.conv1" % li] = (4 * cin * (H + 1) + 4 * F * H + 4 * cmid * H + wbytes, 2.0 * H * F * cin * cmid)
        out["L%d.conv2" % li] = (4 * cmid * H + 4 * cout * H + 4 * cmid * cout * (2 if pipe.nsplit == 3 else 1), 2.0 * H * cmid * cout)
        n = H
    return out


KERNEL_STAGE = (("k_stem_rows", "stem"), ("k_clear", None), ("k_points", "points"), ("k_assign", "points"), ("k_vertices", "vertices"), ("k_zero", "zero"),
                ("k_scatter", "splat"), ("k_normalize", "splat"), ("k_splat", "splat"), ("k_conv_tc", "conv"), ("k_conv", "conv"))


def ncu_sequence_traffic(args, zero_levels=None, timeout=240):
    """DRAM bytes of every kernel of ONE eager launch sequence, measured now with an ncu metrics pass over
    tools/ncu_sequence.py (same pipeline configuration as the bench).  Returns ({stage: bytes}, note)."""
    ncu = None
    for cand in ("ncu", "/usr/local/cuda/bin/ncu"):
        try:
            subprocess.run([cand, "--version"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, check=True)
            ncu = cand
            break
        except Exception:
            continue
    if ncu is None:
        return None, "ncu not found on this box"
    with tempfile.TemporaryDirectory() as td:
        log = os.path.join(td, "seq.csv")
        cmd = [ncu, "--profile-from-start", "off", "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum", "--clock-control", "none",
               "--csv", "--log-file", log, sys.executable, os.path.join(ROOT, "tools", "ncu_sequence.py"), "--scan-batch", str(args.scan_batch),
               "--sensor", args.sensor, "--no-stem"] + (["--int32-only"] if args.int32_only else []) + \
              (["--atomic-splat"] if args.atomic_splat else [])
        env = dict(os.environ)
        for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT"):
            env.pop(k, None)
        try:
            r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=timeout, env=env, text=True)
        except subprocess.TimeoutExpired:
            return None, "ncu pass timed out after %d s" % timeout
        if r.returncode != 0 or not os.path.exists(log):
            return None, "ncu pass failed (rc %d): %s" % (r.returncode, (r.stdout or "")[-200:].replace("\n", " "))
        rows = list(csv.DictReader([l for l in open(log) if not l.startswith("==")]))
    per_launch, order = {}, []
    for row in rows:
        lid = int(row["ID"])
        if lid not in per_launch:
            per_launch[lid] = [row["Kernel Name"], 0.0]
            order.append(lid)
        val = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"].lower()
        mult = {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(unit, 1.0)
        per_launch[lid][1] += val * mult
    stages, seen = {}, {}
    for lid in order:
        name, b = per_launch[lid]
        for pat, st in KERNEL_STAGE:
            if pat in name:
                break
        else:
            continue
        k = seen.get(pat, 0)
        seen[pat] = k + 1
        if st is None:
            st, lvl = "points", k
        elif st == "conv":
            lvl, st = k // 2, "conv%d" % (1 + k % 2)
        elif pat in ("k_scatter", "k_normalize"):
            lvl = 0
        elif pat == "k_zero" and zero_levels:
            lvl = zero_levels[k] if k < len(zero_levels) else k
        elif pat == "k_splat":
            lvl = k + (0 if "k_scatter" not in seen else 1)
        else:
            lvl = k
        key = "L%d.%s" % (lvl, st)
        stages[key] = stages.get(key, 0.0) + b
    return stages, "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum over one eager launch sequence, this run (%d launches)" % len(order)


def tf32_peak_tflops(torch, dev):
    """cuBLAS TF32 GEMM throughput on this GPU, measured now (8192^3, best of 5) - the denominator for the
    convolution's tensor-pipe fraction (the driver's MEASURED_PEAKS.json only holds a bf16 figure)."""
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        a = torch.randn(8192, 8192, device=dev)
        b = torch.randn(8192, 8192, device=dev)
        best = 1e9
        for _ in range(2):
            torch.matmul(a, b)
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            torch.matmul(a, b)
            e1.record()
            torch.cuda.synchronize(dev)
            best = min(best, e0.elapsed_time(e1))
        return 2.0 * 8192 ** 3 / (best * 1e-3) / 1e12
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old


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
    numa = bind_near_gpu(local) if world > 1 else "single rank: not bound"
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    main = torch.cuda.current_stream(dev)

    def timed_region(fn, steps, streams=()):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(main)
        for st in streams:
            st.wait_event(e0)
        for _ in range(steps):
            fn()
        for st in streams:
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

    # ---------------------------------------------------------------------------------------------
    # training leg (BASELINE configs[3]); the headline when --train, an extra key otherwise
    # ---------------------------------------------------------------------------------------------
    def train_leg(steps, warmup):
        from efgh_b200 import sharding, training
        S = args.train_scans
        mine = sharding.scan_indices_for_rank(S * world, rank, world)
        host = [synth.synth_scan(i, TRAIN_SENSOR) for i in mine]
        clouds = [torch.from_numpy(c).to(dev) for c in host]
        impl = args.train_impl
        if impl == "auto":
            impl = "batched" if hasattr(ScanPipeline, "backward") else "module"
        tr = (training.BatchedTrainer if impl == "batched" else training.ModulePathTrainer)(clouds, dev, world)
        for _ in range(max(warmup, 2)):
            loss = tr.step()
        l0 = _capi.lib().efgh_launch_count()
        ms = timed_region(lambda: tr.step(), steps)
        launches = int(_capi.lib().efgh_launch_count() - l0)     # this library's kernels (torch's own stem / optimizer kernels not counted)
        if getattr(tr, "_graph", None) is not None:               # replayed CUDA graph: the captured launches run again every step
            launches = int(tr.launches_per_step * steps)
        # end to end: the step's inputs come from pinned host memory, the loss goes back to the host, every step
        pins = [torch.from_numpy(c).pin_memory() for c in host]
        loss_pin = torch.empty((), dtype=torch.float32).pin_memory()

        def step_e2e():
            for c, p in zip(clouds, pins):
                c.copy_(p, non_blocking=True)
            if impl == "batched":
                tr.pc.copy_(torch.cat(clouds, dim=1))
            loss_pin.copy_(tr.step().float(), non_blocking=True)
        step_e2e()
        e2e_steps = max(2, steps // 2)
        ms_e2e = timed_region(step_e2e, e2e_steps)
        chk = torch.tensor([float(sum(p.detach().double().sum() for p in tr.parameters()))], device=dev, dtype=torch.float64)
        in_sync = True
        if world > 1:
            lo, hi = chk.clone(), chk.clone()
            dist.all_reduce(lo, op=dist.ReduceOp.MIN)
            dist.all_reduce(hi, op=dist.ReduceOp.MAX)
            in_sync = bool((hi - lo).abs() <= 1e-9 * hi.abs().clamp(min=1))
        nparam = sum(p.numel() for p in tr.parameters())
        return {"metric": METRIC_TRAIN, "value": S * world * steps / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / steps, "steps": steps,
                "scans_per_gpu_per_step": S, "points_per_scan": int(host[0].shape[1]), "impl": impl,
                "e2e": {"value": S * world * e2e_steps / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": S * host[0].nbytes, "d2h_bytes_per_step": 4},
                "gpu_launches": launches, "allreduce_calls_per_step": tr.allreduce_calls, "allreduce_bytes_per_step": 4 * nparam if world > 1 else 0,
                "parameters": nparam, "loss": float(loss), "replicas_in_sync": in_sync,
                "optimizer": "Adam", "collective": "nccl all-reduce (mean), %d ranks" % world if world > 1 else "none (1 rank)"}

    if args.train:
        sampler = ClockSampler(local) if rank == 0 else None
        tl = train_leg(args.steps, args.warmup)
        clocks = sampler.stop() if sampler else None
        if rank == 0:
            line = {"metric": METRIC_TRAIN, "value": tl["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 2),
                    "ms_per_step": tl["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                    "data": "synthetic", "config": workload_config(args, True), "clocks": clocks, "e2e": tl["e2e"],
                    "gpu_launches": tl["gpu_launches"], "train": tl, "numa": numa}
            print(json.dumps(line))
        if world > 1:
            dist.destroy_process_group()
        return

    # ---------------------------------------------------------------------------------------------
    # forward leg (BASELINE configs[1])
    # ---------------------------------------------------------------------------------------------
    B = args.batch
    G = max(1, min(args.scan_batch, B))                 # scans per launch sequence
    R = max(G, min(args.resident, B))                   # distinct resident scans
    if B % G or R % G:
        raise SystemExit("bench.py: --batch and --resident must be multiples of --scan-batch")
    NG = B // G                                         # launch sequences per step
    NR = R // G                                         # resident groups (inputs of one launch sequence)
    P = max(1, min(args.streams, NR))                   # compute streams
    NP = max(P, min(NR, 2))                             # pipelines (buffer sets): >= 2 so that copies of one batch overlap kernels of another
    use_stem = not args.no_stem
    weights = make_enet_weights(synth.ENET_BCL)
    # this rank's scans: global scan index = rank + world * j  (scan-index sharding)
    seeds = [rank + world * j for j in range(R)]
    clouds = [synth.synth_scan(sd, args.sensor) for sd in seeds]
    N = clouds[0].shape[1]
    gs_ = torch.Generator().manual_seed(5)   # random-init conv_in (reference nets/enet.py:24-28): 3 -> 32 -> 32 -> 32, LeakyReLU(0.1)
    stem_layers = [(torch.randn(co, ci, 1, generator=gs_) * 0.3, torch.randn(co, generator=gs_) * 0.1) for ci, co in ((3, 32), (32, 32), (32, 32))]

    def make_feats():
        rng = np.random.default_rng(1000 + rank)
        return [rng.standard_normal((32, N)).astype(np.float32) for _ in range(R)]

    def make_pipes(stem_on, count):
        return [ScanPipeline(N, synth.SCALE_MAP, synth.ENET_BCL, weights, dev, vertex_cap_factor=1.0, batch=G,
                             gather_splat=not args.atomic_splat, stem=(stem_layers, True) if stem_on else None,
                             emit_int64=not args.int32_only) for _ in range(count)]

    # `value` leg = the metric's path as BASELINE.json states it: lattice build + 5 BCL forward, its inputs (cloud and the
    # (32, N) stem features that reference nets/enet.py:111 feeds bcn1) resident in HBM.  The e2e leg goes through the
    # public API from what a caller has - the cloud (reference nets/enet.py:103-111) - with the stem fused into the
    # level-0 splat (`--no-stem`: the features come from the host as well).
    feats = make_feats()
    pc_dev = [torch.from_numpy(np.concatenate(clouds[g * G:(g + 1) * G], axis=1)).to(dev) for g in range(NR)]
    ft_dev = [torch.from_numpy(np.concatenate(feats[g * G:(g + 1) * G], axis=1)).to(dev) for g in range(NR)]
    pipes = make_pipes(False, NP)
    streams = [torch.cuda.Stream(dev) for _ in range(P)]
    copy_streams = [torch.cuda.Stream(dev) for _ in range(NP)]
    use_graph = not args.no_graph
    graphs = None
    if use_graph:   # one CUDA graph per resident group (captures the whole 5-level launch sequence on that group's stream)
        graphs = [pipes[j % NP].graph_for(pc_dev[j], ft_dev[j], streams[j % P]) for j in range(NR)]

    def step(timers=None):
        for i in range(NG):
            j = i % NR
            if use_graph and timers is None:
                with torch.cuda.stream(streams[j % P]):
                    graphs[j].replay()
            else:
                pipes[j % NP].enqueue(pc_dev[j], ft_dev[j], stream=streams[j % P], timers=timers)

    # ---- warm-up (also validates capacities: raises if any level overflowed) and per-group result sizes
    group_counts = []
    for j in range(NR):
        l0 = _capi.lib().efgh_launch_count()
        pipes[j % NP].enqueue(pc_dev[j], ft_dev[j], stream=streams[j % P])
        seq_launches = _capi.lib().efgh_launch_count() - l0     # this library's kernels per launch sequence (counted by the C ABI)
        torch.cuda.synchronize(dev)
        group_counts.append(pipes[j % NP].counts())
    vs = pipes[(NR - 1) % NP].vertex_starts()
    counts = group_counts[0]
    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize(dev)
    for p in pipes:
        p.counts()

    # ---- timed region, inputs resident in HBM
    sampler = ClockSampler(local) if rank == 0 else None
    ms = timed_region(step, args.steps, streams + copy_streams)
    clocks = sampler.stop() if sampler else None
    scans = B * world * args.steps
    value = scans / (ms * 1e-3)

    # ---- end to end: pinned host clouds -> H2D -> scan -> D2H of the WHOLE last-level output + level records
    c_last = synth.ENET_BCL[-1][1][-1]
    nlev = len(synth.SCALE_MAP)

    def e2e_leg(pipes_e, stem_on, feats_e, steps):
        pc_pin = [torch.from_numpy(c).pin_memory() for c in clouds]
        ft_pin = [torch.from_numpy(f).pin_memory() for f in feats_e] if not stem_on else [None] * R
        out_pin = [torch.empty((group_counts[j][-1], c_last), dtype=torch.float32).pin_memory() for j in range(NR)]
        st_pin = [torch.empty((nlev, 24), dtype=torch.int32).pin_memory() for _ in range(NR)]
        vs_pin = [torch.empty((nlev, G + 1), dtype=torch.int32).pin_memory() for _ in range(NR)]

        def step_e2e():
            for i in range(NG):
                j = i % NR
                # copies on the pipeline's own stream, kernels on a shared compute stream: H2D of batch j+1 overlaps batch j
                if G == 1:
                    pipes_e[j % NP].forward_host(pc_pin[j], ft_pin[j], out_pin[j], st_pin[j], stream=copy_streams[j % NP],
                                                 use_graph=use_graph, compute_stream=streams[j % P])
                else:
                    pipes_e[j % NP].forward_host(pc_pin[j * G:(j + 1) * G], ft_pin[j * G:(j + 1) * G], out_pin[j], st_pin[j],
                                                 stream=copy_streams[j % NP], use_graph=use_graph, starts_host=vs_pin[j],
                                                 compute_stream=streams[j % P])
        for _ in range(2):
            step_e2e()
        ms_e = timed_region(step_e2e, steps, streams + copy_streams)
        for j in range(NR):                              # the records and the full result really came back
            assert [int(v) for v in st_pin[j][:, 1]] == group_counts[j], "e2e: level records differ from the resident run"
            assert bool(torch.isfinite(out_pin[j][-1]).all()) and float(out_pin[j].abs().sum()) > 0
        h2d = B * (clouds[0].nbytes + (0 if stem_on else feats_e[0].nbytes))
        d2h = sum((out_pin[i % NR].numel() + st_pin[0].numel() + (vs_pin[0].numel() if G > 1 else 0)) * 4 for i in range(NG))
        return {"value": B * world * steps / (ms_e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h}

    e2e_steps = max(2, args.steps // 2)
    pipes_stem = make_pipes(True, NP) if use_stem else None
    e2e = e2e_leg(pipes_stem if use_stem else pipes, use_stem, feats, e2e_steps)
    value_stem = None
    if use_stem:      # the stem-fused pipeline with the clouds resident: what the extra stem work costs on the device
        gr = [pipes_stem[j % NP].graph_for(pc_dev[j], None, streams[j % P]) for j in range(NR)] if use_graph else None

        def step_stem():
            for i in range(NG):
                j = i % NR
                if use_graph:
                    with torch.cuda.stream(streams[j % P]):
                        gr[j].replay()
                else:
                    pipes_stem[j % NP].enqueue(pc_dev[j], None, stream=streams[j % P])
        step_stem()
        vsteps = max(2, args.steps // 4)
        value_stem = B * world * vsteps / (timed_region(step_stem, vsteps, streams + copy_streams) * 1e-3)

    # ---- per-stage table: eager launches with CUDA events on the launching stream (events cannot be read back from
    #      inside a replayed graph), three passes over resident group 0, median
    stage_t = {"*": []}
    for _ in range(3):
        pipes[0].enqueue(pc_dev[0], ft_dev[0], stream=streams[0], timers=stage_t)
    torch.cuda.synchronize(dev)
    stages = {k: float(np.median([a.elapsed_time(b) for a, b in v])) * 1e3 for k, v in stage_t.items() if k != "*"}
    if args.stages and rank == 0:
        for k in sorted(stages, key=lambda k: (k.split(".")[0], -stages[k])):
            print("  %-14s %9.1f us" % (k, stages[k]), file=sys.stderr)
        print("  sum            %9.1f us" % sum(stages.values()), file=sys.stderr)

    # ---- single-scan latency (one scan per launch sequence, CUDA graph)
    pipe1 = pipes[0] if G == 1 else ScanPipeline(N, synth.SCALE_MAP, synth.ENET_BCL, weights, dev, vertex_cap_factor=1.0,
                                                 gather_splat=not args.atomic_splat)
    lat = []
    pc1_dev = pc_dev[0][:, :N].contiguous()
    ft1_dev = ft_dev[0][:, :N].contiguous()
    g1 = pipe1.graph_for(pc1_dev, ft1_dev, streams[0]) if use_graph else None
    for _ in range(7):
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
    del pipe1, g1

    # ---- extras (never inside the timed regions above): feature-input e2e variant, module path, training leg
    extras = {}
    if not args.no_extras:
        try:   # the other input variant: the (32, N) stem features come from the host as well (18.4 MB per scan)
            alt = pipes if use_stem else make_pipes(True, NP)
            alt_steps = max(2, e2e_steps // 4)
            extras["e2e_variants"] = {
                "stem_fused_cloud_only" if use_stem else "stem_features_from_host": e2e,
                "stem_features_from_host" if use_stem else "stem_fused_cloud_only": e2e_leg(alt, not use_stem, feats, alt_steps)}
        except Exception as e:
            extras["e2e_variants"] = {"error": repr(e)}
        try:
            extras["module_path"] = module_path_leg(torch, dev, synth, clouds[:4], weights)
        except Exception as e:
            extras["module_path"] = {"error": repr(e)}
        try:
            extras["aux_stages"] = aux_leg(torch, dev, synth, clouds[0], measured_peaks())
        except Exception as e:
            extras["aux_stages"] = {"error": repr(e)}
        torch.cuda.empty_cache()
        try:
            extras["train"] = train_leg(5, 2)
        except Exception as e:
            extras["train"] = {"error": repr(e)}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- rooflines
    peaks = measured_peaks()
    model = stage_model(pipes[0], counts, N * G)
    tf32_peak = tf32_peak_tflops(torch, dev)
    traffic, traffic_note = (None, "skipped (--no-extras)") if args.no_extras else \
        ((None, "not measured at N > 1 (ncu never wraps a multi-rank command)") if world > 1 else
         ncu_sequence_traffic(args, [li for li, lv in enumerate(pipes[0].levels) if (not lv["gs"]) or (lv["tc"] and lv.get("split0"))]))
    mma_mult = 3.0 if pipes[0].nsplit == 3 else 1.0
    levels = []
    for li in range(nlev):
        for cv in ("conv1", "conv2"):
            k = "L%d.%s" % (li, cv)
            if k not in stages or stages[k] <= 0:
                continue
            by, fl = model[k]
            t = stages[k] * 1e-6
            hbm_frac = by / t / 1e9 / peaks["hbm_gbs"]
            tensor_frac = fl * mma_mult / t / 1e12 / tf32_peak
            bound = "tensor" if tensor_frac >= max(hbm_frac, 0.4) else ("hbm" if hbm_frac >= 0.4 else "issue")
            levels.append({"stage": k, "us": stages[k], "algorithmic_bytes": by, "hbm_gbs": by / t / 1e9, "hbm_frac": hbm_frac,
                           "tflops_fp32_equiv": fl / t / 1e12, "tensor_tflops_issued": fl * mma_mult / t / 1e12, "tf32_frac": tensor_frac,
                           "bound": bound, "traffic": traffic.get(k) if traffic else None})
    dom = max((k for k in stages if k in model), key=lambda k: stages[k])
    dom_bytes, dom_flops = model[dom]
    dom_ms = stages[dom] * 1e-3
    total_bytes, _ = pipes[0].algorithmic_bytes(counts)
    total_bytes /= G                                    # per scan
    scan_ms = ms / (B * args.steps)
    # which ceiling bounds the dominant kernel: the lower one of its two roofline times - algorithmic bytes at the
    # measured copy bandwidth, or the tensor-pipe work it issues (3 MMAs per product with 3xTF32) at the cuBLAS TF32
    # rate measured in this run
    t_hbm = dom_bytes / (peaks["hbm_gbs"] * 1e9)
    t_tensor = dom_flops * mma_mult / (tf32_peak * 1e12) if dom_flops else 0.0
    if t_tensor > t_hbm:
        ach = dom_flops * mma_mult / (dom_ms * 1e-3) / 1e12
        roof = {"kernel": "%s (%d scans per launch)" % (dom, G), "bound": "tensor", "achieved": ach, "peak": tf32_peak, "unit": "TFLOP/s",
                "frac": ach / tf32_peak,
                "peak_source": "cuBLAS TF32 GEMM (torch.matmul fp32, allow_tf32, 8192^3, best of 5) measured in this run - MEASURED_PEAKS.json "
                               "holds only a bf16 figure (%.0f TFLOP/s), TF32 runs at half of that nominally" % peaks["bf16_tflops"],
                "hbm_frac": dom_bytes / (dom_ms * 1e-3) / 1e9 / peaks["hbm_gbs"]}
    else:
        ach = dom_bytes / (dom_ms * 1e-3) / 1e9
        roof = {"kernel": "%s (%d scans per launch)" % (dom, G), "bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": ach / peaks["hbm_gbs"],
                "peak_source": peaks["source"] + " copy bandwidth (MEASURED_PEAKS.json)" if peaks["source"] == "measured" else "fallback"}
    roof.update({"traffic": traffic.get(dom) if traffic else None, "traffic_source": traffic_note,
                 "kernel_ms": dom_ms, "algorithmic_bytes": dom_bytes,
                 "tensor_tflops_fp32_equiv": dom_flops / (dom_ms * 1e-3) / 1e12 if dom_flops else None,
                 "share_of_scan": stages[dom] / max(sum(stages.values()), 1e-9),
                 "limiter": next((l["bound"] for l in levels if l["stage"] == dom), None),
                 "note": "timed with CUDA events on the launching stream over three eager passes (median); `limiter` is what the kernel "
                         "is short of in practice (issue = instruction issue / gather rate, profiles/r2_conv_tc_levels.json); "
                         "roofline_levels lists every convolution"})
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": workload_config(args, False), "clocks": clocks, "e2e": e2e,
        "gpu_launches": int(seq_launches * NG * args.steps),
        "roofline": roof,
        "details": {"scans_per_gpu_per_step": B, "scans_per_launch_sequence": G, "resident_scans": R, "compute_streams": P, "pipelines": NP,
                    "levels_H_scan0": [v[1] - v[0] for v in vs], "levels_H_group0": counts,
                    "splat": "levels 1-4 gather through vertex -> contributions lists; level 0: %s" % pipes[0].level0_splat
                             if pipes[0].gather_splat else "atomic scatter",
                    "lattice_index_dtype": "int64 (reference format) + int32 copies for the BCL kernels" if pipes[0].emit_int64 else "int32 only",
                    "conv_precision": pipes[0].precision, "cuda_graphs": use_graph, "single_scan_latency_ms": float(np.median(lat)),
                    "algorithmic_MB_per_scan": total_bytes / 1e6, "scan_roofline_frac": total_bytes / (scan_ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
                    "e2e_over_value": e2e["value"] / value, "value_stem_fused_resident_cloud": value_stem, "numa": numa, "tf32_peak_tflops_measured": tf32_peak,
                    "timed_region_s": ms * 1e-3},
        "roofline_levels": levels,
        "stages_us": stages,          # per launch sequence (G scans), eager single-stream pass
        "stages_traffic": traffic,    # DRAM bytes per launch sequence and stage (ncu, this run)
    }
    line.update(extras)
    if world == 1 and not args.no_cpu_baseline:
        try:
            arm = CpuArm(weights, True)
            f0 = np.random.default_rng(0).standard_normal((32, N)).astype(np.float32)
            arm.scan_seconds(clouds[0], f0)                       # numba JIT / first touch
            reps = 3 if arm.kind == "reference" else 5
            sp = [arm.scan_seconds(clouds[0], f0) for _ in range(reps)]
            sec = float(np.median([a + b for a, b in sp]))
            line["cpu_baseline"] = {"value": 1.0 / sec, "unit": UNIT, "cores": arm.cores, "kind": arm.kind,
                                    "sample": "1 scan (seed 0), median of %d: %s; lattice build %.3f s + BCL fwd %.3f s"
                                              % (reps, arm.describe(), float(np.median([s[0] for s in sp])), float(np.median([s[1] for s in sp])))}
        except Exception as e:  # the CPU arm is optional infrastructure; the product numbers stand without it
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": "unavailable: %r" % (e,)}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def aux_leg(torch, dev, synth, cloud, peaks):
    """SURVEY.md §8 rows f3 / f4 beside the lattice path: the point -> image projections (reference
    common/torch_utils.py:11-103) and the cloud pre-processing (reference data_loader/loader_utils.py:163-202) on one
    131k-point scan, CUDA-event timed (median of 20, after 3 warm-ups), with their algorithmic bytes (points read once,
    image written once) against the measured copy bandwidth."""
    import numpy as np
    from efgh_b200.projections import range_img_from_cartesian_pc_torch, depth_img_from_cartesian_pc_torch
    from efgh_b200.preproc import preproc_pcd
    pc = torch.from_numpy(cloud)[None].to(dev)
    n = pc.shape[-1]
    K = np.array([[2813.6, 0, 969.3], [0, 2808.3, 624.0], [0, 0, 1.0]])
    R0 = np.array([[0, -1.0, 0], [0, 0, -1.0], [1.0, 0, 0]])
    T = torch.from_numpy((K @ np.concatenate([R0, np.array([[0.03], [-0.1], [-0.12]])], 1)).astype(np.float32))[None].to(dev)
    scan = torch.cat((pc[0].t() * 1.3, torch.rand(n, 1, device=dev)), 1).contiguous()
    gts = {"rand_init_l": np.eye(4)}
    sample = np.random.default_rng(0).permutation(100000)[:65536]

    def timed(fn):
        for _ in range(3):
            fn()
        ts = []
        for _ in range(20):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize(dev)
            ts.append(a.elapsed_time(b))
        return float(np.median(ts)) * 1e3

    out = {}
    for name, fn, by in (
            ("range_image_600x3840", lambda: range_img_from_cartesian_pc_torch(pc, (600, 3840), (0.125, -0.125), "cuda"), 12 * n + 16 * 600 * 3840),
            ("depth_image_1200x1920", lambda: depth_img_from_cartesian_pc_torch(pc, T, (1200, 1920), "cuda"), 12 * n + 16 * 1200 * 1920),
            ("preproc_131k_to_65536", lambda: preproc_pcd(scan, gts, 65536, sample=sample), 16 * n + 8 * 65536 + 32 * 65536)):
        us = timed(fn)
        out[name] = {"us": us, "algorithmic_bytes": by, "hbm_gbs": by / us / 1e3, "hbm_frac": by / us / 1e3 / peaks["hbm_gbs"]}
    out["note"] = "host-timed wrappers (allocation + 3-4 launches each); latency-bound at one scan per call"
    return out


def module_path_leg(torch, dev, synth, clouds, weights):
    """The reference's operator API on this repo's kernels: GenerateData(pc) + 5 x BilateralConvFlex.forward wired as
    reference nets/enet.py:107-141, one scan per call, forward only.  exact=True returns exact-size tensors (one small
    D2H read per level, like the reference's Python ints); exact=False reads the level records once per scan."""
    from efgh_b200.generate_data import GenerateData
    from efgh_b200.bilateralNN import BilateralConvFlex
    pcs = [torch.from_numpy(c).to(dev) for c in clouds]
    feat = torch.randn(1, 32, pcs[0].shape[1], device=dev)
    out = {"unit": "scans/s", "what": "GenerateData + 5 BilateralConvFlex forward per 131k-point scan, one scan per call (host-timed, synchronised)"}
    for exact in (False, True):
        gd = GenerateData(3, synth.SCALE_MAP, "cuda", exact=exact)
        bcls = []
        for li, (cin, nout) in enumerate(synth.ENET_BCL):
            m = BilateralConvFlex(3, 1, cin, nout, "cuda", True, True, True, True, False, False, chunk_size=-1).to(dev)
            with torch.no_grad():
                for c, (W, b) in zip((m.blur_conv[0], m.blur_conv[2]), weights[li]):
                    c.weight.copy_(W.to(dev)); c.bias.copy_(b.to(dev))
            bcls.append(m)

        def fwd(pc):
            with torch.no_grad():
                _, data = gd(pc)
                x = feat
                for d, m in zip(data, bcls):
                    x = m(torch.cat((d["pc1_el_minus_gr"], x), 1), d["pc1_barycentric"], d["pc1_lattice_offset"], d["pc1_blur_neighbors"], None, None)
                return x
        for _ in range(3):
            fwd(pcs[0])
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        n = 24
        for i in range(n):
            fwd(pcs[i % len(pcs)])
        torch.cuda.synchronize(dev)
        out["exact_true" if exact else "exact_false"] = n / (time.perf_counter() - t0)
    return out


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
