#!/usr/bin/env python
"""Benchmark of the Diff-DOPE hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--scaling weak|strong]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[1]): simple_scene, HOPE AlphabetSoup, 64 pose hypotheses per GPU,
640x640 loss window on the 1920x1080 frame, rgb + depth + mask losses (weights 0.7/1/1), SGD with
the reference schedule. One *step* = one optimisation iteration of all hypotheses of a rank:
forward render + losses + backward to the 7 pose parameters + SGD step
(reference: one pass of the loop body at diffdope/diffdope.py:1656-1714).
`--scaling strong` runs BASELINE configs[3] instead (T-LESS stand-in, 256 hypotheses split over the N GPUs).

Prints ONE JSON line (rank 0). `value` = hypothesis-iterations / s with everything resident in
HBM, timed per iteration with CUDA events, L2 flushed between iterations. `e2e` = the same metric
through the public API (`DiffDope.run_optimization`) with the target images and poses coming from
pinned host memory and the result tables read back, inside the timed region. `configs` carries bounded
measurements of the other BASELINE configurations (1, 3, 4, 5) so that they are visible in the same line.
`--impl reference` times the CPU restatement of the reference path (oracle/, the reference itself
has no CPU path and its GPU path needs nvdiffrast + OpenGL).
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "diff-dope_b200"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # NCCL's version banner must not land on stdout next to the JSON line

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "pose-hypotheses x iterations / s"
UNIT = "hyp*iter/s"
B_PER_GPU = 64
WINDOW = 640
LOSSES = dict(use_rgb=True, use_depth=True, use_mask=True, w_rgb=0.7, w_depth=1.0, w_mask=1.0)
HYPER = dict(base_lr=20.0, lr_decay=0.1)
STRONG_B = 256  # BASELINE configs[3]: 256 hypotheses sharded over 2/4/8 GPUs
COUNTERS = os.path.join(ROOT, "profiles", "r02_counters.json")  # ncu counters of the shipped kernels (scripts/ncu_counters.py)


def lr_schedule(n):
    nb = max(n - 1, 1)
    return [HYPER["base_lr"] * HYPER["lr_decay"] ** (it / nb + 1) for it in range(n)]


def workload_config(n_gpus, scaling="weak"):
    if scaling == "strong":
        return {
            "workload": "T-LESS stand-in (procedural 10k-triangle textureless sphere, V=5002, rendered 720x540 depth + mask targets), "
                        "%d hypotheses in total split over the GPUs, l1 depth+mask, SGD (BASELINE.json configs[3])" % STRONG_B,
            "global_hypotheses": STRONG_B,
            "hypotheses_per_gpu": (STRONG_B + n_gpus - 1) // n_gpus,
            "frame": [540, 720],
            "parallelism": "hypotheses sharded over %d GPU(s), one all-gather of the result table" % n_gpus,
            "l2": "flushed between timed iterations (256 MiB write outside the timed events)",
            "timed_step": "one ddope_optimize call of one iteration for all hypotheses of the rank",
        }
    return {
        "workload": "simple_scene HOPE AlphabetSoup (V=8240, T=13860, 2048^2 texture), %d hypotheses/GPU, "
        "%dx%d loss window of the 1920x1080 frame, l1 rgb+depth+mask, SGD (BASELINE.json configs[1])" % (B_PER_GPU, WINDOW, WINDOW),
        "hypotheses_per_gpu": B_PER_GPU,
        "global_hypotheses": B_PER_GPU * n_gpus,
        "window": [WINDOW, WINDOW],
        "frame": [1080, 1920],
        "parallelism": "hypotheses sharded over %d GPU(s), one all-gather of the result table" % n_gpus,
        "l2": "flushed between timed iterations (256 MiB write outside the timed events)",
        "timed_step": "one ddope_optimize call of one iteration (pose, raster, pixel pass, step) for all hypotheses of the rank",
        "scheduling": "3 kernels per iteration with programmatic dependent launch. A call of several iterations (value_l2_warm_single_call, e2e) "
                      "splits the hypotheses into parts on internal streams (raster of one part overlaps the pixel pass of another), joined into "
                      "the caller's stream; a call of one iteration (the timed step) has nothing to pipeline and runs as one part: 4 launches",
    }


# ----------------------------------------------------------------------------------------------
# survey byte model, split by the kernel that moves each term


def survey_bytes_per_hit(V=8240, T=13860, P=WINDOW * WINDOW, c=0.06, rgb=True, depth=True, textured=True):
    """B_alg = P*GT + 2*P*16 + V*VERT + T*12 + TEX. GT = 12 (rgb or edge) + 4 (depth) + 4 (seg); VERT = 12 (+8 uv if a
    textured colour is needed, +12 if a vertex colour is needed); TEX = 2*48*c*P when a textured colour is needed."""
    gt = P * (4.0 + (12.0 if rgb else 0.0) + (4.0 if depth else 0.0))
    rast_w = P * 16.0
    rast_r = P * 16.0
    vert = 12.0 + ((8.0 if textured else 12.0) if rgb else 0.0)
    mesh = V * vert + T * 12.0
    tex = 2 * 48.0 * c * P if (rgb and textured) else 0.0
    return {"total": gt + rast_w + rast_r + mesh + tex, "pixel_kernel": gt + rast_r + tex, "raster_kernel": rast_w + mesh}


def source_sha():
    """sha256 over the kernel sources: a committed ncu capture is only quoted for the sources it was taken from."""
    h = hashlib.sha256()
    d = os.path.join(ROOT, "diff-dope_b200", "csrc")
    for name in sorted(os.listdir(d)):
        if name.endswith((".cu", ".cuh", ".h")) or name == "Makefile":
            h.update(name.encode())
            h.update(open(os.path.join(d, name), "rb").read())
    return h.hexdigest()[:16]


def load_counters():
    """profiles/r02_counters.json: per-kernel ncu counters (warp instructions, DRAM bytes per launch of 64 hypotheses) of the
    bench workload, written by scripts/ncu_counters.py on the GPU box; refused when the kernel sources changed since."""
    try:
        c = json.load(open(COUNTERS))
    except Exception:
        return None, "no capture committed"
    if c.get("source_sha") != source_sha():
        return None, "capture is of other kernel sources (%s, now %s): not quoted" % (c.get("source_sha"), source_sha())
    return c, "profiles/r02_counters.json (ncu --set full, same kernel sources)"


# ----------------------------------------------------------------------------------------------
# clocks


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.path = tempfile.mktemp(suffix=".csv")
        self.proc = None
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def count(self):
        """samples written so far (nvidia-smi takes a second or more to start: the caller keeps the load up until enough exist)"""
        try:
            self.f.flush()
            with open(self.path) as f:
                return sum(1 for line in f if line.count(",") >= 8)
        except Exception:
            return 0

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                p = [x.strip() for x in line.split(",")]
                if len(p) < 9:
                    continue
                try:
                    sm.append(float(p[1]))
                    mx.append(float(p[2]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), samples=len(sm))
        out["reasons"] = sorted(reasons)
        return out


# ----------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle restatement on host cores


class CpuReference:
    """The CPU restatement of the reference path (oracle/refpath.py) on the bench workload."""

    def __init__(self, sample_hyp, threads):
        import scene_util as su
        from oracle import refpath

        torch.set_num_threads(threads)
        self.refpath, self.su = refpath, su
        arr = su.example_mesh_arrays()
        q, t = su.example_pose()
        self.gt = {k: torch.from_numpy(v) for k, v in su.example_targets(1.0).items()}
        self.H, self.W = self.gt["rgb"].shape[:2]
        self.window = su.centred_window(self.gt["segmentation"].numpy(), WINDOW, self.H, self.W)
        self.mesh = refpath.Mesh(arr["pos"], arr["tri"], arr["uv"], arr["tex"])
        self.lr = su.lr_multipliers(sample_hyp)
        self.cfg = dict(l1_rgb_with_mask=True, weight_rgb=0.7, l1_depth_with_mask=True, weight_depth=1.0, l1_mask=True, weight_mask=1.0)
        self.q0, self.t0 = np.tile(q, (sample_hyp, 1)), np.tile(t, (sample_hyp, 1))
        self.n = sample_hyp
        self.proj = su.projection()

    def iterate(self, iters):
        """`iters` iterations (forward, losses, backward, SGD step); returns seconds."""
        hyper = dict(nb_iterations=max(iters - 1, 1), base_lr=HYPER["base_lr"], lr_decay=HYPER["lr_decay"], learning_rate_base=1)
        t0 = time.perf_counter()
        if iters == 1:
            self.refpath.forward_backward(self.mesh, self.proj, self.q0, self.t0, self.gt, self.lr, self.cfg, self.H, self.W, window=self.window)
        else:
            self.refpath.run_optimization(self.mesh, self.proj, self.q0, self.t0, self.gt, self.lr, self.cfg, hyper, self.H, self.W, window=self.window)
        return time.perf_counter() - t0


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    sample_hyp = 2
    ref = CpuReference(sample_hyp, threads)
    for _ in range(min(args.warmup, 1)):
        ref.iterate(1)
    steps = max(1, min(args.steps, 25))  # bounded (~1.5 s of CPU work per step, <= 40 s): each step is one iteration of a 2-hypothesis sample
    dt = sum(ref.iterate(1) for _ in range(steps))
    value = sample_hyp * steps / dt
    sample = ("%d hypotheses x 1 iteration per step of the bench workload (full 1920x1080 frame rendered as the reference "
              "does, loss over the %dx%d window), %d steps, %d torch threads") % (sample_hyp, WINDOW, WINDOW, steps, threads)
    cfg = workload_config(args.gpus)
    # this arm times a bounded SAMPLE of that workload on the host cores: say so next to the workload text
    cfg["reference_arm_sample"] = {"hypotheses_per_step": sample_hyp, "steps_timed": steps, "steps_requested": args.steps,
                                   "note": "per hypothesis-iteration figure; the 64-hypothesis batch is not run on the CPU"}
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "reference example scene (data/example), synthetic learning-rate multipliers (random.seed(0))",
        "config": cfg,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference semantics restated on CPU (oracle/); the reference's own GPU path needs nvdiffrast + OpenGL, absent here",
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------
# workloads on the device


def _loss_cfg(nat, L):
    return nat.make_loss_cfg(L.get("l1_rgb_with_mask", False), L.get("l1_depth_with_mask", False), L.get("l1_mask", False), L.get("weight_rgb", 1),
                             L.get("weight_depth", 1), L.get("weight_mask", 1), L.get("l1_edge", False), L.get("weight_edge", 1))


def standin_scene(nat, w, dev):
    """NativeScene of a tests/workloads.py stand-in with targets rendered by the CUDA renderer at the ground-truth pose."""
    if w.get("tex") is not None:
        sc = nat.NativeScene(w["pos"], w["tri"], uv=w["uv"], tex=w["tex"])
    else:
        sc = nat.NativeScene(w["pos"], w["tri"], vtx_color=w["vtx_color"])
    sc.set_camera(w["P"], w["H"], w["W"])
    out = sc.render(torch.from_numpy(w["q_gt"][None]).to(dev), torch.from_numpy(w["t_gt"][None]).to(dev), want=("rgb", "depth", "rast"))
    cov = (out["rast"][0, ..., 3] > 0).float()
    tgt = dict(rgb=out["rgb"][0].contiguous(), depth=(out["depth"][0] * cov).contiguous(), seg=cov.contiguous())
    sc.set_target(tgt["rgb"], tgt["depth"], tgt["seg"])
    return sc, tgt, float(cov.mean())


def time_call(fn, reps=3):
    """Median CUDA-event milliseconds of `fn()` over `reps` runs after one warm-up run."""
    fn()
    torch.cuda.synchronize()
    ms = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    return float(np.median(ms))


def config_records(nat, dev, rank, world, peak, dist):
    """Bounded measurements of BASELINE configs 1, 3, 4, 5: one ddope_optimize call of the config's iteration
    count per run, median of 3, L2 warm (the natural mode). Configs 3-5 are the labelled STAND-INS of tests/workloads.py
    (BOP assets are not in the tree). Config 4 is the strong-scaling case: 256 hypotheses split over the ranks; config 5 is
    run at its 8-GPU shard size (128 hypotheses per GPU) whatever N is."""
    import scene_util as su
    import workloads as wl

    recs = []

    def sgd(sc, w_q0, w_t0, B, lr, cfg, iters, b_global=None):
        sched = lr_schedule(iters)
        q0 = torch.from_numpy(np.tile(w_q0, (B, 1))).to(dev).contiguous()  # start poses resident on the device: the timed call
        t0 = torch.from_numpy(np.tile(w_t0, (B, 1))).to(dev).contiguous()  # begins with device-side copies, not pageable uploads

        def run():
            q, t = q0.clone(), t0.clone()
            sc.optimize(q, t, lr, sched, cfg, b_global=b_global, keep_history=False)

        return run

    def maxreduce(ms):
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return ms

    # ---- config 1: one hypothesis, 320x320 window, 50 iterations, default losses (mask only) and the full reference stack
    arr = su.example_mesh_arrays()
    q, t = su.example_pose()
    gt = su.example_targets(1.0)
    H, W = gt["rgb"].shape[:2]
    sc = nat.NativeScene(arr["pos"], arr["tri"], arr["uv"], arr["tex"])
    sc.set_camera(su.projection_native(), H, W)
    sc.set_window(*su.centred_window(gt["segmentation"], 320, H, W))
    g = {k: torch.from_numpy(v).to(dev) for k, v in gt.items()}
    sc.set_target(g["rgb"], g["depth"], g["segmentation"][..., 0].contiguous())
    lr1 = torch.from_numpy(np.array([0.01], dtype=np.float32)).to(dev)
    for name, L, kw in (("mask only (default losses)", dict(l1_mask=True, weight_mask=1.0), dict(rgb=False, depth=False)),
                        ("rgb+depth+mask", dict(l1_rgb_with_mask=True, weight_rgb=0.7, l1_depth_with_mask=True, weight_depth=1.0, l1_mask=True, weight_mask=1.0), dict())):
        ms = time_call(sgd(sc, q, t, 1, lr1, _loss_cfg(nat, L), 50))
        model = survey_bytes_per_hit(P=320 * 320, c=0.24, **kw)["total"]
        recs.append({"config": 1, "workload": "simple_scene, 1 hypothesis, 50 iterations, 320x320 window, " + name, "hypotheses_per_gpu": 1, "iters": 50,
                     "ms_per_iter": ms / 50, "value": 50 / (ms * 1e-3), "unit": UNIT, "survey_bytes_per_hit": model,
                     "hbm_frac": model * 50 / (ms * 1e-3) / 1e9 / peak, "bound": "launch latency (one hypothesis): wall time is the figure, not the fraction"})
    del sc

    # ---- config 3: 8 objects x 128 hypotheses x 100 iterations, full stack incl. Sobel edge (stand-in)
    c3 = wl.config3()
    scenes = []
    for o in c3["objects"]:
        w = dict(o, P=c3["P"], H=c3["H"], W=c3["W"])
        s3, tg, cov = standin_scene(nat, w, dev)
        scenes.append((s3, tg, w))
    cfg3 = _loss_cfg(nat, c3["losses"])
    lr3 = torch.from_numpy(su.lr_multipliers(c3["B"], 0.01, 2.0)).to(dev)
    runs = [sgd(s3, w["q0"], w["t0"], c3["B"], lr3, cfg3, c3["iters"]) for s3, _, w in scenes]

    def all_objects():
        for r in runs:
            r()

    ms = time_call(all_objects)
    hits = len(scenes) * c3["B"] * c3["iters"]
    model = survey_bytes_per_hit(P=c3["H"] * c3["W"], c=0.05)["total"]
    recs.append({"config": 3, "workload": c3["name"] + "; 8 objects x 128 hypotheses x 100 iterations, rgb+depth+mask+Sobel edge, sequential object loop on one GPU",
                 "stand_in": True, "hypotheses_per_gpu": c3["B"], "objects": len(scenes), "iters": c3["iters"], "ms_total": ms, "ms_per_iter": ms / c3["iters"],
                 "value": hits / (ms * 1e-3), "unit": UNIT, "survey_bytes_per_hit": model, "hbm_frac": model * hits / (ms * 1e-3) / 1e9 / peak})
    del scenes, runs

    # ---- config 4 (strong scaling): 256 hypotheses split over the ranks, 100 iterations, depth + mask (stand-in)
    c4 = wl.config4()
    s4, _, cov4 = standin_scene(nat, c4, dev)
    per = (STRONG_B + world - 1) // world
    lo = min(rank * per, STRONG_B)
    Bl = min(lo + per, STRONG_B) - lo
    lr4 = torch.from_numpy(su.lr_multipliers(STRONG_B, 0.01, 3.0)[lo:lo + Bl].copy()).to(dev)
    ms = maxreduce(time_call(sgd(s4, c4["q0"], c4["t0"], Bl, lr4, _loss_cfg(nat, c4["losses"]), c4["iters"], b_global=STRONG_B)))
    model = survey_bytes_per_hit(V=5002, T=10000, P=c4["H"] * c4["W"], c=0.05, rgb=False, textured=False)["total"]
    recs.append({"config": 4, "workload": c4["name"] + "; 256 hypotheses in total over %d GPU(s), 100 iterations, depth+mask" % world, "stand_in": True,
                 "scaling": "strong", "global_hypotheses": STRONG_B, "hypotheses_per_gpu": Bl, "iters": c4["iters"], "ms_per_iter": ms / c4["iters"],
                 "value": STRONG_B * c4["iters"] / (ms * 1e-3), "value_per_gpu": Bl * c4["iters"] / (ms * 1e-3), "unit": UNIT, "survey_bytes_per_hit": model,
                 "hbm_frac": model * Bl * c4["iters"] / (ms * 1e-3) / 1e9 / peak, "timing": "max over ranks of the CUDA-event time of one call (median of 3)"})
    del s4

    # ---- config 5 (HBM stress): the 8-GPU shard, 128 hypotheses per GPU at 1024^2, 50k triangles, full stack incl. Sobel (stand-in)
    c5 = wl.config5()
    s5, _, cov5 = standin_scene(nat, c5, dev)
    B5 = c5["B"] // 8
    lr5 = torch.from_numpy(su.lr_multipliers(c5["B"], 0.01, 2.0)[rank * B5 % c5["B"]:][:B5].copy()).to(dev)
    model = survey_bytes_per_hit(V=25002, T=50000, P=c5["H"] * c5["W"], c=0.5)["total"]
    for name, L in (("rgb+depth+mask+Sobel edge", c5["losses"]), ("rgb+depth+mask (no edge loss)", {k: v for k, v in c5["losses"].items() if "edge" not in k})):
        iters5 = 20
        ms = maxreduce(time_call(sgd(s5, c5["q0"], c5["t0"], B5, lr5, _loss_cfg(nat, L), iters5, b_global=c5["B"])))
        recs.append({"config": 5, "workload": c5["name"] + "; 128 hypotheses per GPU (the 8-GPU shard of 1024), %d of the 50 iterations, %s; covered fraction %.2f"
                     % (iters5, name, cov5), "stand_in": True, "hypotheses_per_gpu": B5, "iters": iters5, "ms_per_iter": ms / iters5,
                     "value_per_gpu": B5 * iters5 / (ms * 1e-3), "unit": UNIT, "survey_bytes_per_hit": model,
                     "hbm_frac": model * B5 * iters5 / (ms * 1e-3) / 1e9 / peak})
    del s5
    torch.cuda.empty_cache()
    return recs


# ----------------------------------------------------------------------------------------------
# our arm


def run_ours(args):
    import torch.distributed as dist

    import scene_util as su
    import workloads as wl
    from diffdope import _native as nat

    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the hot path has no CPU fallback")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n_gpus = world
    strong = args.scaling == "strong"
    K, Wm = args.steps, max(args.warmup, 3)

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak = float(json.load(open(peaks_path))["hbm_gbs"])
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"

    if strong:
        w = wl.config4()
        B_global = STRONG_B
        per = (B_global + world - 1) // world
        lo = min(rank * per, B_global)
        B = min(lo + per, B_global) - lo
        sc, tgt, _ = standin_scene(nat, w, dev)
        q, t = w["q0"], w["t0"]
        lr_all = su.lr_multipliers(B_global, 0.01, 3.0)
        lr_host = lr_all[lo:lo + B].copy()
        cfg = _loss_cfg(nat, w["losses"])
        model = survey_bytes_per_hit(V=5002, T=10000, P=w["H"] * w["W"], c=0.05, rgb=False, textured=False)
        gt_host = None
    else:
        B = B_PER_GPU
        B_global = B * n_gpus
        arr = su.example_mesh_arrays()
        q, t = su.example_pose()
        gt_host = su.example_targets(1.0)
        H, Wd = gt_host["rgb"].shape[:2]
        window = su.centred_window(gt_host["segmentation"], WINDOW, H, Wd)
        seg1 = np.ascontiguousarray(gt_host["segmentation"][..., 0])
        lr_all = su.lr_multipliers(B_global)
        lr_host = lr_all[rank * B:(rank + 1) * B].copy()
        sc = nat.NativeScene(arr["pos"], arr["tri"], arr["uv"], arr["tex"])
        sc.set_camera(su.projection_native(), H, Wd)
        sc.set_window(*window)
        g_rgb = torch.from_numpy(gt_host["rgb"]).to(dev)
        g_depth = torch.from_numpy(gt_host["depth"]).to(dev)
        g_seg = torch.from_numpy(seg1).to(dev)
        sc.set_target(g_rgb, g_depth, g_seg)
        cfg = nat.make_loss_cfg(**LOSSES)
        model = survey_bytes_per_hit()
    lr = torch.from_numpy(lr_host).to(dev)
    sched = lr_schedule(K)
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev, dtype=torch.float32)

    def fresh_pose():
        return (torch.from_numpy(np.tile(q, (B, 1))).to(dev).contiguous(), torch.from_numpy(np.tile(t, (B, 1))).to(dev).contiguous())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up ------------------------------------------------------------------------------
    sampler = ClockSampler(local_rank) if rank == 0 else None  # runs across warm-up + every timed region
    qd, td = fresh_pose()
    sc.optimize(qd, td, lr, lr_schedule(Wm), cfg, b_global=B_global, keep_history=False)
    torch.cuda.synchronize()

    # ---- timed: K iterations, one C-ABI call each, L2 flushed between them ----------------------
    qd, td = fresh_pose()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    loss_tab = torch.empty(K, B, nat.NUM_LOSSES, device=dev)
    pose_tab = torch.empty(K, B, 7, device=dev)
    barrier()
    launches = 0
    wall0 = time.perf_counter()
    # every argument of the C-ABI call is prepared outside the timed loop: the loop body is flush, event, call, event
    import ctypes
    optimize = nat.lib().ddope_optimize
    sched_np = np.asarray(sched, dtype=np.float32)
    sched_ptr = [ctypes.c_void_p(sched_np.ctypes.data + 4 * i) for i in range(K)]
    pose_ptr = [ctypes.c_void_p(pose_tab.data_ptr() + 4 * i * B * 7) for i in range(K)]
    loss_ptr = [ctypes.c_void_p(loss_tab.data_ptr() + 4 * i * B * nat.NUM_LOSSES) for i in range(K)]
    q_ptr, t_ptr, lr_ptr, cfg_ref, stream = nat._ptr(qd), nat._ptr(td), nat._ptr(lr), ctypes.byref(cfg), nat._stream()
    rcs = 0
    for i in range(K):
        flush.fill_(0.0)
        ev[i][0].record()
        rcs |= optimize(sc._h, q_ptr, t_ptr, lr_ptr, B, B_global, sched_ptr[i], 1, cfg_ref, pose_ptr[i], loss_ptr[i], stream)
        ev[i][1].record()
    nat._check(rcs)
    launches = K * sc.last_launch_count()
    barrier()
    wall = time.perf_counter() - wall0
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    tms = torch.tensor([dev_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    dev_ms_max = float(tms.item())
    value = B_global * K / (dev_ms_max * 1e-3)

    # ---- the natural mode: one call, all iterations back to back, L2 warm -----------------------
    qd2, td2 = fresh_pose()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    sc.optimize(qd2, td2, lr, sched, cfg, b_global=B_global, keep_history=True)
    e1.record()
    barrier()
    hot_ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(hot_ms, op=dist.ReduceOp.MAX)
    value_hot = B_global * K / (float(hot_ms.item()) * 1e-3)

    # ---- per-kernel times (CUDA events on the launching stream inside the library) --------------
    qd3, td3 = fresh_pose()
    sc.profile_begin()
    sc.optimize(qd3, td3, lr, sched, cfg, b_global=B_global, keep_history=False)
    kms, klaunch = sc.profile_end()
    n_prof = max(klaunch["pixel_kernel"], 1)  # = iterations
    t_load = time.perf_counter()
    # keep the GPU under the same load until the 20 ms clock sampler has delivered >= 10 samples (nvidia-smi needs a second or more
    # to start on a fresh box; bounded at 10 s)
    while time.perf_counter() - t_load < 0.3 or (sampler is not None and sampler.proc is not None and sampler.count() < 10 and time.perf_counter() - t_load < 10.0):
        qd4, td4 = fresh_pose()
        sc.optimize(qd4, td4, lr, sched, cfg, b_global=B_global, keep_history=False)
        torch.cuda.synchronize()
    clocks = sampler.stop() if sampler else None

    # ---- forward only (BASELINE.json's metric also names forward+backward ms/iter: that is ms_per_step) -------------
    # ddope_render of all hypotheses over the loss window, images written to HBM (rgb, depth, mask: 20 B per pixel)
    fwd_ms = None
    try:
        qd5, td5 = fresh_pose()
        fwd_ms = time_call(lambda: [sc.render(qd5, td5, want=("rgb", "depth", "mask")) for _ in range(10)]) / 10
    except Exception as e:  # an extra, never the reason the bench line is missing
        print("forward-only timing skipped:", e, file=sys.stderr)

    # ---- e2e with host buffers ------------------------------------------------------------------
    try:
        if strong:
            e2e = run_e2e_cabi(nat, dist, sc, tgt, dev, K, B, B_global, rank, world, q, t, lr_host, cfg, barrier)
        else:
            e2e = run_e2e(dev, K, B, B_global, rank, world, gt_host, lr_all, barrier, raw=not args.e2e_float)
    except Exception as e:  # keep the device-timed line if the API leg fails; the error is reported, not hidden
        import traceback

        traceback.print_exc(file=sys.stderr)
        e2e = {"value": None, "unit": UNIT, "h2d_bytes_per_step": None, "d2h_bytes_per_step": None, "error": "%s: %s" % (type(e).__name__, e)}

    # ---- the other BASELINE configurations, bounded ---------------------------------------------
    configs = None
    if not args.no_configs and not strong:
        try:
            del flush
            torch.cuda.empty_cache()
            configs = config_records(nat, dev, rank, world, peak, dist)
        except Exception as e:
            import traceback

            traceback.print_exc(file=sys.stderr)
            configs = {"error": "%s: %s" % (type(e).__name__, e)}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel --------------------------------------------------------
    dom = max(("pixel_kernel", "raster_kernel"), key=lambda k: kms[k])
    dom_ms = kms[dom] / max(n_prof, 1)
    achieved = model[dom] * B / (dom_ms * 1e-3) / 1e9
    counters, counters_src = load_counters()
    traffic = None
    if counters and not strong:
        traffic = counters["kernels"].get(dom, {}).get("dram_bytes")
    roofline = {
        "bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
        "peak_source": peak_src, "kernel_ms_per_launch": dom_ms,
        "algorithmic_bytes_per_launch": model[dom] * B,
        "model": "SURVEY.md 8(d) byte model per hypothesis-iteration (%.2f MB), the terms this kernel moves (%.2f MB) x %d hypotheses per launch. "
                 "The model charges a full-window streaming implementation; this design touches only the loss ROI (10 %% of the window here) and "
                 "shares targets / mesh / texture across hypotheses in L2, so `frac` can exceed 1 and is NOT the bound that binds: see `issue`"
                 % (model["total"] / 1e6, model[dom] / 1e6, B),
        "kernel_ms_per_iteration": {k: v / max(n_prof, 1) for k, v in kms.items()},
        "whole_step_frac": model["total"] * value / n_gpus / 1e9 / peak,
        "traffic_source": counters_src,
    }
    if traffic and dom_ms > 0:  # what the counters say the kernel really pulled from HBM (ncu capture) at the live kernel time
        roofline["dram_gbs_from_traffic"] = float(traffic) / (dom_ms * 1e-3) / 1e9
        roofline["dram_frac_from_traffic"] = roofline["dram_gbs_from_traffic"] / peak
    if counters and not strong:
        # the bound that binds: warp-instruction issue. floor = warp instructions of one launch (ncu smsp__inst_executed.sum, committed
        # capture of these sources) / (SMs x 4 schedulers x SM clock under load); frac = floor / measured kernel time
        mhz = (clocks or {}).get("sm_mhz") or counters.get("sm_mhz") or 1965.0
        sms = torch.cuda.get_device_properties(dev).multi_processor_count
        issue = {"bound": "issue", "unit": "us", "sm_mhz": mhz, "sms": sms, "kernels": {}}
        tot_floor = 0.0
        for k, rec in counters["kernels"].items():
            floor_us = rec["inst_executed"] / (sms * 4 * mhz)
            meas_us = 1e3 * kms.get(k, 0.0) / max(n_prof, 1)
            issue["kernels"][k] = {"warp_inst_per_launch": rec["inst_executed"], "floor_us": floor_us, "measured_us": meas_us,
                                   "frac": (floor_us / meas_us) if meas_us > 0 else None}
            tot_floor += floor_us
        it_us = 1e3 * float(hot_ms.item()) / K
        issue.update(floor_us_per_iteration=tot_floor, measured_us_per_iteration_warm=it_us, frac=tot_floor / it_us,
                     measured_us_per_iteration_flushed=1e3 * dev_ms_max / K, frac_flushed=tot_floor / (1e3 * dev_ms_max / K))
        roofline["issue"] = issue

    # ---- cpu baseline (bounded sample) ----------------------------------------------------------
    cpu = None
    if n_gpus == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        ref = CpuReference(2, threads)
        ref.iterate(1)  # warm-up (first-touch of the 1080p buffers, thread pool start)
        dt = ref.iterate(8)
        v = 2 * 8 / dt
        cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": "2 hypotheses x 8 iterations of the simple_scene workload with oracle/refpath.py (full-frame render, %dx%d loss window), %.1f s" % (WINDOW, WINDOW, dt)}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": K, "warmup": Wm,
        "ms_per_step": dev_ms_max / K, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32",
        "data": ("stand-in workload (tests/workloads.py config4): procedural mesh, targets rendered at the ground-truth pose, synthetic multipliers" if strong else
                 "reference example scene (data/example: HOPE AlphabetSoup mesh + rgb/depth/seg images), synthetic learning-rate multipliers (random.seed(0))"),
        "config": workload_config(n_gpus, args.scaling),
        "value_l2_warm_single_call": value_hot,
        "forward_only_ms_per_iter": fwd_ms,
        "wall_s_timed_region": wall,
        "clocks": clocks,
        "e2e": e2e,
        "gpu_launches": launches,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "configs": configs,
        "kernel_source_sha": source_sha(),
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def multi_gpu_check(ddope, dev, K, B, B_global, world, lr_all):
    """Rank 0 recomputes, alone, the first and the last hypothesis of every OTHER rank's shard (same start pose, same multiplier,
    B_global in the mean) and compares the pose and loss histories bit for bit with what the all-gather delivered."""
    from diffdope import _native as nat

    idx = sorted({i for r in range(1, world) for i in (r * B, r * B + B - 1)})
    sc = ddope._native_scene
    L = ddope.cfg.losses
    cfg = nat.make_loss_cfg(True, True, True, L.weight_rgb, L.weight_depth, L.weight_mask)
    q0, t0 = ddope._start_pose
    q = q0[idx].to(dev).contiguous()
    t = t0[idx].to(dev).contiguous()
    lr = torch.from_numpy(lr_all[idx].copy()).to(dev)
    ph, lh = sc.optimize(q, t, lr, ddope._lr_schedule(), cfg, b_global=B_global)
    ph, lh = ph.cpu(), lh.cpu()
    ok_pose = bool(torch.equal(ph, ddope._pose_hist_host[:, idx]))
    ok_loss = all(bool(torch.equal(lh[:, :, c], ddope.losses_values[k][:, idx])) for c, k in ((0, "rgb"), (1, "depth"), (2, "mask_selection")))
    return {"ranks": world, "recomputed_on_rank0": idx, "iterations": K, "pose_history_bitwise_equal": ok_pose, "loss_history_bitwise_equal": ok_loss}


def timed_jobs(job, barrier, dev, world, dist, reps=5):
    """Two warm-up jobs, then `reps` jobs, each bracketed by barrier + synchronize; per job the wall clock (host work is part of end to
    end), max over ranks. Returns (median, all)."""
    job()
    job()
    out = []
    for _ in range(reps):
        barrier()
        t0 = time.perf_counter()
        job()
        barrier()
        ms = torch.tensor([(time.perf_counter() - t0) * 1e3], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        out.append(float(ms[0].item()))
    return float(np.median(out)), out


def run_e2e(dev, K, B, B_global, rank, world, gt_host, lr_all, barrier, raw=True):
    """K iterations through `DiffDope.run_optimization` with the target images, start poses and
    learning-rate multipliers coming from pinned host memory and the result tables read back.
    raw=True (default): the targets cross PCIe as the PNGs' uint8 / uint16 samples and are converted on the device
    (`Image.set_raw` -> ddope_image_from_raw, bit-equal to the host pipeline); the segmentation, whose three channels are
    equal in the file (checked here, outside the timed region), travels as one channel. --e2e-float: float32 images."""
    import torch.distributed as dist
    from omegaconf import OmegaConf

    import diffdope as dd
    import scene_util as su

    cfg = OmegaConf.load(os.path.join(ROOT, "configs", "diffdope.yaml"))
    cfg.scene.image_resize = 1.0
    for k in ("path_img", "path_depth", "path_segmentation"):
        cfg.scene[k] = os.path.join(ROOT, cfg.scene[k])
    cfg.object3d.model_path = os.path.join(ROOT, cfg.object3d.model_path)
    cfg.losses.l1_rgb_with_mask = True
    cfg.losses.l1_depth_with_mask = True
    cfg.losses.l1_mask = True
    cfg.hyperparameters.batchsize = B_global
    cfg.hyperparameters.nb_iterations = K - 1
    ddope = dd.DiffDope(cfg=cfg)
    H, W = gt_host["rgb"].shape[:2]
    ddope.window = su.centred_window(gt_host["segmentation"], WINDOW, H, W)
    pin = {k: torch.from_numpy(v).pin_memory() for k, v in gt_host.items()}
    if raw:
        import cv2

        def samples(a):
            a = a.view(np.int16) if a.dtype == np.uint16 else a
            return torch.from_numpy(np.ascontiguousarray(a)).pin_memory()

        seg = cv2.imread(cfg.scene.path_segmentation, cv2.IMREAD_COLOR)
        seg = seg[..., 0] if (np.array_equal(seg[..., 0], seg[..., 1]) and np.array_equal(seg[..., 0], seg[..., 2])) else seg
        pin = {"rgb": samples(cv2.imread(cfg.scene.path_img, cv2.IMREAD_COLOR)), "depth": samples(cv2.imread(cfg.scene.path_depth, cv2.IMREAD_UNCHANGED)),
               "segmentation": samples(seg)}
    lr_pin = torch.from_numpy(lr_all.copy()).pin_memory()
    q0, t0 = ddope.object3d.pose_tensors()
    q_pin, t_pin = q0.cpu().pin_memory(), t0.cpu().pin_memory()
    ddope._start_pose = (q_pin, t_pin)
    h2d = sum(v.numel() * v.element_size() for v in pin.values()) + lr_pin.numel() * 4 + q_pin.numel() * 4 + t_pin.numel() * 4

    def job():
        if raw:
            ddope.scene.tensor_rgb.set_raw(pin["rgb"], device=dev)
            ddope.scene.tensor_depth.set_raw(pin["depth"], device=dev)
            ddope.scene.tensor_segmentation.set_raw(pin["segmentation"], device=dev)
        else:
            ddope.scene.tensor_rgb.img_tensor = pin["rgb"].to(dev, non_blocking=True)
            ddope.scene.tensor_depth.img_tensor = pin["depth"].to(dev, non_blocking=True)
            ddope.scene.tensor_segmentation.img_tensor = pin["segmentation"].to(dev, non_blocking=True)
        for im in (ddope.scene.tensor_rgb, ddope.scene.tensor_depth, ddope.scene.tensor_segmentation):
            im._batchsize_set = False
            im.set_batchsize(B_global)
        ddope.learning_rates = lr_pin.to(dev, non_blocking=True)
        ddope.object3d.load_pose_tensors(q_pin.to(dev, non_blocking=True), t_pin.to(dev, non_blocking=True))
        ddope.run_optimization()  # ends with the result tables on the host (losses_values, poses)
        best = int(ddope.get_argmin())
        return best, ddope.get_pose(best)

    total_ms, all_ms = timed_jobs(job, barrier, dev, world, dist)
    d2h = K * B_global * (7 + 4) * 4 + B_global * 7 * 4
    out = {"value": B_global * K / (total_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d / K, "d2h_bytes_per_step": d2h / K,
           "ms_per_step": total_ms / K, "api": "diffdope.DiffDope.run_optimization (host -> device -> host)",
           "targets_on_the_wire": "uint8 / uint16 samples, converted on the device (Image.set_raw)" if raw else "float32",
           "job_ms_max_over_ranks": all_ms, "timing": "wall clock of one whole job between barriers, max over ranks; median of %d jobs after 2 warm-up jobs" % len(all_ms)}
    if world > 1:
        chk = multi_gpu_check(ddope, dev, K, B, B_global, world, lr_all) if rank == 0 else None
        barrier()
        if rank == 0:
            out["multi_gpu_check"] = chk
            if not (chk["pose_history_bitwise_equal"] and chk["loss_history_bitwise_equal"]):
                raise RuntimeError("multi-GPU result differs from the single-GPU recomputation: %r" % (chk,))
    return out


def run_e2e_cabi(nat, dist, sc, tgt, dev, K, B, B_global, rank, world, q, t, lr_host, cfg, barrier):
    """Strong-scaling workload end to end through the C ABI with host buffers: targets, start poses and multipliers copied from pinned
    memory, ddope_optimize, one all-gather of the flat result buffers, one copy back to pinned memory."""
    from diffdope import _dist

    pin = {k: v.cpu().pin_memory() for k, v in tgt.items()}
    q_pin = torch.from_numpy(np.tile(q, (B, 1))).pin_memory()
    t_pin = torch.from_numpy(np.tile(t, (B, 1))).pin_memory()
    lr_pin = torch.from_numpy(lr_host).pin_memory()
    sched = lr_schedule(K)
    n, Kl = K, nat.NUM_LOSSES
    per = (B_global + world - 1) // world
    host = torch.empty(world, max(_dist.flat_sizes(n, per, Kl)[2], 1), dtype=torch.float32, pin_memory=True)
    h2d = sum(v.numel() * 4 for v in pin.values()) + (q_pin.numel() + t_pin.numel() + lr_pin.numel()) * 4

    def job():
        d = {k: v.to(dev, non_blocking=True) for k, v in pin.items()}
        sc.set_target(d["rgb"], d["depth"], d["seg"])
        qd, td, lrd = q_pin.to(dev, non_blocking=True), t_pin.to(dev, non_blocking=True), lr_pin.to(dev, non_blocking=True)
        a, b, c = _dist.flat_sizes(n, B, Kl)
        flat = torch.empty(max(c, 1), device=dev)
        sc.optimize(qd, td, lrd, sched, cfg, b_global=B_global, out=(flat[:a].view(n, B, 7), flat[a:b].view(n, B, Kl)))
        fin = flat[b:c].view(B, 7)
        fin[:, :4].copy_(qd)
        fin[:, 4:].copy_(td)
        allf = _dist.gather_flat(flat, host.shape[1])
        host.copy_(allf, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        ph, lh, final = _dist.unpack_flat(host, B_global, n, Kl)
        return int(lh[-1][:, 1:3].mean(-1).argmin())

    total_ms, all_ms = timed_jobs(job, barrier, dev, world, dist)
    d2h = host.numel() * 4
    return {"value": B_global * K / (total_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d / K, "d2h_bytes_per_step": d2h / K, "ms_per_step": total_ms / K,
            "api": "C ABI (ddope_scene_set_target + ddope_optimize) with pinned host buffers, one all_gather_into_tensor, one copy back",
            "job_ms_max_over_ranks": all_ms, "timing": "wall clock of one whole job between barriers, max over ranks; median of %d jobs after 2 warm-up jobs" % len(all_ms)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"], help="strong: BASELINE configs[3], 256 hypotheses split over the GPUs")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the bounded measurements of BASELINE configs 1, 3, 4, 5")
    ap.add_argument("--e2e-float", action="store_true", help="e2e: upload the targets as float32 images instead of integer samples")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
