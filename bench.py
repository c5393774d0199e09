#!/usr/bin/env python
"""Benchmark of the Diff-DOPE hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[1]): simple_scene, HOPE AlphabetSoup, 64 pose hypotheses per GPU,
640x640 loss window on the 1920x1080 frame, rgb + depth + mask losses (weights 0.7/1/1), SGD with
the reference schedule. One *step* = one optimisation iteration of all hypotheses of a rank:
forward render + losses + backward to the 7 pose parameters + SGD step
(reference: one pass of the loop body at diffdope/diffdope.py:1656-1714).

Prints ONE JSON line (rank 0). `value` = hypothesis-iterations / s with everything resident in
HBM, timed per iteration with CUDA events, L2 flushed between iterations. `e2e` = the same metric
through the public API (`DiffDope.run_optimization`) with the target images and poses coming from
pinned host memory and the result tables read back, inside the timed region.
`--impl reference` times the CPU restatement of the reference path (oracle/, the reference itself
has no CPU path and its GPU path needs nvdiffrast + OpenGL).
"""
import argparse
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


def lr_schedule(n):
    nb = max(n - 1, 1)
    return [HYPER["base_lr"] * HYPER["lr_decay"] ** (it / nb + 1) for it in range(n)]


def workload_config(n_gpus):
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
        "scheduling": "3 kernels per iteration with programmatic dependent launch; hypotheses split into parts on internal streams "
                      "(raster of one part overlaps the pixel pass of another), joined into the caller's stream",
    }


# ----------------------------------------------------------------------------------------------
# survey byte model, split by the kernel that moves each term


def survey_bytes_per_hit(V=8240, T=13860, P=WINDOW * WINDOW, c=0.06):
    gt = P * 20.0  # rgb 12 + depth 4 + seg 4
    rast_w = P * 16.0
    rast_r = P * 16.0
    mesh = V * 20.0 + T * 12.0
    tex = 2 * 48.0 * c * P
    return {"total": gt + rast_w + rast_r + mesh + tex, "pixel_kernel": gt + rast_r + tex, "raster_kernel": rast_w + mesh}


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
    steps = max(1, min(args.steps, 10))  # bounded (~15 s of CPU work): each step is one iteration of a 2-hypothesis sample
    dt = sum(ref.iterate(1) for _ in range(steps))
    value = sample_hyp * steps / dt
    sample = ("%d hypotheses x 1 iteration per step of the bench workload (full 1920x1080 frame rendered as the reference "
              "does, loss over the %dx%d window), %d steps, %d torch threads") % (sample_hyp, WINDOW, WINDOW, steps, threads)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "reference example scene (data/example), synthetic learning-rate multipliers (random.seed(0))",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference semantics restated on CPU (oracle/); the reference's own GPU path needs nvdiffrast + OpenGL, absent here",
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------
# our arm


def run_ours(args):
    import torch.distributed as dist

    import scene_util as su
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
    B = B_PER_GPU
    B_global = B * n_gpus
    K, Wm = args.steps, max(args.warmup, 3)

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
    lr = torch.from_numpy(lr_host).to(dev)
    cfg = nat.make_loss_cfg(**LOSSES)
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
    for _ in range(3):  # keep the GPU under the same load long enough for the 20 ms clock sampler
        qd4, td4 = fresh_pose()
        sc.optimize(qd4, td4, lr, sched, cfg, b_global=B_global, keep_history=False)
    torch.cuda.synchronize()
    clocks = sampler.stop() if sampler else None

    # ---- forward only (BASELINE.json's metric also names forward+backward ms/iter: that is ms_per_step) -------------
    # ddope_render of all hypotheses over the loss window, images written to HBM (rgb, depth, mask: 20 B per pixel)
    fwd_ms = None
    try:
        qd5, td5 = fresh_pose()
        for _ in range(3):
            sc.render(qd5, td5, want=("rgb", "depth", "mask"))
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        f0.record()
        for _ in range(10):
            sc.render(qd5, td5, want=("rgb", "depth", "mask"))
        f1.record()
        torch.cuda.synchronize()
        fwd_ms = f0.elapsed_time(f1) / 10
    except Exception as e:  # an extra, never the reason the bench line is missing
        print("forward-only timing skipped:", e, file=sys.stderr)

    # ---- e2e through the public API with host buffers -------------------------------------------
    try:
        e2e = run_e2e(dev, K, B, B_global, rank, world, gt_host, lr_all, barrier, raw=args.e2e_raw)
    except Exception as e:  # keep the device-timed line if the API leg fails; the error is reported, not hidden
        import traceback

        traceback.print_exc(file=sys.stderr)
        e2e = {"value": None, "unit": UNIT, "h2d_bytes_per_step": None, "d2h_bytes_per_step": None, "error": "%s: %s" % (type(e).__name__, e)}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel --------------------------------------------------------
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak = float(json.load(open(peaks_path))["hbm_gbs"])
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    model = survey_bytes_per_hit()
    dom = max(("pixel_kernel", "raster_kernel"), key=lambda k: kms[k])
    dom_ms = kms[dom] / max(n_prof, 1)
    achieved = model[dom] * B / (dom_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "r01_traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get(dom)
        except Exception:
            traffic = None
    roofline = {
        "bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
        "peak_source": peak_src, "kernel_ms_per_launch": dom_ms,
        "algorithmic_bytes_per_launch": model[dom] * B,
        "model": "SURVEY.md 8(d) byte model per hypothesis-iteration (%.2f MB), the terms this kernel moves (%.2f MB) x %d hypotheses per launch; "
                 "the loss-ROI design touches far fewer DRAM bytes than the model (see traffic and DESIGN.md)" % (model["total"] / 1e6, model[dom] / 1e6, B),
        "kernel_ms_per_iteration": {k: v / max(n_prof, 1) for k, v in kms.items()},
    }
    if traffic and dom_ms > 0:  # what the counters say the kernel really pulled from HBM (ncu capture) at the live kernel time
        roofline["dram_gbs_from_traffic"] = float(traffic) / (dom_ms * 1e-3) / 1e9
        roofline["dram_frac_from_traffic"] = roofline["dram_gbs_from_traffic"] / peak

    # ---- cpu baseline (bounded sample) ----------------------------------------------------------
    cpu = None
    if n_gpus == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        ref = CpuReference(2, threads)
        ref.iterate(1)  # warm-up (first-touch of the 1080p buffers, thread pool start)
        dt = ref.iterate(8)
        v = 2 * 8 / dt
        cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": "2 hypotheses x 8 iterations of the bench workload with oracle/refpath.py (full-frame render, %dx%d loss window), %.1f s" % (WINDOW, WINDOW, dt)}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": K, "warmup": Wm,
        "ms_per_step": dev_ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "reference example scene (data/example: HOPE AlphabetSoup mesh + rgb/depth/seg images), synthetic learning-rate multipliers (random.seed(0))",
        "config": workload_config(n_gpus),
        "value_l2_warm_single_call": value_hot,
        "forward_only_ms_per_iter": fwd_ms,
        "wall_s_timed_region": wall,
        "clocks": clocks,
        "e2e": e2e,
        "gpu_launches": launches,
        "roofline": roofline,
        "cpu_baseline": cpu,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_e2e(dev, K, B, B_global, rank, world, gt_host, lr_all, barrier, raw=False):
    """K iterations through `DiffDope.run_optimization` with the target images, start poses and
    learning-rate multipliers coming from pinned host memory and the result tables read back.
    raw=True (--e2e-raw, off by default): the targets cross PCIe as the PNGs' uint8 / uint16 samples and are
    converted on the device (`Image.set_raw`, bit-equal to the host pipeline) instead of as float32."""
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

        def samples(path, flag):
            a = cv2.imread(path, flag)
            a = a.view(np.int16) if a.dtype == np.uint16 else a
            return torch.from_numpy(np.ascontiguousarray(a)).pin_memory()

        pin = {"rgb": samples(cfg.scene.path_img, cv2.IMREAD_COLOR), "depth": samples(cfg.scene.path_depth, cv2.IMREAD_UNCHANGED),
               "segmentation": samples(cfg.scene.path_segmentation, cv2.IMREAD_COLOR)}
    lr_pin = torch.from_numpy(lr_all.copy()).pin_memory()
    q0, t0 = ddope.object3d.pose_tensors()
    q_pin, t_pin = q0.cpu().pin_memory(), t0.cpu().pin_memory()
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

    job()  # warm-up
    barrier()
    t0_ = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    job()
    e1.record()
    barrier()
    wall = time.perf_counter() - t0_
    ms = torch.tensor([max(e0.elapsed_time(e1), 0.0), wall * 1e3], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    total_ms = float(ms[1].item())  # wall clock of the whole call: host work is part of end-to-end
    d2h = K * B_global * (7 + 3) * 4
    return {"value": B_global * K / (total_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d / K, "d2h_bytes_per_step": d2h / K,
            "ms_per_step": total_ms / K, "api": "diffdope.DiffDope.run_optimization (host -> device -> host)",
            "targets_on_the_wire": "uint8 / uint16 samples, converted on the device" if raw else "float32"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-raw", action="store_true", help="e2e: upload the targets as integer samples (Image.set_raw)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
