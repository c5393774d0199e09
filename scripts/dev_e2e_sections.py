"""Host-side wall time of the sections of DiffDope.run_optimization (bench.py's e2e job), after warm-up. Under torchrun (WORLD_SIZE > 1)
every rank runs its shard and rank 0 prints its own sections; K = iterations (env, default 200)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "diff-dope_b200"), os.path.join(ROOT, "tests")]
import numpy as np, torch
import torch.distributed as dist
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0"))))
import diffdope as dd
from omegaconf import OmegaConf
import scene_util as su
K, B = int(os.environ.get("K", "200")), 64 * world
cfg = OmegaConf.load(os.path.join(ROOT, "configs", "diffdope.yaml"))
cfg.scene.image_resize = 1.0
for k in ("path_img", "path_depth", "path_segmentation"): cfg.scene[k] = os.path.join(ROOT, cfg.scene[k])
cfg.object3d.model_path = os.path.join(ROOT, cfg.object3d.model_path)
cfg.losses.l1_rgb_with_mask = True; cfg.losses.l1_depth_with_mask = True; cfg.losses.l1_mask = True
cfg.hyperparameters.batchsize = B; cfg.hyperparameters.nb_iterations = K - 1
d = dd.DiffDope(cfg=cfg)
gt = su.example_targets(1.0)
d.window = su.centred_window(gt["segmentation"], 640, *gt["rgb"].shape[:2])
def T(): torch.cuda.synchronize(); return time.perf_counter()
for rep in range(6):
    if world > 1: dist.barrier()
    t0 = T(); d.losses_values = {}; d.optimization_results = []; d.optimizer = d._make_optimizer(); d._refresh_gt()
    t1 = time.perf_counter(); st = d._fused_enqueue(); t2 = time.perf_counter(); t2s = T()
    d._fused_finish(st); t3 = T(); best = int(d.get_argmin()); p = d.get_pose(best); t4 = T()
    if world > 1: dist.barrier()
    t5 = T()
    if rank == 0:
        print("world %d K %d rep %d: setup %.3f | enqueue (host, async) %.3f | gpu done after %.3f | finish (gather + D2H + host tables) %.3f | argmin/pose %.3f | closing barrier %.3f | total %.3f ms"
              % (world, K, rep, 1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t2s - t1), 1e3 * (t3 - t2s), 1e3 * (t4 - t3), 1e3 * (t5 - t4), 1e3 * (t5 - t0)), flush=True)
if os.environ.get("PROFILE") and rank == 0:
    import cProfile, pstats, io
    d.losses_values = {}; d.optimization_results = []; d.optimizer = d._make_optimizer(); d._refresh_gt()
    st = d._fused_enqueue(); torch.cuda.synchronize()
    pr = cProfile.Profile(); pr.enable()
    d._fused_finish(st); best = int(d.get_argmin()); p = d.get_pose(best)
    pr.disable()
    s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(22); print(s.getvalue()[:4500])
    pr = cProfile.Profile(); pr.enable()
    d.losses_values = {}; d.optimization_results = []; d.optimizer = d._make_optimizer(); d._refresh_gt(); st = d._fused_enqueue()
    pr.disable(); torch.cuda.synchronize()
    s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(14); print(s.getvalue()[:3500])
if world > 1: dist.destroy_process_group()
