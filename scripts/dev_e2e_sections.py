"""Host-side wall time of the sections of DiffDope.run_optimization (bench.py's e2e job), after warm-up."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "diff-dope_b200"), os.path.join(ROOT, "tests")]
import numpy as np, torch
import diffdope as dd
from omegaconf import OmegaConf
import scene_util as su
K, B = 200, 64
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
for rep in range(3):
    t0 = T(); d.losses_values = {}; d.optimization_results = []; d.optimizer = d._make_optimizer(); d._refresh_gt()
    t1 = time.perf_counter(); st = d._fused_enqueue(); t2 = time.perf_counter(); t2s = T()
    d._fused_finish(st); t3 = T(); best = int(d.get_argmin()); p = d.get_pose(best); t4 = T()
    print("rep %d: setup %.2f ms | enqueue (host, async) %.2f ms | gpu done after %.2f ms | finish (D2H + host tables) %.2f ms | argmin/pose %.2f ms | total %.2f ms"
          % (rep, 1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t2s - t1), 1e3 * (t3 - t2s), 1e3 * (t4 - t3), 1e3 * (t4 - t0)))
import cProfile, pstats, io
pr = cProfile.Profile(); pr.enable(); st = d._fused_enqueue(); pr.disable(); torch.cuda.synchronize()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(18); print(s.getvalue()[:3500])
