"""Where does the end-to-end time of DiffDope.run_optimization go? (host-side cProfile of bench.py's e2e job)"""
import cProfile, pstats, sys, os, io, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "diff-dope_b200"), os.path.join(ROOT, "tests")]
import numpy as np, torch
import diffdope  # activates the omegaconf stand-in
import bench, scene_util as su
dev = torch.device("cuda", 0); torch.cuda.set_device(0)
gt_host = su.example_targets(1.0)
K, B = 200, 64
lr_all = su.lr_multipliers(B)
def barrier(): torch.cuda.synchronize()
# monkeypatch run_e2e's inner job by re-running the function twice under a profiler
pr = cProfile.Profile()
t0 = time.perf_counter()
pr.enable()
out = bench.run_e2e(dev, K, B, B, 0, 1, gt_host, lr_all, barrier)
pr.disable()
print("run_e2e wall", time.perf_counter() - t0, out)
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(35); print(s.getvalue()[:6000])
