"""Worker of tests/test_gpu_multi.py: one rank of a sharded `DiffDope.run_optimization` (launched by torch.distributed.run),
or, with WORLD_SIZE unset, the single-process job the ranks are compared with. Saves what the API publishes."""
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "diff-dope_b200"), os.path.join(ROOT, "tests")]

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    out_dir, B, backend = sys.argv[1], int(sys.argv[2]), sys.argv[3]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    ndev = torch.cuda.device_count()
    torch.cuda.set_device(local % ndev)
    if world > 1:
        if backend == "nccl":
            dist.init_process_group("nccl", device_id=torch.device("cuda", local % ndev))
        else:
            dist.init_process_group("gloo")
    import diffdope as dd
    from omegaconf import OmegaConf

    cfg = OmegaConf.load(os.path.join(ROOT, "configs", "diffdope.yaml"))
    for k in ("path_img", "path_depth", "path_segmentation"):
        cfg.scene[k] = os.path.join(ROOT, cfg.scene[k])
    cfg.object3d.model_path = os.path.join(ROOT, cfg.object3d.model_path)
    cfg.losses.l1_rgb_with_mask = True
    cfg.losses.l1_depth_with_mask = True
    cfg.hyperparameters.batchsize = B
    cfg.hyperparameters.nb_iterations = 5
    random.seed(100 + rank)  # ranks draw DIFFERENT multipliers on purpose: rank 0's draw must become the job's
    d = dd.DiffDope(cfg=cfg)
    if world == 1:
        random.seed(100)
        d.set_batchsize(B)  # the single-process job with rank 0's random state
    d.run_optimization()
    res = {"losses": {k: v.clone() for k, v in d.losses_values.items()}, "poses": d._pose_hist_host.clone(), "argmin": int(d.get_argmin()),
           "pose": torch.from_numpy(d.get_pose()), "lr": d.learning_rates.cpu(), "final": torch.stack([p.detach().cpu() for p in d.object3d.parameters()], 1)}
    torch.save(res, os.path.join(out_dir, "rank%d_of%d.pt" % (rank, world)))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
