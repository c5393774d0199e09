"""All objects of one BOP-format frame refined concurrently (one CUDA stream per object) instead of the
sequential per-object loop of examples/run_bop_scene.py (reference: examples/run_bop_scene.py:48-93).
Same command line as run_bop_scene.py; without +bop.* arguments three shifted copies of the example object
of configs/diffdope.yaml are refined to exercise the batched path."""
import json
import os

import cv2
import numpy as np

import diffdope as dd  # first: activates the hydra / omegaconf stand-ins when the real ones are missing

import hydra  # noqa: E402
from omegaconf import DictConfig  # noqa: E402


@hydra.main(version_base=None, config_path="../configs/", config_name="diffdope")
def main(cfg: DictConfig):
    base = dd.DiffDope(cfg=cfg)
    out_dir = hydra.core.hydra_config.HydraConfig.get().runtime.output_dir
    bop = cfg.get("bop")
    B = cfg.hyperparameters.batchsize
    if bop is None:
        p = np.array(cfg.object3d.position, dtype=np.float64)
        objects = [dict(obj_id=0, cam_t_m2c=list(p + d), cam_R_m2c=cfg.object3d.rotation) for d in ([0, 0, 0], [2, -1, 3], [-2, 1, -3])]
        frame, scene = "0", base.scene
    else:
        frame = str(bop.get("frame", "0"))
        with open(bop.poses) as f:
            objects = json.load(f)[frame]
        name = frame.zfill(6)
        scene = dd.Scene(path_img=f"{bop.scene_dir}/rgb/{name}.png", path_depth=f"{bop.scene_dir}/depth/{name}.png",
                         path_segmentation=f"{bop.scene_dir}/rgb/{name}.png", image_resize=cfg.scene.image_resize)
        scene.cuda()
        scene.set_batchsize(B)
    meshes, jobs = {}, []
    for i_obj, obj in enumerate(objects):
        if bop is None:
            mesh_path, mask_path = cfg.object3d.model_path, cfg.scene.path_segmentation
        else:
            mesh_path = f"{bop.models_dir}/obj_{str(obj['obj_id']).zfill(6)}.ply"
            mask_path = f"{bop.scene_dir}/mask_visib/{frame.zfill(6)}_{str(i_obj).zfill(6)}.png"
        if obj["obj_id"] not in meshes:
            m = dd.Mesh(mesh_path, scale=cfg.object3d.scale)
            m.set_batchsize(B)
            m.cuda()
            meshes[obj["obj_id"]] = m
        pose = dd.Object3D(position=obj["cam_t_m2c"], rotation=obj["cam_R_m2c"], scale=cfg.object3d.scale, batchsize=B)
        pose.mesh = meshes[obj["obj_id"]]
        pose.cuda()
        mask = dd.Image(img_path=mask_path, img_resize=cfg.scene.image_resize)
        mask.cuda()
        mask.set_batchsize(B)
        # every object gets its own Scene view: shared rgb / depth images, its own visibility mask
        sc = dd.Scene(tensor_rgb=scene.tensor_rgb, tensor_depth=scene.tensor_depth, tensor_segmentation=mask)
        jobs.append(dd.DiffDope(cfg=cfg, camera=base.camera, object3d=pose, scene=sc))
    dd.run_optimization_batched(jobs)
    for i_obj, job in enumerate(jobs):
        print(f"object {i_obj}: best hypothesis {int(job.get_argmin())}\n{job.get_pose()}")
        cv2.imwrite(os.path.join(out_dir, f"{str(i_obj).zfill(2)}.png"), job.render_img())
    print("wrote", out_dir)


if __name__ == "__main__":
    main()
