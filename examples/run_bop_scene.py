"""Refine every object of one BOP-format frame, one after the other, starting from perturbed poses
(same flow as the reference's examples/run_bop_scene.py:12-96, with the dataset location given on the
command line instead of hard-coded):

  PYTHONPATH=diff-dope_b200 python examples/run_bop_scene.py \
      +bop.scene_dir=/data/hope/val/000001 +bop.models_dir=/data/hope/models \
      +bop.poses=data/hope/val/000001/scene_error_deg_040_trans_016.json +bop.frame=0

A BOP scene dir holds rgb/<frame>.png, depth/<frame>.png, mask_visib/<frame>_<obj>.png; the pose file
is {frame: [{cam_R_m2c[9], cam_t_m2c[3], obj_id}, ...]}. Without +bop.* arguments the example object
of configs/diffdope.yaml is refined through the same per-object code path."""
import json
import os

import cv2

import diffdope as dd  # first: activates the hydra / omegaconf stand-ins when the real ones are missing

import hydra  # noqa: E402
from omegaconf import DictConfig  # noqa: E402


@hydra.main(version_base=None, config_path="../configs/", config_name="diffdope")
def main(cfg: DictConfig):
    ddope = dd.DiffDope(cfg=cfg)
    out_dir = hydra.core.hydra_config.HydraConfig.get().runtime.output_dir
    bop = cfg.get("bop")
    B = cfg.hyperparameters.batchsize
    if bop is None:
        objects = [dict(obj_id=0, cam_t_m2c=cfg.object3d.position, cam_R_m2c=cfg.object3d.rotation)]
        frame = "0"
    else:
        frame = str(bop.get("frame", "0"))
        with open(bop.poses) as f:
            objects = json.load(f)[frame]
        name = frame.zfill(6)
        scene = dd.Scene(path_img=f"{bop.scene_dir}/rgb/{name}.png", path_depth=f"{bop.scene_dir}/depth/{name}.png",
                         path_segmentation=f"{bop.scene_dir}/rgb/{name}.png", image_resize=cfg.scene.image_resize)
        scene.cuda()
        scene.set_batchsize(B)
    meshes = {}
    for i_obj, obj in enumerate(objects):
        if bop is None:
            mesh_path, mask_path, scene = cfg.object3d.model_path, cfg.scene.path_segmentation, ddope.scene
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
        scene.tensor_segmentation = mask
        ddope.scene = scene
        ddope.object3d = pose
        ddope.run_optimization()
        print(f"object {i_obj}: best hypothesis {int(ddope.get_argmin())}\n{ddope.get_pose()}")
        cv2.imwrite(os.path.join(out_dir, f"{str(i_obj).zfill(2)}.png"), ddope.render_img())
    print("wrote", out_dir)


if __name__ == "__main__":
    main()
