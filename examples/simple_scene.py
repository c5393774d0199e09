"""Refine the pose of the example object and write the loss plot and an animation.
Run from the repository root:  PYTHONPATH=diff-dope_b200 python examples/simple_scene.py [key=value ...]
(the reference's own examples/simple_scene.py runs unchanged against this package too)."""
import cv2

import diffdope as dd  # first: activates the in-repo hydra / omegaconf stand-ins when the real ones are missing

import hydra  # noqa: E402
from omegaconf import DictConfig  # noqa: E402


@hydra.main(version_base=None, config_path="../configs/", config_name="diffdope")
def main(cfg: DictConfig):
    ddope = dd.DiffDope(cfg=cfg)
    ddope.run_optimization()
    best = int(ddope.get_argmin())
    print("best hypothesis:", best)
    print("pose (OpenGL camera frame, scaled units):\n", ddope.get_pose())
    plot = ddope.plot_losses()
    if plot is not None:
        cv2.imwrite("plot.png", plot)
    ddope.make_animation(output_file_path="simple_scene.mp4")
    print("wrote plot.png and simple_scene.mp4")


if __name__ == "__main__":
    main()
