"""Extract the BOP YCB-V pose rows BASELINE config 3 uses from the reference's data files:
frame "1" of scene 000048, ground-truth-like poses (`scene_error_deg_001_trans_001.json`) and perturbed
initial poses (`scene_error_deg_010_trans_004.json`). Run in the build container (the reference tree does
not exist on the GPU box); the output is committed next to this script."""
import json
import os

REF = "/root/reference/data/ycbv/test/000048"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ycbv_000048_frame1.json")

if __name__ == "__main__":
    gt = json.load(open(os.path.join(REF, "scene_error_deg_001_trans_001.json")))["1"]
    init = json.load(open(os.path.join(REF, "scene_error_deg_010_trans_004.json")))["1"]
    assert [o["obj_id"] for o in gt] == [o["obj_id"] for o in init]
    json.dump({"source": "NVlabs/diff-dope data/ycbv/test/000048, frame '1'", "gt": gt, "init": init}, open(OUT, "w"), indent=1)
    print("wrote", OUT, len(gt), "objects")
