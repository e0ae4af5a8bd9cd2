"""Writes a reduced copy of a runfile (small frames, few scenes) for quick entry-point checks.
usage: python tools/small_runfile.py runfiles/SonyA7S2/PNNP.yml /tmp/small.yml H W FRAMES"""
import sys, yaml
src, dst, H, W, frames = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
cfg = yaml.load(open(src), Loader=yaml.FullLoader)
for k in ("dst", "dst_train", "dst_eval", "dst_test"):
    cfg[k]["H"], cfg[k]["W"], cfg[k]["synthetic_frames"] = H, W, frames
    if "iso_list" in cfg[k]:
        cfg[k]["iso_list"] = cfg[k]["iso_list"][:1]
open(dst, "w").write(yaml.dump(cfg))
