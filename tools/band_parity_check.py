import os, sys
sys.path.insert(0, 'tools'); sys.path.insert(0, '.')
import torch
import video_c5
from vggsfm_b200 import video
for frames, new in ((320, 128), (320, 128), (480, 256)):
    for band in ("0", "1", "0", "1"):
        os.environ["VGG_BAND"] = band
        out = video_c5.final_problem(frames=frames, new_per_window=new, reps=1)
        s = video.last_joint_summary
        print(frames, new, "band", band, "its", out["lm_iterations"][0], "init %.10e final %.12e" % (s.initial_cost, s.final_cost), "term", getattr(s, "termination", None), flush=True)
