"""One stem + conv1b launch pair for `ncu --set full -k regex:stem_conv` (V2-S, 240x320, batch 64)."""
import contextlib, io, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nano_vs_slam_b200 import tiny_factory
from nano_vs_slam_b200.synthetic import spread_init, synthetic_frames
with contextlib.redirect_stdout(io.StringIO()):
    m = tiny_factory("S", 28, v3=False)
m.load_state_dict(spread_init(m.state_dict(), 1234))
m.eval(); m.training = False
m = m.cuda(); m.cuda_graph_max_batch = 0
x = synthetic_frames(64, 240, 320, 0).cuda()
out = m(x)
torch.cuda.synchronize()
print("ok")
