import numpy as np, torch, sys
sys.path.insert(0, '/root/repo')
import fl_slam_b200 as g
from oracle import gc_oracle as O
rng = np.random.default_rng(0)
for (n, B) in [(640, 256), (512, 48)]:
    P = rng.normal(size=(n, 3)) * 3 + np.array([4.0, 0, 0]); w = rng.uniform(0.1, 1, n)
    d = O.ray_directions(P, np.zeros(3)); resp = O.bin_soft_assign(d, O.create_fibonacci_atlas(B), 0.1)["responsibilities"]
    mm, c, e = g.scan_bin_moment_match(P, np.zeros((n,3,3)), w, resp, direction_origin=np.zeros(3))
    torch.cuda.synchronize()
    ref = O.scan_bin_moment_match(P, None, w, resp, None, np.zeros(3))
    print(n, B, float(np.max(np.abs(mm.N.cpu().numpy() - ref["N"]))))
