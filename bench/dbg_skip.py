import sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from vokselis_b200 import abi, rt
W, H = 1280, 720
cam = rt.Camera(3.0, -0.5, 1.0, (0, 0, 0), W / H).get_proj_view_matrix()
with rt.Context(0, W, H) as ctx:
    ctx.generate_xor(256, 0)
    res = {}
    for skip in (0, 1):
        q = rt.default_params(abi.MODE_M0)
        q.skip_empty, q.count_samples = skip, 1
        ctx.set_params(q)
        ctx.reset_stats()
        ctx.render(cam)
        res[skip] = (ctx.readback(), ctx.readback_aux(), ctx.stats())
    f0, a0, s0 = res[0]; f1, a1, s1 = res[1]
    diff = (f0 != f1).any(axis=-1)
    print("differing pixels", diff.sum(), "aux differing", (a0 != a1).sum(), "iters", s0.samples_reference, s1.samples_reference, "fetched", s0.samples_fetched, s1.samples_fetched)
    ys, xs = np.nonzero(diff)
    for y, x in list(zip(ys, xs))[:8]:
        print((x, y), f0[y, x].view(np.float16), f1[y, x].view(np.float16), a0[y, x] & 0x7fffffff, a1[y, x] & 0x7fffffff)
