"""Forward / backward feature fold at the training batch shape (16 frames 192x640, C = Hd = 64): CUDA-event timings and
effective HBM bandwidth.  Run under gpurun; under ncu use `-k regex:fold` for the source-level profile."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from behindthescenes_b200 import _lib

lib = _lib.load()
n, C, Hd, H, W = int(os.environ.get("N", 16)), 64, 64, 192, 640
g = torch.Generator().manual_seed(0)
feat = torch.randn((n, C, H, W), generator=g).cuda()
w_in = torch.randn((Hd, C + 39), generator=g).cuda()
b_in = torch.randn((Hd,), generator=g).cuda()
dG = torch.randn((n, H, W, Hd), generator=g).cuda()
G = torch.empty((n, H, W, Hd), device="cuda")
d_feat = torch.empty_like(feat)
d_wf = torch.zeros((Hd, C), device="cuda")
d_b = torch.zeros((Hd,), device="cuda")
s = torch.cuda.current_stream().cuda_stream


def fwd():
    _lib.check(lib.bts_fold_features(feat.data_ptr(), n, C, H, W, w_in.data_ptr(), C + 39, b_in.data_ptr(), Hd, None, G.data_ptr(), None, s), "fold")


def bwd():
    _lib.check(lib.bts_fold_features_bwd(dG.data_ptr(), feat.data_ptr(), n, C, H, W, w_in.data_ptr(), C + 39, Hd, d_feat.data_ptr(),
                                         d_wf.data_ptr(), d_b.data_ptr(), s), "fold_bwd")


def timeit(fn, steps):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / steps


steps = int(os.environ.get("STEPS", 20))
px = n * H * W
ms_f, ms_b = timeit(fwd, steps), timeit(bwd, steps)
print(json.dumps({"shape": f"n={n} C={C} Hd={Hd} {H}x{W}", "fold_fwd_ms": ms_f, "fold_fwd_GBps": px * (C + Hd) * 4 / ms_f / 1e6,
                  "fold_bwd_ms": ms_b, "fold_bwd_GBps": px * (2 * C + Hd) * 4 / ms_b / 1e6, "timeouts": lib.bts_debug_tc_timeouts()}))
