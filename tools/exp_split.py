#!/usr/bin/env python
"""Timing split of the bench frame: rays of the encoder view (every sample of a ray projects into one 2x2 block) vs
rays of the stereo view (samples walk along a horizontal epipolar line, ~1 px per sample) vs the whole frame.
usage (GPU box): python tools/exp_split.py [out.json]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from behindthescenes_b200 import _lib, synthetic as syn  # noqa: E402


def timed(fn, n=20, warm=3, flush=None):
    for _ in range(warm):
        fn()
    evs = []
    for _ in range(n):
        if flush is not None:
            flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    return ts[len(ts) // 2]


def main():
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    wl = syn.cfg2_kitti360_eval(H=192, W=640, K=64, stereo=True, learn_empty=True, lin_out_bias=-4.0, seed=0)
    net, wrapped = syn.build_model(wl, dev)
    with torch.no_grad():
        net.encode(wl.images.to(dev), wl.Ks.to(dev), wl.poses_c2w.to(dev), ids_encoder=wl.ids_encoder, ids_render=wl.ids_render)
    rays = wl.rays.to(dev)
    n = rays.shape[1]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    res = {}
    wants = dict(want_weights=True, want_alphas=True)
    subsets = {"all": rays, "centre": rays[:, : n // 2].contiguous(), "stereo": rays[:, n // 2:].contiguous()}
    extra = [a for a in sys.argv[2:]]
    with torch.no_grad():
        for band in (True, False):
            wrapped.renderer.band = band
            wrapped.renderer.precision = 0
            for name, r in subsets.items():
                ms = timed(lambda: wrapped(r, **wants), flush=flush)
                key = f"{name}_{'band' if band else 'gather'}"
                res[key] = {"ms": ms, "rays": r.shape[1], "Mrays_s": r.shape[1] / ms / 1e3,
                            "kernel": _lib.load().bts_last_kernel().decode()}
                if band and getattr(net, "_band_work", None) is not None:
                    res[key]["groups_handed_over"] = int(net._band_work[0].item())
        wrapped.renderer.band = True
        for name, r in subsets.items():
            ms = timed(lambda: wrapped(r), flush=flush)
            res[f"{name}_band_minimal_outputs"] = {"ms": ms}
    for k in extra:
        res.setdefault("args", []).append(k)
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "exp_split.json")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    json.dump(res, open(out, "w"), indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
