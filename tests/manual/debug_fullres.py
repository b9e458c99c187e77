import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from behindthescenes_b200 import synthetic as syn
from tests import parity_util as pu
ALL = dict(want_weights=True, want_alphas=True, want_z_samps=True, want_rgb_samps=True)
wl = syn.cfg3_kitti_raw_train(H=370, W=1226, n=1, lin_out_bias=-4.0, seed=71, ray_subsample=6000)
sd_net, _ = syn.build_model(wl, "cpu")
sd = {k: v.clone() for k, v in sd_net.state_dict().items() if not k.startswith("encoder.")}
ref = pu.oracle_render(wl, sd, **ALL)
for prec, stream in ((0, False), (2, False), (0, True)):
    net, wrapped = syn.build_model(wl, "cuda", state_dict=sd)
    wrapped.renderer.precision = prec
    side = torch.cuda.Stream() if stream else torch.cuda.current_stream()
    with torch.no_grad(), torch.cuda.stream(side):
        net.encode(wl.images.cuda(), wl.Ks.cuda(), wl.poses_c2w.cuda(), ids_encoder=wl.ids_encoder, ids_render=wl.ids_render)
        out = wrapped(wl.rays.cuda(), u=wl.u.cuda(), **ALL)["coarse"]
    torch.cuda.synchronize()
    got = {k: v.cpu() for k, v in out.items()}
    mism = (got["invalid"] != ref["invalid"]).flatten(2).any(-1)
    ok = ~mism
    print("precision", prec, "side stream", stream, "flag-flip rays", int(mism.sum()))
    for k in ("depth", "rgb", "weights", "alphas", "rgb_samps"):
        d = (got[k][ok] - ref[k][ok]).abs()
        print(f"   {k:10s} max {d.max().item():.3e} mean {d.mean().item():.3e}")
    d = (got["alphas"] - ref["alphas"]).abs()[0]
    top = torch.topk(d.flatten(), 5)
    K = d.shape[1]
    for v, i in zip(top.values.tolist(), top.indices.tolist()):
        r, k = divmod(i, K)
        print(f"   ray {r} k {k} dalpha {v:.2e} ours {got['alphas'][0,r,k]:.6f} ref {ref['alphas'][0,r,k]:.6f} z {got['z_samps'][0,r,k]:.3f} flip {bool(mism[0,r])}")
