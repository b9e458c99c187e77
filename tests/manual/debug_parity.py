"""GPU debug: where do the CUDA path and the oracle differ most on the smoke workload? (run under gpurun)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from behindthescenes_b200 import synthetic as syn
from tests import parity_util as pu

wl = syn.cfg2_kitti360_eval(H=48, W=160, K=64, stereo=True, lin_out_bias=-4.0, seed=3, ray_subsample=4096)
net, _ = syn.build_model(wl, "cpu")
sd = {k: v.clone() for k, v in net.state_dict().items() if not k.startswith("encoder.")}
wants = dict(want_weights=True, want_alphas=True, want_z_samps=True)
ref = pu.oracle_render(wl, sd, **wants)
for prec in (0, 2):
    _, _, a = pu.cuda_render(wl, state_dict=sd, precision=prec, **wants)
    _, _, b = pu.cuda_render(wl, state_dict=sd, precision=prec, **wants)
    print("precision", prec, "deterministic:", all(torch.equal(a[k], b[k]) for k in a))
    d = (a["alphas"] - ref["alphas"]).abs()[0]
    top = torch.topk(d.flatten(), 8)
    K = d.shape[1]
    for v, i in zip(top.values.tolist(), top.indices.tolist()):
        r, k = divmod(i, K)
        print(f"  ray {r:5d} k {k:2d} |d alpha| {v:.2e} ours {a['alphas'][0, r, k]:.6f} ref {ref['alphas'][0, r, k]:.6f} z {a['z_samps'][0, r, k]:.4f} "
              f"inv ours {a['invalid'][0, r, k, 0]:.0f} ref {ref['invalid'][0, r, k, 0]:.0f} view {'0' if wl.rays[0, r, 0] == 0 else '1'}")
    ks = top.indices % K
    print("  k histogram of worst 200:", torch.bincount(torch.topk(d.flatten(), 200).indices % K, minlength=K).tolist())
