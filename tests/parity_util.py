"""Shared helpers of the GPU parity tests: run the CUDA path on a fixture / workload and compare with the oracle."""
import torch

from behindthescenes_b200 import synthetic as syn
from oracle import bts_oracle as orc
from tests.golden_io import mlp_params

# ---- stated tolerances of the exact (fp32) tier, SURVEY.md §8d -------------------------------------------------------
DEPTH_ABS_REL_MEAN = 1e-5     # mean |d - d_ref| / d_ref over rays
DEPTH_REL_MAX = 1e-4          # worst ray
VALUE_MAX_ABS = 1e-4          # rgb, weights, alphas, rgb_samps
KNIFE_EDGE = 1e-5             # |u|,|v| within this of 1 (or z of 1e-3): the flag depends on ulp-level rounding


def workload_from_golden(g):
    K = g.renderer_conf["n_coarse"]
    return syn.Workload(name=g.name, model_conf=g.model_conf, renderer_conf=g.renderer_conf, images=g.images, Ks=g.Ks,
                        poses_c2w=g.poses_c2w, feats=g.feats, ids_encoder=g.ids_encoder, ids_render=g.ids_render,
                        rays=g.rays, u=g.u.reshape(-1, K))


def cuda_render(wl, state_dict=None, device="cuda", use_u=True, precision=0, band=True, **wants):
    """band=False keeps single-frame scenes on the gather kernel (render_tc.cu) instead of the band kernel."""
    net, wrapped = syn.build_model(wl, device, state_dict=state_dict)
    wrapped.renderer.precision = precision
    wrapped.renderer.band = band
    with torch.no_grad():
        net.encode(wl.images.to(device), wl.Ks.to(device), wl.poses_c2w.to(device), ids_encoder=wl.ids_encoder,
                   ids_render=wl.ids_render)
        net.set_scale(0)
        out = wrapped(wl.rays.to(device), u=(wl.u.to(device) if use_u else None), **wants)
    torch.cuda.synchronize()
    return net, wrapped, {k: v.cpu() for k, v in out["coarse"].items()}


def oracle_render(wl, sd, **wants):
    st = orc.encode_state(wl.images, wl.Ks, wl.poses_c2w, wl.feats, wl.ids_encoder, wl.ids_render)
    fcfg = orc.FieldConfig.from_model_conf(wl.model_conf)
    rcfg = orc.RenderConfig.from_renderer_conf(wl.renderer_conf)
    return orc.render(st, mlp_params(sd), fcfg, rcfg, wl.rays, u=wl.u, empty_feature=sd.get("empty_feature"), **wants)["coarse"]


def knife_edge_samples(wl, z_samps):
    """(n,R,K,Vc) bool: samples whose invalid flag is decided within KNIFE_EDGE of a frustum border."""
    st = orc.encode_state(wl.images, wl.Ks, wl.poses_c2w, wl.feats, wl.ids_encoder, wl.ids_render)
    n, R, K = z_samps.shape
    pts = (wl.rays[:, :, None, :3] + z_samps.unsqueeze(-1) * wl.rays[:, :, None, 3:6]).reshape(n, R * K, 3)

    def margin(Kmat, M):
        uv, z, _, _ = orc._project(pts, Kmat, M)
        m = torch.minimum((1 - uv.abs()).abs().amin(-1, keepdim=True), (z - orc.EPS).abs())
        return m[..., 0]                                   # (n,V,P)
    mf = margin(st.Kf, st.Mf)[:, 0]                        # (n,P)
    mc = margin(st.Kc, st.Mc)                              # (n,Vc,P)
    m = torch.minimum(mf.unsqueeze(1), mc)                 # (n,Vc,P)
    return (m <= KNIFE_EDGE).permute(0, 2, 1).reshape(n, R, K, -1)


def compare_render(wl, got, ref, label=""):
    """Assert the CUDA outputs `got` match the reference/oracle outputs `ref` at the exact-tier tolerances."""
    assert set(got) == set(ref), (sorted(got), sorted(ref))
    for k in ref:
        assert got[k].shape == ref[k].shape, (k, got[k].shape, ref[k].shape)
        assert torch.isfinite(got[k]).all(), k
    z = ref.get("z_samps")
    if z is not None:
        assert torch.equal(got["z_samps"], z), f"{label}: stratified depths are not bit-exact"
    # invalid: exact away from knife edges
    mism = got["invalid"] != ref["invalid"]
    bad_rays = mism.flatten(2).any(-1)                      # (n,R)
    if mism.any():
        assert z is not None or True
        zz = z if z is not None else oracle_z(wl)
        ke = knife_edge_samples(wl, zz)
        assert not (mism & ~ke).any(), f"{label}: invalid flag differs on {(mism & ~ke).sum().item()} non-knife-edge samples"
        assert bad_rays.float().mean().item() <= 0.02, f"{label}: too many knife-edge rays"
    ok = ~bad_rays
    stats = {}
    d, dr = got["depth"][ok], ref["depth"][ok]
    rel = (d - dr).abs() / dr.abs().clamp_min(1e-6)
    stats["depth_abs_rel"] = rel.mean().item()
    stats["depth_rel_max"] = rel.max().item()
    assert stats["depth_abs_rel"] <= DEPTH_ABS_REL_MEAN, (label, stats)
    assert stats["depth_rel_max"] <= DEPTH_REL_MAX, (label, stats)
    # Colour samples are a bilinear fetch at ix = (u+1)*W/2: a few-ulp difference in the projected u (different but
    # equally valid fp32 summation orders in the projection, e.g. MKL vs FMA chain) moves ix by ~2.5e-7*W pixels, and the
    # synthetic images are white noise (|gradient| ~ 1 per pixel) -> the colour tolerance scales with the image width.
    W_img = wl.images.shape[-1]
    colour_tol = max(VALUE_MAX_ABS, 2.5e-7 * W_img)
    for k in ("rgb", "weights", "alphas", "rgb_samps"):
        if k in ref:
            e = (got[k][ok] - ref[k][ok]).abs().max().item()
            stats[k] = e
            assert e <= (colour_tol if k in ("rgb", "rgb_samps") else VALUE_MAX_ABS), (label, k, e)
    return stats


def oracle_z(wl):
    rcfg = orc.RenderConfig.from_renderer_conf(wl.renderer_conf)
    n, R, _ = wl.rays.shape
    return orc.stratified_depths(wl.rays.reshape(-1, 8), rcfg.n_coarse, rcfg.lindisp, wl.u).view(n, R, -1)
