"""GPU: the CUDA render path (through the drop-in classes -> ctypes -> C ABI) against
(a) the committed fixtures produced by the UNMODIFIED reference and (b) the oracle at the BASELINE sizes."""
import pytest
import torch

from behindthescenes_b200 import _lib
from behindthescenes_b200 import synthetic as syn
from oracle import bts_oracle as orc
from tests import parity_util as pu
from tests.golden_io import golden_names, load_golden, mlp_params

pytestmark = pytest.mark.gpu
ALL = dict(want_weights=True, want_alphas=True, want_z_samps=True, want_rgb_samps=True)


@pytest.mark.parametrize("name", golden_names())
def test_render_matches_reference_fixture(name):
    g = load_golden(name)
    wl = pu.workload_from_golden(g)
    _, _, got = pu.cuda_render(wl, state_dict=g.sd, **ALL)
    stats = pu.compare_render(wl, got, g.out, label=name)
    print(name, stats)


@pytest.mark.parametrize("name", golden_names())
def test_point_query_matches_reference_fixture(name):
    g = load_golden(name)
    wl = pu.workload_from_golden(g)
    net, _ = syn.build_model(wl, "cuda", state_dict=g.sd)
    with torch.no_grad():
        net.encode(wl.images.cuda(), wl.Ks.cuda(), wl.poses_c2w.cuda(), ids_encoder=wl.ids_encoder, ids_render=wl.ids_render)
        pts = g.query["pts"].cuda()
        rgb, inv, sig = (t.cpu() for t in net(pts))
        d_rgb, d_inv, d_sig = (t.cpu() for t in net(pts, only_density=True))
    assert rgb.shape == g.query["q_rgb"].shape and inv.shape == g.query["q_invalid"].shape and sig.shape == g.query["q_sigma"].shape
    assert d_rgb.shape == g.query["d_rgb"].shape and d_inv.shape == g.query["d_invalid"].shape
    same = (inv == g.query["q_invalid"]).all(-1)
    assert same.float().mean().item() >= 0.98
    assert torch.allclose(sig[same], g.query["q_sigma"][same], rtol=1e-4, atol=1e-5)
    assert (rgb[same] - g.query["q_rgb"][same]).abs().max().item() <= 1e-5
    assert d_rgb.abs().max().item() == 0
    assert torch.allclose(d_sig[same], g.query["d_sigma"][same], rtol=1e-4, atol=1e-5)
    assert torch.equal(d_sig, sig)


def _last_kernel():
    return _lib.load().bts_last_kernel().decode()


def _handed_over(net):
    """8-ray groups the band kernel passed on to the gather kernel in the last render (scratch[0])."""
    return int(net._band_work[0].item())


def test_tensor_core_kernels_are_dispatched_and_match_general_kernel():
    """cfg2 takes the band kernel (exact tier; every tile of the stereo / same-view geometry fits the band), with
    ``band=False`` the gather kernel; both must agree with the fp32 CUDA-core kernel and the oracle, and no bounded
    barrier wait may expire."""
    wl = syn.cfg2_kitti360_eval(H=96, W=320, stereo=True, lin_out_bias=-4.0, seed=31, ray_subsample=20001)
    sd_net, _ = syn.build_model(wl, "cpu")
    sd = {k: v.clone() for k, v in sd_net.state_dict().items() if not k.startswith("encoder.")}
    net, _, band = pu.cuda_render(wl, state_dict=sd, precision=0, **ALL)
    assert _last_kernel() == "band"
    print("groups handed to the gather kernel:", _handed_over(net), "of", (wl.rays.shape[1] + 7) // 8)
    _, _, tc = pu.cuda_render(wl, state_dict=sd, precision=0, band=False, **ALL)
    assert _last_kernel() == "tc"
    _, _, gen = pu.cuda_render(wl, state_dict=sd, precision=2, **ALL)
    assert _last_kernel() == "general"
    assert _lib.load().bts_debug_tc_timeouts() == 0
    ref = pu.oracle_render(wl, sd, **ALL)
    print("band vs oracle", pu.compare_render(wl, band, ref, label="band"))
    print("tc vs oracle", pu.compare_render(wl, tc, ref, label="tc"))
    print("general vs oracle", pu.compare_render(wl, gen, ref, label="general"))
    print("band vs tc", pu.compare_render(wl, band, tc, label="band-vs-tc"))
    print("tc vs general", pu.compare_render(wl, tc, gen, label="tc-vs-general"))


def test_band_kernel_covers_the_whole_frame_of_the_bench_geometry():
    """Full consecutive-ray frame (the layout the band kernel is built for): encoder-view rays + stereo rays, no group
    handed over, outputs within the exact tier of the oracle; in-kernel jitter is the same stream in both kernels."""
    wl = syn.cfg2_kitti360_eval(H=64, W=208, stereo=True, lin_out_bias=-4.0, seed=33)
    sd_net, _ = syn.build_model(wl, "cpu")
    sd = {k: v.clone() for k, v in sd_net.state_dict().items() if not k.startswith("encoder.")}
    wants = dict(want_weights=True, want_alphas=True)
    net, _, band = pu.cuda_render(wl, state_dict=sd, **wants)
    assert _last_kernel() == "band" and _handed_over(net) == 0
    ref = pu.oracle_render(wl, sd, **wants)
    print("band vs oracle", pu.compare_render(wl, band, ref, label="band-frame"))
    _, _, b2 = pu.cuda_render(wl, state_dict=sd, use_u=False, **wants)
    _, _, t2 = pu.cuda_render(wl, state_dict=sd, use_u=False, band=False, **wants)
    print("band vs tc, in-kernel jitter", pu.compare_render(wl, b2, t2, label="band-vs-tc-seeded"))


def test_band_kernel_full_width_rows_fit_the_band():
    """At the full 640-pixel width the stereo footprint of a tile reaches 30 pixels now and then (6 tiles for this seed,
    simulated on the CPU from the injected jitter); with the three shifted copies of G16 every such span still fits 4
    blocks: nothing is handed over, everything matches the oracle."""
    wl = syn.cfg2_kitti360_eval(H=16, W=640, K=64, stereo=True, learn_empty=True, lin_out_bias=-4.0, seed=35)
    sd_net, _ = syn.build_model(wl, "cpu")
    sd = {k: v.clone() for k, v in sd_net.state_dict().items() if not k.startswith("encoder.")}
    net, _, got = pu.cuda_render(wl, state_dict=sd, **ALL)
    assert _last_kernel() == "band" and _handed_over(net) == 0
    ref = pu.oracle_render(wl, sd, **ALL)
    print("full-width band vs oracle", pu.compare_render(wl, got, ref, label="full-width"))


def test_band_kernel_hands_steep_epipolar_groups_to_the_gather_kernel():
    """Rays of a camera displaced vertically and rolled: epipolar lines cross many rows, most tiles do not fit the band
    -> the groups go through the work list and the gather kernel; results must be the oracle's either way."""
    poses = [syn._pose(), syn._pose(tx=0.2, ty=0.45, tz=-0.3, yaw_deg=2.0, pitch_deg=4.0)]
    wl = syn.make_workload("steep", H=64, W=208, poses=poses, model_conf=syn.kitti_model_conf(True),
                           renderer_conf=syn.kitti_renderer_conf(64), ids_encoder=[0], ids_render=[0, 1], ray_views=[0, 1],
                           seed=34, lin_out_bias=-4.0)
    sd_net, _ = syn.build_model(wl, "cpu")
    sd = {k: v.clone() for k, v in sd_net.state_dict().items() if not k.startswith("encoder.")}
    net, _, got = pu.cuda_render(wl, state_dict=sd, **ALL)
    assert _last_kernel() == "band"
    n_groups = (wl.rays.shape[1] + 7) // 8
    handed = _handed_over(net)
    print("groups handed to the gather kernel:", handed, "of", n_groups)
    assert 0 < handed < n_groups                # view-0 rays fit, most rays of the displaced camera do not
    ref = pu.oracle_render(wl, sd, **ALL)
    print("mixed band/gather vs oracle", pu.compare_render(wl, got, ref, label="steep"))
    assert _lib.load().bts_debug_tc_timeouts() == 0


def test_fast_tier_within_its_stated_tolerance():
    """precision=1 (single-pass fp16/tf32 operands): depth abs_rel <= 1e-3 (SURVEY.md §8d fast tier), in practice ~1e-5."""
    wl = syn.cfg2_kitti360_eval(H=96, W=320, stereo=True, lin_out_bias=-4.0, seed=32, ray_subsample=20000)
    sd_net, _ = syn.build_model(wl, "cpu")
    sd = {k: v.clone() for k, v in sd_net.state_dict().items() if not k.startswith("encoder.")}
    _, _, fast = pu.cuda_render(wl, state_dict=sd, precision=1, want_weights=True)
    assert _last_kernel() == "tc"
    ref = pu.oracle_render(wl, sd, want_weights=True)
    same = (fast["invalid"] == ref["invalid"]).flatten(2).all(-1)
    rel = ((fast["depth"] - ref["depth"]).abs() / ref["depth"])[same]
    print("fast tier: abs_rel", rel.mean().item(), "max", rel.max().item(), "weights", (fast["weights"] - ref["weights"])[same].abs().max().item())
    assert rel.mean().item() <= 1e-3 and (fast["weights"] - ref["weights"])[same].abs().max().item() <= 5e-3


def _full_size_case(wl, wants, lin_out_bias):
    sd_net, _ = syn.build_model(wl, "cpu")
    sd = {k: v.clone() for k, v in sd_net.state_dict().items() if not k.startswith("encoder.")}
    _, _, got = pu.cuda_render(wl, state_dict=sd, **wants)
    ref = pu.oracle_render(wl, sd, **wants)
    return pu.compare_render(wl, got, ref, label=wl.name)


def test_cfg2_full_frame_vs_oracle():
    """BASELINE config 2 at its full size (192x640, K=64, one view of rays), evaluator outputs (weights, alphas)."""
    wl = syn.cfg2_kitti360_eval(stereo=False, lin_out_bias=-4.0, seed=11)
    stats = _full_size_case(wl, dict(want_weights=True, want_alphas=True), -4.0)
    print(stats)


def test_cfg2_stereo_rays_vs_oracle():
    wl = syn.cfg2_kitti360_eval(H=96, W=320, stereo=True, lin_out_bias=0.0, seed=12)
    print(_full_size_case(wl, dict(want_weights=True, want_alphas=True, want_z_samps=True), 0.0))


def test_cfg3_training_outputs_vs_oracle():
    wl = syn.cfg3_kitti_raw_train(H=92, W=306, n=2, lin_out_bias=-4.0, seed=13, ray_subsample=8192)
    print(_full_size_case(wl, ALL, -4.0))


def test_cfg4_re10k_vs_oracle():
    wl = syn.cfg4_re10k_nvs(H=128, W=192, K=128, lin_out_bias=-4.0, seed=14, ray_subsample=8192)
    print(_full_size_case(wl, ALL, -4.0))


def test_stress_width128_K256_vs_oracle():
    """BASELINE config 5 shape (Hd=128, K=256) on a ray subset."""
    wl = syn.cfg2_kitti360_eval(H=48, W=160, K=256, stereo=True, learn_empty=False, lin_out_bias=-4.0, seed=15, ray_subsample=2048)
    wl.model_conf["mlp_coarse"]["d_hidden"] = 128
    print(_full_size_case(wl, dict(want_weights=True), -4.0))
    assert _last_kernel() == "tc"                       # the 128-wide variant of the tensor-core kernel


def test_ragged_K_and_empty_rays():
    wl = syn.cfg2_kitti360_eval(H=16, W=48, K=48, stereo=True, lin_out_bias=-3.0, seed=16, ray_subsample=77)
    print(_full_size_case(wl, ALL, -3.0))
    wl.renderer_conf["n_coarse"] = 5
    wl.u = wl.u[:, :5].contiguous()
    print(_full_size_case(wl, ALL, -3.0))
    net, wrapped = syn.build_model(wl, "cuda")
    rgb, depth = wrapped(torch.zeros((0, 10, 8), device="cuda"))          # nerf.py:20-24
    assert rgb.shape == (0, 3) and depth.shape == (0,)


def test_size_independent_properties_full_size():
    """Properties that hold at any size (checked at BASELINE config 2 with the stereo view = 245 760 rays, in-kernel
    jitter): alpha in [0,1], weights sum to 1 under hard_alpha_cap, depth within [near, far], sorted samples,
    0/1 flags, determinism for a fixed seed, colour linearity (rgb scales with the images)."""
    wl = syn.cfg2_kitti360_eval(stereo=True, lin_out_bias=-4.0, seed=21)
    net, wrapped = syn.build_model(wl, "cuda")
    dev = "cuda"
    with torch.no_grad():
        net.encode(wl.images.to(dev), wl.Ks.to(dev), wl.poses_c2w.to(dev), ids_encoder=wl.ids_encoder, ids_render=wl.ids_render)
        rays = wl.rays.to(dev)
        torch.manual_seed(5)
        a = wrapped(rays, want_weights=True, want_alphas=True, want_z_samps=True)["coarse"]
        torch.manual_seed(5)
        b = wrapped(rays, want_weights=True, want_alphas=True, want_z_samps=True)["coarse"]
        c = wrapped(rays, want_z_samps=True)["coarse"]
    for k in a:
        assert torch.equal(a[k], b[k]), k
    assert not torch.equal(a["z_samps"], c["z_samps"])        # a new call draws new jitter
    z = a["z_samps"]
    assert (z[..., 1:] > z[..., :-1]).all() and (z >= 3.0 - 1e-4).all() and (z <= 80.0 + 1e-3).all()
    assert ((a["alphas"] >= 0) & (a["alphas"] <= 1)).all()
    assert (a["weights"] >= 0).all()
    assert (a["weights"].sum(-1) - 1).abs().max().item() <= 2e-5
    assert ((a["depth"] >= 3.0 - 1e-3) & (a["depth"] <= 80.0 + 1e-2)).all()
    inv = a["invalid"]
    assert ((inv == 0) | (inv == 1)).all()
    assert (a["rgb"] >= -1e-6).all() and (a["rgb"] <= 1 + 1e-5).all()
    # the bin structure of stratified sampling: sample k lies in disparity bin k
    disp = 1 / z
    lo, hi = 1 / 80.0, 1 / 3.0
    s = (hi - disp) / (hi - lo) * 64
    kk = torch.arange(64, device=dev).view(1, 1, -1)
    assert ((s >= kk - 1e-3) & (s <= kk + 1 + 1e-3)).all()
    # linearity in the colour images: halving the images halves rgb (same jitter)
    with torch.no_grad():
        net.encode(wl.images.to(dev), wl.Ks.to(dev), wl.poses_c2w.to(dev), ids_encoder=wl.ids_encoder,
                   ids_render=wl.ids_render, images_alt=(wl.images.to(dev) * 0.5 + 0.5) * 0.5)
        torch.manual_seed(5)
        h = wrapped(rays)["coarse"]
    assert torch.allclose(h["rgb"], a["rgb"] * 0.5, atol=2e-6)
    assert torch.equal(h["depth"], a["depth"])


def test_errors_are_loud():
    wl = syn.cfg2_kitti360_eval(H=16, W=48, ray_subsample=32)
    net, wrapped = syn.build_model(wl, "cuda")
    with pytest.raises(RuntimeError):
        wrapped(wl.rays.cuda())                                  # encode() not called
    net.encode(wl.images.cuda(), wl.Ks.cuda(), wl.poses_c2w.cuda(), ids_encoder=[0], ids_render=[0])
    with pytest.raises(_lib.BtsCudaError):
        wrapped(wl.rays)                                         # CPU rays: no CPU fallback
    wrapped.renderer.n_fine = 16
    with pytest.raises(NotImplementedError):
        wrapped(wl.rays.cuda())
    wrapped.renderer.n_fine = 0
    # several encoder views: the features stand-in must return one map per encoder view
    wl2 = syn.cfg2_kitti360_eval(H=16, W=48, ray_subsample=32)
    wl2.feats = torch.cat((wl2.feats, wl2.feats), 0)
    net2, wrapped2 = syn.build_model(wl2, "cuda")
    net2.encode(wl2.images.cuda(), wl2.Ks.cuda(), wl2.poses_c2w.cuda(), ids_encoder=[0, 1], ids_render=[0])
    with pytest.raises(NotImplementedError):
        wrapped2(wl2.rays.cuda())


def test_non_default_stream_and_odd_frame_size():
    """The ABI launches on torch's *current* stream (SURVEY.md §8b "Threading / streams"); KITTI-Raw's full-resolution
    frame is 370x1226 (not a multiple of anything): 116 MB folded map, rays from two views."""
    wl = syn.cfg3_kitti_raw_train(H=370, W=1226, n=1, lin_out_bias=-4.0, seed=71, ray_subsample=6000)
    sd_net, _ = syn.build_model(wl, "cpu")
    sd = {k: v.clone() for k, v in sd_net.state_dict().items() if not k.startswith("encoder.")}
    net, wrapped = syn.build_model(wl, "cuda", state_dict=sd)
    side = torch.cuda.Stream()
    with torch.no_grad(), torch.cuda.stream(side):
        net.encode(wl.images.cuda(), wl.Ks.cuda(), wl.poses_c2w.cuda(), ids_encoder=wl.ids_encoder, ids_render=wl.ids_render)
        out = wrapped(wl.rays.cuda(), u=wl.u.cuda(), **ALL)["coarse"]
    side.synchronize()
    got = {k: v.cpu() for k, v in out.items()}
    ref = pu.oracle_render(wl, sd, **ALL)
    print(pu.compare_render(wl, got, ref, label="370x1226 on a side stream"))


def test_batched_scene_n16_patch_rays():
    """Training-shaped batch: n=16 scenes in one launch (per-ray batch index), 2 colour views."""
    wl = syn.cfg3_kitti_raw_train(H=24, W=80, n=16, lin_out_bias=-4.0, seed=72, ray_subsample=512)
    print(_full_size_case(wl, dict(want_weights=True, want_alphas=True, want_rgb_samps=True), -4.0))


def test_cfg4_full_size_vs_oracle():
    """BASELINE config 4 at its full size: RealEstate10K shape 256x384 (98 304 rays of a novel view), K=128, C=32,
    d_hidden=32 + one ResnetFC block, distance code, 2 colour views."""
    wl = syn.cfg4_re10k_nvs(H=256, W=384, K=128, lin_out_bias=-4.0, seed=17)
    print(_full_size_case(wl, dict(want_weights=True, want_alphas=True), -4.0))
    assert _last_kernel() == "tc"          # d_hidden 32 + one ResnetFC block: the block runs as two more MMA rounds


def test_cfg5_slice_one_million_rays():
    """BASELINE config 5 (stress: K=256, d_hidden=128, random rays) on a 2^20-ray slice: size-independent properties on
    every ray (weights sum to 1 under the alpha cap, sorted samples inside [near, far], depth in range, finite colours) and
    the oracle on a 2048-ray subset of the same launch (in-kernel jitter: z_samps of the launch are fed to the oracle)."""
    wl = syn.cfg5_stress(R=1 << 20, K=256, d_hidden=128, seed=5)
    sd_net, _ = syn.build_model(wl, "cpu")
    sd = {k: v.clone() for k, v in sd_net.state_dict().items() if not k.startswith("encoder.")}
    net, wrapped = syn.build_model(wl, "cuda", state_dict=sd)
    with torch.no_grad():
        net.encode(wl.images.cuda(), wl.Ks.cuda(), wl.poses_c2w.cuda(), ids_encoder=wl.ids_encoder, ids_render=wl.ids_render)
        rays = wl.rays.cuda()
        out = wrapped(rays, want_weights=True, want_z_samps=True, want_invalid=False)["coarse"]
    assert _last_kernel() == "tc"
    z, w, d = out["z_samps"], out["weights"], out["depth"]
    assert torch.isfinite(out["rgb"]).all() and torch.isfinite(d).all()
    assert (z[..., 1:] > z[..., :-1]).all() and (z >= 3.0 - 1e-4).all() and (z <= 80.0 + 1e-3).all()
    assert (w >= 0).all() and (w.sum(-1) - 1).abs().max().item() <= 5e-5
    assert ((d >= 3.0 - 1e-3) & (d <= 80.0 + 1e-2)).all()
    # oracle on a subset, with the jitter of this launch recovered from its depths (u is what maps bin k to z)
    n_s = 2048
    sub = syn.Workload(**{**wl.__dict__, "rays": wl.rays[:, :n_s].contiguous()})
    disp = 1.0 / z[0, :n_s].cpu().double()
    lo, hi = 1.0 / 80.0, 1.0 / 3.0
    s = (hi - disp) / (hi - lo)                                   # nerf.py:117-120 inverted: s = z_steps + u / K
    sub.u = ((s - torch.arange(256, dtype=torch.float64) / 256) * 256).clamp(0, 1).float()
    with torch.no_grad():
        got = wrapped(rays[:, :n_s].contiguous(), u=sub.u.cuda(), want_weights=True)["coarse"]
    got = {k: v.cpu() for k, v in got.items()}
    ref = pu.oracle_render(sub, sd, want_weights=True)
    print(pu.compare_render(sub, got, ref, label="cfg5 slice"))


@pytest.mark.parametrize("name,kernel", [("cfg4_small", "tc"), ("dist_block1", "tc"), ("cfg2_small_bias4", "band"), ("cfg3_small_n2", "tc")])
def test_fixture_kernel_dispatch(name, kernel):
    """Which kernel family renders which reference fixture: the RealEstate10K field (d_hidden 32, one ResnetFC block, distance
    code, K = 128) and the 64-wide block variant run on the tensor-core gather kernel (block = two more MMA rounds), single-frame
    KITTI scenes on the band kernel, batched scenes on the gather kernel.  Fast tier of the block variants within its tolerance."""
    g = load_golden(name)
    wl = pu.workload_from_golden(g)
    _, _, got = pu.cuda_render(wl, state_dict=g.sd, **ALL)
    assert _last_kernel() == kernel
    print(name, pu.compare_render(wl, got, g.out, label=name))
    if kernel == "tc":
        _, _, fast = pu.cuda_render(wl, state_dict=g.sd, precision=1, want_weights=True)
        assert _last_kernel() == "tc"
        same = (fast["invalid"] == g.out["invalid"]).flatten(2).all(-1)
        rel = ((fast["depth"] - g.out["depth"]).abs() / g.out["depth"])[same]
        assert rel.mean().item() <= 1e-3, rel.mean().item()
    assert _lib.load().bts_debug_tc_timeouts() == 0
