"""GPU: bts_fold_features (called through the raw C ABI with device pointers) -- TMA-fed tcgen05 path vs the CUDA-core
kernel vs an fp64 reference of G = W_in[:, :C] F + b_in."""
import os

import pytest
import torch

from behindthescenes_b200 import _lib

pytestmark = pytest.mark.gpu


def _fold(feat, w_in, b_in, Hd, empty=None, force_simt=False):
    lib = _lib.load()
    n, C, H, W = feat.shape
    G = torch.empty((n, H, W, Hd), device="cuda")
    ge = torch.empty((Hd,), device="cuda") if empty is not None else None
    old = os.environ.get("BTS_FOLD_SIMT")
    os.environ["BTS_FOLD_SIMT"] = "1" if force_simt else "0"
    try:
        rc = lib.bts_fold_features(feat.data_ptr(), n, C, H, W, w_in.data_ptr(), w_in.shape[1], b_in.data_ptr(), Hd,
                                   None if empty is None else empty.data_ptr(), G.data_ptr(),
                                   None if ge is None else ge.data_ptr(), torch.cuda.current_stream().cuda_stream)
    finally:
        if old is None:
            os.environ.pop("BTS_FOLD_SIMT", None)
        else:
            os.environ["BTS_FOLD_SIMT"] = old
    _lib.check(rc, "bts_fold_features")
    torch.cuda.synchronize()
    return G, ge


@pytest.mark.parametrize("n,C,Hd,H,W", [(2, 64, 64, 192, 640), (1, 32, 32, 24, 36), (3, 64, 128, 13, 20), (1, 64, 32, 7, 12),
                                        (1, 64, 64, 13, 21), (1, 48, 64, 16, 16)])
def test_fold_matches_fp64_and_both_kernels_agree(n, C, Hd, H, W):
    g = torch.Generator().manual_seed(n * 1000 + C + Hd + H)
    feat = (torch.randn((n, C, H, W), generator=g) * 3).cuda()
    w_in = torch.randn((Hd, C + 39), generator=g).cuda()
    b_in = torch.randn((Hd,), generator=g).cuda()
    empty = torch.randn((C,), generator=g).cuda()
    ref = torch.einsum("kc,nchw->nhwk", w_in[:, :C].double(), feat.double()) + b_in.double()
    scale = ref.abs().max().item()
    G_tc, ge_tc = _fold(feat, w_in, b_in, Hd, empty)                  # tensor-core path where the frame qualifies
    G_si, ge_si = _fold(feat, w_in, b_in, Hd, empty, force_simt=True)
    assert _lib.load().bts_debug_tc_timeouts() == 0
    for G in (G_tc, G_si):
        assert torch.isfinite(G).all()
        assert (G.double() - ref).abs().max().item() <= 3e-6 * scale
    assert (G_tc - G_si).abs().max().item() <= 3e-6 * scale
    ge_ref = w_in[:, :C].double() @ empty.double() + b_in.double()
    assert (ge_tc.double() - ge_ref).abs().max().item() <= 1e-5 and torch.equal(ge_tc, ge_si)


def test_fold_of_values_that_stress_the_tf32_split():
    """Large dynamic range, exact powers of two and values whose low mantissa bits are all set."""
    C, Hd, H, W = 64, 64, 8, 16
    feat = torch.zeros((1, C, H, W))
    feat[0, :, 0, :] = torch.tensor([1.0 + 2.0 ** -23, 65504.0, 1e-20, -3.999999, 0.1, 1e6, -1e-6, 7.0]).repeat(2)[:W].view(1, W)
    feat[0, :, 1:, :] = torch.randn((C, H - 1, W)) * torch.logspace(-6, 6, C).view(C, 1, 1)
    w_in = torch.randn((Hd, C + 39))
    b_in = torch.zeros(Hd)
    ref = torch.einsum("kc,nchw->nhwk", w_in[:, :C].double(), feat.double())
    G, _ = _fold(feat.cuda(), w_in.cuda(), b_in.cuda(), Hd)
    err = (G.double().cpu() - ref).abs()
    bound = 4e-6 * torch.einsum("kc,nchw->nhwk", w_in[:, :C].abs().double(), feat.abs().double()) + 1e-30
    assert (err <= bound).all()


def _fold_bwd(dG, feat, w_in, Hd, want_df=True):
    lib = _lib.load()
    n, C, H, W = feat.shape
    d_feat = torch.full_like(feat, float("nan")) if want_df else None
    d_wf = torch.zeros((Hd, C), device="cuda")
    d_b = torch.zeros((Hd,), device="cuda")
    rc = lib.bts_fold_features_bwd(dG.data_ptr(), feat.data_ptr(), n, C, H, W, w_in.data_ptr(), w_in.shape[1], Hd,
                                   None if d_feat is None else d_feat.data_ptr(), d_wf.data_ptr(), d_b.data_ptr(),
                                   torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    return rc, d_feat, d_wf, d_b


@pytest.mark.parametrize("n,C,Hd,H,W", [(2, 64, 64, 192, 640), (3, 64, 64, 13, 20), (1, 64, 32, 25, 36), (2, 32, 64, 8, 16),
                                        (1, 32, 32, 3, 4), (5, 64, 64, 64, 157 * 4)])
def test_fold_backward_matches_fp64(n, C, Hd, H, W):
    """bts_fold_features_bwd (3xTF32 tcgen05, hi/lo halves stacked along M) vs fp64 einsums, incl. ragged last tiles
    (H*W % 64 != 0), more tiles than SMs and fewer tiles than SMs."""
    g = torch.Generator().manual_seed(n * 77 + C + Hd + H)
    feat = (torch.randn((n, C, H, W), generator=g) * 2).cuda()
    w_in = torch.randn((Hd, C + 39), generator=g).cuda()
    dG = torch.randn((n, H, W, Hd), generator=g).cuda()
    dG[:, ::3] = 0                                                  # the real dG is sparse: rows no ray touched
    rc, d_feat, d_wf, d_b = _fold_bwd(dG, feat, w_in, Hd)
    assert rc == 0, _lib.load().bts_last_error()
    assert _lib.load().bts_debug_tc_timeouts() == 0
    W_F = w_in[:, :C].double()
    ref_df = torch.einsum("nhwk,kc->nchw", dG.double(), W_F)
    ref_wf = torch.einsum("nhwk,nchw->kc", dG.double(), feat.double())
    ref_b = dG.double().sum((0, 1, 2))
    assert torch.isfinite(d_feat).all()
    assert (d_feat.double() - ref_df).abs().max().item() <= 3e-6 * ref_df.abs().max().item()
    # the K = n*H*W reductions: error relative to the sum of magnitudes (fp32 accumulation of up to 2M terms)
    mag_wf = torch.einsum("nhwk,nchw->kc", dG.abs().double(), feat.abs().double())
    assert ((d_wf.double() - ref_wf).abs() <= 2e-6 * mag_wf + 1e-30).all()
    assert ((d_b.double() - ref_b).abs() <= 2e-6 * dG.abs().double().sum((0, 1, 2)) + 1e-30).all()
    # frozen encoder: no dF, same weight gradients (up to atomic ordering)
    rc2, none, d_wf2, d_b2 = _fold_bwd(dG, feat, w_in, Hd, want_df=False)
    assert rc2 == 0 and none is None
    assert ((d_wf2 - d_wf).abs().double() <= 1e-6 * mag_wf + 1e-30).all()


@pytest.mark.parametrize("n,C,Hd,H,W", [(1, 48, 64, 8, 8), (2, 64, 96, 7, 9), (1, 16, 32, 5, 13)])
def test_fold_backward_other_shapes_take_the_cuda_core_kernel(n, C, Hd, H, W):
    """Shapes outside the tensor-core fold backward (C not in {32, 64}, d_hidden 96, H*W % 4 != 0) run the CUDA-core kernel
    of fold.cu -- there is no torch fallback in the autograd node any more; same fp64 yardstick."""
    g = torch.Generator().manual_seed(n + C + Hd + H + W)
    feat = torch.randn((n, C, H, W), generator=g).cuda()
    w_in = torch.randn((Hd, C + 39), generator=g).cuda()
    dG = torch.randn((n, H, W, Hd), generator=g).cuda()
    rc, d_feat, d_wf, d_b = _fold_bwd(dG, feat, w_in, Hd)
    assert rc == 0, _lib.load().bts_last_error()
    W_F = w_in[:, :C].double()
    ref_df = torch.einsum("nhwk,kc->nchw", dG.double(), W_F)
    ref_wf = torch.einsum("nhwk,nchw->kc", dG.double(), feat.double())
    ref_b = dG.double().sum((0, 1, 2))
    assert (d_feat.double() - ref_df).abs().max().item() <= 3e-6 * ref_df.abs().max().item()
    mag_wf = torch.einsum("nhwk,nchw->kc", dG.abs().double(), feat.abs().double())
    assert ((d_wf.double() - ref_wf).abs() <= 2e-6 * mag_wf + 1e-30).all()
    assert ((d_b.double() - ref_b).abs() <= 2e-6 * dG.abs().double().sum((0, 1, 2)) + 1e-30).all()
    rc2, none, d_wf2, _ = _fold_bwd(dG, feat, w_in, Hd, want_df=False)
    assert rc2 == 0 and none is None and ((d_wf2 - d_wf).abs().double() <= 1e-6 * mag_wf + 1e-30).all()


def test_row_copy_loads_agree_with_tensor_map_loads(monkeypatch):
    """BTS_FOLD_TMA_ROWS=1 forces the per-row bulk copies (the path used when the driver has no tensor-map encoder):
    same shared-memory layout, so results are bit-identical for the forward and equal up to atomic order backward."""
    n, C, Hd, H, W = 2, 64, 64, 25, 36                             # ragged: 900 px = 14 tiles of 64 + 4
    g = torch.Generator().manual_seed(11)
    feat = torch.randn((n, C, H, W), generator=g).cuda()
    w_in = torch.randn((Hd, C + 39), generator=g).cuda()
    b_in = torch.randn((Hd,), generator=g).cuda()
    dG = torch.randn((n, H, W, Hd), generator=g).cuda()
    G_map, _ = _fold(feat, w_in, b_in, Hd)
    _, df_map, dwf_map, db_map = _fold_bwd(dG, feat, w_in, Hd)
    monkeypatch.setenv("BTS_FOLD_TMA_ROWS", "1")
    G_rows, _ = _fold(feat, w_in, b_in, Hd)
    rc, df_rows, dwf_rows, db_rows = _fold_bwd(dG, feat, w_in, Hd)
    assert rc == 0 and _lib.load().bts_debug_tc_timeouts() == 0
    assert torch.equal(G_map, G_rows) and torch.equal(df_map, df_rows)
    assert torch.allclose(dwf_map, dwf_rows, rtol=1e-5, atol=1e-4) and torch.allclose(db_map, db_rows, rtol=1e-5, atol=1e-4)


# ---- sparse fold (training): only the 64-pixel tiles a ray batch samples are folded / differentiated ---------------------

def _patch_workload(H=48, W=160, n=2, patches=6, seed=5):
    """cfg3-shaped training workload whose rays are real 8x8 pixel patches (PatchRaySampler), different per frame."""
    from behindthescenes_b200 import synthetic as syn
    wl = syn.cfg3_kitti_raw_train(H=H, W=W, n=n, lin_out_bias=-4.0, seed=seed)
    return syn.with_patch_rays(wl, H, W, 2, patches, seed=seed)


def test_sparse_fold_marks_and_folds_exactly_the_sampled_tiles():
    """bts_mark_tiles + bts_fold_features_sparse: marked tiles equal the dense fold bit for bit, a superset of the tiles the
    renderer's taps touch (checked through the render: sparse and dense outputs are identical), unmarked tiles untouched."""
    import ctypes as C
    from behindthescenes_b200 import synthetic as syn
    wl = _patch_workload(H=96, W=320, patches=3)
    net, wrapped = syn.build_model(wl, "cuda")
    with torch.no_grad():
        net.encode(wl.images.cuda(), wl.Ks.cuda(), wl.poses_c2w.cuda(), ids_encoder=wl.ids_encoder, ids_render=wl.ids_render)
        feats = net.grid_f_features[0]
        assert net._sparse_ok(feats)
        rays, u = wl.rays.cuda(), wl.u.cuda()
        dense = net._render_rays_raw(wrapped.renderer, rays, u=u, want_weights=True, want_alphas=True)
        Gd = net._scene_struct()[1][3].clone()
        sparse = net._render_rays_raw(wrapped.renderer, rays, u=u, want_weights=True, want_alphas=True, sparse=True)
        scene, keep = net._scene_struct(sparse=True)
        mask, Gs = scene._sparse["mask"], scene._sparse["G"]
        torch.cuda.synchronize()
    for k in dense:
        assert torch.equal(dense[k], sparse[k]), k
    n, H, W, Hd = Gd.shape
    per = (H * W + 63) // 64
    m = mask.view(n, per)
    assert set(m.unique().tolist()) <= {0, 2}
    frac = (m != 0).float().mean().item()
    assert 0.0 < frac < 0.8, frac
    tile_of_pixel = (torch.arange(H * W, device="cuda") // 64)
    folded = (m != 0)[:, tile_of_pixel].view(n, H, W)
    assert torch.equal(Gs[folded], Gd[folded])                 # same kernel, same operands: identical values
    print("tiles folded:", frac)


def test_sparse_training_step_gradients_equal_the_dense_path():
    """Autograd through the drop-in classes with the sparse fold (default) and with it switched off: outputs identical,
    gradients equal up to the order of the atomic accumulation."""
    from behindthescenes_b200 import synthetic as syn
    wl = _patch_workload(seed=9)
    res = {}
    for mode in (True, False):
        net, wrapped = syn.build_model(wl, "cuda")
        net.sparse_fold = mode
        net.encoder.feats.requires_grad_(True)
        net.encode(wl.images.cuda(), wl.Ks.cuda(), wl.poses_c2w.cuda(), ids_encoder=wl.ids_encoder, ids_render=wl.ids_render)
        out = wrapped(wl.rays.cuda(), u=wl.u.cuda(), want_weights=True, want_alphas=True)["coarse"]
        loss = out["depth"].mean() * 0.1 + (out["rgb"] ** 2).mean() + out["weights"].sum(-1).mean() * 0.01
        loss.backward()
        torch.cuda.synchronize()
        grads = {"feats": net.encoder.feats.grad.clone()}
        grads.update({k: p.grad.clone() for k, p in net.mlp_coarse.named_parameters()})
        res[mode] = (out["depth"].detach().clone(), out["rgb"].detach().clone(), grads)
    assert torch.equal(res[True][0], res[False][0]) and torch.equal(res[True][1], res[False][1])
    for k, gs in res[True][2].items():
        gd = res[False][2][k]
        assert torch.isfinite(gs).all(), k
        err = (gs - gd).abs().max().item() / max(gd.abs().max().item(), 1e-12)
        assert err <= 2e-5, (k, err)
    # untouched pixels of the feature gradient are exact zeros in both paths
    assert (res[True][2]["feats"] == 0).float().mean().item() > 0.1


def test_marked_tiles_agree_with_the_oracle_taps():
    """bts_mark_tiles vs the taps the oracle's sampling + projection + grid_sample arithmetic produce for the same batch
    (stratified depths nerf.py:103-123, projection models_bts.py:145-155, align_corners=False bilinear taps with border
    padding): every tile the oracle's taps fall into is marked, and what is marked beyond that is at most the handful of
    tiles where a tap sits within rounding distance of a pixel boundary."""
    import ctypes as C
    from behindthescenes_b200 import synthetic as syn
    from oracle import bts_oracle as orc
    H, W = 96, 320
    wl = _patch_workload(H=H, W=W, n=2, patches=3, seed=11)
    net, wrapped = syn.build_model(wl, "cuda")
    lib = _lib.load()
    with torch.no_grad():
        net.encode(wl.images.cuda(), wl.Ks.cuda(), wl.poses_c2w.cuda(), ids_encoder=wl.ids_encoder, ids_render=wl.ids_render)
        scene, keep = net._scene_struct(sparse=True)
        cfg, keep3 = net._cfg_struct(wrapped.renderer)
        rays, u = wl.rays.cuda().contiguous(), wl.u.cuda().contiguous()
        n, R, _ = rays.shape
        mask = scene._sparse["mask"]
        rc = lib.bts_mark_tiles(C.byref(scene), C.byref(cfg), rays.data_ptr(), n, R, u.data_ptr(), 0, mask.data_ptr(),
                                torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, "bts_mark_tiles")
        torch.cuda.synchronize()
    got = mask.view(n, -1).cpu() != 0
    # oracle: depths, world points, projection into the encoder view, taps
    K = wl.renderer_conf["n_coarse"]
    st = orc.encode_state(wl.images, wl.Ks, wl.poses_c2w, wl.feats, wl.ids_encoder, wl.ids_render)
    z = orc.stratified_depths(wl.rays.reshape(-1, 8), K, bool(wl.renderer_conf.get("lindisp", True)), wl.u).view(n, R, K)
    pts = wl.rays[:, :, None, :3] + z.unsqueeze(-1) * wl.rays[:, :, None, 3:6]
    uv, _, _, _ = orc._project(pts.reshape(n, R * K, 3), st.Kf, st.Mf)
    uv = uv[:, 0].double()
    ix = (((uv[..., 0] + 1) * W - 1) / 2).clamp(0, W - 1)
    iy = (((uv[..., 1] + 1) * H - 1) / 2).clamp(0, H - 1)
    want = torch.zeros_like(got)
    near = torch.zeros_like(got)                 # tiles reachable when a tap coordinate is within 1e-3 px of an integer
    for dx in (0, 1):
        for dy in (0, 1):
            x = (ix.floor().long() + dx).clamp(max=W - 1)
            y = (iy.floor().long() + dy).clamp(max=H - 1)
            t = (y * W + x) // 64
            for b in range(n):
                want[b, t[b]] = True
    for ex in (-1e-3, 1e-3):
        for ey in (-1e-3, 1e-3):
            fx, fy = (ix + ex).clamp(0, W - 1).floor().long(), (iy + ey).clamp(0, H - 1).floor().long()
            for dx in (0, 1):
                for dy in (0, 1):
                    t = ((fy + dy).clamp(max=H - 1) * W + (fx + dx).clamp(max=W - 1)) // 64
                    for b in range(n):
                        near[b, t[b]] = True
    missing = (want & ~got).sum().item()
    extra = (got & ~want & ~near).sum().item()
    print("tiles: oracle", int(want.sum()), "marked", int(got.sum()), "oracle-only", missing, "marked-only beyond rounding", extra)
    assert missing == 0 and extra == 0      # float32 (kernel) vs float64 (here) taps may differ only inside `near`
