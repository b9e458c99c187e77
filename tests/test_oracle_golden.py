"""CPU: the oracle port (oracle/bts_oracle.py) against the fixtures produced by the UNMODIFIED reference."""
import pytest
import torch

from oracle import bts_oracle as orc
from tests.golden_io import golden_names, load_golden, mlp_params

# fp32, same torch ops in the same order -> agreement is at rounding level; invalid must be exact.
ATOL = 2e-6


def _oracle_render(g, **wants):
    st = orc.encode_state(g.images, g.Ks, g.poses_c2w, g.feats, g.ids_encoder, g.ids_render)
    fcfg = orc.FieldConfig.from_model_conf(g.model_conf)
    rcfg = orc.RenderConfig.from_renderer_conf(g.renderer_conf)
    return orc.render(st, mlp_params(g.sd), fcfg, rcfg, g.rays, u=g.u, empty_feature=g.sd.get("empty_feature"),
                      **wants)["coarse"], st, fcfg


@pytest.mark.parametrize("name", golden_names())
def test_render_matches_reference(name):
    g = load_golden(name)
    out, _, _ = _oracle_render(g, want_weights=True, want_alphas=True, want_z_samps=True, want_rgb_samps=True)
    assert set(out) == set(g.out)
    assert torch.equal(out["z_samps"], g.out["z_samps"])
    assert torch.equal(out["invalid"], g.out["invalid"])
    for k in ("rgb", "weights", "alphas", "rgb_samps"):
        assert out[k].shape == g.out[k].shape
        assert (out[k] - g.out[k]).abs().max().item() <= ATOL, k
    rel = ((out["depth"] - g.out["depth"]).abs() / g.out["depth"].abs().clamp_min(1e-6)).max().item()
    assert rel <= 2e-6


@pytest.mark.parametrize("name", golden_names())
def test_point_query_matches_reference(name):
    g = load_golden(name)
    st = orc.encode_state(g.images, g.Ks, g.poses_c2w, g.feats, g.ids_encoder, g.ids_render)
    fcfg = orc.FieldConfig.from_model_conf(g.model_conf)
    ef = g.sd.get("empty_feature")
    rgb, inv, sig = orc.field_forward(st, mlp_params(g.sd), fcfg, g.query["pts"], ef)
    assert torch.equal(inv, g.query["q_invalid"])
    assert (rgb - g.query["q_rgb"]).abs().max().item() <= ATOL
    assert torch.allclose(sig, g.query["q_sigma"], rtol=2e-6, atol=ATOL)
    rgb, inv, sig = orc.field_forward(st, mlp_params(g.sd), fcfg, g.query["pts"], ef, only_density=True)
    assert rgb.shape == g.query["d_rgb"].shape and rgb.abs().max().item() == 0
    assert torch.equal(inv, g.query["d_invalid"])
    assert torch.allclose(sig, g.query["d_sigma"], rtol=2e-6, atol=ATOL)


def test_default_outputs_only():
    g = load_golden(golden_names()[0])
    out, _, _ = _oracle_render(g)
    assert set(out) == {"rgb", "depth", "invalid"}
