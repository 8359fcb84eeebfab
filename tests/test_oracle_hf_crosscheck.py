"""Second opinion on the timm restatement: transformers' Dinov2WithRegistersModel / SiglipVisionModel are an
independent implementation of the same published architectures.  Reduced depth (3 blocks) at the real widths keeps
this in CPU-seconds; the per-block code path is what is being pinned.  The same comparison is frozen
into tests/golden/hf_towers_depth3.npz (tests/test_oracle_golden.py) so that it survives a transformers upgrade."""

import pytest
import torch

from bridgelang_b200.config import DINOV2_L14_REG4, SIGLIP_SO400M_14
from bridgelang_b200.weights import make_vit_state_dict
from oracle import vit_oracle

transformers = pytest.importorskip("transformers")
import hf_mapping  # noqa: E402  (tests/hf_mapping.py)

DEPTH = 3


def _pixels(seed=0, batch=1):
    return torch.randn(batch, 3, 224, 224, generator=torch.Generator().manual_seed(seed))


def test_dinov2_reg4_against_hf():
    cfg = DINOV2_L14_REG4.with_depth(DEPTH)
    sd = make_vit_state_dict(cfg, seed=11, init="stress")
    m = hf_mapping.build_hf_dinov2_reg4(sd, cfg, DEPTH)
    x = _pixels()
    ref = hf_mapping.hf_dinov2_penultimate(m, x)
    with torch.no_grad():
        got = vit_oracle.vit_intermediate(sd, cfg, x)
    assert got.shape == (1, 256, 1024)
    err = ((got - ref).abs().max() / ref.abs().max()).item()
    assert err < 1e-5, err


def test_siglip_against_hf():
    cfg = SIGLIP_SO400M_14.with_depth(DEPTH)
    sd = make_vit_state_dict(cfg, seed=12, init="stress")
    m = hf_mapping.build_hf_siglip(sd, cfg, DEPTH)
    x = _pixels(1)
    ref = hf_mapping.hf_siglip_penultimate(m, x)
    with torch.no_grad():
        got = vit_oracle.vit_intermediate(sd, cfg, x)
    assert got.shape == (1, 256, 1152)
    err = ((got - ref).abs().max() / ref.abs().max()).item()
    assert err < 1e-5, err


def test_gelu_choice_is_detectable():
    """erf vs tanh GELU must be visible to this cross-check (otherwise it pins nothing)."""
    cfg = SIGLIP_SO400M_14.with_depth(DEPTH)
    sd = make_vit_state_dict(cfg, seed=12, init="stress")
    x = _pixels(1)
    a = vit_oracle.vit_intermediate(sd, cfg, x)
    orig = torch.nn.functional.gelu
    try:
        torch.nn.functional.gelu = lambda t: orig(t, approximate="tanh")
        vit_oracle.F.gelu = torch.nn.functional.gelu
        b = vit_oracle.vit_intermediate(sd, cfg, x)
    finally:
        torch.nn.functional.gelu = orig
        vit_oracle.F.gelu = orig
    assert ((a - b).abs().max() / a.abs().max()).item() > 2e-5


def test_openai_clip_336_against_hf():
    """the CLIP tower of `dinoclip-vit-l-336px` (clip_vit.py:15-27) — class token first, pos_embed on every
    token, norm_pre, bias-free conv, quick-GELU — against transformers' CLIPVisionModel."""
    from bridgelang_b200.config import CLIP_L14_336
    cfg = CLIP_L14_336.with_depth(DEPTH)
    sd = make_vit_state_dict(cfg, seed=13, init="stress")
    assert "patch_embed.proj.bias" not in sd and "norm_pre.weight" in sd and sd["pos_embed"].shape == (1, 577, 1024)
    m = hf_mapping.build_hf_clip(sd, cfg, DEPTH)
    x = torch.randn(1, 3, 336, 336, generator=torch.Generator().manual_seed(2))
    ref = hf_mapping.hf_clip_penultimate(m, x)
    with torch.no_grad():
        got = vit_oracle.vit_intermediate(sd, cfg, x)
    assert got.shape == (1, 576, 1024)
    err = ((got - ref).abs().max() / ref.abs().max()).item()
    assert err < 1e-5, err


def test_384px_towers_against_hf():
    """`dinosiglip-vit-so-384px`: a 384 px frame gives 27 x 27 = 729 patches (the strided conv never reads the last 6
    pixel rows / columns) in both towers."""
    from bridgelang_b200.config import DINOV2_L14_REG4_384, SIGLIP_SO400M_14_384
    from transformers import Dinov2WithRegistersConfig, SiglipVisionConfig  # noqa: F401
    x = torch.randn(1, 3, 384, 384, generator=torch.Generator().manual_seed(4))
    for cfg0, build, run, seed in ((SIGLIP_SO400M_14_384, hf_mapping.build_hf_siglip, hf_mapping.hf_siglip_penultimate, 14),
                                   (DINOV2_L14_REG4_384, hf_mapping.build_hf_dinov2_reg4, hf_mapping.hf_dinov2_penultimate, 15)):
        cfg = cfg0.with_depth(DEPTH)
        sd = make_vit_state_dict(cfg, seed=seed, init="stress")
        with torch.no_grad():
            got = vit_oracle.vit_intermediate(sd, cfg, x)
        assert got.shape == (1, 729, cfg.dim)
        # the HF models take image_size = 378 (= 27·14): crop the frame to what the conv reads
        import hf_mapping as hm
        m = build(sd, _Cfg378(cfg), DEPTH)
        ref = run(m, x[:, :, :378, :378])
        err = ((got - ref).abs().max() / ref.abs().max()).item()
        assert err < 1e-5, (cfg.timm_id, err)


class _Cfg378:
    """view of a VitConfig whose HF twin is built at image_size 378 (hf_mapping reads .dim/.heads/.mlp_hidden/.img_size)"""
    def __init__(self, cfg):
        self._c = cfg
    def __getattr__(self, k):
        return 378 if k == "img_size" else getattr(self._c, k)
