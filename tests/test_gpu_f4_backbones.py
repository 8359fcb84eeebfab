"""SURVEY §8f.4 on a B200: the other fused backbones of the reference registry (materialize.py:48-49) —
`dinosiglip-vit-so-384px` (729 patches; DINOv2 734 tokens hd 64, SigLIP 729 tokens hd 72) and `dinoclip-vit-l-336px`
(576 patches; OpenAI CLIP ViT-L/14-336: quick-GELU, norm_pre, bias-free conv, pos-embed with a class row) — against the
fp32 oracle (itself cross-checked against transformers' CLIPVisionModel / Siglip / Dinov2 at these sizes on the CPU)."""

import pytest
import torch
import torch.nn.functional as F

import bridgelang_b200 as blb
from bridgelang_b200 import ops
from bridgelang_b200.config import CLIP_L14_336, DINOV2_L14_REG4_336, DINOV2_L14_REG4_384, SIGLIP_SO400M_14_384
from bridgelang_b200.weights import make_projector_state_dict, make_vit_state_dict, normalize_for, synthetic_frames
from oracle import vit_oracle

pytestmark = pytest.mark.gpu
TOL = 2e-2


def _rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).abs().max() / b.abs().max()).item()


@pytest.mark.parametrize("B,T,H,hd", [(2, 577, 16, 64), (2, 734, 16, 64), (2, 729, 16, 72), (1, 321, 4, 72), (3, 1000, 2, 64)])
def test_streaming_attention_long_sequences(B, T, H, hd):
    g = torch.Generator(device="cuda").manual_seed(T)
    D = H * hd
    qkv = torch.randn(B * T, 3 * D, device="cuda", generator=g).bfloat16()
    out = ops.attention(qkv, B, T, H, hd)
    q, k, v = qkv.float().view(B, T, 3, H, hd).permute(2, 0, 3, 1, 4).unbind(0)
    ref = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B * T, D)
    assert _rel(out, ref) < 6e-3
    # scale sensitivity: the default scale is hd^-0.5
    bad = F.scaled_dot_product_attention(q, k, v, scale=1.0 / hd).transpose(1, 2).reshape(B * T, D)
    assert _rel(bad, ref) > 3e-2


def test_quick_gelu_epilogue():
    g = torch.Generator(device="cuda").manual_seed(1)
    a = (torch.randn(300, 256, device="cuda", generator=g) * 0.5).bfloat16()
    w = (torch.randn(512, 256, device="cuda", generator=g) * 0.2).bfloat16()
    bias = torch.randn(512, device="cuda", generator=g) * 0.1
    out = ops.gemm(a, w, ops.EPI_BIAS_QGELU, bias=bias).float()
    x = a.float() @ w.float().t() + bias
    ref = x * torch.sigmoid(1.702 * x)
    assert ((out - ref).abs().max() / ref.abs().max()).item() < 6e-3
    # ... and it is NOT the erf GELU: for x in (-3, -1) the two differ by ~0.02 at |value| ~0.05-0.15
    m = (x < -1) & (x > -3)
    assert (out[m] - ref[m]).abs().mean() < 0.1 * (F.gelu(x)[m] - ref[m]).abs().mean()


@pytest.mark.parametrize("cfg0", [CLIP_L14_336, SIGLIP_SO400M_14_384, DINOV2_L14_REG4_384, DINOV2_L14_REG4_336],
                         ids=["clip-336", "siglip-384", "dinov2-384", "dinov2-336"])
@pytest.mark.parametrize("folded", [True, False])
def test_tower_vs_oracle(cfg0, folded):
    cfg = cfg0.with_depth(3)
    sd = make_vit_state_dict(cfg, seed=61, init="stress")
    vit = blb.VisionTransformer(cfg)
    vit.ln_folded = folded
    vit.load_state_dict(sd)
    vit.cuda()
    frames = synthetic_frames(2, seed=3, size=cfg.img_size)
    px = normalize_for(cfg, frames).bfloat16()
    got = vit(px.cuda())
    assert got.shape == (2, cfg.num_patches, cfg.dim)
    ref = vit_oracle.vit_intermediate(sd, cfg, px.float())
    err = _rel(got, ref)
    print(f"{cfg.timm_id} @ {cfg.img_size}px ({cfg.tokens} tokens) ln_folded={folded}: {err:.3e}")
    assert err < TOL, err
    # the uint8 entry (Normalize folded into the patch-embed weights) lands on the same answer
    got8 = vit.forward_uint8(frames.cuda())
    ref8 = vit_oracle.vit_intermediate(sd, cfg, normalize_for(cfg, frames))
    assert _rel(got8, ref8) < TOL
    # sensitivity: a broken block must be visible
    bad = {k: v.clone() for k, v in sd.items()}
    bad["blocks.0.mlp.fc2.weight"].zero_()
    assert _rel(vit_oracle.vit_intermediate(bad, cfg, px.float()), ref) > 2 * TOL


@pytest.mark.parametrize("ident,size,cls,keys,cfgs", [
    ("dinosiglip-vit-so-384px", 384, blb.DinoSigLIPViTBackbone, ("dino", "siglip"), (DINOV2_L14_REG4_384, SIGLIP_SO400M_14_384)),
    ("dinoclip-vit-l-336px", 336, blb.DinoCLIPViTBackbone, ("dino", "clip"), (DINOV2_L14_REG4_336, CLIP_L14_336)),
])
def test_fused_backbone_and_projector_vs_oracle(ident, size, cls, keys, cfgs):
    c0, c1 = cfgs[0].with_depth(3), cfgs[1].with_depth(3)
    sd0, sd1 = make_vit_state_dict(c0, seed=71, init="stress"), make_vit_state_dict(c1, seed=72, init="stress")
    bb = cls(ident, "resize-naive", default_image_size=size)
    assert bb.num_patches == (size // 14) ** 2 and bb.default_image_resolution == (3, size, size)
    setattr(bb, f"{keys[0]}_featurizer", blb.VisionTransformer(c0))
    setattr(bb, f"{keys[1]}_featurizer", blb.VisionTransformer(c1))
    getattr(bb, f"{keys[0]}_featurizer").load_state_dict(sd0)
    getattr(bb, f"{keys[1]}_featurizer").load_state_dict(sd1)
    fused = c0.dim + c1.dim
    psd = make_projector_state_dict(fused_dim=fused, seed=73)
    proj = blb.FusedMLPProjector(fused, 4096)
    proj.load_state_dict(psd)
    enc = blb.VisualPrefixEncoder(bb, proj).cuda()
    frames = synthetic_frames(2, seed=9, size=size)
    px = {keys[0]: normalize_for(c0, frames).bfloat16(), keys[1]: normalize_for(c1, frames).bfloat16()}
    out, feats = enc({k: v.cuda() for k, v in px.items()}, return_features=True)
    assert out.shape == (2, bb.num_patches, 4096) and feats.shape == (2, bb.num_patches, fused)
    r0 = vit_oracle.vit_intermediate(sd0, c0, px[keys[0]].float())
    r1 = vit_oracle.vit_intermediate(sd1, c1, px[keys[1]].float())
    ref_feats = torch.cat([r0, r1], dim=2)
    ref = vit_oracle.projector_forward(psd, ref_feats)
    e0, e1, ep = _rel(feats[..., :c0.dim], r0), _rel(feats[..., c0.dim:], r1), _rel(out, ref)
    print(f"{ident}: {keys[0]} {e0:.3e}  {keys[1]} {e1:.3e}  projected {ep:.3e}")
    assert e0 < TOL and e1 < TOL and ep < TOL
    # reference-shaped two-step call and the folded uint8 entry
    assert torch.equal(enc.projector(enc.vision_backbone({k: v.cuda() for k, v in px.items()})), out)
    out8 = enc.forward_uint8(frames.cuda())
    assert _rel(out8, ref) < TOL
    with pytest.raises(ValueError):
        cls(ident, "resize-naive", default_image_size=224)


def test_registry_rejects_unknown_ids():
    with pytest.raises(ValueError):
        blb.DinoCLIPViTBackbone("dinoclip-vit-b", "resize-naive", default_image_size=336)
    with pytest.raises(ValueError):
        blb.CLIPViTBackbone("clip-vit-b", "resize-naive")
    bb = blb.SigLIPViTBackbone("siglip-vit-so400m-384px", "resize-naive", default_image_size=384)
    assert bb.num_patches == 729 and bb.embed_dim == 1152
