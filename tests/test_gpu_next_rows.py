"""SURVEY.md §8f "next" rows on a B200: device-side image preprocessing, ActionTokenizer encode ids, and the training
loops' action-token accuracy / L1 — each against the reference's own formulation (torchvision transforms, NumPy
digitize via the oracle + golden vectors, the base_strategy.py:314-329 restatement)."""

import json

import numpy as np
import pytest
import torch

import bridgelang_b200 as blb
from bridgelang_b200.weights import DINO_MEAN, DINO_STD, SIGLIP_MEAN, SIGLIP_STD, normalize_frames, synthetic_frames
from oracle import action_oracle

pytestmark = pytest.mark.gpu


class _Tok:
    vocab_size = 32000


@pytest.fixture(scope="module")
def at():
    return blb.ActionTokenizer(_Tok())


# ---- preprocessing -------------------------------------------------------------------------------------------------
def test_preprocess_uint8_bit_identical_to_torchvision_transforms():
    from PIL import Image
    from torchvision import transforms as T

    bb = blb.DinoSigLIPViTBackbone("dinosiglip-vit-so-224px", "resize-naive")
    frames = synthetic_frames(5, seed=11)
    frames[0] = 0
    frames[1] = 255
    frames[2, :, :, 0], frames[2, :, :, 1], frames[2, :, :, 2] = 17, 128, 254        # channel order must survive
    got = bb.preprocess_uint8(frames.cuda())
    assert got["dino"].dtype == torch.bfloat16 and got["dino"].shape == (5, 3, 224, 224)
    # (1) the tensor formulation used everywhere else in this repo
    want = normalize_frames(frames)
    for k in ("dino", "siglip"):
        assert torch.equal(got[k].cpu(), want[k].to(torch.bfloat16)), k
    # (2) the reference's literal transform objects on PIL images (ToTensor + Normalize, processing_prismatic.py:128-145)
    for name, mean, std in (("dino", DINO_MEAN, DINO_STD), ("siglip", SIGLIP_MEAN, SIGLIP_STD)):
        tf = T.Compose([T.ToTensor(), T.Normalize(mean=torch.tensor(mean), std=torch.tensor(std))])
        for i in (2, 3):
            ref = tf(Image.fromarray(frames[i].numpy())).to(torch.bfloat16)
            assert torch.equal(got[name][i].cpu(), ref), (name, i)


def test_forward_uint8_equals_forward_on_host_normalized_frames():
    from bridgelang_b200.config import DINOV2_L14_REG4, SIGLIP_SO400M_14
    from bridgelang_b200.weights import make_projector_state_dict, make_vit_state_dict

    bb = blb.DinoSigLIPViTBackbone("dinosiglip-vit-so-224px", "resize-naive")
    bb.dino_featurizer = blb.VisionTransformer(DINOV2_L14_REG4.with_depth(3))
    bb.siglip_featurizer = blb.VisionTransformer(SIGLIP_SO400M_14.with_depth(3))
    bb.dino_featurizer.load_state_dict(make_vit_state_dict(DINOV2_L14_REG4.with_depth(3), seed=1))
    bb.siglip_featurizer.load_state_dict(make_vit_state_dict(SIGLIP_SO400M_14.with_depth(3), seed=2))
    proj = blb.FusedMLPProjector(2176, 4096)
    proj.load_state_dict(make_projector_state_dict(seed=3))
    enc = blb.VisualPrefixEncoder(bb, proj).cuda()
    frames = synthetic_frames(3, seed=4)
    a = enc.forward_uint8(frames.cuda(), folded=False)      # LUT route: bit-identical to the host transform
    b = enc({k: v.to(torch.bfloat16).cuda() for k, v in normalize_frames(frames).items()})
    assert torch.equal(a, b)


# ---- encode --------------------------------------------------------------------------------------------------------
def test_encode_ids_golden_and_oracle(at, golden_dir):
    g = json.loads((golden_dir / "action_tokenizer.json").read_text())
    enc = g["encode"]
    a = torch.tensor(enc["actions"], dtype=torch.float64, device="cuda")
    assert at.encode_on_device(a).cpu().tolist() == enc["ids"]
    rng = np.random.default_rng(0)
    for dtype in (np.float64, np.float32):
        x = rng.uniform(-1.3, 1.3, size=(64, 7)).astype(dtype)
        x[0, :4] = [-1.0, 1.0, 0.0, np.nan]
        x[1] = at.bins[100:107].astype(dtype)                  # exactly on bin edges (digitize is right-open)
        x[2] = np.nextafter(at.bins[100:107], -np.inf).astype(dtype)
        want = action_oracle.encode_actions_to_token_ids(x, 32000)
        got = at.encode_on_device(torch.from_numpy(x).cuda())
        assert got.shape == x.shape
        assert np.array_equal(got.cpu().numpy(), want), dtype
    # round trip through the device decode: |decode(encode(a)) - a| <= one bin width
    a = torch.linspace(-1, 1, 1001, dtype=torch.float64, device="cuda")
    norm, _ = at.decode_on_device(at.encode_on_device(a))
    assert float((norm - a).abs().max()) <= 2.0 / 255 + 1e-12


# ---- training-side metrics -----------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_action_token_metrics_match_the_training_loop(at, dtype):
    B, P, L, V = 5, 256, 24, 32064
    g = torch.Generator().manual_seed(3)
    logits = torch.randn(B, P + L, V, generator=g).to(dtype)
    labels = torch.full((B, L), -100, dtype=torch.int64)                 # IGNORE_INDEX everywhere ...
    for b in range(B):
        start = 8 + b
        labels[b, start:start + 7] = torch.randint(31744, 32000, (7,), generator=g)   # ... except 7 action tokens
        labels[b, start + 7] = 2                                                       # EOS: not > begin_idx
        for j in range(7):                                                             # make ~half the predictions right
            pos = P + start + j - 1                                                    # logits position predicting label j
            tok = int(labels[b, start + j]) if (b + j) % 2 == 0 else 31744 + (b * 7 + j) % 256
            logits[b, pos, tok] = 50.0
    labels[0, 3] = 31743                                                  # == begin_idx: excluded by the strict '>'
    want_acc, want_l1, want_preds, want_mask = action_oracle.action_token_metrics(logits.float(), labels, P, 32000)
    acc, l1 = at.action_metrics(logits.cuda(), labels.cuda(), P)
    r = blb.ops.action_token_metrics(logits.cuda(), labels.cuda(), P, at.action_token_begin_idx, 32000, at.tables(None))
    assert int(r["counts"][1]) == int(want_mask.sum()) == 35
    assert int(r["counts"][0]) == int(((want_preds == labels[:, 1:]) & want_mask).sum())
    assert torch.equal(r["preds"].cpu()[want_mask], want_preds[want_mask])           # ids bit-exact where they count
    assert bool((r["preds"].cpu()[~want_mask] == -1).all())                           # skipped rows
    assert acc.dtype == torch.float32 and float(acc) == float(want_acc)
    assert l1.dtype == torch.float64 and abs(float(l1) - float(want_l1)) <= 1e-12


def test_hf_processor_device_path_bit_identical():
    """PrismaticImageProcessor.preprocess_to_device (one uint8 frame + LUT kernel) == preprocess(...).to(cuda, bf16)."""
    from PIL import Image
    proc = blb.PrismaticImageProcessor(True, "resize-naive", [(3, 224, 224)] * 2, ["bicubic"] * 2,
                                       [DINO_MEAN, SIGLIP_MEAN], [DINO_STD, SIGLIP_STD])
    rng = np.random.default_rng(5)
    imgs = [Image.fromarray((rng.random((256, 256, 3)) * 255).astype(np.uint8)),
            Image.fromarray((rng.random((480, 640, 3)) * 255).astype(np.uint8))]
    want = proc.preprocess(imgs, return_tensors="pt")["pixel_values"].to("cuda", dtype=torch.bfloat16)
    got = proc.preprocess_to_device(imgs)
    assert got.shape == (2, 6, 224, 224) and got.dtype == torch.bfloat16
    assert torch.equal(got, want)


# ---------------------------------------------------------------------------------------------------------------
# Round 2, SURVEY §8f.2 done properly: mean/std folded into the patch-embed weights, ONE patch-major uint8 matrix
# feeding both towers (im2col as a permutation, TMA-loaded A operand), coalesced EPI_PATCH stores
# ---------------------------------------------------------------------------------------------------------------
def test_u8_to_patches_is_the_hwc_permutation():
    import torch.nn.functional as F
    from bridgelang_b200 import ops
    frames = synthetic_frames(3, seed=4).cuda()
    cols = ops.u8_to_patches(frames)
    assert cols.shape == (3 * 256, 592) and cols.dtype == torch.bfloat16
    # reference: unfold of the NCHW view gives (c, kh, kw) order → reorder to (kh, kw, c)
    x = frames.permute(0, 3, 1, 2).float()
    ref = F.unfold(x, kernel_size=14, stride=14).transpose(1, 2).reshape(3 * 256, 3, 14, 14).permute(0, 2, 3, 1).reshape(-1, 588)
    assert torch.equal(cols[:, :588].float(), ref) and bool((cols[:, 588:] == 0).all())


def test_folded_uint8_entry_matches_the_normalised_path_and_the_oracle():
    """forward_uint8 (folded patch weights, shared patch matrix) vs forward on the normalised frames vs the fp32
    oracle: the fold moves the bf16 rounding from the normalised pixel to the weight, so the two native paths agree to
    bf16 accuracy (not bit for bit) and both sit inside the gate against the oracle."""
    import bridgelang_b200 as blb
    from bridgelang_b200.config import DINOV2_L14_REG4, SIGLIP_SO400M_14
    from bridgelang_b200.weights import make_projector_state_dict, make_vit_state_dict, normalize_frames
    from oracle import vit_oracle
    dcfg, scfg = DINOV2_L14_REG4.with_depth(4), SIGLIP_SO400M_14.with_depth(4)
    dsd, ssd = make_vit_state_dict(dcfg, seed=41, init="stress"), make_vit_state_dict(scfg, seed=42, init="stress")
    psd = make_projector_state_dict(seed=43)
    bb = blb.DinoSigLIPViTBackbone("dinosiglip-vit-so-224px", "resize-naive")
    bb.dino_featurizer, bb.siglip_featurizer = blb.VisionTransformer(dcfg), blb.VisionTransformer(scfg)
    bb.dino_featurizer.load_state_dict(dsd)
    bb.siglip_featurizer.load_state_dict(ssd)
    proj = blb.FusedMLPProjector(2176, 4096)
    proj.load_state_dict(psd)
    enc = blb.VisualPrefixEncoder(bb, proj).cuda()
    frames = synthetic_frames(3, seed=8)
    px = normalize_frames(frames)
    ref = vit_oracle.featurize_project(dsd, dcfg, ssd, scfg, psd, px)
    rel = lambda a, b: ((a.float().cpu() - b.float().cpu()).abs().max() / b.float().cpu().abs().max()).item()
    out_u8, feats_u8 = enc.forward_uint8(frames.cuda(), return_features=True)
    out_px = enc({k: v.bfloat16().cuda() for k, v in px.items()})
    out_lut = enc.forward_uint8(frames.cuda(), folded=False)
    assert torch.equal(out_lut, out_px)                         # the LUT route is bit-identical to the host transform
    e_u8, e_px = rel(out_u8, ref), rel(out_px, ref)
    print(f"uint8 folded {e_u8:.3e}   normalised bf16 {e_px:.3e}   folded vs normalised {rel(out_u8, out_px):.3e}")
    assert e_u8 < 2e-2 and e_px < 2e-2 and e_u8 < 1.5 * e_px + 1e-3
    # single-tower uint8 entries
    d8 = bb.dino_featurizer.forward_uint8(frames.cuda())
    s8 = bb.siglip_featurizer.forward_uint8(frames.cuda())
    assert torch.equal(d8, feats_u8[..., :1024]) and torch.equal(s8, feats_u8[..., 1024:])
    # stream() with uint8 frames uses the folded entry
    got = list(enc.stream([frames.pin_memory()], uint8=True))
    assert torch.equal(got[0], out_u8)
    with pytest.raises(ValueError):
        enc.forward_uint8(torch.zeros(1, 3, 224, 224, dtype=torch.uint8, device="cuda"))       # CHW is rejected
