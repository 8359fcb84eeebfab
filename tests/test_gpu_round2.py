"""Round-2 parity cases on a B200 (VERDICT r01 "next round" item 1): BASELINE configs[2] shard sizes (512 / 1024 images
per GPU: 32-bit index overflow territory), configs[3] single-encoder backbones at full depth against the oracle, images
of a B=256 batch against the oracle, the windowed argmax mode, and the NCCL all-gather of prefixes on >= 2 GPUs."""

import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

import bridgelang_b200 as blb
from bridgelang_b200 import ops
from bridgelang_b200.config import DINOV2_L14_REG4, SIGLIP_SO400M_14
from bridgelang_b200.weights import (make_projector_state_dict, make_vit_state_dict, normalize_frames,
                                     synthetic_frames)
from oracle import action_oracle, vit_oracle

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent
TOL = 2e-2


def _rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).abs().max() / b.abs().max()).item()


# ---------------------------------------------------------------------------------------------------------------
# BASELINE configs[2]: global batch 2048 at 2 / 4 GPUs = 1024 / 512 images per GPU
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("cfg,key", [(DINOV2_L14_REG4, "dino"), (SIGLIP_SO400M_14, "siglip")])
@pytest.mark.parametrize("batch", [512, 1024])
def test_shard_sizes_of_global_batch_2048(cfg, key, batch):
    """B_local = 512 / 1024: M = B·T rows x up to 4352 columns crosses 2^31 BYTES (DINOv2 B=1024: 267 264 x 4096 x 2 B =
    2.19 GB for the MLP hidden) — every row/column offset in the kernels must be 64-bit.  Depth 2 keeps it fast; images
    at both ends and in the middle of the batch must equal, bit for bit, what they get in a batch of their own, and
    one of them is compared with the oracle."""
    cfg = cfg.with_depth(2)
    sd = make_vit_state_dict(cfg, seed=77, init="stress")
    vit = blb.VisionTransformer(cfg)
    vit.load_state_dict(sd)
    vit.cuda()
    gen = torch.Generator(device="cuda").manual_seed(batch)
    px = torch.randn((batch, 3, 224, 224), device="cuda", generator=gen).bfloat16()
    full = vit(px)
    assert full.shape == (batch, 256, cfg.dim)
    assert bool(torch.isfinite(full.float()).all())
    for lo, hi in ((0, 2), (batch // 2 - 1, batch // 2 + 1), (batch - 2, batch)):
        assert torch.equal(vit(px[lo:hi].contiguous()), full[lo:hi]), (lo, hi)
    last = px[batch - 1:batch].float().cpu()
    ref = vit_oracle.vit_intermediate(sd, cfg, last)
    assert _rel(full[batch - 1:batch], ref) < TOL
    del full
    torch.cuda.empty_cache()


# ---------------------------------------------------------------------------------------------------------------
# BASELINE configs[3]: single-encoder variants, full depth
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("cls,ident,cfg,key,seed", [
    (blb.DinoV2ViTBackbone, "dinov2-vit-l", DINOV2_L14_REG4, "dino", 1234),
    (blb.SigLIPViTBackbone, "siglip-vit-so400m", SIGLIP_SO400M_14, "siglip", 1235),
])
def test_single_encoder_backbones_full_depth_vs_oracle(cls, ident, cfg, key, seed):
    """`dinov2-vit-l` (24 blocks, hd 64, LayerScale, 261 tokens) and `siglip-vit-so400m` (27 blocks, hd 72) through the
    reference-named backbone classes (dinov2_vit.py:9-19, siglip_vit.py:8-24), stress-init weights, B=2, every block."""
    sd = make_vit_state_dict(cfg, seed=seed, init="stress")
    bb = cls(ident, "resize-naive")
    bb.featurizer.load_state_dict(sd)
    bb.cuda()
    px = normalize_frames(synthetic_frames(2, seed=11))[key].bfloat16()
    got = bb(px.cuda())
    assert got.shape == (2, 256, cfg.dim) and bb.embed_dim == cfg.dim and bb.num_patches == 256
    ref = vit_oracle.vit_intermediate(sd, cfg, px.float())
    err = _rel(got, ref)
    print(f"{ident}: full-depth parity {err:.3e}")
    assert err < TOL, err


def test_images_of_a_256_batch_vs_oracle():
    """Full-depth fused path at B=256 (BASELINE configs[1]): the first, a middle and the last image of THE SAME batch are
    compared with the fp32 oracle (the oracle needs ~1 s per image, so three images, not 256)."""
    dsd = make_vit_state_dict(DINOV2_L14_REG4, seed=1234, init="stress")
    ssd = make_vit_state_dict(SIGLIP_SO400M_14, seed=1235, init="stress")
    psd = make_projector_state_dict(seed=4321)
    bb = blb.DinoSigLIPViTBackbone("dinosiglip-vit-so-224px", "resize-naive")
    bb.dino_featurizer.load_state_dict(dsd)
    bb.siglip_featurizer.load_state_dict(ssd)
    proj = blb.FusedMLPProjector(bb.embed_dim, 4096)
    proj.load_state_dict(psd)
    enc = blb.VisualPrefixEncoder(bb, proj).cuda()
    px = {k: v.bfloat16() for k, v in normalize_frames(synthetic_frames(256, seed=3)).items()}
    out = enc({k: v.cuda() for k, v in px.items()})
    assert out.shape == (256, 256, 4096)
    for i in (0, 131, 255):
        one = {k: v[i:i + 1].float() for k, v in px.items()}
        ref = vit_oracle.featurize_project(dsd, DINOV2_L14_REG4, ssd, SIGLIP_SO400M_14, psd, one)
        err = _rel(out[i:i + 1], ref)
        print(f"image {i} of the B=256 batch: projected parity {err:.3e}")
        assert err < TOL, (i, err)


# ---------------------------------------------------------------------------------------------------------------
# windowed argmax (north_star: "argmax over the last 256 vocab bins"): an explicit, separately tested mode
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
def test_windowed_argmax_mode(dtype):
    g = torch.Generator(device="cuda").manual_seed(5)
    logits = torch.randn((9, 32064), device="cuda", generator=g).to(dtype)
    lo, hi = 31744, 32000                       # the 256 action bins of the 32000-token vocabulary
    logits[0, 31999] = 50.0                     # winner inside the window → both modes agree
    logits[1, 17] = 50.0                        # full-row winner is a text token → the modes differ
    logits[2, 32010] = 50.0                     # ... or a padding row beyond the tokenizer vocabulary
    logits[3, 31800] = logits[3, 31900] = 40.0  # tie inside the window: first index wins
    logits[4, lo:hi] = float("-inf")
    logits[4, 31777] = -1e4                     # everything else -inf
    logits[5, 31750] = float("nan")             # NaN is maximal, as in torch.argmax
    ids = ops.argmax_window(logits, lo, hi)
    want = torch.argmax(logits[:, lo:hi], dim=-1) + lo
    assert torch.equal(ids, want)
    full = ops.argmax(logits)
    assert torch.equal(full, torch.argmax(logits, dim=-1))
    assert ids[0] == full[0] and ids[1] != full[1] and ids[2] != full[2] and ids[3] == 31800
    # arbitrary (also wide: block-reduced) windows
    for a, b in ((0, 32064), (5, 6), (100, 1124), (100, 1125), (31000, 32064)):
        assert torch.equal(ops.argmax_window(logits, a, b), torch.argmax(logits[:, a:b], dim=-1) + a)
    # fused with the bin-centre lookup and the un-normalize
    stats = {"q01": [-0.03, -0.04, -0.05, -0.08, -0.10, -0.20, 0.0], "q99": [0.03, 0.04, 0.05, 0.08, 0.10, 0.20, 1.0],
             "mask": [True] * 6 + [False]}
    bins = np.linspace(-1, 1, 256)
    tables = ops.DecodeTables((bins[:-1] + bins[1:]) / 2.0, stats["q01"], stats["q99"], stats["mask"])
    ids2, norm, act = ops.argmax_window_detokenize_unnormalize(logits[:7], lo, hi, 32000, tables)
    assert torch.equal(ids2, want[:7])
    want_norm = action_oracle.decode_token_ids_to_actions(want[:7].cpu().numpy(), 32000)
    assert np.array_equal(norm.cpu().numpy(), want_norm)
    assert np.array_equal(act.cpu().numpy(), action_oracle.unnormalize(want_norm, stats))
    with pytest.raises(RuntimeError):
        ops.argmax_window(logits, 10, 10)
    with pytest.raises(RuntimeError):
        ops.argmax_window(logits, 0, 40000)


# ---------------------------------------------------------------------------------------------------------------
# device guard: a model on cuda:1 while cuda:0 is current
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_operands_on_a_non_current_device():
    cfg = SIGLIP_SO400M_14.with_depth(2)
    sd = make_vit_state_dict(cfg, seed=3)
    px = normalize_frames(synthetic_frames(2, seed=1))["siglip"].bfloat16()
    vit0, vit1 = blb.VisionTransformer(cfg), blb.VisionTransformer(cfg)
    vit0.load_state_dict(sd)
    vit1.load_state_dict(sd)
    vit0.to("cuda:0")
    vit1.to("cuda:1")
    torch.cuda.set_device(0)
    a = vit0(px.to("cuda:0"))
    b = vit1(px.to("cuda:1"))                   # current device is still cuda:0
    torch.cuda.synchronize(0)
    torch.cuda.synchronize(1)
    assert b.device.index == 1 and torch.equal(a.cpu(), b.cpu())
    with pytest.raises(RuntimeError):
        ops.gemm(torch.zeros(128, 64, device="cuda:0", dtype=torch.bfloat16),
                 torch.zeros(128, 64, device="cuda:1", dtype=torch.bfloat16), ops.EPI_BIAS)


# ---------------------------------------------------------------------------------------------------------------
# NCCL all-gather of projected prefixes over NVLink — only where the box has >= 2 GPUs
# ---------------------------------------------------------------------------------------------------------------
_NCCL_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.environ["BLB_ROOT"])
import bridgelang_b200 as blb
from bridgelang_b200.config import DINOV2_L14_REG4, SIGLIP_SO400M_14
from bridgelang_b200.weights import make_projector_state_dict, make_vit_state_dict, normalize_frames, synthetic_frames
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", init_method="tcp://127.0.0.1:" + os.environ["BLB_PORT"], rank=rank, world_size=world,
                        device_id=torch.device("cuda", rank))
dcfg, scfg = DINOV2_L14_REG4.with_depth(3), SIGLIP_SO400M_14.with_depth(3)
bb = blb.DinoSigLIPViTBackbone("dinosiglip-vit-so-224px", "resize-naive")
bb.dino_featurizer = blb.VisionTransformer(dcfg); bb.siglip_featurizer = blb.VisionTransformer(scfg)
bb.dino_featurizer.load_state_dict(make_vit_state_dict(dcfg, seed=1)); bb.siglip_featurizer.load_state_dict(make_vit_state_dict(scfg, seed=2))
proj = blb.FusedMLPProjector(2176, 4096); proj.load_state_dict(make_projector_state_dict(seed=3))
enc = blb.VisualPrefixEncoder(bb, proj).cuda()
for global_batch in (6, 5):                      # even and ragged shards
    px = {k: v.bfloat16().cuda() for k, v in normalize_frames(synthetic_frames(global_batch, seed=9)).items()}
    full = enc(px)                               # every rank computes the whole batch as the reference answer
    local = enc(blb.shard_pixel_values(px, rank, world))
    out = blb.gather_prefixes(local, global_batch)
    assert out.shape == full.shape and torch.equal(out, full), (global_batch, rank)
    # overlapped variants: the gather of step i runs on a side stream under the encode of step i+1, as NCCL or as
    # copy-engine pushes into the peers' symmetric buffers; three rounds exercise the double buffering
    for mode in ("nccl", "p2p"):
        g = blb.PrefixGatherer(global_batch, mode=mode)
        for rnd in range(3):
            h1 = g.launch(local); local2 = enc(blb.shard_pixel_values(px, rank, world)); o1 = g.wait(h1)
            assert torch.equal(o1, full) and torch.equal(local2, local), (mode, rnd)
torch.cuda.synchronize(); dist.barrier(); dist.destroy_process_group()
print("ok", rank)
"""


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_nccl_all_gather_of_prefixes(tmp_path):
    script = tmp_path / "nccl_worker.py"
    script.write_text(_NCCL_WORKER)
    port = str(29600 + os.getpid() % 2000)
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", BLB_PORT=port, BLB_ROOT=str(ROOT))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=600)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert all("ok" in o for o in outs)
