"""CPU-side checks: the C-ABI library loads and exports every declared symbol, the reference-facing classes keep
the reference's names / keys / error behaviour, sharding + all-gather logic (gloo, world_size 2)."""

import ctypes
import os
import re
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

import bridgelang_b200 as blb
from bridgelang_b200 import _lib
from bridgelang_b200.config import DINOV2_L14_REG4, SIGLIP_SO400M_14
from bridgelang_b200.weights import make_projector_state_dict, make_vit_state_dict

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def lib():
    from bridgelang_b200.build import build_library
    build_library()
    return _lib.load()


def test_library_exports_every_declared_symbol(lib):
    header = (ROOT / "include" / "bridgelang_b200.h").read_text()
    declared = set(re.findall(r"\b(blb_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_lib.EXPORTED_SYMBOLS), declared ^ set(_lib.EXPORTED_SYMBOLS)
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.blb_abi_version() == 3
    assert lib.blb_status_string(0) == b"ok"
    assert b"workspace" in lib.blb_status_string(-5)


def test_struct_layouts_match_header(lib):
    # ask the C compiler: sizeof + the offset of the last field of every struct in include/bridgelang_b200.h
    import subprocess, tempfile
    src = r'''#include <stdio.h>
#include <stddef.h>
#include "bridgelang_b200.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(blb_epilogue), offsetof(blb_epilogue, ld_xb),
         sizeof(blb_block_weights), offsetof(blb_block_weights, fc1_colsum), sizeof(blb_vit_weights),
         offsetof(blb_vit_weights, ln_folded), sizeof(blb_projector_weights), offsetof(blb_projector_weights, fc3_b));
  return 0;
}'''
    with tempfile.TemporaryDirectory() as d:
        (Path(d) / "t.c").write_text(src)
        subprocess.run(["gcc", "-I", str(ROOT / "include"), str(Path(d) / "t.c"), "-o", str(Path(d) / "t")], check=True)
        got = [int(x) for x in subprocess.run([str(Path(d) / "t")], capture_output=True, text=True, check=True).stdout.split()]
    want = [ctypes.sizeof(_lib.Epilogue), _lib.Epilogue.ld_xb.offset,
            ctypes.sizeof(_lib.BlockWeights), _lib.BlockWeights.fc1_colsum.offset,
            ctypes.sizeof(_lib.VitWeights), _lib.VitWeights.ln_folded.offset,
            ctypes.sizeof(_lib.ProjectorWeights), _lib.ProjectorWeights.fc3_b.offset]
    assert got == want, (got, want)


def test_bad_arguments_return_negative_status_without_gpu(lib):
    e = _lib.Epilogue()
    assert lib.blb_gemm_bf16(None, 0, None, 0, 0, 0, 0, 0, ctypes.byref(e), None) < 0
    assert lib.blb_layernorm(None, 0, None, None, None, 0, 0, 1024, 1e-6, None) == -1
    assert lib.blb_attention(None, None, 1, 1, 1, 64, None) == -1
    assert lib.blb_vit_workspace_bytes(None, 4) == 0


def test_cpu_tensors_fail_loudly():
    from bridgelang_b200 import ops
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.layernorm(torch.zeros(8, 1024), torch.ones(1024), torch.zeros(1024))
    proj = blb.FusedMLPProjector(256, 256)
    with pytest.raises(RuntimeError, match="CUDA"):
        proj(torch.zeros(1, 4, 256))
    vit = blb.VisionTransformer(SIGLIP_SO400M_14.with_depth(2))
    with pytest.raises(RuntimeError, match="CUDA"):
        vit(torch.zeros(1, 3, 224, 224))


def test_state_dict_keys_follow_the_reference():
    for cfg in (DINOV2_L14_REG4.with_depth(3), SIGLIP_SO400M_14.with_depth(3)):
        vit = blb.VisionTransformer(cfg)
        sd = make_vit_state_dict(cfg, init="timm")
        assert set(vit.state_dict().keys()) == set(sd.keys())
        vit.load_state_dict(sd, strict=True)
    # HF twin spells LayerScale `scale_factor` (modeling_prismatic.py:52-59) and accepts timm's `gamma`
    twin = blb.VisionTransformer(DINOV2_L14_REG4.with_depth(3), ls_param_name="scale_factor")
    assert "blocks.0.ls1.scale_factor" in twin.state_dict()
    twin.load_state_dict(make_vit_state_dict(DINOV2_L14_REG4.with_depth(3)), strict=True)
    proj = blb.FusedMLPProjector(2176, 4096)
    assert set(proj.state_dict()) == set(make_projector_state_dict())
    assert proj.initial_projection_dim == 8704
    hf = blb.PrismaticProjector(True, 2176, 4096)
    assert set(hf.state_dict()) == {f"fc{i}.{p}" for i in (1, 2, 3) for p in ("weight", "bias")}


def test_backbone_contract():
    bb = blb.DinoSigLIPViTBackbone("dinosiglip-vit-so-224px", "resize-naive", default_image_size=224)
    assert bb.embed_dim == 2176 and bb.num_patches == 256
    assert bb.default_image_resolution == (3, 224, 224)
    assert bb.half_precision_dtype == torch.bfloat16
    assert bb.identifier == "dinosiglip-vit-so-224px" and bb.image_resize_strategy == "resize-naive"
    names = dict(bb.named_children())
    assert {"dino_featurizer", "siglip_featurizer"} <= set(names)
    assert len(bb.dino_featurizer.blocks) == 24 and len(bb.siglip_featurizer.blocks) == 27
    assert callable(bb.get_fsdp_wrapping_policy())
    with pytest.raises(ValueError):
        blb.DinoSigLIPViTBackbone("dinosiglip-vit-so-384px", "resize-naive")
    with pytest.raises(ValueError):
        blb.DinoSigLIPViTBackbone("dinosiglip-vit-so-224px", "no-such-strategy")


def test_image_transform_shapes_and_normalisation():
    from PIL import Image
    bb = blb.DinoSigLIPViTBackbone("dinosiglip-vit-so-224px", "resize-naive")
    arr = np.full((256, 320, 3), 128, dtype=np.uint8)
    out = bb.get_image_transform()(Image.fromarray(arr))
    assert set(out) == {"dino", "siglip"}
    assert out["dino"].shape == (3, 224, 224) and out["siglip"].shape == (3, 224, 224)
    v = 128 / 255.0
    assert torch.allclose(out["siglip"], torch.full((3, 224, 224), (v - 0.5) / 0.5), atol=1e-6)
    want = torch.tensor([(v - m) / s for m, s in zip((0.485, 0.456, 0.406), (0.229, 0.224, 0.225))])
    assert torch.allclose(out["dino"].mean(dim=(1, 2)), want, atol=1e-5)


class _Tok:
    vocab_size = 32000

    def decode(self, ids):
        return " ".join(str(int(i)) for i in ids)

    def batch_decode(self, rows):
        return [self.decode(r) for r in rows]


def test_action_tokenizer_host_side(golden_dir):
    import json
    g = json.loads((golden_dir / "action_tokenizer.json").read_text())
    at = blb.ActionTokenizer(_Tok())
    assert at.action_token_begin_idx == 31743 and at.n_bins == 256 and at.vocab_size == 256
    assert at.bin_centers.shape == (255,)
    assert [float(x).hex() for x in at.bin_centers] == g["bin_centers_hex"]
    assert at(np.array(g["encode"]["actions"])) == " ".join(str(i) for i in g["encode"]["ids"])
    assert at(np.array([g["encode"]["actions"][:3]])) == [" ".join(str(i) for i in g["encode"]["ids"][:3])]


def test_unnorm_key_checks_match_reference_messages():
    stats = {"a": {"action": {"q01": [0.0] * 7, "q99": [1.0] * 7}}, "b": {"action": {"q01": [0.0] * 6, "q99": [1.0] * 6}}}
    with pytest.raises(AssertionError, match="trained on more than one dataset"):
        blb.OpenVLA._check_unnorm_key(stats, None)
    with pytest.raises(AssertionError, match="not in the set of available statistics"):
        blb.OpenVLA._check_unnorm_key(stats, "c")
    assert blb.OpenVLA._check_unnorm_key({"a": stats["a"]}, None) == "a"
    # HF twin (modeling_prismatic.py:538-552): its own wording
    with pytest.raises(AssertionError, match="trained on more than one dataset"):
        blb.OpenVLAForActionPrediction._check_unnorm_key(stats, None)
    with pytest.raises(AssertionError, match="not in the set of available dataset statistics"):
        blb.OpenVLAForActionPrediction._check_unnorm_key(stats, "c")


def test_prompt_builder_format():
    pb = blb.PurePromptBuilder("openvla")
    pb.add_turn(role="human", message="What action should the robot take to pick up the cup?")
    assert pb.get_prompt() == "In: What action should the robot take to pick up the cup?\nOut:"


def test_shard_bounds_cover_the_batch():
    for n, w in ((2048, 8), (2048, 4), (256, 1), (10, 4), (3, 8), (0, 2)):
        spans = [blb.shard_bounds(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [hi - lo for lo, hi in spans]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        blb.shard_bounds(8, 2, 2)


_GLOO_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.environ["BLB_ROOT"])
from bridgelang_b200.pipeline import PrefixGatherer, gather_prefixes, shard_bounds, shard_pixel_values
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:" + os.environ["BLB_PORT"],
                        rank=int(os.environ["RANK"]), world_size=2)
rank = dist.get_rank()
for global_batch in (6, 5):
    full = torch.arange(global_batch * 4 * 3, dtype=torch.float32).view(global_batch, 4, 3)
    px = shard_pixel_values({"dino": full, "siglip": full + 1}, rank, 2)
    lo, hi = shard_bounds(global_batch, rank, 2)
    assert torch.equal(px["dino"], full[lo:hi]) and torch.equal(px["siglip"], full[lo:hi] + 1)
    local = px["dino"] * 2            # stand-in for the per-rank featurize+project result
    out = gather_prefixes(local, global_batch)
    assert torch.equal(out, full * 2), (global_batch, rank)
    g = PrefixGatherer(global_batch)              # launch / wait form (in line on CPU tensors)
    assert torch.equal(g.wait(g.launch(local)), full * 2), (global_batch, rank)
dist.barrier()
dist.destroy_process_group()
print("ok", rank)
"""


def test_all_gather_of_prefixes_world_size_2_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER)
    port = str(29500 + os.getpid() % 2000)
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), BLB_PORT=port, BLB_ROOT=str(ROOT))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=180)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert all("ok" in o for o in outs)


def test_hf_image_processor_twin_matches_native_transform():
    """PrismaticImageProcessor.apply_transform (processing_prismatic.py:128-145) == the native dict transform,
    channel-stacked dino first; attributes parsed like the reference; unsupported strategy raises like the reference."""
    import numpy as np
    from PIL import Image
    from bridgelang_b200.weights import DINO_MEAN, DINO_STD, SIGLIP_MEAN, SIGLIP_STD
    img = Image.fromarray((np.random.default_rng(0).random((300, 260, 3)) * 255).astype(np.uint8))
    for strategy in ("resize-naive", "resize-crop", "letterbox"):
        proc = blb.PrismaticImageProcessor(True, strategy, [(3, 224, 224), (3, 224, 224)], ["bicubic", "bicubic"],
                                           [DINO_MEAN, SIGLIP_MEAN], [DINO_STD, SIGLIP_STD])
        t = proc.apply_transform(img)
        assert t.shape == (6, 224, 224) and t.dtype == torch.float32
        if strategy != "letterbox":      # (the HF twin's letterbox fill uses the LAST tower's mean, :118; native: per tower)
            d = blb.DinoSigLIPViTBackbone("dinosiglip-vit-so-224px", strategy).get_image_transform()(img)
            assert torch.equal(t[:3], d["dino"]) and torch.equal(t[3:], d["siglip"])
    proc = blb.PrismaticImageProcessor(True, "resize-naive", [(3, 224, 224)] * 2, ["bicubic"] * 2,
                                       [DINO_MEAN, SIGLIP_MEAN], [DINO_STD, SIGLIP_STD])
    assert proc.tvf_resize_params[0]["size"] == (224, 224) and proc.tvf_crop_params[0] == {"output_size": (224, 224)}
    out = proc([img, img], return_tensors="pt")["pixel_values"]
    assert out.shape == (2, 6, 224, 224) and torch.equal(out[0], out[1])
    assert isinstance(proc(img)["pixel_values"], np.ndarray)
    with pytest.raises(ValueError, match="is not supported"):
        blb.PrismaticImageProcessor(True, "stretch", [(3, 224, 224)] * 2, ["bicubic"] * 2)


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU restatement of the reference path on the host cores) must print one JSON line
    with the keys the round driver reads; it never touches a GPU."""
    import json
    import subprocess
    import sys
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, env={**os.environ, "CUDA_VISIBLE_DEVICES": ""})
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
    assert line["impl"] == "reference" and line["unit"] == "images/s" and line["higher_is_better"] is True
    assert line["metric"].startswith("images/sec DinoSigLIP-224px featurize+project")
    assert line["value"] > 0 and line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in line["config"] and "model" not in line["config"]


def test_openvla_reference_construction_and_from_pretrained(tmp_path):
    """openvla.py:23-33 + prismatic.py:40-123 + load.py:214-224: OpenVLA(model_id, vision_backbone, llm_backbone,
    arch_specifier=..., norm_stats=..., action_tokenizer=...) builds the projector from the arch specifier;
    from_pretrained loads ["model"]["projector" | "llm_backbone" | "vision_backbone"] and freezes."""
    import torch.nn as nn

    class _Cfg:
        hidden_size = 64

    class _TinyLM(nn.Module):
        config = _Cfg()

        def __init__(self):
            super().__init__()
            self.emb = nn.Embedding(10, 64)

    class _Tok:
        vocab_size = 32000

    stats = {"bridge_orig": {"action": {"q01": [0.0] * 7, "q99": [1.0] * 7}}}
    dcfg, scfg = DINOV2_L14_REG4.with_depth(2), SIGLIP_SO400M_14.with_depth(2)

    def backbone():
        bb = blb.DinoSigLIPViTBackbone("dinosiglip-vit-so-224px", "resize-naive")
        bb.dino_featurizer, bb.siglip_featurizer = blb.VisionTransformer(dcfg), blb.VisionTransformer(scfg)
        return bb

    src_bb = backbone()
    src_bb.dino_featurizer.load_state_dict(make_vit_state_dict(dcfg, seed=5))
    src_bb.siglip_featurizer.load_state_dict(make_vit_state_dict(scfg, seed=6))
    llm_backbone = blb.LLMBackbone("llama2-7b-pure", _TinyLM(), _Tok())
    assert llm_backbone.embed_dim == 64 and llm_backbone.half_precision_dtype == torch.bfloat16
    at = blb.ActionTokenizer(llm_backbone.get_tokenizer())
    src = blb.OpenVLA("openvla-7b", src_bb, llm_backbone, arch_specifier="no-align+fused-gelu-mlp", norm_stats=stats,
                      action_tokenizer=at)
    assert isinstance(src.projector, blb.FusedMLPProjector) and src.projector.projector[0].in_features == 2176
    assert src.projector.projector[4].out_features == 64
    assert src.all_module_keys == ["vision_backbone", "llm_backbone", "projector"] and src.model_family == "prismatic"
    assert src.get_action_dim() == 7 and src.norm_stats is stats and src.action_tokenizer is at
    for bad in ("linear", "gelu-mlp", "something-else"):
        with pytest.raises(ValueError):
            blb.OpenVLA("x", src_bb, llm_backbone, arch_specifier=bad, norm_stats=stats, action_tokenizer=at)
    with pytest.raises(TypeError):                       # norm_stats / action_tokenizer are required keywords
        blb.OpenVLA("x", src_bb, llm_backbone, arch_specifier="fused-gelu-mlp")
    ckpt = tmp_path / "step-000001.pt"
    torch.save({"model": {"projector": src.projector.state_dict(), "llm_backbone": llm_backbone.state_dict(),
                          "vision_backbone": src_bb.state_dict()}}, ckpt)
    dst = blb.OpenVLA.from_pretrained(ckpt, "openvla-7b", backbone(), blb.LLMBackbone("llama2-7b-pure", _TinyLM(), _Tok()),
                                      arch_specifier="no-align+fused-gelu-mlp", freeze_weights=True, norm_stats=stats,
                                      action_tokenizer=at)
    assert not any(p.requires_grad for p in dst.parameters()) and not dst.training
    for (ka, va), (kb, vb) in zip(src.state_dict().items(), dst.state_dict().items()):
        assert ka == kb and torch.equal(va, vb), ka
    torch.save({"model": {"projector": src.projector.state_dict()}}, ckpt)
    with pytest.raises(AssertionError, match="expects checkpoint with keys"):
        blb.OpenVLA.from_pretrained(ckpt, "m", backbone(), llm_backbone, arch_specifier="fused-gelu-mlp",
                                    norm_stats=stats, action_tokenizer=at)
    # generate() kwargs policy is host logic
    blb.OpenVLA._check_generate_kwargs({"do_sample": False, "use_cache": False})
    with pytest.raises(ValueError):
        blb.OpenVLA._check_generate_kwargs({"num_beams": 4})
