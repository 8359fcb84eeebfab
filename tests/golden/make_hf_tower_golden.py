"""make_hf_tower_golden.py — freeze the outputs of transformers' Dinov2WithRegistersModel / SiglipVisionModel
(an implementation of the two towers that is independent of this repo) on seeded weights and inputs:

    python tests/golden/make_hf_tower_golden.py      → tests/golden/hf_towers_depth3.npz, hf_towers_full_depth.npz

timm 0.9.10 — the library the reference actually calls (pyproject.toml:45) — cannot be installed offline, so this is
the strongest available pin for oracle/vit_oracle.py: depth-3 towers at the real widths (1024 / 1152, 16 heads, 261 / 256
tokens), stress-init weights from bridgelang_b200.weights.make_vit_state_dict (seeds 11 / 12), one 224 px frame.
Stored: the second-to-last block's patch tokens, every 8th token x every 16th channel (fp32), plus the full-tensor
sum and absolute sum.  transformers version used is recorded in the file.
Round 2 adds hf_towers_full_depth.npz: the FULL-depth towers (24 / 27 blocks, the weights of the parity gates: seeds
1234 / 1235, stress-init), a batch of TWO frames, same sub-sampling per image — this pins both token orders (261-token
cls+reg+patch DINOv2, 256-token SigLIP), the prefix drop and every block index at the depth the product runs."""

import sys
from pathlib import Path

import numpy as np
import torch
import transformers

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import hf_mapping  # noqa: E402
from bridgelang_b200.config import DINOV2_L14_REG4, SIGLIP_SO400M_14  # noqa: E402
from bridgelang_b200.weights import make_vit_state_dict  # noqa: E402

DEPTH = 3


def main():
    out = {"transformers_version": np.array(transformers.__version__), "depth": np.array(DEPTH)}
    for name, cfg0, seed, xseed, build, run in (
            ("dino", DINOV2_L14_REG4, 11, 0, hf_mapping.build_hf_dinov2_reg4, hf_mapping.hf_dinov2_penultimate),
            ("siglip", SIGLIP_SO400M_14, 12, 1, hf_mapping.build_hf_siglip, hf_mapping.hf_siglip_penultimate)):
        cfg = cfg0.with_depth(DEPTH)
        sd = make_vit_state_dict(cfg, seed=seed, init="stress")
        x = torch.randn(1, 3, 224, 224, generator=torch.Generator().manual_seed(xseed))
        ref = run(build(sd, cfg, DEPTH), x)
        out[f"{name}_slice"] = ref[0, ::8, ::16].numpy().astype(np.float32)
        out[f"{name}_sum"] = np.array(ref.double().sum().item())
        out[f"{name}_abs_sum"] = np.array(ref.double().abs().sum().item())
        out[f"{name}_weight_seed"], out[f"{name}_pixel_seed"] = np.array(seed), np.array(xseed)
    np.savez_compressed(Path(__file__).resolve().parent / "hf_towers_depth3.npz", **out)
    print({k: (v.shape if v.ndim else v.item()) for k, v in out.items()})

    full = {"transformers_version": np.array(transformers.__version__)}
    for name, cfg, seed, xseed, build, run in (
            ("dino", DINOV2_L14_REG4, 1234, 20, hf_mapping.build_hf_dinov2_reg4, hf_mapping.hf_dinov2_penultimate),
            ("siglip", SIGLIP_SO400M_14, 1235, 21, hf_mapping.build_hf_siglip, hf_mapping.hf_siglip_penultimate)):
        sd = make_vit_state_dict(cfg, seed=seed, init="stress")
        x = torch.randn(2, 3, 224, 224, generator=torch.Generator().manual_seed(xseed))
        ref = run(build(sd, cfg, cfg.depth), x)
        assert ref.shape == (2, 256, cfg.dim)
        full[f"{name}_slice"] = ref[:, ::8, ::16].numpy().astype(np.float32)
        full[f"{name}_sum"] = np.array(ref.double().sum().item())
        full[f"{name}_abs_sum"] = np.array(ref.double().abs().sum().item())
        full[f"{name}_weight_seed"], full[f"{name}_pixel_seed"] = np.array(seed), np.array(xseed)
        full[f"{name}_depth"] = np.array(cfg.depth)
    np.savez_compressed(Path(__file__).resolve().parent / "hf_towers_full_depth.npz", **full)
    print({k: (v.shape if v.ndim else v.item()) for k, v in full.items()})


if __name__ == "__main__":
    main()
