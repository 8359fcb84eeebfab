"""make_golden.py — generate golden vectors by EXECUTING the reference's own code (run in the build container,
where /root/reference exists; the GPU box only sees the committed outputs).

  python tests/golden/make_golden.py        → tests/golden/action_tokenizer.json, tests/golden/projector_small.npz

Loaded by file path (importing the `prismatic` package itself fails on missing draccus/timm/tensorflow):
  /root/reference/prismatic/vla/action_tokenizer.py   ActionTokenizer
  /root/reference/prismatic/util/nn_utils.py          FusedMLPProjector
The un-normalize lines are the literal expression of /root/reference/prismatic/models/vlas/openvla.py:94-101.
"""

import importlib.util
import json
from pathlib import Path

import numpy as np
import torch

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent


def load(path: Path, name: str):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class StubTokenizer:
    """Llama-2 tokenizer stand-in: ActionTokenizer only reads .vocab_size on the decode path."""
    vocab_size = 32000


def main() -> None:
    at_mod = load(REF / "prismatic/vla/action_tokenizer.py", "ref_action_tokenizer")
    nn_mod = load(REF / "prismatic/util/nn_utils.py", "ref_nn_utils")

    tok = at_mod.ActionTokenizer(StubTokenizer())
    stats = {  # bridge_orig-style statistics: 6 normalised dims + un-normalised gripper
        "q01": [-0.03, -0.04, -0.05, -0.08, -0.10, -0.20, 0.0],
        "q99": [0.03, 0.04, 0.05, 0.08, 0.10, 0.20, 1.0],
        "mask": [True] * 6 + [False],
    }

    def unnorm(normalized_actions, action_norm_stats):   # openvla.py:94-101, verbatim expression
        mask = action_norm_stats.get("mask", np.ones_like(action_norm_stats["q01"], dtype=bool))
        action_high, action_low = np.array(action_norm_stats["q99"]), np.array(action_norm_stats["q01"])
        return np.where(mask, 0.5 * (normalized_actions + 1) * (action_high - action_low) + action_low,
                        normalized_actions)

    rng = np.random.default_rng(7)
    cases = []
    id_sets = [
        [31872, 31999, 31744, 31745, 31900, 31810, 31750],      # the SURVEY KAT
        [32000, 32063, 31743, 5, 0, 31998, 31746],              # pad rows, below the window, ordinary text ids
        *rng.integers(31744, 32000, size=(6, 7)).tolist(),
        *rng.integers(0, 32064, size=(4, 7)).tolist(),
    ]
    stats_nomask = {k: v for k, v in stats.items() if k != "mask"}
    for ids in id_sets:
        ids_np = np.array(ids, dtype=np.int64)
        norm = tok.decode_token_ids_to_actions(ids_np)
        cases.append({
            "ids": ids,
            "normalized_hex": [float(x).hex() for x in norm],
            "actions_hex": [float(x).hex() for x in unnorm(norm, stats)],
            "actions_nomask_hex": [float(x).hex() for x in unnorm(norm, stats_nomask)],
        })
    enc_in = [-1.0, -0.999, 0.0, 0.5, 0.999, 1.0, 1.5, -3.0, 0.00390625]
    clipped = np.clip(np.array(enc_in), a_min=float(tok.min_action), a_max=float(tok.max_action))
    enc_ids = (tok.tokenizer.vocab_size - np.digitize(clipped, tok.bins)).tolist()   # action_tokenizer.py:40-45
    golden = {
        "source": "prismatic/vla/action_tokenizer.py + prismatic/models/vlas/openvla.py:94-101 executed by make_golden.py",
        "vocab_size": 32000,
        "n_bins": tok.n_bins,
        "action_token_begin_idx": tok.action_token_begin_idx,
        "bin_centers_hex": [float(x).hex() for x in tok.bin_centers],
        "stats": stats,
        "decode_cases": cases,
        "encode": {"actions": enc_in, "ids": enc_ids},
    }
    (OUT / "action_tokenizer.json").write_text(json.dumps(golden, indent=1))

    torch.manual_seed(0)
    proj = nn_mod.FusedMLPProjector(fused_vision_dim=64, llm_dim=96).eval()
    x = torch.randn(2, 5, 64)
    with torch.no_grad():
        y = proj(x)
    arrays = {k.replace(".", "__"): v.numpy() for k, v in proj.state_dict().items()}
    np.savez_compressed(OUT / "projector_small.npz", x=x.numpy(), y=y.numpy(), **arrays)
    print("wrote", OUT / "action_tokenizer.json", OUT / "projector_small.npz")


if __name__ == "__main__":
    main()
