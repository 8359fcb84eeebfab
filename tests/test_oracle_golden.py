"""The oracle against the golden vectors produced by the reference's own code (tests/golden/make_golden.py)."""

import json

import numpy as np
import torch

from oracle import action_oracle, vit_oracle


def _hex(xs):
    return np.array([float.fromhex(x) for x in xs], dtype=np.float64)


def test_bin_centers_and_begin_idx(golden_dir):
    g = json.loads((golden_dir / "action_tokenizer.json").read_text())
    _, centers = action_oracle.make_bins(g["n_bins"])
    assert centers.shape == (255,)
    assert np.array_equal(centers, _hex(g["bin_centers_hex"]))
    assert action_oracle.action_token_begin_idx(g["vocab_size"], g["n_bins"]) == g["action_token_begin_idx"] == 31743


def test_decode_and_unnormalize_bit_exact(golden_dir):
    g = json.loads((golden_dir / "action_tokenizer.json").read_text())
    stats = g["stats"]
    nomask = {k: v for k, v in stats.items() if k != "mask"}
    for case in g["decode_cases"]:
        ids = np.array(case["ids"], dtype=np.int64)
        norm = action_oracle.decode_token_ids_to_actions(ids, g["vocab_size"])
        assert np.array_equal(norm, _hex(case["normalized_hex"]))
        assert np.array_equal(action_oracle.unnormalize(norm, stats), _hex(case["actions_hex"]))
        assert np.array_equal(action_oracle.unnormalize(norm, nomask), _hex(case["actions_nomask_hex"]))


def test_survey_known_answers(golden_dir):
    # ids → bin idx [127,0,254,254,99,189,249]; ids 31744 and 31745 collide; pad id → idx 0
    ids = np.array([31872, 31999, 31744, 31745, 31900, 31810, 31750])
    _, centers = action_oracle.make_bins()
    got = action_oracle.decode_token_ids_to_actions(ids, 32000)
    assert np.array_equal(got, centers[[127, 0, 254, 254, 99, 189, 249]])
    assert action_oracle.decode_token_ids_to_actions(np.array([32000, 32063]), 32000).tolist() == [centers[0]] * 2
    assert action_oracle.decode_token_ids_to_actions(np.array([5, 31745]), 32000).tolist() == [centers[254]] * 2


def test_encode_ids(golden_dir):
    g = json.loads((golden_dir / "action_tokenizer.json").read_text())
    got = action_oracle.encode_actions_to_token_ids(np.array(g["encode"]["actions"]), g["vocab_size"])
    assert got.tolist() == g["encode"]["ids"]


def test_greedy_first_index_and_nan():
    x = np.zeros((3, 10))
    x[0, [3, 7]] = 5.0
    x[1, 9] = 1.0
    x[2, 4] = np.nan
    x[2, 2] = 100.0
    assert action_oracle.greedy_token_ids(x).tolist() == [3, 9, 4]
    t = torch.from_numpy(x)
    assert torch.argmax(t, dim=-1).tolist() == [3, 9, 4]


def test_projector_matches_reference_module(golden_dir):
    z = np.load(golden_dir / "projector_small.npz")
    sd = {k.replace("__", "."): torch.from_numpy(z[k]) for k in z.files if k not in ("x", "y")}
    y = vit_oracle.projector_forward(sd, torch.from_numpy(z["x"]))
    assert torch.equal(y, torch.from_numpy(z["y"]))


def test_action_token_metrics_restatement_on_a_hand_computed_case():
    """base_strategy.py:314-329: 2 samples, 3 patches, 5 text positions; vocab 40 with 8 bins → begin_idx = 31."""
    import torch
    P, L, V = 3, 5, 40
    logits = torch.zeros(2, P + L, V)
    labels = torch.tensor([[-100, 33, 39, 2, -100], [-100, -100, 32, 31, 36]])     # 31 == begin_idx → not an action token
    # position P+j-1 predicts label j (labels[:, 1:] vs logits[:, P:-1])
    logits[0, P + 0, 33] = 9.0          # correct
    logits[0, P + 1, 38] = 9.0          # wrong: 38 vs 39 → adjacent bins
    logits[1, P + 1, 32] = 9.0          # correct
    logits[1, P + 3, 36] = 9.0          # correct
    acc, l1, preds, mask = action_oracle.action_token_metrics(logits, labels, P, V, n_bins=8)
    assert mask.tolist() == [[True, True, False, False], [False, True, False, True]]
    assert float(acc) == 0.75
    _, centers = action_oracle.make_bins(8)
    want = abs(centers[min(max(V - 38 - 1, 0), 6)] - centers[min(max(V - 39 - 1, 0), 6)]) / 4
    assert abs(float(l1) - want) < 1e-15


def test_preprocess_lut_matches_torchvision_on_every_uint8_value():
    """The table `blb_preprocess_u8` gathers from == ToTensor + Normalize + bf16 cast of the reference transform."""
    import numpy as np
    import torch
    from PIL import Image
    from torchvision import transforms as T

    from bridgelang_b200.vision import make_preprocess_lut
    from bridgelang_b200.weights import DINO_MEAN, DINO_STD, SIGLIP_MEAN, SIGLIP_STD

    lut = make_preprocess_lut("cpu")
    assert lut.shape == (2, 3, 256) and lut.dtype == torch.bfloat16
    ramp = np.tile(np.arange(256, dtype=np.uint8)[None, :, None], (2, 1, 3))          # [2,256,3] HWC image
    for t, (mean, std) in enumerate(((DINO_MEAN, DINO_STD), (SIGLIP_MEAN, SIGLIP_STD))):
        tf = T.Compose([T.ToTensor(), T.Normalize(mean=torch.tensor(mean), std=torch.tensor(std))])
        ref = tf(Image.fromarray(ramp)).to(torch.bfloat16)                            # [3,2,256]
        assert torch.equal(lut[t], ref[:, 0, :])


def test_tower_oracle_against_frozen_hf_outputs(golden_dir):
    """oracle/vit_oracle.py vs tests/golden/hf_towers_depth3.npz: outputs of transformers' independent DINOv2-reg4 and
    SigLIP implementations, frozen by tests/golden/make_hf_tower_golden.py (timm itself cannot be installed offline)."""
    import numpy as np
    import torch
    from bridgelang_b200.config import DINOV2_L14_REG4, SIGLIP_SO400M_14
    from bridgelang_b200.weights import make_vit_state_dict
    from oracle import vit_oracle

    g = np.load(golden_dir / "hf_towers_depth3.npz")
    depth = int(g["depth"])
    for name, cfg0 in (("dino", DINOV2_L14_REG4), ("siglip", SIGLIP_SO400M_14)):
        cfg = cfg0.with_depth(depth)
        sd = make_vit_state_dict(cfg, seed=int(g[f"{name}_weight_seed"]), init="stress")
        x = torch.randn(1, 3, 224, 224, generator=torch.Generator().manual_seed(int(g[f"{name}_pixel_seed"])))
        with torch.no_grad():
            got = vit_oracle.vit_intermediate(sd, cfg, x)
        ref = torch.from_numpy(g[f"{name}_slice"])
        scale = ref.abs().max()
        assert ((got[0, ::8, ::16] - ref).abs().max() / scale).item() < 2e-5, name
        assert abs(got.double().sum().item() - float(g[f"{name}_sum"])) < 1e-4 * float(g[f"{name}_abs_sum"])
        assert abs(got.double().abs().sum().item() / float(g[f"{name}_abs_sum"]) - 1) < 1e-5


def test_tower_oracle_against_frozen_hf_outputs_full_depth_batch2(golden_dir):
    """Round 2 (VERDICT r01 "next" 1e): the same pin at the depth the product runs — 24 / 27 blocks, the weights of the
    parity gates (seeds 1234 / 1235, stress-init) and a batch of two frames — against transformers' independent
    implementations frozen in tests/golden/hf_towers_full_depth.npz.  The GPU parity tests compare the kernels with this
    very oracle on these very weights, so the chain  kernels == oracle == HF  closes at full depth."""
    import numpy as np
    import torch
    from bridgelang_b200.config import DINOV2_L14_REG4, SIGLIP_SO400M_14
    from bridgelang_b200.weights import make_vit_state_dict
    from oracle import vit_oracle

    g = np.load(golden_dir / "hf_towers_full_depth.npz")
    for name, cfg in (("dino", DINOV2_L14_REG4), ("siglip", SIGLIP_SO400M_14)):
        assert int(g[f"{name}_depth"]) == cfg.depth
        sd = make_vit_state_dict(cfg, seed=int(g[f"{name}_weight_seed"]), init="stress")
        x = torch.randn(2, 3, 224, 224, generator=torch.Generator().manual_seed(int(g[f"{name}_pixel_seed"])))
        with torch.no_grad():
            got = vit_oracle.vit_intermediate(sd, cfg, x)
        ref = torch.from_numpy(g[f"{name}_slice"])
        assert got.shape == (2, 256, cfg.dim) and ref.shape == got[:, ::8, ::16].shape
        assert ((got[:, ::8, ::16] - ref).abs().max() / ref.abs().max()).item() < 5e-5, name
        assert abs(got.double().sum().item() - float(g[f"{name}_sum"])) < 1e-4 * float(g[f"{name}_abs_sum"])
        assert abs(got.double().abs().sum().item() / float(g[f"{name}_abs_sum"]) - 1) < 1e-5
        # the two images differ (a batch-index bug would not hide behind identical frames)
        assert (ref[0] - ref[1]).abs().max() > 0.05 * ref.abs().max()
