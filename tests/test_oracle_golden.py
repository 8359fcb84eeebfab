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
