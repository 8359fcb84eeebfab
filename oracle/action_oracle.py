"""action_oracle.py — NumPy restatement of the decode tail.  TEST INFRASTRUCTURE ONLY.

Pinned: tests/golden/action_tokenizer.json was produced by executing the reference's own
prismatic/vla/action_tokenizer.py (loaded by file path) — tests/test_oracle_golden.py checks this file against it.

  ActionTokenizer.__init__                 prismatic/vla/action_tokenizer.py:28-36
  ActionTokenizer.__call__ (ids only)      prismatic/vla/action_tokenizer.py:38-47
  decode_token_ids_to_actions              prismatic/vla/action_tokenizer.py:49-68
  un-normalize                             prismatic/models/vlas/openvla.py:94-101
                                           (HF twin: prismatic/extern/hf/modeling_prismatic.py:521-534)
  greedy step                              torch.argmax over the full logits row (transformers GenerationMixin,
                                           called at openvla.py:81-86): first maximal index wins
"""

from __future__ import annotations

from typing import Dict, Optional

import numpy as np


def make_bins(n_bins: int = 256, min_action: float = -1, max_action: float = 1):
    bins = np.linspace(min_action, max_action, n_bins)
    return bins, (bins[:-1] + bins[1:]) / 2.0


def action_token_begin_idx(vocab_size: int, n_bins: int = 256) -> int:
    return int(vocab_size - (n_bins + 1))


def encode_actions_to_token_ids(action: np.ndarray, vocab_size: int, n_bins: int = 256, min_action: float = -1,
                                max_action: float = 1) -> np.ndarray:
    bins, _ = make_bins(n_bins, min_action, max_action)
    action = np.clip(action, a_min=float(min_action), a_max=float(max_action))
    return vocab_size - np.digitize(action, bins)


def decode_token_ids_to_actions(action_token_ids: np.ndarray, vocab_size: int, n_bins: int = 256) -> np.ndarray:
    _, bin_centers = make_bins(n_bins)
    discretized = vocab_size - action_token_ids
    discretized = np.clip(discretized - 1, a_min=0, a_max=bin_centers.shape[0] - 1)
    return bin_centers[discretized]


def unnormalize(normalized_actions: np.ndarray, action_norm_stats: Dict) -> np.ndarray:
    mask = action_norm_stats.get("mask", np.ones_like(action_norm_stats["q01"], dtype=bool))
    action_high, action_low = np.array(action_norm_stats["q99"]), np.array(action_norm_stats["q01"])
    return np.where(mask, 0.5 * (normalized_actions + 1) * (action_high - action_low) + action_low,
                    normalized_actions)


def greedy_token_ids(logits: np.ndarray) -> np.ndarray:
    """argmax over the last axis; NaN counts as maximal and the first maximal index wins (torch.argmax)."""
    logits = np.asarray(logits, dtype=np.float64)
    out = np.empty(logits.shape[0], dtype=np.int64)
    for r in range(logits.shape[0]):
        row = logits[r]
        nan = np.isnan(row)
        out[r] = int(np.argmax(nan)) if nan.any() else int(np.argmax(row))
    return out


def decode_tail(logits: np.ndarray, vocab_size: int, stats: Optional[Dict]) -> Dict[str, np.ndarray]:
    ids = greedy_token_ids(logits)
    norm = decode_token_ids_to_actions(ids, vocab_size)
    act = unnormalize(norm, stats) if stats is not None else norm
    return {"ids": ids, "normalized": norm, "actions": act}


def action_token_metrics(logits, labels, num_patches: int, vocab_size: int, n_bins: int = 256):
    """prismatic/training/strategies/base_strategy.py:314-329 (== vla-scripts/finetune.py:270-286), torch on CPU.
    logits [B, P+L, V] float tensor, labels int64 [B, L].  Returns (accuracy float32, l1 float64, preds, mask)."""
    import torch

    begin = action_token_begin_idx(vocab_size, n_bins)
    action_preds = logits[:, num_patches:-1].argmax(dim=2)
    action_gt = labels[:, 1:]
    mask = action_gt > begin
    correct_preds = (action_preds == action_gt) & mask
    action_accuracy = correct_preds.sum().float() / mask.sum().float()
    continuous_actions_pred = torch.tensor(decode_token_ids_to_actions(action_preds[mask].cpu().numpy(), vocab_size, n_bins))
    continuous_actions_gt = torch.tensor(decode_token_ids_to_actions(action_gt[mask].cpu().numpy(), vocab_size, n_bins))
    action_l1_loss = torch.nn.functional.l1_loss(continuous_actions_pred, continuous_actions_gt)
    return action_accuracy, action_l1_loss, action_preds, mask
