"""action_tokenizer.py — mirror of prismatic/vla/action_tokenizer.py:13-72 with the decode path on the device.

Same constructor, attributes (`bins`, `bin_centers`, `action_token_begin_idx`, `n_bins`, `vocab_size`) and call
signatures.  `decode_token_ids_to_actions` keeps the reference's NumPy-in / NumPy-out contract but the arithmetic
(id → bin → centre, clip semantics included) runs in the CUDA kernel of csrc/decode_tail.cu and is bit-identical in
float64; `decode_on_device` is the no-copy variant `predict_action` uses.  `__call__` (np.digitize + tokenizer.decode
to a string) is host-side text processing exactly as in the reference; `encode_on_device` gives its token ids on the
GPU and `action_metrics` the training loops' accuracy / L1.
"""

from __future__ import annotations

from typing import Dict, List, Optional, Tuple, Union

import numpy as np
import torch

from . import ops


class ActionTokenizer:
    def __init__(self, tokenizer, bins: int = 256, min_action: int = -1, max_action: int = 1) -> None:
        self.tokenizer, self.n_bins, self.min_action, self.max_action = tokenizer, bins, min_action, max_action
        self.bins = np.linspace(min_action, max_action, self.n_bins)
        self.bin_centers = (self.bins[:-1] + self.bins[1:]) / 2.0
        self.action_token_begin_idx: int = int(self.tokenizer.vocab_size - (self.n_bins + 1))
        self._tables: Dict[Tuple, ops.DecodeTables] = {}

    def __call__(self, action: np.ndarray) -> Union[str, List[str]]:
        action = np.clip(action, a_min=float(self.min_action), a_max=float(self.max_action))
        discretized_action = np.digitize(action, self.bins)
        if len(discretized_action.shape) == 1:
            return self.tokenizer.decode(list(self.tokenizer.vocab_size - discretized_action))
        else:
            return self.tokenizer.batch_decode((self.tokenizer.vocab_size - discretized_action).tolist())

    # -- device decode -------------------------------------------------------------------------------
    def tables(self, stats: Optional[Dict] = None, device: str = "cuda") -> ops.DecodeTables:
        """Device tables for one dataset's statistics (`norm_stats[key]["action"]`), cached."""
        if stats is None:
            key = (str(device), None)
            q01 = q99 = mask = None
        else:
            q01, q99 = np.array(stats["q01"], dtype=np.float64), np.array(stats["q99"], dtype=np.float64)
            mask = np.array(stats.get("mask", np.ones_like(stats["q01"], dtype=bool)), dtype=bool)
            key = (str(device), q01.tobytes(), q99.tobytes(), mask.tobytes())
        if key not in self._tables:
            self._tables[key] = ops.DecodeTables(self.bin_centers, q01, q99, mask, device=device)
        return self._tables[key]

    def decode_on_device(self, action_token_ids: torch.Tensor, stats: Optional[Dict] = None):
        """int64 CUDA ids [n] → (normalized, actions) float64 CUDA tensors; actions == normalized if stats is None."""
        t = self.tables(stats, device=action_token_ids.device)
        return ops.detokenize_unnormalize(action_token_ids.contiguous().view(-1), int(self.tokenizer.vocab_size), t)

    # -- device encode / training metrics --------------------------------------------
    def encode_on_device(self, action: torch.Tensor) -> torch.Tensor:
        """`__call__` up to the token ids (clip → np.digitize → vocab_size − index) for a CUDA float32/float64 tensor of
        any shape; the id → string step (`tokenizer.decode`) is text processing and stays on the host."""
        if getattr(self, "_bins_dev", None) is None or self._bins_dev.device != action.device:
            self._bins_dev = torch.from_numpy(np.ascontiguousarray(self.bins, dtype=np.float64)).to(action.device)
        return ops.encode_actions(action, self._bins_dev, float(self.min_action), float(self.max_action),
                                  int(self.tokenizer.vocab_size))

    def action_metrics(self, logits: torch.Tensor, labels: torch.Tensor, num_patches: int):
        """Action-token accuracy and L1 of the training loops (base_strategy.py:314-329, finetune.py:270-286) without the
        `.cpu().numpy()` round trips.  Returns (accuracy float32 0-d, l1 float64 0-d) CUDA tensors."""
        r = ops.action_token_metrics(logits, labels, num_patches, self.action_token_begin_idx,
                                     int(self.tokenizer.vocab_size), self.tables(None, device=logits.device))
        correct, n = r["counts"][0], r["counts"][1]
        return correct.float() / n.float(), r["l1_sum"][0] / n.to(torch.float64)

    def decode_token_ids_to_actions(self, action_token_ids: np.ndarray) -> np.ndarray:
        ids_np = np.asarray(action_token_ids)
        ids = torch.from_numpy(np.ascontiguousarray(ids_np.reshape(-1), dtype=np.int64)).cuda()
        norm, _ = self.decode_on_device(ids, None)
        return norm.cpu().numpy().reshape(ids_np.shape)

    @property
    def vocab_size(self) -> int:
        return self.n_bins
