"""vla.py — `OpenVLA.predict_action` with the visual prefix and the decode tail on the B200-native kernels.

Mirrors prismatic/models/vlas/openvla.py:23-131 (native) and the HF twin
prismatic/extern/hf/modeling_prismatic.py:492-562: same `predict_action(image, instruction, unnorm_key)` contract,
same assertion messages in `_check_unnorm_key`, same `get_action_dim` / `get_action_stats`.

What changed underneath (and only that):
  * vision_backbone + projector run as one fused native call and the projector's fc3 epilogue stores the 256
    prefix rows straight into the LLM's `inputs_embeds` buffer at token offset 1 (the `torch.cat` splice of
    prismatic.py:389-396 disappears).  The prefix is computed ONCE per action — the fork's demo runs
    `use_cache=False` (run_openvla_demo.py:43) and re-runs the towers for each of the 7 decode steps.
  * greedy decoding uses a device argmax over the FULL logits row (first max index wins, like torch.argmax inside
    transformers' GenerationMixin), and the final ids → bin centres → q01/q99 un-normalize happen in one kernel in
    float64; a single device→host copy returns the action vector.
The language model is untouched: any HF-style causal LM (`get_input_embeddings()`, `forward(inputs_embeds=...,
past_key_values=..., use_cache=True)` → `.logits`, `.past_key_values`) is driven as a black box.
"""

from __future__ import annotations

from typing import Any, Dict, List, Optional

import numpy as np
import torch
import torch.nn as nn

from . import ops
from .action_tokenizer import ActionTokenizer
from .config import NUM_PATCHES
from .pipeline import VisualPrefixEncoder


class PurePromptBuilder:
    """prismatic/models/backbones/llm/prompting/base_prompter.py:28-75 (`llama2-7b-pure`)."""

    def __init__(self, model_family: str = "openvla", system_prompt: Optional[str] = None) -> None:
        self.model_family, self.system_prompt = model_family, system_prompt
        self.bos, self.eos = "<s>", "</s>"
        self.prompt, self.turn_count = "", 0

    def add_turn(self, role: str, message: str) -> str:
        assert (role == "human") if (self.turn_count % 2 == 0) else (role == "gpt")
        message = message.replace("<image>", "").strip()
        if (self.turn_count % 2) == 0:
            wrapped = f"In: {message}\nOut: "
        else:
            wrapped = f"{message if message != '' else ' '}{self.eos}"
        self.prompt += wrapped
        self.turn_count += 1
        return wrapped

    def get_prompt(self) -> str:
        return self.prompt.removeprefix(self.bos).rstrip()


class LLMBackbone(nn.Module):
    """What this path needs of prismatic/models/backbones/llm/base_llm.py::LLMBackbone (the LLM itself is out of
    scope and is driven as a black box): `.llm` (an HF-style causal LM), `.tokenizer` / `get_tokenizer()`, `.embed_dim`,
    `.half_precision_dtype`, `.prompt_builder_fn`, `.identifier`."""

    def __init__(self, llm_backbone_id: str, llm: nn.Module, tokenizer: Any,
                 half_precision_dtype: torch.dtype = torch.bfloat16) -> None:
        super().__init__()
        self.identifier = llm_backbone_id
        self.llm = llm
        self.tokenizer = tokenizer
        self._half_precision_dtype = half_precision_dtype

    def get_tokenizer(self):
        return self.tokenizer

    @property
    def embed_dim(self) -> int:
        return int(self.llm.config.hidden_size)

    @property
    def half_precision_dtype(self) -> torch.dtype:
        return self._half_precision_dtype

    @property
    def prompt_builder_fn(self):
        return PurePromptBuilder


def _is_llama_tokenizer_fast(tokenizer: Any) -> bool:
    try:
        from transformers import LlamaTokenizerFast
    except Exception:                               # transformers absent: nothing can be a LlamaTokenizerFast
        return False
    return isinstance(tokenizer, LlamaTokenizerFast)


class OpenVLA(nn.Module):
    """prismatic/models/vlas/openvla.py:23-33 on top of prismatic/models/vlms/prismatic.py:40-92 — same construction:

        OpenVLA(model_id, vision_backbone, llm_backbone, enable_mixed_precision_training=True, arch_specifier=...,
                norm_stats=..., action_tokenizer=...)
        OpenVLA.from_pretrained(checkpoint_pt, model_id, vision_backbone, llm_backbone, arch_specifier=...,
                                freeze_weights=True, norm_stats=..., action_tokenizer=...)      # load.py:214-224

    `vision_backbone` is a (native) VisionBackbone, `llm_backbone` anything with the LLMBackbone contract above.  The
    projector is built from `arch_specifier` exactly where the reference builds it (prismatic.py:60-68); this path
    implements the fused three-layer projector, i.e. every `*fused-gelu-mlp` specifier (OpenVLA's is
    `no-align+fused-gelu-mlp`) — `linear` / `gelu-mlp` raise.  `from_components` assembles the same object from an
    already built projector and a bare HF causal LM (tests, the HF twin)."""

    def __init__(self, model_id: str, vision_backbone: nn.Module, llm_backbone: nn.Module,
                 enable_mixed_precision_training: bool = True, arch_specifier: str = "gelu-mlp", *,
                 norm_stats: Dict[str, Dict[str, Dict[str, List[float]]]], action_tokenizer: ActionTokenizer,
                 empty_token_id: int = 29871, **kwargs: Any) -> None:
        super().__init__()
        from .projector import FusedMLPProjector
        self.model_family, self.model_id = "prismatic", model_id
        self.vision_backbone, self.llm_backbone = vision_backbone, llm_backbone
        self.enable_mixed_precision_training = enable_mixed_precision_training
        torch.manual_seed(vision_backbone.embed_dim)            # prismatic.py:57 (projector init consistency)
        self.arch_specifier = arch_specifier
        if arch_specifier.endswith("fused-gelu-mlp"):
            self.projector = FusedMLPProjector(vision_backbone.embed_dim, llm_backbone.embed_dim)
        elif arch_specifier == "linear" or arch_specifier.endswith("gelu-mlp"):
            raise ValueError(f"PrismaticVLM with `{arch_specifier = }` is not on the B200-native path (it implements "
                             f"the fused three-layer projector of the fused DINOv2+SigLIP backbones)")
        else:
            raise ValueError(f"PrismaticVLM with `{arch_specifier = }` is not supported!")
        self.vision_backbone_requires_grad = False
        self.all_module_keys = ["vision_backbone", "llm_backbone", "projector"]
        self.trainable_module_keys: List[str] = []
        self.norm_stats = norm_stats
        self.action_tokenizer = action_tokenizer
        self.empty_token_id = empty_token_id
        self._prefix_encoder: Optional[VisualPrefixEncoder] = None

    # -- construction paths ------------------------------------------------------------------------------
    @classmethod
    def from_pretrained(cls, pretrained_checkpoint, model_id: str, vision_backbone: nn.Module, llm_backbone: nn.Module,
                        enable_mixed_precision_training: bool = True, arch_specifier: str = "gelu-mlp",
                        freeze_weights: bool = True, **kwargs: Any) -> "OpenVLA":
        """prismatic.py:85-123: build, then load `["model"]["projector" | "llm_backbone" | ("vision_backbone")]`."""
        vlm = cls(model_id, vision_backbone, llm_backbone, enable_mixed_precision_training=enable_mixed_precision_training,
                  arch_specifier=arch_specifier, **kwargs)
        model_state_dict = torch.load(pretrained_checkpoint, map_location="cpu")["model"]
        assert (
            "projector" in model_state_dict and "llm_backbone" in model_state_dict
        ), "PrismaticVLM `from_pretrained` expects checkpoint with keys for `projector` AND `llm_backbone`!"
        vlm.projector.load_state_dict(model_state_dict["projector"])
        vlm.llm_backbone.load_state_dict(model_state_dict["llm_backbone"])
        if "vision_backbone" in model_state_dict.keys():
            vlm.vision_backbone.load_state_dict(model_state_dict["vision_backbone"])
        if freeze_weights:
            vlm.requires_grad_(False)
            vlm.eval()
        return vlm

    @classmethod
    def from_components(cls, vision_backbone: nn.Module, projector: nn.Module, llm: nn.Module, tokenizer: Any,
                        norm_stats: Dict, action_tokenizer: Optional[ActionTokenizer] = None,
                        model_id: str = "openvla", llm_backbone_id: str = "llama2-7b-pure") -> "OpenVLA":
        backbone = llm if isinstance(llm, LLMBackbone) else LLMBackbone(llm_backbone_id, llm, tokenizer)
        vla = cls.__new__(cls)
        nn.Module.__init__(vla)
        vla.model_family, vla.model_id = "prismatic", model_id
        vla.vision_backbone, vla.llm_backbone, vla.projector = vision_backbone, backbone, projector
        vla.enable_mixed_precision_training, vla.arch_specifier = True, "no-align+fused-gelu-mlp"
        vla.vision_backbone_requires_grad = False
        vla.all_module_keys, vla.trainable_module_keys = ["vision_backbone", "llm_backbone", "projector"], []
        vla.norm_stats = norm_stats
        vla.action_tokenizer = action_tokenizer if action_tokenizer is not None else ActionTokenizer(tokenizer)
        vla.empty_token_id = 29871
        vla._prefix_encoder = None
        return vla

    # -- reference attribute names -----------------------------------------------------------------------
    @property
    def llm(self) -> nn.Module:
        return self.llm_backbone.llm

    @property
    def tokenizer(self):
        return self.llm_backbone.tokenizer

    @property
    def device(self) -> torch.device:
        return next(self.llm.parameters()).device

    @property
    def prefix_encoder(self) -> VisualPrefixEncoder:
        if self._prefix_encoder is None:
            object.__setattr__(self, "_prefix_encoder", VisualPrefixEncoder(self.vision_backbone, self.projector))
        return self._prefix_encoder

    def get_prompt_builder(self, system_prompt: Optional[str] = None) -> PurePromptBuilder:
        return self.llm_backbone.prompt_builder_fn("prismatic", system_prompt=system_prompt)

    # -- the part of predict_action between the tokenizer and generate() --------------------------------
    def _prepare_input_ids(self, instruction: str, device: torch.device) -> torch.Tensor:
        tokenizer = self.llm_backbone.tokenizer
        prompt_builder = self.get_prompt_builder()
        prompt_builder.add_turn(role="human", message=f"What action should the robot take to {instruction.lower()}?")
        prompt_text = prompt_builder.get_prompt()
        input_ids = tokenizer(prompt_text, truncation=True, return_tensors="pt").input_ids.to(device)
        if _is_llama_tokenizer_fast(tokenizer):                      # openvla.py:57-66
            if not torch.all(input_ids[:, -1] == self.empty_token_id):
                extra = torch.tensor([[self.empty_token_id]], dtype=torch.long, device=device)
                input_ids = torch.cat((input_ids, extra), dim=1)
        else:
            raise ValueError(f"Unsupported `tokenizer` type = {type(tokenizer)}")
        return input_ids

    @torch.inference_mode()
    def generate_action_token_ids(self, input_ids: torch.Tensor, pixel_values, max_new_tokens: int) -> torch.Tensor:
        """Greedy decode of `max_new_tokens` ids (batch size 1, as modeling_prismatic.py:460-463 requires)."""
        assert input_ids.shape[0] == 1, "predict_action supports batch size 1 (as the reference does)"
        emb = self.llm.get_input_embeddings()
        tok = emb(input_ids)                                              # [1, T, llm_dim]
        T, dim = tok.shape[1], tok.shape[2]
        n_patch = int(getattr(self.vision_backbone, "num_patches", NUM_PATCHES))   # 256 / 576 / 729
        embeds = torch.empty((1, 1 + n_patch + (T - 1), dim), dtype=torch.bfloat16, device=tok.device)
        embeds[:, :1] = tok[:, :1]                                        # <BOS>
        embeds[:, 1 + n_patch:] = tok[:, 1:]
        # visual prefix, written by fc3's epilogue at token offset 1 (prismatic.py:389-396 splice)
        feats = self.vision_backbone(pixel_values)      # dict (native backbone) or packed [1,6,224,224] (HF twin)
        self.projector.project(feats, out=embeds, tok_in=n_patch, tok_out=embeds.shape[1], tok_shift=1)
        out = self.llm(inputs_embeds=embeds.to(tok.dtype), use_cache=True)
        ids: List[torch.Tensor] = []
        for step in range(max_new_tokens):
            logits = out.logits[:, -1, :]
            nxt = ops.argmax(logits if logits.stride(-1) == 1 else logits.contiguous())   # full-row, first max wins
            ids.append(nxt)
            if step + 1 < max_new_tokens:
                out = self.llm(inputs_embeds=emb(nxt.view(1, 1)), past_key_values=out.past_key_values, use_cache=True)
        return torch.cat(ids)

    # generate() arguments the greedy device loop honours (everything else the reference would forward to
    # GenerationMixin.generate is rejected instead of being dropped on the floor)
    _GENERATE_KWARGS = {"do_sample": False, "use_cache": None, "num_beams": 1, "temperature": None, "top_p": None,
                        "top_k": None, "pad_token_id": None, "eos_token_id": None}

    @classmethod
    def _check_generate_kwargs(cls, kwargs: Dict[str, Any]) -> None:
        for k, v in kwargs.items():
            if k not in cls._GENERATE_KWARGS:
                raise ValueError(f"predict_action: unsupported generate() argument `{k}` on the native greedy decode path")
            want = cls._GENERATE_KWARGS[k]
            if want is not None and v is not None and v != want:
                raise ValueError(f"predict_action: `{k}={v}` is not supported (the native decode tail is greedy: "
                                 f"`{k}={want}`, as the reference's scripts run it)")

    @torch.inference_mode()
    def predict_action(self, image, instruction: str, unnorm_key: Optional[str] = None, **kwargs: str) -> np.ndarray:
        """PIL image + instruction → un-normalized continuous action, np.float64[action_dim]."""
        self._check_generate_kwargs(kwargs)
        device = self.device
        input_ids = self._prepare_input_ids(instruction, device)
        pixel_values = self.vision_backbone.get_image_transform()(image)
        if isinstance(pixel_values, torch.Tensor):
            pixel_values = pixel_values[None, ...].to(device)
        elif isinstance(pixel_values, dict):
            pixel_values = {k: v[None, ...].to(device) for k, v in pixel_values.items()}
        else:
            raise ValueError(f"Unsupported `pixel_values` type = {type(pixel_values)}")
        action_dim = self.get_action_dim(unnorm_key)
        ids = self.generate_action_token_ids(input_ids, pixel_values, action_dim)
        _, actions = self.action_tokenizer.decode_on_device(ids, self.get_action_stats(unnorm_key))
        return actions.cpu().numpy()

    # -- openvla.py:105-131, verbatim semantics --------------------------------------------------------
    @staticmethod
    def _check_unnorm_key(norm_stats: Dict, unnorm_key: Optional[str]) -> str:
        if unnorm_key is None:
            assert len(norm_stats) == 1, (
                f"Your model was trained on more than one dataset, please pass a `unnorm_key` from the following "
                f"options to choose the statistics used for un-normalizing actions: {norm_stats.keys()}"
            )
            unnorm_key = next(iter(norm_stats.keys()))
        assert (
            unnorm_key in norm_stats
        ), f"The `unnorm_key` you chose is not in the set of available statistics; choose from: {norm_stats.keys()}"
        return unnorm_key

    def get_action_dim(self, unnorm_key: Optional[str] = None) -> int:
        unnorm_key = self._check_unnorm_key(self.norm_stats, unnorm_key)
        return len(self.norm_stats[unnorm_key]["action"]["q01"])

    def get_action_stats(self, unnorm_key: Optional[str] = None) -> Dict:
        unnorm_key = self._check_unnorm_key(self.norm_stats, unnorm_key)
        return self.norm_stats[unnorm_key]["action"]


class OpenVLAForActionPrediction(nn.Module):
    """HF twin, prismatic/extern/hf/modeling_prismatic.py:492-562 — the class every script of the fork drives
    (run_openvla_demo.py:37-44, vla-scripts/deploy.py:104-105, experiments/robot/openvla_utils.py:166-169):

        inputs = processor(prompt, image).to(device, dtype=torch.bfloat16)
        action = vla.predict_action(**inputs, unnorm_key="bridge_orig", do_sample=False)

    Same call: `predict_action(input_ids, unnorm_key=None, **kwargs)` with `pixel_values` [1,6,224,224] (dino channels
    first, processing_prismatic.py:143) among the kwargs; `attention_mask` / `do_sample=False` / `use_cache` are
    accepted (greedy decoding is what the reference runs, modeling_prismatic.py:518 with do_sample=False).  Attributes
    `norm_stats`, `bins`, `bin_centers`, `vocab_size` (= text vocab − pad_to_multiple_of, :504) as in the reference;
    same assertion messages.  The visual prefix is computed once per action whatever `use_cache` says."""

    def __init__(self, vision_backbone: nn.Module, projector: nn.Module, language_model: nn.Module,
                 norm_stats: Dict[str, Dict[str, Any]], text_vocab_size: int = 32064, pad_to_multiple_of: int = 64,
                 n_action_bins: int = 256) -> None:
        super().__init__()
        self.vision_backbone, self.projector, self.language_model = vision_backbone, projector, language_model
        self.norm_stats = norm_stats
        self.bins = np.linspace(-1, 1, n_action_bins)
        self.bin_centers = (self.bins[:-1] + self.bins[1:]) / 2.0
        self.vocab_size = text_vocab_size - pad_to_multiple_of

        class _Vocab:                               # ActionTokenizer only needs `.vocab_size` for the decode path
            vocab_size = self.vocab_size
        self._action_tokenizer = ActionTokenizer(_Vocab(), bins=n_action_bins)
        # shares the greedy loop with the native class; kept out of the module tree so that state_dict() has the
        # reference's keys only (vision_backbone.*, projector.*, language_model.*)
        object.__setattr__(self, "_core", OpenVLA.from_components(vision_backbone, projector, language_model, _Vocab(),
                                                                  norm_stats, self._action_tokenizer))

    @torch.inference_mode()
    def predict_action(self, input_ids: Optional[torch.LongTensor] = None, unnorm_key: Optional[str] = None,
                       **kwargs: Any) -> np.ndarray:
        pixel_values = kwargs.get("pixel_values")
        if input_ids is None or pixel_values is None:
            raise ValueError("predict_action needs `input_ids` and `pixel_values` (the processor's outputs)")
        OpenVLA._check_generate_kwargs({k: v for k, v in kwargs.items() if k not in ("pixel_values", "attention_mask")})
        if not torch.all(input_ids[:, -1] == 29871):                      # modeling_prismatic.py:512-515
            input_ids = torch.cat(
                (input_ids, torch.unsqueeze(torch.Tensor([29871]).long(), dim=0).to(input_ids.device)), dim=1)
        action_dim = self.get_action_dim(unnorm_key)
        ids = self._core.generate_action_token_ids(input_ids, pixel_values, action_dim)
        _, actions = self._action_tokenizer.decode_on_device(ids, self.get_action_stats(unnorm_key))
        return actions.cpu().numpy()

    @staticmethod
    def _check_unnorm_key(norm_stats: Dict[str, Dict[str, Any]], unnorm_key: Optional[str]) -> str:
        if unnorm_key is None:
            assert len(norm_stats) == 1, (
                f"Your model was trained on more than one dataset, "
                f"please pass a `unnorm_key` from the following options to choose the statistics "
                f"used for un-normalizing actions: {norm_stats.keys()}"
            )
            unnorm_key = next(iter(norm_stats.keys()))

        assert unnorm_key in norm_stats, (
            f"The `unnorm_key` you chose is not in the set of available dataset statistics, "
            f"please choose from: {norm_stats.keys()}"
        )
        return unnorm_key

    def get_action_dim(self, unnorm_key: Optional[str] = None) -> int:
        unnorm_key = self._check_unnorm_key(self.norm_stats, unnorm_key)
        return len(self.norm_stats[unnorm_key]["action"]["q01"])

    def get_action_stats(self, unnorm_key: Optional[str] = None) -> Dict[str, Any]:
        unnorm_key = self._check_unnorm_key(self.norm_stats, unnorm_key)
        return self.norm_stats[unnorm_key]["action"]


def decode_tail_from_logits(logits: torch.Tensor, action_tokenizer: ActionTokenizer, stats: Optional[Dict]):
    """[steps, vocab_rows] logits → (ids, normalized, actions) in one launch (argmax → de-tokenize → un-normalize).
    Equivalent to torch.argmax per step followed by openvla.py:89-101 on identical logits."""
    tables = action_tokenizer.tables(stats, device=logits.device)
    return ops.argmax_detokenize_unnormalize(logits, int(action_tokenizer.tokenizer.vocab_size), tables)
