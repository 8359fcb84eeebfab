"""BASELINE config 5: end-to-end `predict_action` with a random-init Llama-style LLM on a synthetic frame.

New path: native visual prefix (fc3 epilogue stores straight into `inputs_embeds`), device argmax greedy loop,
de-tokenize + un-normalize kernel.  Reference-style path on the SAME logits source: torch.cat splice + HF
`generate(do_sample=False)` + the NumPy decode tail (oracle).  Action token ids and actions must be bit-exact."""

import numpy as np
import pytest
import torch

import bridgelang_b200 as blb
from bridgelang_b200.config import DINOV2_L14_REG4, SIGLIP_SO400M_14
from bridgelang_b200.weights import make_projector_state_dict, make_vit_state_dict
from oracle import action_oracle

pytestmark = pytest.mark.gpu

STATS = {"bridge_orig": {"action": {"q01": [-0.03, -0.04, -0.05, -0.08, -0.10, -0.20, 0.0],
                                    "q99": [0.03, 0.04, 0.05, 0.08, 0.10, 0.20, 1.0],
                                    "mask": [True] * 6 + [False]}}}


from transformers import LlamaTokenizerFast  # noqa: E402


class _Tok(LlamaTokenizerFast):
    """Llama-tokenizer stand-in: deterministic ids in [3, 31000), BOS=1 first, vocab_size 32000.  It IS-A
    LlamaTokenizerFast because predict_action, like the reference (openvla.py:57-66), refuses anything else; no tokenizer
    files exist offline, so the base constructor is bypassed."""
    vocab_size = 32000

    def __init__(self):
        pass

    def __call__(self, text, truncation=True, return_tensors="pt"):
        ids = [1] + [3 + (sum(map(ord, w)) * 7919) % 30000 for w in text.split()]

        class _Out:
            input_ids = torch.tensor([ids], dtype=torch.long)
        return _Out()


@pytest.fixture(scope="module")
def vla():
    transformers = pytest.importorskip("transformers")
    torch.manual_seed(0)
    cfg = transformers.LlamaConfig(vocab_size=32064, hidden_size=4096, intermediate_size=2048, num_hidden_layers=2,
                                   num_attention_heads=32, num_key_value_heads=32, max_position_embeddings=512)
    llm = transformers.LlamaForCausalLM(cfg).to(torch.bfloat16).cuda().eval()
    bb = blb.DinoSigLIPViTBackbone("dinosiglip-vit-so-224px", "resize-naive")
    # shallow towers keep the test fast; the arithmetic path is identical
    bb.dino_featurizer = blb.VisionTransformer(DINOV2_L14_REG4.with_depth(3))
    bb.siglip_featurizer = blb.VisionTransformer(SIGLIP_SO400M_14.with_depth(3))
    bb.dino_featurizer.load_state_dict(make_vit_state_dict(DINOV2_L14_REG4.with_depth(3), seed=1))
    bb.siglip_featurizer.load_state_dict(make_vit_state_dict(SIGLIP_SO400M_14.with_depth(3), seed=2))
    proj = blb.FusedMLPProjector(2176, 4096)
    proj.load_state_dict(make_projector_state_dict(seed=3))
    bb.cuda(), proj.cuda()
    # the reference's construction path (openvla.py:23-33 / load.py:214-224): model_id, vision_backbone, llm_backbone,
    # arch_specifier; norm_stats + action_tokenizer by keyword; the projector is built inside and then loaded
    llm_backbone = blb.LLMBackbone("llama2-7b-pure", llm, _Tok())
    vla = blb.OpenVLA("openvla-test", bb, llm_backbone, arch_specifier="no-align+fused-gelu-mlp", norm_stats=STATS,
                      action_tokenizer=blb.ActionTokenizer(llm_backbone.get_tokenizer()))
    vla.projector.load_state_dict(proj.state_dict())
    vla.projector.cuda()
    assert vla.llm is llm and vla.tokenizer is llm_backbone.tokenizer and vla.model_id == "openvla-test"
    return vla


def test_predict_action_matches_reference_style_pipeline(vla):
    from PIL import Image
    rng = np.random.default_rng(0)
    image = Image.fromarray((rng.random((256, 256, 3)) * 255).astype(np.uint8))   # verify_openvla.py:74
    instruction = "Pick up the red block"
    actions = vla.predict_action(image, instruction)             # unnorm_key=None: exactly one dataset
    assert isinstance(actions, np.ndarray) and actions.dtype == np.float64 and actions.shape == (7,)

    # ---- reference-style evaluation of the same model -------------------------------------------------
    with torch.inference_mode():
        ids_in = vla._prepare_input_ids(instruction, torch.device("cuda"))
        assert ids_in[0, -1].item() == 29871                      # openvla.py:59-64 empty-token rule
        px = {k: v[None].cuda() for k, v in vla.vision_backbone.get_image_transform()(image).items()}
        projected = vla.projector(vla.vision_backbone(px))         # [1,256,4096]
        emb = vla.llm.get_input_embeddings()(ids_in)
        spliced = torch.cat([emb[:, :1], projected.to(emb.dtype), emb[:, 1:]], dim=1)   # prismatic.py:389-396
        # (1) greedy loop with torch.argmax on the very same forward calls → identical logits → ids must be bit-equal
        out = vla.llm(inputs_embeds=spliced, use_cache=True)
        ref = []
        for step in range(7):
            nxt = torch.argmax(out.logits[:, -1, :], dim=-1)
            ref.append(nxt)
            if step < 6:
                out = vla.llm(inputs_embeds=vla.llm.get_input_embeddings()(nxt.view(1, 1)),
                              past_key_values=out.past_key_values, use_cache=True)
        ref_ids = torch.cat(ref).cpu().numpy()
        new_ids = vla.generate_action_token_ids(ids_in, px, 7).cpu().numpy()
        # (2) transformers' own GenerationMixin (what openvla.py:81-86 calls)
        gen = vla.llm.generate(inputs_embeds=spliced, max_new_tokens=7, do_sample=False, use_cache=True,
                               pad_token_id=0)
        hf_ids = gen[0, -7:].cpu().numpy()
    assert np.array_equal(new_ids, ref_ids)
    assert np.array_equal(new_ids, hf_ids)
    # the fc3-epilogue splice equals the torch.cat splice
    with torch.inference_mode():
        T = spliced.shape[1]
        embeds = torch.zeros((1, T, 4096), dtype=torch.bfloat16, device="cuda")
        vla.projector.project(vla.vision_backbone(px), out=embeds, tok_in=256, tok_out=T, tok_shift=1)
        assert torch.equal(embeds[:, 1:257], projected)
    want = action_oracle.unnormalize(action_oracle.decode_token_ids_to_actions(ref_ids, 32000),
                                     STATS["bridge_orig"]["action"])
    assert np.array_equal(actions, want)


def test_unnorm_key_errors(vla):
    from PIL import Image
    img = Image.fromarray(np.zeros((224, 224, 3), dtype=np.uint8))
    with pytest.raises(AssertionError, match="not in the set of available statistics"):
        vla.predict_action(img, "do something", unnorm_key="no_such_dataset")
    # generate() arguments: greedy ones are accepted, anything else is rejected instead of silently ignored
    a = vla.predict_action(img, "do something", do_sample=False, use_cache=True)
    assert a.shape == (7,)
    with pytest.raises(ValueError, match="do_sample"):
        vla.predict_action(img, "do something", do_sample=True)
    with pytest.raises(ValueError, match="unsupported generate"):
        vla.predict_action(img, "do something", repetition_penalty=1.2)
    # the reference refuses tokenizers that are not LlamaTokenizerFast (openvla.py:66)
    class _Other:
        vocab_size = 32000
        def __call__(self, text, truncation=True, return_tensors="pt"):
            class _Out:
                input_ids = torch.tensor([[1, 5, 6]], dtype=torch.long)
            return _Out()
    other = blb.OpenVLA.from_components(vla.vision_backbone, vla.projector, vla.llm, _Other(), STATS)
    with pytest.raises(ValueError, match="Unsupported `tokenizer` type"):
        other.predict_action(img, "do something")


def test_hf_twin_predict_action_equals_native(vla):
    """OpenVLAForActionPrediction.predict_action(input_ids, pixel_values=[1,6,224,224], unnorm_key, do_sample=False)
    — the call of run_openvla_demo.py:37-44 / deploy.py:104-105 — on the same weights gives the native class's action."""
    from PIL import Image
    bb = blb.PrismaticVisionBackbone(True, [224, 224], [DINOV2_L14_REG4.timm_id, SIGLIP_SO400M_14.timm_id], [None, None])
    bb.featurizer = blb.VisionTransformer(DINOV2_L14_REG4.with_depth(3), ls_param_name="scale_factor")
    bb.fused_featurizer = blb.VisionTransformer(SIGLIP_SO400M_14.with_depth(3), ls_param_name="scale_factor")
    bb.featurizer.load_state_dict(make_vit_state_dict(DINOV2_L14_REG4.with_depth(3), seed=1))      # timm's `gamma` accepted
    bb.fused_featurizer.load_state_dict(make_vit_state_dict(SIGLIP_SO400M_14.with_depth(3), seed=2))
    psd = make_projector_state_dict(seed=3)
    proj = blb.PrismaticProjector(True, 2176, 4096)
    proj.load_state_dict({f"fc{i + 1}.{p}": psd[f"projector.{2 * i}.{p}"] for i in range(3) for p in ("weight", "bias")})
    bb.cuda(), proj.cuda()
    twin = blb.OpenVLAForActionPrediction(bb, proj, vla.llm, STATS, text_vocab_size=32064, pad_to_multiple_of=64)
    assert twin.vocab_size == 32000 and twin.bin_centers.shape == (255,)
    assert not any(k.startswith("_core") for k in twin.state_dict())
    rng = np.random.default_rng(0)
    image = Image.fromarray((rng.random((256, 256, 3)) * 255).astype(np.uint8))
    instruction = "Pick up the red block"
    want = vla.predict_action(image, instruction, unnorm_key="bridge_orig")
    px = vla.vision_backbone.get_image_transform()(image)
    packed = torch.cat([px["dino"], px["siglip"]], dim=0)[None].to("cuda", dtype=torch.bfloat16)   # processing_prismatic.py:143
    input_ids = vla._prepare_input_ids(instruction, torch.device("cuda"))[:, :-1]                  # without the 29871 token
    got = twin.predict_action(input_ids=input_ids, pixel_values=packed, unnorm_key="bridge_orig", do_sample=False)
    assert got.dtype == np.float64 and np.array_equal(got, want)
    with pytest.raises(AssertionError, match="not in the set of available dataset statistics"):
        twin.predict_action(input_ids=input_ids, pixel_values=packed, unnorm_key="nope")
