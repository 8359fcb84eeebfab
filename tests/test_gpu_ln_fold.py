"""LayerNorm-fold stress on a B200 (VERDICT r01 weak #1c): rows with large means (3 σ … 200 σ) and with
outlier channels at 100x — the regime of real DINOv2 / SigLIP residual streams — through the folded path WITH the
rolling per-row shift, the round-1 fold without it, the explicit LayerNorm kernel, and fp32.

Metric: per-row relative error  ||y_m − ref_m||₂ / ||ref_m||₂, worst row (max|err|/max|ref| over the whole tensor is
dominated by the outlier rows and hides everything else)."""

import pytest
import torch
import torch.nn.functional as F

import bridgelang_b200 as blb
from bridgelang_b200 import ops
from bridgelang_b200.config import DINOV2_L14_REG4, SIGLIP_SO400M_14
from bridgelang_b200.weights import make_vit_state_dict, normalize_frames, synthetic_frames
from oracle import vit_oracle

pytestmark = pytest.mark.gpu


def _row_err(y, ref):
    y, ref = y.float(), ref.float()
    return ((y - ref).norm(dim=-1) / ref.norm(dim=-1).clamp_min(1e-20)).max().item()


def _stress_rows(M, D, offset_sigma, gen):
    """std-1 rows + a common offset of `offset_sigma` standard deviations + 6 outlier channels at 100x."""
    x = torch.randn(M, D, device="cuda", generator=gen)
    x[:, torch.tensor([3, 77, 300, 511, 800, D - 1], device="cuda")] *= 100.0
    sigma = x.std(dim=1, keepdim=True)
    sign = torch.where(torch.arange(M, device="cuda") % 2 == 0, 1.0, -1.0)[:, None]
    return x + sign * offset_sigma * sigma


@pytest.mark.parametrize("D,N", [(1024, 3072), (1152, 4352)])
@pytest.mark.parametrize("offset_sigma", [0.0, 3.0, 5.0, 50.0, 200.0])
def test_folded_consumer_is_independent_of_the_row_mean(D, N, offset_sigma):
    M = 512
    g = torch.Generator(device="cuda").manual_seed(int(D + offset_sigma))
    x = _stress_rows(M, D, offset_sigma, g)
    ln_w = torch.rand(D, device="cuda", generator=g) + 0.5
    ln_b = torch.randn(D, device="cuda", generator=g) * 0.1
    W = torch.randn(N, D, device="cuda", generator=g) * 0.03
    b = torch.randn(N, device="cuda", generator=g) * 0.1
    ref = F.linear(F.layer_norm(x.double(), (D,), ln_w.double(), ln_b.double(), 1e-6), W.double(), b.double()).float()
    wf = (W * ln_w[None, :]).bfloat16()
    colsum, bf = wf.float().sum(dim=1), b + W @ ln_b
    parts = ops.gemm_stats_parts(D)
    # explicit LayerNorm kernel, then a plain GEMM (the BLB_LN_EXPLICIT path)
    explicit = ops.gemm(ops.layernorm(x, ln_w, ln_b, 1e-6), W.bfloat16(), ops.EPI_BIAS, bias=b)
    # folded, with the per-row shift (default) and without (round-1 behaviour)
    xs, st_s, shift = ops.rowstats_cast(x, parts, with_shift=True)
    assert torch.allclose(shift, x.mean(dim=1), rtol=1e-5, atol=1e-4)
    shifted = ops.gemm(xs, wf, ops.EPI_BIAS, bias=bf, ln_stats=st_s, ln_colsum=colsum, ln_eps=1e-6)
    x0, st_0 = ops.rowstats_cast(x, parts)
    unshifted = ops.gemm(x0, wf, ops.EPI_BIAS, bias=bf, ln_stats=st_0, ln_colsum=colsum, ln_eps=1e-6)
    e_exp, e_sh, e_un = _row_err(explicit, ref), _row_err(shifted, ref), _row_err(unshifted, ref)
    print(f"D={D} offset {offset_sigma:5.1f} sigma: explicit {e_exp:.3e}  folded+shift {e_sh:.3e}  folded(no shift) {e_un:.3e}")
    assert e_sh < 2.0 * e_exp + 1e-3, (e_sh, e_exp)          # the bar the verdict set: within 2x of the explicit path
    assert e_sh < 1.5e-2
    if offset_sigma >= 50.0:                                   # the test can see the failure mode it guards against
        assert e_un > 5.0 * e_sh, (e_un, e_sh)


@pytest.mark.parametrize("ctas", [1, 2])
@pytest.mark.parametrize("N,K", [(1024, 512), (1152, 4352)])
def test_producer_and_consumer_roll_the_shift_forward(ctas, N, K):
    """EPI_RESIDUAL with shift_in: xb_out = bf16(x_new − c), stats_out = partial (Σ, Σ²) of x_new − c per epilogue-warp
    column span, the fp32 stream itself stays unshifted.  The folded consumer that follows hands over the next shift:
    shift_out = c + mean(x_new − c) = the row mean of x_new."""
    ops.set_gemm_cta_group(ctas)
    M = 3 * 261
    g = torch.Generator(device="cuda").manual_seed(N + K)
    a = (torch.randn(M, K, device="cuda", generator=g) * 0.5).bfloat16()
    w = (torch.randn(N, K, device="cuda", generator=g) * 0.05).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g) * 0.1
    gamma = torch.rand(N, device="cuda", generator=g) + 0.5
    resid0 = torch.randn(M, N, device="cuda", generator=g) + 40.0 * torch.randn(M, 1, device="cuda", generator=g)
    parts = ops.gemm_stats_parts(N)
    c = resid0.mean(dim=1).contiguous()                       # "the row mean one update ago"
    resid = resid0.clone()
    stats = torch.full((parts, M, 2), float("nan"), device="cuda")
    xb = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16)
    ops.gemm(a, w, ops.EPI_RESIDUAL, bias=bias, gamma=gamma, resid=resid, stats_out=stats, xb_out=xb, shift_in=c)
    ref = resid0 + gamma * (a.float() @ w.float().t() + bias)
    assert ((resid - ref).abs().max() / ref.abs().max()).item() < 2e-5
    assert torch.equal(xb, (resid - c[:, None]).bfloat16())
    span = N // parts
    d = resid - c[:, None]
    want = torch.stack([torch.stack([d[:, p * span:(p + 1) * span].sum(dim=1),
                                     (d[:, p * span:(p + 1) * span] ** 2).sum(dim=1)], dim=1) for p in range(parts)])
    assert torch.allclose(stats, want, rtol=1e-4, atol=1e-2)
    # consumer: same output with or without the hand-over, and shift_out = the new row mean
    W2 = (torch.randn(256, N, device="cuda", generator=g) * 0.03).bfloat16()
    colsum, b2 = W2.float().sum(dim=1), torch.zeros(256, device="cuda")
    nxt = torch.full((M,), float("nan"), device="cuda")
    y1 = ops.gemm(xb, W2, ops.EPI_BIAS, bias=b2, ln_stats=stats, ln_colsum=colsum, shift_in=c, shift_out=nxt)
    y0 = ops.gemm(xb, W2, ops.EPI_BIAS, bias=b2, ln_stats=stats, ln_colsum=colsum)
    assert torch.equal(y0, y1)
    assert torch.allclose(nxt, resid.mean(dim=1), rtol=1e-4, atol=1e-3)
    lnref = F.linear(F.layer_norm(resid, (N,), None, None, 1e-6), W2.float())
    assert _row_err(y1, lnref) < 1e-2
    with pytest.raises(RuntimeError):                                       # ping-pong buffers are mandatory
        ops.gemm(xb, W2, ops.EPI_BIAS, bias=b2, ln_stats=stats, ln_colsum=colsum, shift_in=c, shift_out=c)
    ops.set_gemm_cta_group(0)


def _massive_activation_state_dict(cfg, seed):
    """stress-init weights whose residual stream looks like a trained ViT's: every token carries a common offset of
    ~40 (patch-embed bias) on top of std ~1 content, and 6 channels sit at ±100 (position embedding)."""
    sd = make_vit_state_dict(cfg, seed=seed, init="stress")
    sd["patch_embed.proj.bias"] = sd["patch_embed.proj.bias"] + 40.0
    pe = sd["pos_embed"].clone()
    pe[:, :, [5, 130, 400, 640, 900, cfg.dim - 2]] += torch.tensor([100.0, -100.0, 100.0, -100.0, 100.0, -100.0])
    sd["pos_embed"] = pe
    return sd


@pytest.mark.parametrize("cfg,key", [(DINOV2_L14_REG4, "dino"), (SIGLIP_SO400M_14, "siglip")])
def test_depth4_tower_with_massive_activations_folded_vs_explicit_vs_fp32(cfg, key):
    cfg = cfg.with_depth(5)                                    # 4 blocks run (get_intermediate_layers n = depth-2)
    sd = _massive_activation_state_dict(cfg, seed=91)
    px = normalize_frames(synthetic_frames(2, seed=17))[key].bfloat16()
    ref = vit_oracle.vit_intermediate(sd, cfg, px.float())
    rowmean = (ref.mean(dim=-1).abs() / ref.std(dim=-1)).median().item()
    outs = {}
    for folded in (True, False):
        vit = blb.VisionTransformer(cfg)
        vit.ln_folded = folded
        vit.load_state_dict(sd)
        vit.cuda()
        outs[folded] = vit(px.cuda())
    e_f, e_x = _row_err(outs[True].cpu(), ref), _row_err(outs[False].cpu(), ref)
    print(f"{key}: |row mean|/std (median) {rowmean:.2f}; worst-row error: folded+shift {e_f:.3e}  explicit {e_x:.3e}")
    assert e_f < 2.0 * e_x + 1e-3, (e_f, e_x)
    assert e_f < 2e-2 and e_x < 2e-2
