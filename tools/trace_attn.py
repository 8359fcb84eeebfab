"""trace_attn.py — print the pipeline timeline (SM cycles) of the tcgen05 attention kernel, CTA 0, tiles 8..15, per warp."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from bridgelang_b200 import _lib, ops

MMA = {0: "s_free(g) seen", 1: "S(g+1) issued", 2: "o_empty(g-1) seen", 3: "PV(g) issued", 4: "p_full[0] seen",
       5: "p_full[1] seen", 8: "k_full seen", 9: "q_full seen"}
SM = {8: "s_full", 9: "S in regs", 10: "max read", 4: "P0 st", 5: "P1 st", 12: "P published"}
HP = {11: "max(g) published", 13: "O(g) in regs", 14: "epi(g) stored"}
for (B, T, H, hd) in ((256, 261, 16, 64), (256, 256, 16, 72)):
    qkv = torch.randn(B * T, 3 * H * hd, device="cuda").bfloat16()
    buf = torch.zeros(8 * 32 * 16, dtype=torch.int64, device="cuda")
    ops.attention(qkv, B, T, H, hd)
    _lib.load().blb_debug_attention_trace(buf.data_ptr())
    ops.attention(qkv, B, T, H, hd)
    torch.cuda.synchronize()
    _lib.load().blb_debug_attention_trace(None)
    t = buf.cpu().view(8, 32, 16)
    t0 = int(t[0][t[0] > 0].min())
    print(f"==== T={T} hd={hd}")
    for gi in range(1, 5):
        print(f"-- tile g={gi + 8}")
        order = [8, 9, 10, 4, 5, 12]
        print("   warp " + " ".join(f"{SM[e]:>14s}" for e in order))
        nsw = 16 if int(t[gi, 20:24].abs().sum()) > 0 else 8      # softmax warps of this build (BLB_ATTN_NSG = 4 | 2)
        for w in range(nsw):
            print(f"   {w:4d} " + " ".join(f"{int(t[gi, w, e]) - t0:14d}" for e in order))
        for w in range(nsw + 4, nsw + 8):
            print(f"   help {w} " + "  ".join(f"{HP[e]}={int(t[gi, w, e]) - t0}" for e in (11, 13, 14)))
        print("   mma  " + "  ".join(f"{MMA[e]}={int(t[gi, nsw + 1, e]) - t0}" for e in (0, 8, 9, 1, 2, 4, 5, 3)))
