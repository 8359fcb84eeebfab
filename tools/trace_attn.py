"""trace_attn.py — print the pipeline timeline (SM cycles) of the tcgen05 attention kernel, CTA 0, tiles 8..15."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from bridgelang_b200 import _lib, ops

NAMES = {0: "mma:s_free(g) seen", 1: "mma:S(g+1) issued", 2: "mma:p_full(g)+o_empty seen", 3: "mma:PV(g) issued",
         8: "sm:s_full(g) seen", 9: "sm:S in regs / s_free", 10: "sm:max exchanged", 11: "sm:p_empty(g-1) seen",
         12: "sm:P stored / p_full(g)", 13: "sm:o_full(g-1) seen", 14: "sm:epilogue(g-1) done"}
for (B, T, H, hd) in ((256, 261, 16, 64), (256, 256, 16, 72)):
    qkv = torch.randn(B * T, 3 * H * hd, device="cuda").bfloat16()
    buf = torch.zeros(256, dtype=torch.int64, device="cuda")
    ops.attention(qkv, B, T, H, hd)
    _lib.load().blb_debug_attention_trace(buf.data_ptr())
    ops.attention(qkv, B, T, H, hd)
    torch.cuda.synchronize()
    _lib.load().blb_debug_attention_trace(None)
    t = buf.cpu().view(-1, 16)
    t0 = int(t[0][t[0] > 0].min())
    print(f"==== T={T} hd={hd}")
    for gi in range(8):
        ev = [(int(t[gi, e]) - t0, NAMES[e]) for e in NAMES if t[gi, e] > 0]
        for c, n in sorted(ev):
            print(f"  g={gi + 8}  {c:8d}  {n}")
        print()
