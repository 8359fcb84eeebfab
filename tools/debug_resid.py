import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from bridgelang_b200 import ops
for ctas in (1, 2):
    ops.set_gemm_cta_group(ctas)
    M, N, K = 256, 256, 64
    a = torch.zeros(M, K, device="cuda", dtype=torch.bfloat16)
    w = torch.zeros(N, K, device="cuda", dtype=torch.bfloat16)
    resid0 = (torch.arange(M, device="cuda").float()[:, None] * 1000 + torch.arange(N, device="cuda").float()[None, :])
    resid = resid0.clone()
    ops.gemm(a, w, ops.EPI_RESIDUAL, bias=None, gamma=None, resid=resid)
    torch.cuda.synchronize()
    bad = (resid != resid0).nonzero()
    print("ctas", ctas, "mismatches", bad.shape[0])
    for r, c in bad[:24].tolist():
        print(f"  out[{r},{c}] = {resid[r, c].item():.0f} (want {resid0[r, c].item():.0f})")
