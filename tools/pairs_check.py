"""A/B of the long-K Gram forward: CTA pairs (gemm_2sm) vs single CTAs (gemm_tc)."""
import os, sys, torch
sys.path.insert(0, ".")
from carel_vae_b200 import functional as F, _lib
torch.manual_seed(0)
for n in (300, 1000, 8192):
    x = torch.tanh(torch.randn(n, 768, device="cuda")); y = torch.tanh(torch.randn(n + 7, 768, device="cuda"))
    x, y = x / x.norm(dim=1, keepdim=True), y / y.norm(dim=1, keepdim=True)
    res = {}
    for mode in ("pairs", "single"):
        if mode == "single": os.environ["CAREL_MMD_NO_PAIRS"] = "1"
        else: os.environ.pop("CAREL_MMD_NO_PAIRS", None)
        for full in (False, True):
            for _ in range(3): v = F.mmd(x, y, [0.1], full_gram=full)
            _lib.timing_enable(True)
            for _ in range(10): v = F.mmd(x, y, [0.1], full_gram=full)
            rep = _lib.timing_report(); _lib.timing_enable(False)
            res[(mode, full)] = (float(v), rep["mmd_gram_fwd"][0] / rep["mmd_gram_fwd"][1] * 1e3)
    print(n, {k: (f"{v[0]:.9e}", round(v[1], 1)) for k, v in res.items()}, flush=True)
