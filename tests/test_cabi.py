"""CPU: the C-ABI shared library loads and exports every symbol include/carel_head.h declares
(no compute calls — there is no GPU in the build container), and the product path refuses CPU tensors."""
import ctypes
import re
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent


def _declared_symbols():
    text = (ROOT / "include" / "carel_head.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(carel_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_path():
    syms = _declared_symbols()
    for op in ("mmd", "hsic", "latent", "bow"):
        assert f"carel_{op}_fwd" in syms and f"carel_{op}_bwd" in syms and f"carel_{op}_workspace_bytes" in syms
    assert {"carel_version", "carel_last_error", "carel_launch_count"} <= set(syms)


def test_library_exports_every_declared_symbol():
    from carel_vae_b200 import _lib
    if not _lib.LIB_PATH.exists():
        pytest.skip("libcarel_head.so not built (python -m carel_vae_b200.build)")
    handle = ctypes.CDLL(str(_lib.LIB_PATH))
    missing = [s for s in _declared_symbols() if not hasattr(handle, s)]
    assert not missing, f"symbols declared in include/carel_head.h but not exported: {missing}"
    handle.carel_version.restype = ctypes.c_int
    assert handle.carel_version() == 100
    # every ctypes signature the Python layer binds refers to a real export
    assert not [s for s in _lib._SIGNATURES if not hasattr(handle, s)]


def test_ctypes_signatures_match_the_header():
    """Every entry point bound in _lib._SIGNATURES takes as many arguments as include/carel_head.h declares (a stale
    binding would otherwise shift the stream pointer into garbage)."""
    import re
    from carel_vae_b200 import _lib
    text = re.sub(r"/\*.*?\*/", "", (ROOT / "include" / "carel_head.h").read_text(), flags=re.S)
    decl = {m.group(1): m.group(2) for m in re.finditer(r"\b(carel_\w+)\s*\(([^;{]*?)\)\s*;", text)}
    for name, (_, args) in _lib._SIGNATURES.items():
        assert name in decl, f"{name} is bound but not declared in the header"
        params = [p for p in decl[name].split(",") if p.strip() and p.strip() != "void"]
        assert len(params) == len(args), f"{name}: header has {len(params)} parameters, ctypes binding {len(args)}"


def test_no_cpu_fallback():
    """The product path must fail loudly on CPU tensors instead of silently computing elsewhere."""
    import carel_vae_b200 as cv
    x, y = torch.randn(8, 4), torch.randn(8, 4)
    with pytest.raises(cv.CarelError):
        cv.MMDStatistic(8, 8)(x, y, [0.1])
    with pytest.raises(cv.CarelError):
        cv.HSIC(x, y)
    with pytest.raises(ZeroDivisionError):
        cv.MMDStatistic(1, 1)          # reference behaviour for a batch of one row (mmd:543)
    from carel_vae_b200 import functional as F
    with pytest.raises(cv.CarelError):
        F.weighted_sum([torch.ones(()), torch.ones(2)], [1.0, (0.5, 0.5)])
    with pytest.raises(cv.CarelError):
        F.bow_pack_ell(torch.zeros(4, 16), 8)
    with pytest.raises(cv.CarelError):
        F.club_aprx_loss(torch.zeros(4, 8), torch.zeros(4, 8), torch.zeros(4, 8))
    with pytest.raises(cv.CarelError):
        F.club_upper_bound(torch.zeros(4, 8), torch.zeros(4, 8), torch.zeros(4, 8), torch.randperm(4))


def test_product_does_not_import_oracle():
    for f in (ROOT / "carel_vae_b200").glob("*.py"):
        assert "oracle" not in f.read_text().replace("no CPU / PyTorch fallback", ""), f


def test_state_dict_keys_match_reference(golden_dir):
    """Checkpoint compatibility: same keys and shapes as the reference's state_dict (minus encoder.*)."""
    import carel_vae_b200 as cv
    seen = set()
    for c in torch.load(golden_dir / "head.pt"):
        if (c["variant"], c["L"]) in seen:
            continue
        seen.add((c["variant"], c["L"]))
        opt = cv.default_opt(c["variant"], ec_dim=c["L"], pair_bow_dim=c["V"], bert_dim=c["D"], dropout=c["dropout"])
        model = cv.VARIANTS[c["variant"]](opt, encoder=cv.IdentityEncoder())
        ref = {k: v.shape for k, v in c["state"].items()}
        mine = {k: v.shape for k, v in model.state_dict().items()}
        assert mine == ref
        model.load_state_dict(c["state"])
        groups = model.get_params()
        groups = [groups] if c["variant"] in ("mmd", "hsic") else [list(g) for g in groups]
        assert [sum(p.numel() for p in g) for g in groups] == c["params_in_get_params"]
