"""A few eager MMD-VAE head train steps at the bench configuration (for ncu launch lists).
usage: python tools/step_once.py [steps] [dense|ell]"""
import sys
import torch
sys.path.insert(0, ".")
import bench
import carel_vae_b200 as cv
from carel_vae_b200.parallel import FlatGradBucket
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
mode = sys.argv[2] if len(sys.argv) > 2 else "dense"
dev = torch.device("cuda", 0)
opt = cv.default_opt("mmd", ec_dim=bench.LATENT, pair_bow_dim=bench.VOCAB, bert_dim=bench.D_EMB, dropout=0.5)
torch.manual_seed(1234)
model = cv.DrlClassifierMMD(opt, encoder=cv.IdentityEncoder()).to(dev).train()
b = {k: v.to(dev) for k, v in bench.synthetic_batch(dev).items()}
params = model.get_params()
bucket = FlatGradBucket(params)
optim = torch.optim.Adam(params, lr=1e-5, capturable=True, fused=True)
target = b["bow"] if mode == "dense" else (b["ell_cols"], b["ell_vals"], b["ell_nnz"])
for i in range(steps):
    bucket.zero_()
    for p in model.parameters():
        if p.grad is not None and not any(p is q for q in params):
            p.grad = None
    loss = model(b["X"], None, None, b["emo"], b["cau"], b["pair"], target, i)
    loss.backward()
    optim.step()
torch.cuda.synchronize()
print("ok", float(loss.detach()))
