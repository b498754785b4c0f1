"""How much of the step do the PyTorch classifier heads cost? (graph replay, ELL target)"""
import sys, torch
sys.path.insert(0, ".")
import bench
import carel_vae_b200 as cv
from carel_vae_b200.parallel import FlatGradBucket
dev = torch.device("cuda", 0)
def run(mode):
    opt = cv.default_opt("mmd", ec_dim=bench.LATENT, pair_bow_dim=bench.VOCAB, bert_dim=bench.D_EMB, dropout=0.5)
    opt.overlap_heads = mode != "serial"
    torch.manual_seed(1234)
    model = cv.DrlClassifierMMD(opt, encoder=cv.IdentityEncoder()).to(dev).train()
    if mode == "noheads":
        z0 = lambda *a, **k: a[0].sum() * 0.0
        model.get_emotion_mul_loss = z0; model.get_cause_mul_loss = z0; model.get_pair_mul_loss = z0
    b = {k: v.to(dev) for k, v in bench.synthetic_batch(dev).items()}
    params = model.get_params()
    bucket = FlatGradBucket(params)
    optim = torch.optim.Adam(params, lr=1e-5, capturable=True, fused=True)
    target = (b["ell_cols"], b["ell_vals"], b["ell_nnz"])
    def step():
        bucket.zero_()
        for p in model.parameters():
            if p.grad is not None and not any(p is q for q in params):
                p.grad = None
        loss = model(b["X"], None, None, b["emo"], b["cau"], b["pair"], target, 0)
        loss.backward()
        optim.step()
    for _ in range(3): step()
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s): step()
    torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g): step()
    for _ in range(3): g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(20): g.replay()
    e1.record(); torch.cuda.synchronize()
    print(mode, round(e0.elapsed_time(e1) / 20, 4), "ms/step")
for m in ("overlap", "serial", "noheads"):
    run(m)
