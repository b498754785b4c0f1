"""Is the classifier-head branch (side stream) hidden behind the main stream in the graph-replayed MMD step?
Times the step with opt.overlap_heads = True / False, and with the heads' loss weights set to zero work removed
(no heads at all: a lower bound).  usage: python tools/heads_overlap.py"""
import sys, torch
sys.path.insert(0, ".")
import bench
import carel_vae_b200 as cv
dev = torch.device("cuda", 0)
b = {k: v.to(dev) for k, v in bench.synthetic_batch(dev).items()}
target = (b["ell_cols"], b["ell_vals"], b["ell_nnz"])


def run(overlap, no_heads=False):
    opt = cv.default_opt("mmd", ec_dim=bench.LATENT, pair_bow_dim=bench.VOCAB, bert_dim=bench.D_EMB, dropout=0.5)
    opt.overlap_heads = overlap
    torch.manual_seed(1234)
    model = cv.DrlClassifierMMD(opt, encoder=cv.IdentityEncoder()).to(dev).train()
    if no_heads:
        model._heads = lambda z, *a, **k: z.new_zeros(())
    params = model.get_params()
    optim = torch.optim.Adam(params, lr=1e-5, capturable=True, fused=True)

    def step():
        optim.zero_grad(set_to_none=True)
        model(b["X"], None, None, b["emo"], b["cau"], b["pair"], target, 0).backward()
        optim.step()
    for _ in range(3):
        step()
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        step()
    torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        step()
    for _ in range(3):
        g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(30):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 30


for name, args in (("heads on the side stream", (True,)), ("heads on the main stream", (False,)),
                   ("no heads at all", (True, True))):
    print(f"{name:28s} {run(*args) * 1e3:8.1f} us/step", flush=True)
