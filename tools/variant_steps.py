"""Graph-replayed train-step time of the four head variants at N = 8192, L = 128, V = 30000 (ELL BoW target,
Adam over every parameter group): BASELINE configs[2] (HSIC) and [3] (VI) next to the MMD headline."""
import json, sys, torch
sys.path.insert(0, ".")
import bench
import carel_vae_b200 as cv
dev = torch.device("cuda", 0)
b = {k: v.to(dev) for k, v in bench.synthetic_batch(dev).items()}
target = (b["ell_cols"], b["ell_vals"], b["ell_nnz"])
out = {}
for variant in ("mmd", "hsic", "vi", "gan"):
    opt = cv.default_opt(variant, ec_dim=bench.LATENT, pair_bow_dim=bench.VOCAB, bert_dim=bench.D_EMB, dropout=0.5)
    torch.manual_seed(1234)
    model = cv.VARIANTS[variant](opt, encoder=cv.IdentityEncoder()).to(dev).train()
    groups = model.get_params()
    if variant in ("mmd", "hsic"):
        groups = [groups]
    params = [p for g in groups for p in g]
    optim = torch.optim.Adam(params, lr=1e-5, capturable=True, fused=True)
    emo = b["emo"] if variant in ("mmd", "vi") else b["pair"]
    def step():
        optim.zero_grad(set_to_none=True)
        r = model(b["X"], None, None, emo, b["cau"], b["pair"], target, 0)
        if variant == "vi":
            z_e, z_c, aprx, loss = r
            loss = loss + aprx + 0.1 * model.get_ec_upper_loss(z_e, z_c)
        elif variant == "gan":
            loss = r[0] + r[1] + r[2]
        else:
            loss = r
        loss.backward()
        optim.step()
    for _ in range(3):
        step()
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        step()
    torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    try:
        with torch.cuda.graph(g):
            step()
        if variant == "vi":          # next CLUB permutation into the pinned slot the captured copy reads
            def run():
                model.redraw_negative_index(bench.N_ROWS)
                g.replay()
        else:
            run = g.replay
        mode = "graph"
    except Exception as e:  # noqa: BLE001  (kept as a fallback)
        torch.cuda.synchronize()
        run, mode = step, "eager (" + type(e).__name__ + ")"
    for _ in range(3):
        run()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(20):
        run()
    e1.record(); torch.cuda.synchronize()
    out[variant] = {"ms_per_step": round(e0.elapsed_time(e1) / 20, 4), "mode": mode}
    print(variant, out[variant], flush=True)
print(json.dumps(out))
