"""BASELINE configs[4]: large-N MMD Gram sweep (N = 4k..131k rows per sample, d = 768) with the Gram row blocks sharded
over the ranks: every rank holds both samples, computes the block sums of its row tiles, one NCCL all-reduce
of 3 doubles, identical value on every rank.

    python tools/gram_sweep.py                         # 1 GPU
    python -m torch.distributed.run --nproc-per-node G --master-addr 127.0.0.1 tools/gram_sweep.py

Prints one JSON line per N (rank 0): device time of the whole sharded evaluation (CUDA events, max over ranks),
algorithmic TFLOP/s = 2 (2N)^2 d / time (full Gram, no symmetry credit, as in bench.py) and the value."""
import json, os, sys
import torch
import torch.distributed as dist
sys.path.insert(0, ".")
from carel_vae_b200 import parallel as par

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
D = 768
hsic_mode = "--hsic" in sys.argv
sizes = [int(a) for a in sys.argv[1:] if not a.startswith("--")] or ([4096, 8192, 16384, 32768, 65536] if hsic_mode else
                                                                    [4096, 8192, 16384, 32768, 65536, 131072])
ref = {}
for N in sizes:
    g = torch.Generator(device="cpu").manual_seed(N)
    # unit-norm rows (squared distances in [0, 4], alphas non-degenerate); chunked to bound host memory
    x = torch.empty(N, D, device=dev)
    y = torch.empty(N, D, device=dev)
    for t in (x, y):
        for r0 in range(0, N, 16384):
            blk = torch.tanh(torch.randn(min(16384, N - r0), D, generator=g))
            t[r0:r0 + blk.shape[0]] = (blk / blk.norm(dim=1, keepdim=True)).to(dev)
    sig = 2.0                      # unit-norm rows: squared distances in [0, 4]
    fn = (lambda: par.sharded_hsic(x, y, sig, sig)) if hsic_mode else (lambda: par.sharded_mmd(x, y, [0.1]))
    for _ in range(2):
        v = fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier(device_ids=[local])
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5 if N <= 32768 else 2
    e0.record()
    for _ in range(reps):
        v = fn()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / reps], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        flops = 2.0 * 2.0 * N ** 2 * D if hsic_mode else 2.0 * (2 * N) ** 2 * D     # HSIC: two N x N Grams
        print(json.dumps({"metric": "hsic_gram_sweep" if hsic_mode else "mmd_gram_sweep", "N": N, "d": D, "n_gpus": world, "ms": round(float(ms), 3),
                          "algorithmic_tflops": round(flops / (float(ms) * 1e-3) / 1e12, 1), "value": float(v),
                          "includes": "pack of both samples to bf16 + Gram tiles (j>i schedule) + all-reduce + finalize"}),
              flush=True)
    del x, y
if world > 1:
    dist.destroy_process_group()
