#!/usr/bin/env python
"""Tensor-parallel persistent decode kernel (csrc/decode_step.cu, tp_world > 1) under torchrun, one rank per GPU:
parity against the per-op tensor-parallel path (q4_attn / decode_attn / q4_attn_2_tp / q4_mlp_tp + NCCL all-reduce, itself
checked against the float64 oracle by tools/tp_check.py) on the same sharded synthetic stack, agreement between ranks,
several consecutive tokens, and timing of both.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tools/tp_step_check.py [--model 7b] [--layers 4]"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from exllama_b200.stack import SHAPES, DecodeStack  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="7b"); ap.add_argument("--layers", type=int, default=4); ap.add_argument("--ctx", type=int, default=1920)
ap.add_argument("--groupsize", type=int, default=128); ap.add_argument("--reps", type=int, default=20)
args = ap.parse_args()
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{lr}"))
dev = torch.device("cuda", lr)
shape = SHAPES[args.model]
st = DecodeStack(shape, groupsize=args.groupsize, device=str(dev), max_seq=2048, layers=args.layers, tp_rank=rank, tp_size=world)
g = torch.Generator(device=dev); g.manual_seed(1234 + rank)
for kc, vc in zip(st.key_cache, st.value_cache):
    kc.copy_((torch.randn(kc.shape, device=dev, generator=g) * 0.5).half()); vc.copy_((torch.randn(vc.shape, device=dev, generator=g) * 0.5).half())
g0 = torch.Generator(device=dev); g0.manual_seed(7)          # the same input on every rank
xs = [(torch.randn((1, 1, shape.hidden), device=dev, generator=g0) * 0.5).half() for _ in range(3)]
snap = [(kc.clone(), vc.clone()) for kc, vc in zip(st.key_cache, st.value_cache)]

# per-op tensor-parallel path with NCCL all-reduce (eager)
ref = [st.decode_step(x.clone(), args.ctx + i).clone() for i, x in enumerate(xs)]
torch.cuda.synchronize(); dist.barrier()
for (kc, vc), (k0, v0) in zip(zip(st.key_cache, st.value_cache), snap):
    kc.copy_(k0); vc.copy_(v0)
st.make_plan()
dist.barrier()
got = [st.decode_step_fused(x, args.ctx + i).clone() for i, x in enumerate(xs)]
torch.cuda.synchronize(); dist.barrier()
ok = True
for i, (a, b) in enumerate(zip(got, ref)):
    rms = float(b.pow(2).mean().sqrt()); d = float((a - b).abs().max())
    gl = [torch.empty_like(a) for _ in range(world)]
    dist.all_gather(gl, a)
    spread = max(float((t - gl[0]).abs().max()) for t in gl)
    good = d <= 2e-2 * rms + 2e-2 * float(b.abs().max()) and spread <= 1e-2 * rms
    ok = ok and good
    if rank == 0:
        print(f"token {i}: |fused_tp - per_op_tp| max {d:.3e} (logit rms {rms:.3f}); max difference between ranks {spread:.3e} -> {'ok' if good else 'BAD'}", flush=True)


def timed(fn, n):
    for _ in range(3): fn()
    torch.cuda.synchronize(); dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / n], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


x = xs[0]
ms_fused = timed(lambda: st.decode_step_fused(x, args.ctx), args.reps)
h = x.clone()
ms_perop = timed(lambda: (h.copy_(x), st.decode_step(h, args.ctx)), max(4, args.reps // 4))
if rank == 0 and os.environ.get("EXL_DS_TRACE"):
    import numpy as np
    tr = st.dplan.trace().astype(np.int64)
    L = min(2, len(st.layers) - 1)
    med = lambda a_, b_: round(float(np.median(tr[:, L, b_] - tr[:, L, a_])) / 1e3, 2)
    print(json.dumps({"layer_us": med(0, 14), "QKV": med(0, 2), "B1": med(2, 3), "ATT": med(3, 4), "B2": med(4, 5), "O": med(5, 7),
                      "B3_local": med(7, 16), "B3_push": med(16, 17), "B3_cross": med(17, 8), "GU": med(8, 10), "B4": med(10, 11), "DOWN": med(11, 13),
                      "B5_local": med(13, 18), "B5_push": med(18, 19), "B5_cross": med(19, 14)}), flush=True)
if rank == 0:
    print(json.dumps({"world": world, "model": args.model, "layers": len(st.layers), "ctx": args.ctx, "fused_tp_ms": round(ms_fused, 4),
                      "per_op_tp_nccl_eager_ms": round(ms_perop, 4), "parity_ok": ok, "plan": st.dplan.info()}), flush=True)
    print(f"tp_step_check world={world} -> {'OK' if ok else 'FAIL'}", flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
