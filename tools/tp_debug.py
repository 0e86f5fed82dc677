#!/usr/bin/env python
"""Debug: per-op tensor-parallel decode step, finite check after every op (act-order / group-size combinations)."""
import argparse, os, sys
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from exllama_b200.stack import SHAPES, DecodeStack
from exllama_b200 import cuda_ext, tp as tpmod
ap = argparse.ArgumentParser(); ap.add_argument("--model", default="33b"); ap.add_argument("--groupsize", type=int, default=32)
ap.add_argument("--act-order", action="store_true"); ap.add_argument("--layers", type=int, default=2); ap.add_argument("--fused-ar", action="store_true")
args = ap.parse_args()
rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
if world > 1: dist.init_process_group("nccl", device_id=torch.device(f"cuda:{lr}"))
ext = cuda_ext.exllama_ext
st = DecodeStack(SHAPES[args.model], groupsize=args.groupsize, act_order=args.act_order, device=f"cuda:{lr}", max_seq=512, layers=args.layers, tp_rank=rank, tp_size=world)
if args.fused_ar and world > 1: tpmod.init_fused_allreduce(ext, lr)
for kc, vc in zip(st.key_cache, st.value_cache): kc.normal_(0, 0.5); vc.normal_(0, 0.5)
s = st.shape; none = st.none
x = (torch.randn((1, 1, s.hidden), device=f"cuda:{lr}") * 0.5).half()
hq = st.local_heads * s.head_dim
def chk(name, t):
    ok = bool(torch.isfinite(t).all()); mx = float(t.float().abs().max())
    if rank == 0 or not ok: print(f"rank {rank} {name}: finite={ok} absmax={mx:.3g}", flush=True)
past = 100
for i, L in enumerate(st.layers):
    q = torch.empty((1, 1, hq), dtype=torch.float16, device=x.device); k = torch.empty_like(q); v = torch.empty_like(q)
    ext.q4_attn(x, L.ln1, s.eps, q, k, v, L.q.q4, L.k.q4, L.v.q4, st.sin, st.cos, 1, past, st.local_heads, st.local_heads, s.head_dim, st.key_cache[i], st.value_cache[i], st.max_seq, none, none, none, none, none, none, none)
    chk(f"L{i} q", q); chk(f"L{i} k", k); chk(f"L{i} v", v)
    attn = torch.empty_like(q)
    ext.decode_attn(q, st.key_cache[i], st.value_cache[i], attn, st.local_heads, st.local_heads, s.head_dim, past + 1, st.max_seq)
    chk(f"L{i} attn", attn)
    x2 = x.view(-1, s.hidden)
    if world == 1:
        ext.q4_attn_2(x2, attn.view(-1, hq), L.o.q4, none, none, none); chk(f"L{i} x after o", x2)
        x_before = x2.clone()
        ext.q4_mlp(x2, L.ln2, s.eps, L.gate.q4, L.up.q4, L.down.q4, none, none, none, none, none, none, none); chk(f"L{i} x after mlp", x2)
        if os.environ.get("ORACLE"):
            import numpy as np
            from oracle import oracle as O
            def mm(xin, lin, acc=None):
                xm = None if lin.g_idx is None else O.make_x_map(lin.g_idx.numpy(), lin.qzeros.shape[0])
                return O.q4_matmul_f64(xin, lin.qweight.cpu().numpy(), lin.qzeros.cpu().numpy(), lin.scales.cpu().numpy(), xm, acc)   # qweight is already sequential
            xb = x_before.cpu().numpy()
            xn, _ = O.rms_norm(xb, L.ln2.cpu().numpy(), s.eps)
            act = O.silu_mul(mm(xn, L.gate).astype(np.float16), mm(xn, L.up).astype(np.float16))
            want = mm(act, L.down, acc=xb)
            got = x2.cpu().numpy().astype(np.float64)
            print(f"L{i} mlp vs oracle: max err {np.abs(got - want).max():.4g}, |want| max {np.abs(want).max():.4g}, rms {np.sqrt((want ** 2).mean()):.4g}", flush=True)
    else:
        oin = st._o_input(attn.view(-1, hq)); chk(f"L{i} o_input", oin)
        tpmod.row_parallel_residual(ext, x2, oin, L.o.q4, rank, None); chk(f"L{i} x after o", x2)
        tpmod.mlp_tp(ext, cuda_ext, x2, L, s.eps, rank, None); chk(f"L{i} x after mlp", x2)
torch.cuda.synchronize()
if world > 1: dist.barrier(); dist.destroy_process_group()
