"""decode_attn kernel vs the torch ops it replaces (model.py:395-409), one layer's KV cache at several context lengths.
32 distinct caches (>> L2) rotated inside a CUDA graph; prints one JSON line per context."""
import json
import math
import sys, os

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from exllama_b200 import cuda_ext  # noqa: E402

ext = cuda_ext.exllama_ext
heads, hd, max_seq, L = 32, 128, 2048, 32
dev = torch.device("cuda:0")
kcs = [torch.randn((1, heads, max_seq, hd), device=dev).half() for _ in range(L)]
vcs = [torch.randn((1, heads, max_seq, hd), device=dev).half() for _ in range(L)]
q = torch.randn((1, 1, heads * hd), device=dev).half()
out = torch.empty_like(q)


def ours(seq):
    for kc, vc in zip(kcs, vcs):
        ext.decode_attn(q, kc, vc, out, heads, heads, hd, seq, max_seq)


def torch_ops(seq):
    for kc, vc in zip(kcs, vcs):
        qq = q.view(1, 1, heads, hd).transpose(1, 2)
        w = torch.matmul(qq, kc.narrow(2, 0, seq).transpose(2, 3))
        w /= math.sqrt(hd)
        w = torch.nn.functional.softmax(w, dim=-1, dtype=torch.float16)
        torch.matmul(w, vc.narrow(2, 0, seq)).transpose(1, 2).reshape(1, 1, heads * hd)


def timed(fn, seq, reps=10):
    fn(seq); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn(seq)
    g.replay(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): g.replay()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps / L * 1e3


for seq in (16, 256, 512, 1024, 1921, 2048):
    uo, ut = timed(ours, seq), timed(torch_ops, seq)
    by = 2 * heads * seq * hd * 2
    print(json.dumps({"seq": seq, "decode_attn_us": round(uo, 2), "torch_ops_us": round(ut, 2), "kv_bytes": by,
                      "GBps": round(by / uo / 1e3, 1), "speedup": round(ut / uo, 2)}), flush=True)
