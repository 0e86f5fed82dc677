#!/usr/bin/env python
"""Time the whole-token persistent kernel (csrc/decode_step.cu) next to the per-op path on the synthetic stack.
    python tools/step_bench.py [--model 7b] [--ctx 1920] [--groupsize 128] [--layers N] [--reps 20]"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from exllama_b200.stack import SHAPES, DecodeStack


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="7b"); ap.add_argument("--ctx", type=int, default=1920); ap.add_argument("--seq", type=int, default=2048)
    ap.add_argument("--groupsize", type=int, default=128); ap.add_argument("--layers", type=int, default=None)
    ap.add_argument("--reps", type=int, default=20); ap.add_argument("--no-per-op", action="store_true")
    args = ap.parse_args()
    shape = SHAPES[args.model]
    st = DecodeStack(shape, groupsize=args.groupsize, device="cuda:0", max_seq=args.seq, layers=args.layers)
    for kc, vc in zip(st.key_cache, st.value_cache):
        kc.normal_(0, 0.5); vc.normal_(0, 0.5)
    st.make_plan()
    x = (torch.randn((1, 1, shape.hidden), device="cuda") * 0.5).half()

    def timed(fn):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side): fn()
        torch.cuda.current_stream().wait_stream(side); torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g): fn()
        for _ in range(3): g.replay()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(args.reps): g.replay()
        b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b) / args.reps
    res = {"model": args.model, "ctx": args.ctx, "groupsize": args.groupsize, "layers": len(st.layers), "plan": st.dplan.info()}
    ms = timed(lambda: st.decode_step_fused(x, args.ctx))
    wbytes = st.q4_weight_bytes_per_token() + shape.vocab * shape.hidden * 2
    kv = 2 * len(st.layers) * args.ctx * shape.hidden * 2
    res["fused_ms"] = round(ms, 4); res["fused_tok_s"] = round(1000 / ms, 1); res["bytes"] = wbytes + kv
    res["fused_GBps"] = round((wbytes + kv) / ms / 1e6, 1)
    lf = st._plan_logits.clone()
    if not args.no_per_op:
        h = x.clone()
        def per_op():
            h.copy_(x); return st.decode_step(h, args.ctx)
        ms2 = timed(per_op)
        res["per_op_ms"] = round(ms2, 4); res["per_op_tok_s"] = round(1000 / ms2, 1)
        lp = per_op()
        d = (lf - lp).abs().max().item(); r = lp.pow(2).mean().sqrt().item()
        res["max_abs_diff_vs_per_op"] = d; res["logit_rms"] = r
    print(json.dumps(res))


if __name__ == "__main__":
    main()
