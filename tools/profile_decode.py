#!/usr/bin/env python
"""Small driver for ncu: builds a few synthetic 7B layers, warms up, then runs decode steps between
cudaProfilerStart/Stop (use `ncu --profile-from-start off`).  Not a benchmark: numbers under ncu are not bench values."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from exllama_b200.stack import SHAPES, DecodeStack  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--layers", type=int, default=4)
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--ctx", type=int, default=1920)
ap.add_argument("--model", default="7b")
ap.add_argument("--prefill", type=int, default=0)
args = ap.parse_args()
shape = SHAPES[args.model]
stack = DecodeStack(shape, layers=args.layers, max_seq=2048)
hidden = (torch.randn((1, 1, shape.hidden), device="cuda") * 0.5).half()
for _ in range(3):
    stack.decode_step(hidden.clone(), args.ctx)
if args.prefill:
    hp = (torch.randn((1, args.prefill, shape.hidden), device="cuda") * 0.5).half()
    stack.prefill(hp)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
for _ in range(args.steps):
    stack.decode_step(hidden.clone(), args.ctx)
if args.prefill:
    stack.prefill(hp)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("done")
