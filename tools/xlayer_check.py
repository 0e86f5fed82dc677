"""Bring-up check of the experimental persistent decode-layer kernel (exllama_b200/csrc/experimental/decode_layer.cu).

  python -m exllama_b200._build --experimental
  EXL_B200_LIB=$PWD/exllama_b200/libexl_b200_x.so timeout 120 python tools/xlayer_check.py [--layers 2] [--time]

Runs L synthetic Llama-7B layers for one decode token twice on identical tensors -- through the five-launch product path
(q4_attn, decode_attn, q4_attn_2, q4_mlp via the C ABI) and through exl_x_decode_layer -- and compares the hidden state and
the K/V rows written.  Everything goes through ctypes against ONE library so both paths share the handle registry.
Not a test of the product: the kernel it exercises is not part of the default build."""
import argparse
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from exllama_b200 import capi  # noqa: E402
from exllama_b200.stack import SHAPES, synth_q4_device  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--layers", type=int, default=2)
ap.add_argument("--ctx", type=int, default=1920)
ap.add_argument("--time", action="store_true")
args = ap.parse_args()
assert "libexl_b200_x" in capi.LIB_PATH, "point EXL_B200_LIB at libexl_b200_x.so (python -m exllama_b200._build --experimental)"
lib = capi.lib()
lib.exl_x_decode_layer.restype = C.c_int
lib.exl_x_decode_layer.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p,
                                   C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]

s = SHAPES["7b"]
dev = torch.device("cuda:0")
gen = torch.Generator(device=dev); gen.manual_seed(1)
h, inter, heads, hd, max_seq = s.hidden, s.inter, s.heads, s.head_dim, 2048
mk = lambda K, N: capi.Q4(*synth_q4_device(K, N, 128, dev, gen)[:3])
layers = []
for _ in range(args.layers):
    L = dict(q=mk(h, h), k=mk(h, h), v=mk(h, h), o=mk(h, h), gate=mk(h, inter), up=mk(h, inter), down=mk(inter, h),
             ln1=(1 + 0.05 * torch.randn(h, device=dev, generator=gen)).half(),
             ln2=(1 + 0.05 * torch.randn(h, device=dev, generator=gen)).half(),
             kc=(torch.randn((1, heads, max_seq, hd), device=dev, generator=gen) * 0.5).half(),
             vc=(torch.randn((1, heads, max_seq, hd), device=dev, generator=gen) * 0.5).half())
    layers.append(L)
inv_freq = 1.0 / (10000.0 ** (torch.arange(0, hd, 2, device=dev).float() / hd))
emb = torch.cat((torch.einsum("i,j->ij", torch.arange(max_seq, device=dev, dtype=torch.float32), inv_freq),) * 2, dim=-1)
sin, cos = emb.sin()[None, None].half().contiguous(), emb.cos()[None, None].half().contiguous()
temp_state = torch.zeros((max_seq, inter), dtype=torch.float16, device=dev)
temp_mlp = torch.zeros((4, inter), dtype=torch.float16, device=dev)
capi.prepare_buffers(0, temp_state, temp_mlp, torch.zeros((1, 65536), dtype=torch.float32, device=dev),
                     torch.zeros((1, 8), dtype=torch.float16, device=dev))
x0 = (torch.randn((1, 1, h), device=dev, generator=gen) * 0.5).half()
past = args.ctx


def five_launch(x, caches):
    for L, (kc, vc) in zip(layers, caches):
        q = torch.empty((1, 1, h), dtype=torch.float16, device=dev); k = torch.empty_like(q); v = torch.empty_like(q)
        capi.q4_attn(x, L["ln1"], s.eps, q, k, v, L["q"], L["k"], L["v"], sin, cos, 1, past, heads, heads, hd, kc, vc, max_seq)
        attn = capi.decode_attn(q, kc, vc, heads, heads, hd, past + 1, max_seq)
        capi.q4_attn_2(x.view(-1, h), attn.view(-1, h), L["o"])
        capi.q4_mlp(x.view(-1, h), L["ln2"], s.eps, L["gate"], L["up"], L["down"])


scratch = torch.zeros(2 * (2 * h + inter) + 4 * (2 * h // 128) + 64, dtype=torch.uint8, device=dev)


def one_kernel_per_layer(x, caches):
    for i, (L, (kc, vc)) in enumerate(zip(layers, caches)):
        hs = (C.c_void_p * 7)(*[L[n].handle for n in ("q", "k", "v", "o", "gate", "up", "down")])
        capi.check(lib.exl_x_decode_layer(x.data_ptr(), hs, L["ln1"].data_ptr(), L["ln2"].data_ptr(), s.eps, sin.data_ptr(),
                                          cos.data_ptr(), past, max_seq, heads, kc.data_ptr(), vc.data_ptr(), scratch.data_ptr(),
                                          1 if i == 0 else 0, C.c_void_p(torch.cuda.current_stream().cuda_stream)))


ca = [(L["kc"].clone(), L["vc"].clone()) for L in layers]
cb = [(L["kc"].clone(), L["vc"].clone()) for L in layers]
xa, xb = x0.clone(), x0.clone()
five_launch(xa, ca)
torch.cuda.synchronize()
one_kernel_per_layer(xb, cb)
torch.cuda.synchronize()
dx = (xa.float() - xb.float()).abs().max().item()
dk = max((a[0][0, :, past] .float() - b[0][0, :, past].float()).abs().max().item() for a, b in zip(ca, cb))
dv = max((a[1][0, :, past].float() - b[1][0, :, past].float()).abs().max().item() for a, b in zip(ca, cb))
scale = xa.float().abs().max().item()
print(f"layers {args.layers}: max |dx| {dx:.3e} (scale {scale:.3f}), new K row {dk:.3e}, new V row {dv:.3e}")
ok = dx <= 4e-3 * max(1.0, scale) and dk <= 4e-3 and dv <= 4e-3
print("OK" if ok else "MISMATCH")
if args.time and ok:
    for fn, name in ((five_launch, "five launches"), (one_kernel_per_layer, "one kernel / layer")):
        x = x0.clone()
        g = torch.cuda.CUDAGraph()
        fn(x, ca); torch.cuda.synchronize()
        with torch.cuda.graph(g):
            fn(x, ca)
        g.replay(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20): g.replay()
        b.record(); torch.cuda.synchronize()
        print(f"{name}: {a.elapsed_time(b) / 20 / args.layers * 1e3:.1f} us per layer")
sys.exit(0 if ok else 1)
