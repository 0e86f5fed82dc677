#!/usr/bin/env python
"""Write a synthetic Llama-shaped GPTQ model directory the reference's own loader accepts (model level).

No checkpoint, tokenizer or network exists on the GPU box, so the directory is generated there:
  config.json        -- the keys ExLlamaConfig requires (model.py:54-60)
  model.safetensors  -- model.layers.{i}.{self_attn.{q,k,v,o}_proj,mlp.{gate,up,down}_proj}.{qweight,qzeros,scales[,g_idx]},
                        the norm weights, model.embed_tokens.weight, lm_head.weight  (key handling: model.py:696-716,826-831)
  tokenizer.model    -- sentencepiece BPE trained on the reference's datasets/wikitext2_val_sample.jsonl (small vocab; every id
                        is < vocab_size, so the embedding lookup is safe)

Random GPTQ tensors: uniform nibbles, small positive scales (a 32-layer stack stays finite in fp16), optional act-order
g_idx = a random permutation of the rows' groups.

    python tools/make_synth_model.py --model 7b --out /tmp/synth7b [--groupsize 128] [--act-order] [--layers N]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

DIMS = {  # hidden, inter, layers, heads
    "tiny": (1024, 2816, 2, 8),
    "7b": (4096, 11008, 32, 32),
    "13b": (5120, 13824, 40, 40),
    "33b": (6656, 17920, 60, 52),
    "65b": (8192, 22016, 80, 64),
}


def train_tokenizer(out_path, dataset, vocab=4000):
    import sentencepiece as spm
    txt = out_path + ".train.txt"
    with open(dataset) as f, open(txt, "w") as g:
        for line in f:
            t = json.loads(line)["text"].strip()
            if t:
                g.write(t.replace("\n", " ") + "\n")
    prefix = out_path[:-len(".model")]
    spm.SentencePieceTrainer.train(input=txt, model_prefix=prefix, vocab_size=vocab, model_type="bpe", character_coverage=1.0,
                                   bos_id=1, eos_id=2, unk_id=0, pad_id=-1, num_threads=8, input_sentence_size=20000,
                                   shuffle_input_sentence=False, minloglevel=2)
    os.remove(txt)
    if os.path.exists(prefix + ".vocab"):
        os.remove(prefix + ".vocab")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="7b", choices=sorted(DIMS))
    ap.add_argument("--out", required=True)
    ap.add_argument("--groupsize", type=int, default=128)
    ap.add_argument("--act-order", action="store_true")
    ap.add_argument("--layers", type=int, default=None)
    ap.add_argument("--vocab", type=int, default=32000)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--dataset", default=os.path.join(ROOT, "baseline", "_ref", "exllama", "datasets", "wikitext2_val_sample.jsonl"))
    args = ap.parse_args()

    import torch
    from safetensors.torch import save_file
    hidden, inter, layers, heads = DIMS[args.model]
    if args.layers:
        layers = args.layers
    os.makedirs(args.out, exist_ok=True)
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    gen = torch.Generator(device=dev); gen.manual_seed(args.seed)
    gs = args.groupsize
    t = {}

    def q4(key, K, N):
        G = K // gs
        t[key + ".qweight"] = torch.randint(-2**31, 2**31 - 1, (K // 8, N), dtype=torch.int32, device=dev, generator=gen).cpu()
        # zero points clustered mid-range (stored nibble 6 or 7: zero-mean weights) and gain-normalised scales: see exllama_b200/stack.py synth_q4_device
        zn = torch.randint(6, 8, (G, N // 8, 8), dtype=torch.int64, device=dev, generator=gen)
        qz = (zn << (torch.arange(8, device=dev, dtype=torch.int64) * 4)).sum(-1)
        t[key + ".qzeros"] = torch.where(qz >= 2**31, qz - 2**32, qz).to(torch.int32).cpu()
        hi = 2.3e-3 * (4096.0 / K) ** 0.5
        t[key + ".scales"] = (torch.rand((G, N), device=dev, generator=gen) * 0.9 * hi + 0.1 * hi).half().cpu()
        if args.act_order:
            perm = torch.randperm(K, device=dev, generator=gen)
            g_idx = torch.empty(K, dtype=torch.int32, device=dev)
            g_idx[perm] = (torch.arange(K, device=dev) // gs).int()
            t[key + ".g_idx"] = g_idx.cpu()

    for i in range(layers):
        p = f"model.layers.{i}."
        for n in ("q", "k", "v", "o"):
            q4(p + f"self_attn.{n}_proj", hidden, hidden)
        q4(p + "mlp.gate_proj", hidden, inter)
        q4(p + "mlp.up_proj", hidden, inter)
        q4(p + "mlp.down_proj", inter, hidden)
        t[p + "input_layernorm.weight"] = (1 + 0.05 * torch.randn(hidden, device=dev, generator=gen)).half().cpu()
        t[p + "post_attention_layernorm.weight"] = (1 + 0.05 * torch.randn(hidden, device=dev, generator=gen)).half().cpu()
    t["model.norm.weight"] = (1 + 0.05 * torch.randn(hidden, device=dev, generator=gen)).half().cpu()
    t["model.embed_tokens.weight"] = (torch.randn((args.vocab, hidden), device=dev, generator=gen) * 0.5).half().cpu()
    t["lm_head.weight"] = (torch.randn((args.vocab, hidden), device=dev, generator=gen) * 0.02).half().cpu()
    save_file(t, os.path.join(args.out, "model.safetensors"))
    cfg = {"architectures": ["LlamaForCausalLM"], "bos_token_id": 1, "eos_token_id": 2, "pad_token_id": 0, "hidden_act": "silu",
           "hidden_size": hidden, "initializer_range": 0.02, "intermediate_size": inter, "max_position_embeddings": 2048,
           "model_type": "llama", "num_attention_heads": heads, "num_hidden_layers": layers, "rms_norm_eps": 1e-6,
           "vocab_size": args.vocab, "torch_dtype": "float16"}
    json.dump(cfg, open(os.path.join(args.out, "config.json"), "w"), indent=1)
    train_tokenizer(os.path.join(args.out, "tokenizer.model"), args.dataset)
    nbytes = sum(v.numel() * v.element_size() for v in t.values())
    print(f"synthetic {args.model} (g{gs}{' act-order' if args.act_order else ''}, {layers} layers): {nbytes / 2**30:.2f} GiB -> {args.out}")


if __name__ == "__main__":
    main()
