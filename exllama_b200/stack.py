"""Synthetic Llama-shaped decoder stack over the operator surface.

Host-side mirror of the reference's per-layer call sequence -- ExLlamaAttention.fused (model.py:322-417),
ExLlamaMLP.fused (model.py:238-263) for decode, and the unfused path (model.py:419-506, 265-273, 524-552) for
prefill -- expressed only in the reference-facing plugin API (cuda_ext.ext_* / exllama_ext.*) plus the same torch
ops model.py uses for attention.  It exists so that the hot path can be driven and timed end to end on synthetic
GPTQ tensors (no checkpoint, tokenizer or model.py travel to the GPU box); it is not a model implementation:
there is no loader, no sampling, no tokenizer.

Tensor parallelism (new functionality): q/k/v and gate/up are column-sharded by whole heads /
group-aligned column blocks, o_proj and down_proj are row-sharded on group boundaries, and the [M, hidden] partial
of each row-parallel projection is summed with ONE all-reduce (exllama_b200/tp.py).
"""
from __future__ import annotations

import math
import torch

from . import cuda_ext
from . import tp as tpmod

ext = cuda_ext.exllama_ext


from .shapes import SHAPES, LlamaShape  # noqa: E402,F401  (re-exported)


class _Arena:
    """One device allocation handed out in 256-byte aligned slices, in request order (a loader that maps one checkpoint file
    region): the GPTQ tensors of a stack then sit back to back in HBM instead of wherever the caching allocator had room."""

    def __init__(self, nbytes, device):
        self.buf = torch.empty(nbytes, dtype=torch.uint8, device=device)
        self.off = 0

    def take(self, shape, dtype):
        n = 1
        for d in shape:
            n *= d
        nb = n * torch.empty((), dtype=dtype).element_size()
        v = self.buf[self.off:self.off + nb].view(dtype).view(*shape)
        self.off += (nb + 255) & ~255
        return v

    @staticmethod
    def q4_bytes(K, N, groupsize):
        G = K // groupsize
        r = lambda b: (b + 255) & ~255
        return r(K // 8 * N * 4) + r(G * (N // 8) * 4) + r(G * N * 2)


def synth_q4_device(K, N, groupsize, device, gen, act_order=False, scale_lo=None, scale_hi=None, arena=None):
    """Random GPTQ tensor set on the device, shaped like a real checkpoint: uniform 4-bit weights, zero points in the middle of
    the range (stored nibble 6 or 7, i.e. z + 1 = 7 or 8, mean 7.5 = the mean nibble: zero-mean weights, as GPTQ produces for
    symmetric weight groups.  A mean offset of even one quantisation step gives every matrix a DC gain of ~scale * K >> 1 and a
    deep synthetic stack then blows up through its DC mode -- measured: 33B overflowed fp16 within a few layers), and scales sized
    for a per-layer gain below one whatever K is (std(q - z) ~ 4.6, so gain = 4.6 * mean(scale) * sqrt(K) ~ 0.4)."""
    G = K // groupsize
    if scale_hi is None:
        scale_hi = 2.3e-3 * (4096.0 / K) ** 0.5
    if scale_lo is None:
        scale_lo = 0.1 * scale_hi
    if arena is not None:
        qweight = arena.take((K // 8, N), torch.int32)
        torch.randint(-2**31, 2**31 - 1, (K // 8, N), dtype=torch.int32, device=device, generator=gen, out=qweight)
    else:
        qweight = torch.randint(-2**31, 2**31 - 1, (K // 8, N), dtype=torch.int32, device=device, generator=gen)
    zn = torch.randint(6, 8, (G, N // 8, 8), dtype=torch.int64, device=device, generator=gen)
    shifts = (torch.arange(8, device=device, dtype=torch.int64) * 4)
    qz = (zn << shifts).sum(-1)                                   # 8 nibbles per word, nibble n % 8 of word [g, n / 8]
    qzeros = torch.where(qz >= 2**31, qz - 2**32, qz).to(torch.int32)
    scales = (torch.rand((G, N), device=device, generator=gen) * (scale_hi - scale_lo) + scale_lo).half()
    if arena is not None:
        qzeros = arena.take((G, N // 8), torch.int32).copy_(qzeros)
        scales = arena.take((G, N), torch.float16).copy_(scales)
    g_idx = None
    if act_order:
        perm = torch.randperm(K, device=device, generator=gen)
        g_idx = torch.empty(K, dtype=torch.int32, device=device)
        g_idx[perm] = (torch.arange(K, device=device) // groupsize).int()
        g_idx = g_idx.cpu()
    return qweight, qzeros, scales, g_idx


class Q4Linear:
    """Counterpart of Ex4bitLinear (model.py:130-177): owns the tensors, holds the ext handle."""

    def __init__(self, qweight, qzeros, scales, g_idx, device_index):
        self.qweight, self.qzeros, self.scales, self.g_idx = qweight, qzeros, scales, g_idx
        self.height, self.width = qweight.shape[0] * 8, qweight.shape[1]
        self.q4 = cuda_ext.ext_make_q4(qweight, qzeros, scales, g_idx, device_index)

    def nbytes(self):
        return self.qweight.numel() * 4 + self.qzeros.numel() * 4 + self.scales.numel() * 2

    def forward(self, x):
        return cuda_ext.ext_q4_matmul(x, self.q4, self.width)


class Layer:
    pass


class DecodeStack:
    def __init__(self, shape: LlamaShape, groupsize=128, act_order=False, device="cuda:0", max_seq=2048,
                 layers=None, seed=0, tp_rank=0, tp_size=1, tp_group=None, with_head=True):
        self.shape, self.groupsize, self.act_order = shape, groupsize, act_order
        self.device = torch.device(device)
        self.dev_index = self.device.index or 0
        self.max_seq = max_seq
        self.n_layers = layers if layers is not None else shape.layers
        self.tp_rank, self.tp_size, self.tp_group = tp_rank, tp_size, tp_group
        self.fused_decode_attn = True          # csrc/decode_attn.cu instead of the torch ops of model.py:395-409
        # act-order + tensor parallel (tp.py "act-order checkpoints"): row-parallel shards are cut on group ranges; down's
        # input is local (gate/up column-gathered to match), o_proj's input is the all-gathered attention output indexed
        # by the shard's rows.  The synthetic stack draws every rank's shard directly (a valid local act-order matrix) and
        # a random row set for o_proj, so the work and the exchange are those of a real sharded checkpoint.
        self.o_rows = None
        plan = tpmod.plan_shards(shape.hidden, shape.inter, shape.heads, shape.head_dim, groupsize, tp_size)
        self.plan = plan
        h = shape.hidden
        self.local_heads = plan.heads[tp_rank]
        hq = self.local_heads * shape.head_dim
        il = plan.inter_cols[tp_rank][1] - plan.inter_cols[tp_rank][0]
        gen = torch.Generator(device=self.device)
        gen.manual_seed(seed * 1000 + tp_rank)                 # this rank's shards
        rep = torch.Generator(device=self.device)
        rep.manual_seed(seed * 1000 + 977)                     # replicated tensors (norms, head): identical on every rank
        with torch.cuda.device(self.device):
            ext.set_tuning_params(8, 2, 8, False, False, False, False, False, False)
            self.layers = []
            per_layer = (3 * _Arena.q4_bytes(h, hq, groupsize) + _Arena.q4_bytes(hq, h, groupsize) + 2 * _Arena.q4_bytes(h, il, groupsize) +
                         _Arena.q4_bytes(il, h, groupsize))
            self.arena = _Arena(per_layer * self.n_layers, self.device)
            for _ in range(self.n_layers):
                L = Layer()
                mk = lambda K, N: Q4Linear(*synth_q4_device(K, N, groupsize, self.device, gen, act_order, arena=self.arena), self.dev_index)
                L.q, L.k, L.v = mk(h, hq), mk(h, hq), mk(h, hq)
                L.o = mk(hq, h)
                L.gate, L.up = mk(h, il), mk(h, il)
                L.down = mk(il, h)
                L.ln1 = (1 + 0.05 * torch.randn(h, device=self.device, generator=rep)).half()
                L.ln2 = (1 + 0.05 * torch.randn(h, device=self.device, generator=rep)).half()
                self.layers.append(L)
            if act_order and tp_size > 1:
                g0 = torch.Generator(device="cpu"); g0.manual_seed(seed * 7919 + 13)            # same on every rank
                perm = torch.randperm(h, generator=g0)
                self.o_rows = perm[tp_rank * hq:(tp_rank + 1) * hq].sort().values.to(self.device)
            self.norm = (1 + 0.05 * torch.randn(h, device=self.device, generator=rep)).half()
            self.lm_head = (torch.randn((shape.vocab, h), device=self.device, generator=rep) * 0.02).half() if with_head else None
            # sin/cos tables as model.py:864-877
            inv_freq = 1.0 / (10000.0 ** (torch.arange(0, shape.head_dim, 2, device=self.device).float() / shape.head_dim))
            t = torch.arange(max_seq, device=self.device, dtype=torch.float32)
            freqs = torch.einsum("i,j->ij", t, inv_freq)
            emb = torch.cat((freqs, freqs), dim=-1)
            self.sin = emb.sin()[None, None, :, :].half().contiguous()
            self.cos = emb.cos()[None, None, :, :].half().contiguous()
            # KV cache [bsz, kv_heads_local, max_seq, head_dim] per layer (model.py:557-585)
            self.key_cache = [torch.zeros((1, self.local_heads, max_seq, shape.head_dim), dtype=torch.float16, device=self.device)
                              for _ in range(self.n_layers)]
            self.value_cache = [torch.zeros_like(self.key_cache[0]) for _ in range(self.n_layers)]
            # scratch, as model.py:897-917
            self.temp_state = torch.zeros((max_seq, shape.inter), dtype=torch.float16, device=self.device)
            self.temp_mlp = torch.zeros((4, shape.inter), dtype=torch.float16, device=self.device)
            self.temp_zeros_float = torch.zeros((1, 65536), dtype=torch.float32, device=self.device)
            max_dq = max(max(l.gate.qweight.numel(), l.q.qweight.numel(), l.down.qweight.numel()) for l in self.layers) * 8
            self.temp_dq = torch.zeros((1, max_dq), dtype=torch.float16, device=self.device)
            ext.prepare_buffers(self.device, self.temp_state, self.temp_mlp, self.temp_zeros_float, self.temp_dq)
        self.none = cuda_ext.none_tensor

    # ---- bookkeeping for rooflines -------------------------------------------------------------------------
    def q4_weight_bytes_per_token(self):
        return sum(l.q.nbytes() + l.k.nbytes() + l.v.nbytes() + l.o.nbytes() + l.gate.nbytes() + l.up.nbytes() + l.down.nbytes()
                   for l in self.layers)

    # ---- decode: fused path, rows == 1 (model.py:528-547) ----------------------------------------------------
    def decode_step(self, hidden, past_len):
        """hidden: [1, 1, hidden] fp16 on the device, updated in place layer by layer; returns logits [1, vocab]."""
        s, none = self.shape, self.none
        bsz, q_len = 1, hidden.shape[1]
        hq = self.local_heads * s.head_dim
        for i, L in enumerate(self.layers):
            q = torch.empty((bsz, q_len, hq), dtype=torch.float16, device=self.device)
            k = torch.empty_like(q)
            v = torch.empty_like(q)
            ext.q4_attn(hidden, L.ln1, s.eps, q, k, v, L.q.q4, L.k.q4, L.v.q4, self.sin, self.cos, q_len, past_len,
                        self.local_heads, self.local_heads, s.head_dim, self.key_cache[i], self.value_cache[i], self.max_seq,
                        none, none, none, none, none, none, none)
            if self.fused_decode_attn and q_len == 1 and s.head_dim == 128:
                # one kernel over the KV cache instead of the five torch launches below (csrc/decode_attn.cu)
                attn = torch.empty_like(q)
                ext.decode_attn(q, self.key_cache[i], self.value_cache[i], attn, self.local_heads, self.local_heads,
                                s.head_dim, past_len + 1, self.max_seq)
            else:
                q = q.view(bsz, q_len, self.local_heads, s.head_dim).transpose(1, 2)
                keys = self.key_cache[i].narrow(2, 0, past_len + q_len)
                values = self.value_cache[i].narrow(2, 0, past_len + q_len)
                attn = torch.matmul(q, keys.transpose(2, 3))
                attn /= math.sqrt(s.head_dim)
                attn = torch.nn.functional.softmax(attn, dim=-1, dtype=torch.float16)
                attn = torch.matmul(attn, values).transpose(1, 2).reshape(bsz, q_len, hq)
            x2 = hidden.view(-1, s.hidden)
            if self.tp_size == 1:
                ext.q4_attn_2(x2, attn.view(-1, hq), L.o.q4, none, none, none)
                ext.q4_mlp(x2, L.ln2, s.eps, L.gate.q4, L.up.q4, L.down.q4, none, none, none, none, none, none, none)
            else:
                tpmod.row_parallel_residual(ext, x2, self._o_input(attn.view(-1, hq)), L.o.q4, self.tp_rank, self.tp_group)
                tpmod.mlp_tp(ext, cuda_ext, x2, L, s.eps, self.tp_rank, self.tp_group)
        if self.lm_head is None:
            return hidden
        hn = cuda_ext.ext_rms_norm(hidden, self.norm, s.eps)
        return torch.matmul(hn.view(-1, s.hidden), self.lm_head.t()).float()

    # ---- decode: the whole token as ONE persistent kernel (csrc/decode_step.cu) -------------------
    def make_plan(self, with_head=True):
        """Build the exl_decode_plan over this stack's handles, norms, caches and tables (borrowed)."""
        from . import capi
        hs = [[L.q.q4, L.k.q4, L.v.q4, L.o.q4, L.gate.q4, L.up.q4, L.down.q4] for L in self.layers]
        head = self.lm_head if with_head else None
        self.dplan = capi.DecodePlan(hs, [L.ln1 for L in self.layers], [L.ln2 for L in self.layers], self.key_cache, self.value_cache,
                                    self.sin, self.cos, self.local_heads, self.shape.head_dim, self.max_seq, self.shape.eps,
                                    final_norm=self.norm if head is not None else None, lm_head=head,
                                    tp_rank=self.tp_rank, tp_world=self.tp_size)
        if self.tp_size > 1:
            # exchange the cudaIpc handles of the regions the peers reduce into (acc_o, acc_d, cross-barrier counter)
            import torch.distributed as dist
            handles = [None] * self.tp_size
            dist.all_gather_object(handles, self.dplan.ipc_export(), group=self.tp_group)
            self.dplan.ipc_import(handles)
            dist.barrier(self.tp_group)
        self._plan_xout = torch.empty(self.shape.hidden, dtype=torch.float16, device=self.device)
        self._plan_logits = torch.empty((1, self.shape.vocab), dtype=torch.float32, device=self.device) if head is not None else None
        return self.dplan

    def decode_step_fused(self, hidden, past_len):
        """hidden: [1, 1, hidden] fp16 (read only).  Returns logits [1, vocab] fp32 (or the final hidden state without a head);
        the final pre-norm hidden state is left in self._plan_xout."""
        self.dplan.step(hidden, past_len, x_out=self._plan_xout, logits=self._plan_logits)
        return self._plan_logits if self._plan_logits is not None else self._plan_xout

    def _o_input(self, attn_local):
        """o_proj input of this rank: its own heads' output, or -- act-order + TP -- the rows of its group range out of the
        all-gathered attention output (the one extra exchange act-order costs, SURVEY.md 8e hazard 1)."""
        if self.o_rows is None:
            return attn_local
        full = tpmod.all_gather_columns(attn_local, [attn_local.shape[1]] * self.tp_size, self.tp_group)
        return full.index_select(1, self.o_rows)

    # ---- prefill: unfused path (model.py:532-550) --------------------------------------------------------------
    def prefill(self, hidden, past_len=0, last_only=True):
        s = self.shape
        bsz, q_len, _ = hidden.shape
        hq = self.local_heads * s.head_dim
        for i, L in enumerate(self.layers):
            residual = hidden
            x = cuda_ext.ext_rms_norm(hidden, L.ln1, s.eps)
            q = L.q.forward(x)
            k = L.k.forward(x)
            ext.rope_(q, self.sin, self.cos, past_len, self.local_heads, s.head_dim)
            ext.rope_(k, self.sin, self.cos, past_len, self.local_heads, s.head_dim)
            q = q.view(bsz, q_len, self.local_heads, s.head_dim).transpose(1, 2)
            k = k.view(bsz, q_len, self.local_heads, s.head_dim).transpose(1, 2)
            v = L.v.forward(x).view(bsz, q_len, self.local_heads, s.head_dim).transpose(1, 2)
            self.key_cache[i].narrow(2, past_len, q_len).copy_(k)
            self.value_cache[i].narrow(2, past_len, q_len).copy_(v)
            keys = self.key_cache[i].narrow(2, 0, past_len + q_len)
            values = self.value_cache[i].narrow(2, 0, past_len + q_len)
            if past_len > 0:
                raise NotImplementedError("chunked prefill with past is outside the benchmarked path")
            attn = torch.nn.functional.scaled_dot_product_attention(q, keys, values, attn_mask=None, is_causal=True)
            attn = attn.transpose(1, 2).reshape(bsz, q_len, hq)
            o = L.o.forward(self._o_input(attn.view(-1, hq)).view(bsz, q_len, hq))
            if self.tp_size > 1:
                tpmod.all_reduce(o, self.tp_group)
            hidden = residual + o
            residual = hidden
            x = cuda_ext.ext_rms_norm(hidden, L.ln2, s.eps)
            y = torch.nn.functional.silu(L.gate.forward(x))
            y *= L.up.forward(x)
            y = L.down.forward(y)
            if self.tp_size > 1:
                tpmod.all_reduce(y, self.tp_group)
            hidden = residual + y
        if self.lm_head is None:
            return hidden
        if last_only:
            hidden = hidden[:, -1:, :].contiguous()
        hn = cuda_ext.ext_rms_norm(hidden, self.norm, s.eps)
        return torch.matmul(hn.view(-1, s.hidden), self.lm_head.t()).float()

    def close(self):
        ext.cleanup()
