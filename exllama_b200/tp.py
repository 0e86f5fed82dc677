"""Tensor-parallel sharding of a GPTQ Llama layer and the per-projection all-reduce.

The reference has no tensor parallelism (only a sequential layer split, model.py:636-668); this is new
functionality on top of the operator surface:

  column-parallel (no exchange):  q, k, v  -> whole heads per rank;   gate, up -> column blocks per rank
  row-parallel (one all-reduce):  o_proj (rows = local heads' channels),  down_proj (rows = local gate/up columns)

Row shards must start on quantisation-group boundaries so qzeros/scales slice cleanly; when the number of groups
does not divide by the TP degree (65B: 22016 / 128 = 172 groups over 8 ranks) ranks get floor/ceil whole groups.

`plan_shards` and `shard_q4_*` are pure index arithmetic (tested on the CPU with numpy + gloo);
`row_parallel_residual` / `mlp_tp` issue the GPU ops through the plugin API and NCCL.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass
class ShardPlan:
    tp: int
    heads: list            # heads per rank
    head_cols: list        # (col0, col1) of q/k/v columns == o_proj rows per rank
    inter_cols: list       # (col0, col1) of gate/up columns == down_proj rows per rank


def _split_even(total_units, tp):
    base, rem = divmod(total_units, tp)
    sizes = [base + (1 if r < rem else 0) for r in range(tp)]
    offs = np.concatenate([[0], np.cumsum(sizes)]).tolist()
    return sizes, offs


def plan_shards(hidden, inter, heads, head_dim, groupsize, tp) -> ShardPlan:
    if heads % tp != 0:
        raise ValueError(f"{heads} heads do not divide over {tp} ranks (whole heads per rank are required)")
    hp = heads // tp
    if (hp * head_dim) % groupsize != 0 and tp > 1:
        raise ValueError("o_proj row shard is not aligned to the quantisation group size")
    head_cols = [(r * hp * head_dim, (r + 1) * hp * head_dim) for r in range(tp)]
    # intermediate dim: whole groups per rank, also a multiple of 128 columns for the kernels' tiles
    unit = int(np.lcm(groupsize, 128))
    if inter % unit != 0:
        unit = groupsize
    sizes, offs = _split_even(inter // unit, tp)
    inter_cols = [(offs[r] * unit, offs[r + 1] * unit) for r in range(tp)]
    return ShardPlan(tp, [hp] * tp, head_cols, inter_cols)


def shard_q4_columns(qweight, qzeros, scales, c0, c1):
    """Column (N) shard of a GPTQ tensor set: works on numpy arrays or torch tensors."""
    assert c0 % 8 == 0 and c1 % 8 == 0
    return qweight[:, c0:c1], qzeros[:, c0 // 8:c1 // 8], scales[:, c0:c1]


def shard_q4_rows(qweight, qzeros, scales, k0, k1, groupsize):
    """Row (K) shard on group boundaries (no act-order)."""
    assert k0 % groupsize == 0 and k1 % groupsize == 0 and k0 % 8 == 0
    return qweight[k0 // 8:k1 // 8], qzeros[k0 // groupsize:k1 // groupsize], scales[k0 // groupsize:k1 // groupsize]


# ---- act-order (g_idx) checkpoints ---------------------------------------------------------------------------------
# Column-parallel projections need nothing special: every rank holds all K rows, hence the whole g_idx.
# Row-parallel projections cannot be cut on contiguous k: the rows of one quantisation group are scattered over k
# (SURVEY.md 8e hazard 1).  They are cut on GROUP ranges instead -- rank r owns the rows whose group lies in
# [g0, g1) -- which keeps every local group complete (exactly `groupsize` rows), so the shard is itself a valid
# act-order GPTQ matrix with local g_idx = g_idx[rows] - g0 that make_q4 / the kernels take unchanged.  The rank's
# input is then x[:, rows]:
#   * down_proj: gate/up are column-GATHERED with the same `rows`, so each rank's silu(gate)*up already is x[:, rows];
#   * o_proj:    `rows` spans all heads, so the local heads' attention output is all-gathered once and indexed.

def unpack_rows(qweight):
    """int32 [K/8, N] (8 nibbles along K per word, exllama_ext.cpp:166-176) -> uint8 [K, N]."""
    qw = np.ascontiguousarray(qweight).view(np.uint32)
    out = np.empty((qw.shape[0], 8, qw.shape[1]), dtype=np.uint8)
    for i in range(8):
        out[:, i, :] = (qw >> np.uint32(4 * i)) & np.uint32(15)
    return out.reshape(qw.shape[0] * 8, qw.shape[1])


def pack_rows(w):
    """uint8 [K, N] -> int32 [K/8, N]."""
    K, N = w.shape
    assert K % 8 == 0
    w3 = w.reshape(K // 8, 8, N).astype(np.uint32)
    out = np.zeros((K // 8, N), dtype=np.uint32)
    for i in range(8):
        out |= w3[:, i, :] << np.uint32(4 * i)
    return out.view(np.int32)


def unpack_cols(qzeros):
    """int32 [G, N/8] (8 nibbles along N per word) -> uint8 [G, N]."""
    qz = np.ascontiguousarray(qzeros).view(np.uint32)
    out = np.empty((qz.shape[0], qz.shape[1], 8), dtype=np.uint8)
    for i in range(8):
        out[:, :, i] = (qz >> np.uint32(4 * i)) & np.uint32(15)
    return out.reshape(qz.shape[0], qz.shape[1] * 8)


def pack_cols(z):
    G, N = z.shape
    assert N % 8 == 0
    z3 = z.reshape(G, N // 8, 8).astype(np.uint32)
    out = np.zeros((G, N // 8), dtype=np.uint32)
    for i in range(8):
        out |= z3[:, :, i] << np.uint32(4 * i)
    return out.view(np.int32)


def plan_group_ranges(groups, tp):
    """[(g0, g1)] per rank: whole quantisation groups, floor/ceil when they do not divide."""
    _, offs = _split_even(groups, tp)
    return [(offs[r], offs[r + 1]) for r in range(tp)]


def act_order_row_shard(qweight, qzeros, scales, g_idx, g0, g1, groupsize):
    """Row shard of an act-order GPTQ matrix on the group range [g0, g1) (numpy, host side, load time).
    Returns (qweight_local, qzeros_local, scales_local, g_idx_local, rows): rows = ascending original k of the shard;
    the shard multiplies x[:, rows]."""
    g_idx = np.asarray(g_idx)
    rows = np.nonzero((g_idx >= g0) & (g_idx < g1))[0]
    assert rows.size == (g1 - g0) * groupsize, "every quantisation group must have exactly `groupsize` rows"
    qw = pack_rows(unpack_rows(qweight)[rows])
    return (qw, np.ascontiguousarray(qzeros[g0:g1]), np.ascontiguousarray(scales[g0:g1]),
            (g_idx[rows] - g0).astype(np.int32), rows)


def gather_q4_columns(qweight, qzeros, scales, cols):
    """Column gather (arbitrary column list, len % 8 == 0) of a GPTQ tensor set; g_idx (a K property) is unaffected."""
    cols = np.asarray(cols)
    assert cols.size % 8 == 0
    return (np.ascontiguousarray(np.asarray(qweight)[:, cols]), pack_cols(unpack_cols(qzeros)[:, cols]),
            np.ascontiguousarray(np.asarray(scales)[:, cols]))


def all_gather_columns(t_local, sizes, group=None):
    """[M, n_local] per rank -> [M, sum(sizes)] on every rank (the o_proj input exchange of the act-order case)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    if len(set(sizes)) == 1:
        parts = [torch.empty_like(t_local) for _ in range(world)]
    else:
        parts = [t_local.new_empty((t_local.shape[0], n)) for n in sizes]
    dist.all_gather(parts, t_local.contiguous(), group=group)
    return torch.cat(parts, dim=1)


def all_reduce(t, group=None):
    import torch.distributed as dist
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def row_parallel_residual(ext, x, inp, q4, rank, group):
    """x (replicated, [M, hidden]) += all_reduce(inp_local . W_rowshard).
    Rank 0 folds the residual into its partial (no_zero accumulate, exactly q4_attn_2); the other ranks
    overwrite their copy of x with their partial, so one in-place all-reduce leaves x_old + sum(partials) everywhere."""
    if _fused_ready and x.shape[0] <= 8:
        ext.q4_attn_2_ar(x, inp, q4)           # GEMV + one-shot all-reduce + residual in ONE kernel (no NCCL)
        return
    ext.q4_attn_2_tp(x, inp, q4, rank == 0)
    all_reduce(x, group)


def mlp_tp(ext, cuda_ext, x, L, eps, rank, group):
    """Tensor-parallel MLP block: the same two fused launches as the single-GPU q4_mlp ([norm -> gate,up -> silu*mul],
    [down]) on this rank's column / row shards, then ONE all-reduce of the [M, hidden] partial."""
    if _fused_ready and x.shape[0] <= 8 and not getattr(L, "_no_fused_mlp", False):
        try:
            ext.q4_mlp_ar(x, L.ln2, eps, L.gate.q4, L.up.q4, L.down.q4)
            return
        except RuntimeError:
            # act-order (gate / up carry different x_maps) or LoRA: the single-launch configuration does not apply; the
            # entry point refuses BEFORE launching anything, so falling back to the two-step variant + NCCL is safe
            L._no_fused_mlp = True
    ext.q4_mlp_tp(x, L.ln2, eps, L.gate.q4, L.up.q4, L.down.q4, rank == 0)
    all_reduce(x, group)


# ---------------------------------------------------------------------------------------------------------------------
# fused projection + all-reduce over NVLink peer memory (exl_q4_attn_2_ar / exl_q4_mlp_ar)
# ---------------------------------------------------------------------------------------------------------------------

_fused_ready = False


def init_fused_allreduce(ext, device_index, group=None):
    """Exchange the cudaIpc handles of the per-rank workspaces (through the process group) and install the peer table.
    Afterwards the row-parallel projections need no NCCL call: their GEMV epilogue does the one-shot all-reduce itself."""
    global _fused_ready
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    local_ptr, handle = ext.tp_workspace_alloc(device_index)
    handles = [None] * world
    dist.all_gather_object(handles, bytes(handle), group=group)
    ptrs = [local_ptr if r == rank else ext.tp_workspace_open(device_index, handles[r]) for r in range(world)]
    ext.tp_init(device_index, rank, world, ptrs)
    dist.barrier(group)
    _fused_ready = True
    return ptrs


def fused_ready():
    return _fused_ready


def reset_fused_allreduce():
    """Call after ext.cleanup(): the workspace behind the fused all-reduce is gone."""
    global _fused_ready
    _fused_ready = False
