"""CPU tests of the tensor-parallel host logic (exllama_b200/tp.py) with world_size 2 over gloo: shard planning,
GPTQ tensor slicing, and the 'rank 0 keeps the residual, one in-place all-reduce' protocol -- with the oracle standing
in for the GPU kernels."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def test_plan_shards_shapes():
    from exllama_b200 import tp
    # 7B / 65B shapes from SURVEY.md 8e; 65B: 172 groups of 128 do not divide by 8 -> floor/ceil whole units
    p = tp.plan_shards(4096, 11008, 32, 128, 128, 4)
    assert p.heads == [8] * 4 and p.head_cols[1] == (1024, 2048)
    assert [b - a for a, b in p.inter_cols] == [2816, 2816, 2688, 2688] and p.inter_cols[-1][1] == 11008
    p = tp.plan_shards(8192, 22016, 64, 128, 128, 8)
    sizes = [b - a for a, b in p.inter_cols]
    assert sum(sizes) == 22016 and set(sizes) == {2816, 2688} and all(s % 128 == 0 for s in sizes)
    assert all(a % 128 == 0 for a, _ in p.inter_cols)
    with pytest.raises(ValueError):
        tp.plan_shards(6656, 17920, 52, 128, 128, 8)      # 52 heads do not split over 8 ranks
    p = tp.plan_shards(6656, 17920, 52, 128, 32, 4)
    assert sum(b - a for a, b in p.inter_cols) == 17920


def test_shard_slicing_matches_dense(oracle):
    from exllama_b200 import tp
    K, N, gs = 512, 256, 64
    qw, qz, sc, _ = oracle.synth_q4(K, N, gs, seed=4)
    W = oracle.dequant_f64(qw, qz, sc)
    a, b, c = tp.shard_q4_columns(qw, qz, sc, 64, 192)
    np.testing.assert_array_equal(oracle.dequant_f64(np.ascontiguousarray(a), np.ascontiguousarray(b), np.ascontiguousarray(c)), W[:, 64:192])
    a, b, c = tp.shard_q4_rows(qw, qz, sc, 128, 384, gs)
    np.testing.assert_array_equal(oracle.dequant_f64(np.ascontiguousarray(a), np.ascontiguousarray(b), np.ascontiguousarray(c)), W[128:384])


def _worker(rank, world, port, hidden, inter, heads, gs, out_q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from exllama_b200 import tp
    from oracle import oracle as O
    plan = tp.plan_shards(hidden, inter, heads, hidden // heads, gs, world)
    # identical full tensors on every rank (same seeds), each rank slices its shard
    o_full = O.synth_q4(hidden, hidden, gs, seed=10)[:3]
    gate_full = O.synth_q4(hidden, inter, gs, seed=11)[:3]
    up_full = O.synth_q4(hidden, inter, gs, seed=12)[:3]
    down_full = O.synth_q4(inter, hidden, gs, seed=13)[:3]
    x = O.synth_x(2, hidden, seed=14)
    attn = O.synth_x(2, hidden, seed=15)
    c0, c1 = plan.head_cols[rank]
    i0, i1 = plan.inter_cols[rank]
    cont = lambda t: tuple(np.ascontiguousarray(v) for v in t)
    # row-parallel o_proj: local heads' channels of attn, row shard of W_o; rank 0 keeps the residual
    o_sh = cont(tp.shard_q4_rows(*o_full, c0, c1, gs))
    part = O.q4_matmul_f64(np.ascontiguousarray(attn[:, c0:c1]), *o_sh, acc_in=x if rank == 0 else None)
    t = torch.from_numpy(part)
    tp.all_reduce(t)                                   # the ONE exchange of the projection
    x1 = t.numpy().astype(np.float16)
    # MLP: column shards of gate/up, row shard of down
    g_sh, u_sh = cont(tp.shard_q4_columns(*gate_full, i0, i1)), cont(tp.shard_q4_columns(*up_full, i0, i1))
    d_sh = cont(tp.shard_q4_rows(*down_full, i0, i1, gs))
    g = O.q4_matmul_f64(x1, *g_sh).astype(np.float16); u = O.q4_matmul_f64(x1, *u_sh).astype(np.float16)
    act = O.silu_mul(g, u)
    part = O.q4_matmul_f64(act, *d_sh, acc_in=x1 if rank == 0 else None)
    t = torch.from_numpy(part)
    tp.all_reduce(t)
    if rank == 0:
        out_q.put((x1, t.numpy()))
    dist.destroy_process_group()


def test_tp2_matches_single_rank_gloo(oracle):
    hidden, inter, heads, gs = 256, 512, 2, 64
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, hidden, inter, heads, gs, q)) for r in range(2)]
    for p in procs: p.start()
    x1_tp, x2_tp = q.get(timeout=120)
    for p in procs: p.join(timeout=60); assert p.exitcode == 0
    O = oracle
    o_full = O.synth_q4(hidden, hidden, gs, seed=10)[:3]; gate_full = O.synth_q4(hidden, inter, gs, seed=11)[:3]
    up_full = O.synth_q4(hidden, inter, gs, seed=12)[:3]; down_full = O.synth_q4(inter, hidden, gs, seed=13)[:3]
    x = O.synth_x(2, hidden, seed=14); attn = O.synth_x(2, hidden, seed=15)
    x1 = O.q4_matmul_f64(attn, *o_full, acc_in=x)
    np.testing.assert_allclose(x1_tp.astype(np.float64), x1, rtol=2e-3, atol=2e-3 * np.abs(x1).max())
    x1h = x1_tp                                       # continue from the same fp16 state as the TP run
    g = O.q4_matmul_f64(x1h, *gate_full).astype(np.float16); u = O.q4_matmul_f64(x1h, *up_full).astype(np.float16)
    x2 = O.q4_matmul_f64(O.silu_mul(g, u), *down_full, acc_in=x1h)
    np.testing.assert_allclose(x2_tp, x2, rtol=1e-9, atol=1e-9 * np.abs(x2).max())
