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


# ---- act-order checkpoints: row shards on group ranges, gate/up column-gathered to match, o_proj input all-gathered ----
def test_pack_unpack_roundtrip(oracle):
    from exllama_b200 import tp
    qw, qz, sc, _ = oracle.synth_q4(256, 128, 32, seed=2)
    np.testing.assert_array_equal(tp.pack_rows(tp.unpack_rows(qw)), qw)
    np.testing.assert_array_equal(tp.pack_cols(tp.unpack_cols(qz)), qz)
    # unpack_rows agrees with the oracle's dequantiser: W = s * (q - (z + 1))
    W = oracle.dequant_f64(qw, qz, sc)
    q = tp.unpack_rows(qw).astype(np.float64); z = tp.unpack_cols(qz).astype(np.float64)
    g = np.arange(256) // 32
    np.testing.assert_allclose(sc.astype(np.float64)[g] * (q - (z[g] + 1)), W, rtol=0, atol=0)


def test_act_order_row_shard_matches_dense(oracle):
    from exllama_b200 import tp
    K, N, gs = 512, 256, 64
    qw, qz, sc, g_idx = oracle.synth_q4(K, N, gs, act_order=True, seed=21)
    x = oracle.synth_x(3, K, seed=22)
    dense = oracle.ref64_with_act_order(x, qw.copy(), qz, sc, g_idx)
    total = np.zeros_like(dense)
    seen = []
    for g0, g1 in tp.plan_group_ranges(K // gs, 3):          # 8 groups over 3 ranks: 3 / 3 / 2
        a, b, c, gl, rows = tp.act_order_row_shard(qw, qz, sc, g_idx, g0, g1, gs)
        assert a.shape == ((g1 - g0) * gs // 8, N) and gl.min() == 0 and gl.max() == g1 - g0 - 1
        total += oracle.ref64_with_act_order(np.ascontiguousarray(x[:, rows]), a.copy(), b, c, gl)
        seen.append(rows)
    np.testing.assert_array_equal(np.sort(np.concatenate(seen)), np.arange(K))
    np.testing.assert_allclose(total, dense, rtol=1e-12, atol=1e-12 * np.abs(dense).max())


def test_gather_columns_matches_dense(oracle):
    from exllama_b200 import tp
    K, N, gs = 256, 512, 64
    qw, qz, sc, _ = oracle.synth_q4(K, N, gs, seed=23)
    cols = np.random.default_rng(0).permutation(N)[:192]
    a, b, c = tp.gather_q4_columns(qw, qz, sc, cols)
    np.testing.assert_array_equal(oracle.dequant_f64(a, b, c), oracle.dequant_f64(qw, qz, sc)[:, cols])


def _worker_act(rank, world, port, hidden, inter, heads, gs, out_q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from exllama_b200 import tp
    from oracle import oracle as O
    plan = tp.plan_shards(hidden, inter, heads, hidden // heads, gs, world)
    o_full = O.synth_q4(hidden, hidden, gs, act_order=True, seed=30)
    gate_full = O.synth_q4(hidden, inter, gs, act_order=True, seed=31)
    up_full = O.synth_q4(hidden, inter, gs, act_order=True, seed=32)
    up_full = up_full[:3] + (gate_full[3],)               # gate and up share the input, hence the g_idx
    down_full = O.synth_q4(inter, hidden, gs, act_order=True, seed=33)
    x = O.synth_x(2, hidden, seed=34)
    attn = O.synth_x(2, hidden, seed=35)
    c0, c1 = plan.head_cols[rank]
    # o_proj: group-range row shard; its input is the all-gathered attention output indexed by `rows`
    g0, g1 = tp.plan_group_ranges(hidden // gs, world)[rank]
    oa, ob, oc, ogl, orows = tp.act_order_row_shard(*o_full, g0, g1, gs)
    attn_full = tp.all_gather_columns(torch.from_numpy(np.ascontiguousarray(attn[:, c0:c1])),
                                      [b - a for a, b in plan.head_cols]).numpy()
    part = O.ref64_with_act_order(np.ascontiguousarray(attn_full[:, orows]), oa.copy(), ob, oc, ogl,
                                  acc_in=x if rank == 0 else None)
    t = torch.from_numpy(part); tp.all_reduce(t)
    x1 = t.numpy().astype(np.float16)
    # MLP: down is cut on group ranges, gate/up are column-gathered with the same rows -> no exchange before down
    d0, d1 = tp.plan_group_ranges(inter // gs, world)[rank]
    da, db, dc, dgl, drows = tp.act_order_row_shard(*down_full, d0, d1, gs)
    ga, gb, gc = tp.gather_q4_columns(*gate_full[:3], drows)
    ua, ub, uc = tp.gather_q4_columns(*up_full[:3], drows)
    g = O.ref64_with_act_order(x1, ga.copy(), gb, gc, gate_full[3]).astype(np.float16)
    u = O.ref64_with_act_order(x1, ua.copy(), ub, uc, up_full[3]).astype(np.float16)
    part = O.ref64_with_act_order(O.silu_mul(g, u), da.copy(), db, dc, dgl, acc_in=x1 if rank == 0 else None)
    t = torch.from_numpy(part); tp.all_reduce(t)
    if rank == 0:
        out_q.put((x1, t.numpy()))
    dist.destroy_process_group()


def test_tp2_act_order_matches_single_rank_gloo(oracle):
    hidden, inter, heads, gs = 256, 512, 2, 64
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_act, args=(r, 2, port, hidden, inter, heads, gs, q)) for r in range(2)]
    for p in procs: p.start()
    x1_tp, x2_tp = q.get(timeout=120)
    for p in procs: p.join(timeout=60); assert p.exitcode == 0
    O = oracle
    o_full = O.synth_q4(hidden, hidden, gs, act_order=True, seed=30)
    gate_full = O.synth_q4(hidden, inter, gs, act_order=True, seed=31)
    up_full = O.synth_q4(hidden, inter, gs, act_order=True, seed=32); up_full = up_full[:3] + (gate_full[3],)
    down_full = O.synth_q4(inter, hidden, gs, act_order=True, seed=33)
    x = O.synth_x(2, hidden, seed=34); attn = O.synth_x(2, hidden, seed=35)
    cp = lambda f: (f[0].copy(),) + tuple(f[1:])
    x1 = O.ref64_with_act_order(attn, *cp(o_full), acc_in=x)
    np.testing.assert_allclose(x1_tp.astype(np.float64), x1, rtol=2e-3, atol=2e-3 * np.abs(x1).max())
    g = O.ref64_with_act_order(x1_tp, *cp(gate_full)).astype(np.float16); u = O.ref64_with_act_order(x1_tp, *cp(up_full)).astype(np.float16)
    x2 = O.ref64_with_act_order(O.silu_mul(g, u), *cp(down_full), acc_in=x1_tp)
    np.testing.assert_allclose(x2_tp, x2, rtol=1e-9, atol=1e-9 * np.abs(x2).max())


def test_plan_baseline_configs():
    """Shard plans of the tensor-parallel BASELINE configs (33B g32 over 2 and 4 ranks, 65B g128 over 8): whole heads, whole
    groups, kernel granularities (row shards % 32, column shards % 32), and the act-order group ranges tile the groups."""
    from exllama_b200 import tp
    for hidden, inter, heads, gs, worlds in ((6656, 17920, 52, 32, (2, 4)), (8192, 22016, 64, 128, (2, 4, 8)), (4096, 11008, 32, 128, (2, 4, 8))):
        for w in worlds:
            p = tp.plan_shards(hidden, inter, heads, 128, gs, w)
            assert sum(p.heads) == heads and p.head_cols[-1][1] == hidden and p.inter_cols[-1][1] == inter
            for (a, b) in p.head_cols + p.inter_cols:
                assert a % gs == 0 and b % gs == 0 and (b - a) % 32 == 0 and b > a
            for total in (hidden // gs, inter // gs):
                r = tp.plan_group_ranges(total, w)
                assert r[0][0] == 0 and r[-1][1] == total and all(x[1] == y[0] for x, y in zip(r, r[1:]))
                assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1
                assert all(((b - a) * gs) % 32 == 0 for a, b in r)
