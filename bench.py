#!/usr/bin/env python
"""bench.py -- headline measurement of the GPTQ-4bit hot path on B200 (contract: see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--model 7b] [--ctx 1920]

One "step" = one decode token through the synthetic Llama-shaped decoder stack (exllama_b200/stack.py): for each of
the 32 layers  q4_attn -> attention over the KV cache (torch, as model.py) -> q4_attn_2 -> q4_mlp, then final norm +
fp16 lm_head.  Weights are random GPTQ tensors of the named architecture (no checkpoints exist offline).

  value  : decode tok/s, whole step replayed as a CUDA graph, inputs resident in HBM  (K steps timed with CUDA events)
  e2e    : same metric through the reference-facing plugin API called eagerly from Python, with the step's input
           hidden state copied from pinned host memory and the logits copied back inside the timed region
  roofline : the dominant kernel (fused gate+up q4 GEMV launch), algorithmic bytes / mean launch time over all layers
  cpu_baseline : the oracle's CPU port of dequant + GEMV on the box's host cores (bounded sample)
  prefill: prompt tok/s for a (seq - 128)-token forward, reported beside decode

--impl reference times the CPU restatement (the reference ships no CPU path).
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=64)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--model", default="7b")
    ap.add_argument("--groupsize", type=int, default=128)
    ap.add_argument("--ctx", type=int, default=1920, help="KV-cache length the decode step attends over")
    ap.add_argument("--seq", type=int, default=2048)
    ap.add_argument("--no-prefill", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--torch-attn", action="store_true", help="attention between q4_attn and q4_attn_2 with the reference's torch ops instead of csrc/decode_attn.cu")
    return ap.parse_args()


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons through NVML during the timed region."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
            nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
        }
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.05)

    def result(self):
        self.stop_flag = True
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops", 1590.0), "measured"
    return 6650.0, 1590.0, "fallback"


def q4_bytes(K, N, gs, M=1, act=False, accumulate=False):
    """Algorithmic bytes of one q4_matmul."""
    return K * N // 2 + 2 * (K // gs) * N + 4 * (K // gs) * (N // 8) + 2 * M * K + 2 * M * N * (2 if accumulate else 1) + (4 * K if act else 0)


def cpu_baseline_sample(shape, gs, seconds_budget=12.0):
    """Oracle CPU port timed on the host: one decoder layer's seven q4 matmuls at M=1 (bounded sample)."""
    import numpy as np
    from oracle import oracle as O
    O.build()
    dims = [(shape.hidden, shape.hidden)] * 4 + [(shape.hidden, shape.inter)] * 2 + [(shape.inter, shape.hidden)]
    tensors = []
    for i, (K, N) in enumerate(dims):
        qw, qz, sc, _ = O.synth_q4(K, N, gs, seed=i)
        tensors.append((O.synth_x(1, K, seed=i), qw, qz, sc))
    def layer():
        for x, qw, qz, sc in tensors:
            O.q4_matmul_cpu_f32(x, qw, qz, sc)
    layer()
    t0 = time.perf_counter(); n = 0
    while True:
        layer(); n += 1
        if time.perf_counter() - t0 > seconds_budget or n >= 20:
            break
    ms_layer = (time.perf_counter() - t0) * 1e3 / n
    return ms_layer, O.num_threads(), n


def workload_name(shape, gs, ctx, seq):
    return f"{shape.name}-gptq4-g{gs}-noact decode token at ctx {ctx} of seq {seq}"


def run_reference(args):
    """--impl reference: the CPU restatement of the path (the reference has no CPU implementation)."""
    from exllama_b200.stack import SHAPES
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    shape = SHAPES[args.model]
    samples = []
    ms_layer, threads, _ = cpu_baseline_sample(shape, args.groupsize, seconds_budget=6.0)      # warm-up
    for _ in range(max(1, min(args.steps, 3))):
        ms_layer, threads, n = cpu_baseline_sample(shape, args.groupsize, seconds_budget=8.0)
        samples.append(ms_layer)
    ms_layer = sorted(samples)[len(samples) // 2]
    ms_tok = ms_layer * shape.layers
    val = 1000.0 / ms_tok
    sample = (f"oracle port (dequant + fp32 GEMV, OpenMP) of one {shape.name} decoder layer's 7 q4 matmuls at M=1, "
              f"x{shape.layers} layers; attention/lm_head not included")
    line = {
        "impl": "reference", "metric": "decode tok/s Llama-7B 4b GPTQ g128 (q4 hot path)", "value": round(val, 4), "unit": "tok/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_tok, 3),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(shape, args.groupsize, args.ctx, args.seq),
                   "detail": "CPU restatement of the q4 matmuls of the step, bounded sample (the reference has no CPU path)"},
        "cpu_baseline": {"value": round(val, 4), "unit": "tok/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": round(val, 4), "unit": "tok/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from exllama_b200 import capi
    from exllama_b200.stack import SHAPES, DecodeStack

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"          # keep stdout to the one JSON line
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    else:
        torch.cuda.set_device(0)
    dev = torch.device(f"cuda:{local_rank}")
    shape = SHAPES[args.model]
    gs = args.groupsize
    hbm_peak, tf_peak, peak_kind = measured_peaks()

    stack = DecodeStack(shape, groupsize=gs, act_order=False, device=str(dev), max_seq=args.seq,
                        tp_rank=rank, tp_size=world, tp_group=None)
    stack.fused_decode_attn = not args.torch_attn
    fused_ar = False
    if world > 1 and os.environ.get("EXL_TP_FUSED", "1") == "1":
        from exllama_b200 import tp as tpmod, cuda_ext as _ce
        tpmod.init_fused_allreduce(_ce.exllama_ext, dev.index)      # row-parallel projections: GEMV + peer-memory all-reduce in one kernel
        fused_ar = True
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ------------------------------------------------------------------ decode, device-resident, CUDA graph
    hidden0 = (torch.randn((1, 1, shape.hidden), device=dev) * 0.5).half()
    hidden = hidden0.clone()
    past = args.ctx
    # fill the cache so attention reads real data
    for kc, vc in zip(stack.key_cache, stack.value_cache):
        kc.normal_(0, 0.5); vc.normal_(0, 0.5)

    def step_eager():
        hidden.copy_(hidden0)
        return stack.decode_step(hidden, past)

    launches0 = capi.launch_count()
    logits = step_eager()
    torch.cuda.synchronize()
    launches_per_step = capi.launch_count() - launches0
    assert torch.isfinite(logits).all(), "non-finite logits in the synthetic stack"
    for _ in range(2):
        step_eager()
    torch.cuda.synchronize()

    graph = None
    # NCCL inside stream capture hung on this stack; the fused all-reduce kernels are plain launches and capture fine
    use_graph = not args.no_graph and (world == 1 or fused_ar or os.environ.get("EXL_BENCH_TP_GRAPH", "0") == "1")
    if use_graph:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            step_eager()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, capture_error_mode="thread_local"):
            glogits = step_eager()

    def step_device():
        if graph is not None:
            graph.replay()
        else:
            step_eager()

    for _ in range(max(args.warmup, 3)):
        step_device()
    barrier()
    sampler = ClockSampler(local_rank); sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.cudart().cudaProfilerStart()          # no-op unless run under `ncu --profile-from-start off`
    e0.record()
    for _ in range(args.steps):
        step_device()
    e1.record()
    barrier()
    torch.cuda.cudart().cudaProfilerStop()
    ms = e0.elapsed_time(e1)
    clocks = sampler.result()
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t.item()) / args.steps
    value = 1000.0 / ms_per_step

    # the same step with the reference's torch attention ops (model.py:395-409) between our launches, for comparison
    torch_attn_ms = None
    if world == 1 and graph is not None and stack.fused_decode_attn:
        stack.fused_decode_attn = False
        for _ in range(2): step_eager()
        torch.cuda.synchronize()
        g2 = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g2, capture_error_mode="thread_local"):
            step_eager()
        for _ in range(3): g2.replay()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(16): g2.replay()
        b.record(); torch.cuda.synchronize()
        torch_attn_ms = a.elapsed_time(b) / 16
        del g2
        stack.fused_decode_attn = True

    # "best" case of the reference's benchmark: nearly empty context
    def time_ctx(p, n=16):
        nonlocal past
        keep = past; past = p
        for _ in range(3): step_eager()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n): step_eager()
        b.record(); torch.cuda.synchronize()
        past = keep
        return a.elapsed_time(b) / n

    # ------------------------------------------------------------------ e2e: eager plugin API + host copies
    host_in = torch.randn((1, 1, shape.hidden)).half().pin_memory()
    host_out = torch.empty((1, shape.vocab), dtype=torch.float32).pin_memory()

    def step_e2e():
        hidden.copy_(host_in, non_blocking=True)
        lg = stack.decode_step(hidden, past)
        host_out.copy_(lg, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    for _ in range(3):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    n_e2e = max(8, min(args.steps, 32))
    for _ in range(n_e2e):
        step_e2e()
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / n_e2e
    te = torch.tensor([e2e_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_ms = float(te.item())

    # ------------------------------------------------------------------ roofline of the dominant kernel
    # fused gate+up launch (2 x [hidden -> inter]) over all layers: 32 distinct weight sets (>> L2), CUDA events.
    il = stack.layers[0].gate.width
    x1 = (torch.randn((1, shape.hidden), device=dev) * 0.5).half()
    none = stack.none
    from exllama_b200 import cuda_ext
    ext = cuda_ext.exllama_ext
    kern_rows = []

    def time_launches(fn, reps=8):
        """mean device time of one pass of fn, replayed as a CUDA graph (no host overhead inside the timed region)"""
        fn(); torch.cuda.synchronize()
        side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            fn()
        torch.cuda.current_stream().wait_stream(side); torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            fn()
        g.replay(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            g.replay()
        b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b) * 1e3 / reps       # us per pass

    nl = len(stack.layers)
    xm = x1.clone()
    def gateup_down():
        for L in stack.layers:
            ext.q4_mlp(xm, L.ln2, shape.eps, L.gate.q4, L.up.q4, L.down.q4, none, none, none, none, none, none, none)
    def down_only():
        tmp = stack.temp_mlp[:1, :il]
        for L in stack.layers:
            ext.q4_attn_2(xm, tmp, L.down.q4, none, none, none)
    def o_only():
        tmp = x1[:, :stack.layers[0].o.height]
        for L in stack.layers:
            ext.q4_attn_2(xm, tmp, L.o.q4, none, none, none)
    us_mlp = time_launches(gateup_down) / nl
    us_down = time_launches(down_only) / nl
    us_o = time_launches(o_only) / nl
    us_gateup = us_mlp - us_down
    b_gateup = 2 * (q4_bytes(shape.hidden, il, gs) - 2 * shape.hidden) + 2 * shape.hidden - 2 * il   # x once, one fp16 out
    b_down = q4_bytes(il, shape.hidden, gs, accumulate=True)
    b_o = q4_bytes(stack.layers[0].o.height, shape.hidden, gs, accumulate=True)
    for name, us, by in (("gate+up fused (norm prologue, silu*mul epilogue)", us_gateup, b_gateup),
                         ("down (+residual)", us_down, b_down), ("o_proj (+residual)", us_o, b_o)):
        kern_rows.append({"kernel": name, "us": round(us, 3), "bytes": by, "GBps": round(by / us / 1e3, 1), "frac": round(by / us / 1e3 / hbm_peak, 4)})
    dom = kern_rows[0]
    traffic = None
    tp_path = os.path.join(ROOT, "profiles", "traffic_r1.json")
    if world == 1 and args.model == "7b" and gs == 128 and os.path.exists(tp_path):
        traffic = json.load(open(tp_path)).get("traffic")      # dram read+write of this kernel from the committed ncu --set full capture
    roofline = {"bound": "hbm", "achieved": dom["GBps"], "peak": hbm_peak, "unit": "GB/s", "frac": dom["frac"],
                "traffic": traffic, "kernel": "q4_gemv_kernel<RMSNORM,SILU_MUL> (fused gate+up)", "peak_kind": peak_kind,
                "algorithmic_bytes_per_launch": dom["bytes"], "us_per_launch": dom["us"]}

    # ------------------------------------------------------------------ prefill
    prefill = None
    if not args.no_prefill:
        T = args.seq - 128
        hp = (torch.randn((1, T, shape.hidden), device=dev) * 0.5).half()
        try:
            for _ in range(2):
                stack.prefill(hp)
            barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            nrep = 3
            for _ in range(nrep):
                stack.prefill(hp)
            b.record(); barrier()
            pms = a.elapsed_time(b) / nrep
            flops = 2.0 * T * (stack.q4_weight_bytes_per_token() * 2 / 1.0) * 0 + 2.0 * T * sum(
                (l.q.height * l.q.width * 3 + l.o.height * l.o.width + l.gate.height * l.gate.width * 2 + l.down.height * l.down.width)
                for l in stack.layers)
            prefill = {"value": round(T / pms * 1e3, 1), "unit": "tok/s", "tokens": T, "ms": round(pms, 3),
                       "q4_linear_tflops": round(flops / pms / 1e9, 1), "q4_path": capi.last_q4_path()}
        except Exception as ex:  # noqa: BLE001
            prefill = {"error": repr(ex)[:200]}

    # ------------------------------------------------------------------ cpu baseline (rank 0, N == 1)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        ms_layer, threads, n = cpu_baseline_sample(shape, gs)
        cpu = {"value": round(1000.0 / (ms_layer * shape.layers), 4), "unit": "tok/s", "cores": threads, "kind": "port",
               "sample": f"oracle CPU port (dequant + fp32 GEMV, OpenMP) of one decoder layer's 7 q4 matmuls, M=1, x{shape.layers}; {n} reps"}

    best_ms = time_ctx(4) if world == 1 else None
    tp_timeouts = None
    if fused_ar:
        tt = torch.tensor([_ce.exllama_ext.tp_status(dev.index)], device=dev, dtype=torch.int64)
        dist.all_reduce(tt, op=dist.ReduceOp.SUM)
        tp_timeouts = int(tt.item())          # flag waits of the fused all-reduce that gave up, over all ranks: must be 0

    if rank == 0:
        line = {
            "metric": "decode tok/s Llama-7B 4b GPTQ g128 (q4 hot path)" if args.model == "7b" else f"decode tok/s {shape.name} 4b GPTQ",
            "value": round(value, 2), "unit": "tok/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": round(ms_per_step, 4), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f16", "data": "synthetic",
            "config": {"workload": workload_name(shape, gs, past, args.seq),
                       "detail": f"{shape.layers} layers x (q4_attn, " + ("decode_attn kernel" if stack.fused_decode_attn else "torch attention ops") +
                                 " over the KV cache, q4_attn_2, q4_mlp) + final norm + fp16 lm_head",
                       "parallelism": f"tp{world}", "cuda_graph": graph is not None, "allreduce": ("fused GEMV epilogue over NVLink peer memory" if fused_ar else ("nccl" if world > 1 else None)),
                       "l2": "weights (3.6 GB/token) >> L2, every step streams them from HBM"},
            "e2e": {"value": round(1000.0 / e2e_ms, 2), "unit": "tok/s", "h2d_bytes_per_step": host_in.numel() * 2,
                    "d2h_bytes_per_step": host_out.numel() * 4, "ms_per_step": round(e2e_ms, 4), "mode": "eager plugin API"},
            "gpu_launches": int(launches_per_step * args.steps),
            "launches_per_step": int(launches_per_step),
            "clocks": clocks,
            "roofline": roofline,
            "kernels": kern_rows,
            "cpu_baseline": cpu,
            "prefill": prefill,
            "tp_allreduce_timeouts": tp_timeouts,
            "decode_torch_attention": {"value": round(1000.0 / torch_attn_ms, 2), "unit": "tok/s", "mode": "cuda graph",
                                       "note": "same step with the reference's torch attention ops instead of decode_attn"} if torch_attn_ms else None,
            "decode_best_ctx4": {"value": round(1000.0 / best_ms, 2), "unit": "tok/s", "mode": "eager"} if best_ms else None,
            "q4_weight_bytes_per_token": stack.q4_weight_bytes_per_token(),
            "weights_only_bound_tok_s": round(hbm_peak * 1e9 / (stack.q4_weight_bytes_per_token() + shape.vocab * shape.hidden * 2), 1),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
