#!/usr/bin/env python
"""bench.py — decode tok/s of the sharded transformer-block forward path (BASELINE.json's metric):
Llama-3-8B bf16, bs=1, 2k context, layers sharded over N GPUs of one box (N=1 -> all local).

    python bench.py [--gpus N] [--steps K] [--warmup W]          # this repo's CUDA path
    python bench.py --impl reference [...]                        # the reference's CPU path (oracle port)
    torchrun --nproc-per-node N bench.py --gpus N ...             # N > 1, one rank per GPU

A "step" is one decoded token: one pass of the hot path (embed -> 32 blocks -> ln_f/lm_head/argmax) at
batch 1.  tok/s follows the reference's definition (master.rs:131-166): prefill excluded.
Prints ONE JSON line (rank 0).  Timing: CUDA events on the library's stream, barrier + synchronize on
both sides, max over ranks.  Each token streams ~15 GB of weights, far beyond the 126 MB L2, so no
explicit L2 flush is needed between steps (stated in config.l2).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# Before numpy/torch load an OpenMP runtime: pin OpenMP threads.  The CPU legs (oracle port) otherwise lose
# >10x to thread migration between torch's and the system's libgomp (1.2 GB/s unbound vs 68 GB/s bound).
NCORES = len(os.sched_getaffinity(0))  # read before an OpenMP runtime pins the main thread to one core
os.environ.setdefault("OMP_PROC_BIND", "true")
os.environ.setdefault("OMP_WAIT_POLICY", "passive")
os.environ.setdefault("OMP_NUM_THREADS", str(NCORES))

METRIC = "decode tok/s (Llama-3-8B bf16, bs=1, 2k ctx)"
E2E_WARM = 3
CTX_LEN = 2048


def bytes_per_token(cfg, L: float, es: int = 2) -> float:
    """weights + lm_head + norms + embed row + KV read (length L) + KV write."""
    H, I, nh, nkv, hd, nl, V = (cfg.hidden_size, cfg.intermediate_size, cfg.num_attention_heads,
                                cfg.num_key_value_heads, cfg.hd, cfg.num_hidden_layers, cfg.vocab_size)
    per_layer = H * (nh + 2 * nkv) * hd + nh * hd * H + 3 * H * I + 2 * H
    w = nl * es * per_layer + es * V * H + es * H + es * H
    kv = nl * 2 * nkv * hd * es * L + nl * 2 * nkv * hd * es
    return float(w + kv)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples SM clock + throttle reasons during the timed region (nvidia-smi, 200 ms)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device, self.rows, self.proc = device, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.device), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
def host_info() -> dict:
    model = ""
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    model = line.split(":", 1)[1].strip()
                    break
    except Exception:
        pass
    return {"cpu": model, "logical_cores": NCORES}


_BEST_THREADS = None


def pick_threads() -> int:
    """Thread count for the CPU legs: the fastest of {all, 1/2, 1/4} logical cores on a short GEMV probe (on a
    2-socket box with SMT, all logical cores can be several times slower than one thread per physical core)."""
    global _BEST_THREADS
    if _BEST_THREADS is not None:
        return _BEST_THREADS
    import numpy as np
    import torch
    from oracle import oracle as O
    W = torch.randn(16384, 4096).to(torch.bfloat16)  # 128 MB
    x = O.round_to(np.random.default_rng(0).standard_normal((1, 4096)).astype(np.float32), "bf16")
    best, best_t = NCORES, float("inf")
    for n in sorted({NCORES, max(1, NCORES // 2), max(1, NCORES // 4)}, reverse=True):
        O.lib().ora_set_num_threads(n)
        O.linear(x, W, None, "bf16")
        t0 = time.perf_counter()
        for _ in range(4):
            O.linear(x, W, None, "bf16")
        dt = time.perf_counter() - t0
        if dt < best_t:
            best, best_t = n, dt
    O.lib().ora_set_num_threads(best)
    _BEST_THREADS = best
    return best


def cpu_reference_tok_s(cfg, n_tokens: int, budget_s: float, layers_cap=None) -> dict:
    """The reference's CPU path (oracle port, all host threads) on the same workload: bs=1 decode at KV
    length CTX_LEN.  Bounded sample: `n_tokens` tokens over `n_sample` of the 32 layers (+ the full
    ln_f/lm_head tail), scaled to the whole model — per-layer work is identical across layers."""
    import numpy as np
    import torch
    from cake_b200.synth import make_head, make_layer
    from oracle import oracle as O

    pick_threads()  # the best-performing thread count on this host (of all / half / quarter of the logical cores)
    t_build = time.perf_counter()
    nl = cfg.num_hidden_layers
    n_sample = min(nl, layers_cap or nl)
    # one random layer + distinct-memory clones: CPU timing is data independent; avoids minutes of host RNG
    base = make_layer(cfg, 0, "bf16", seed=1234)
    sd = make_head(cfg, "bf16", seed=1234)
    for i in range(n_sample):
        for k, v in base.items():
            sd[k.replace(".layers.0.", f".layers.{i}.")] = v if i == 0 else v.clone()
    om = O.OracleModel(cfg, sd, "bf16", max_seq=CTX_LEN + 64, layers=range(n_sample))
    cache = om.new_cache(CTX_LEN + 64)
    rng = np.random.default_rng(0)
    for l in range(n_sample):
        k, v = cache.kv(l)
        k[:, :CTX_LEN] = O.round_to(rng.standard_normal((k.shape[0], CTX_LEN, k.shape[2]), dtype=np.float32), "bf16")
        v[:, :CTX_LEN] = k[:, :CTX_LEN]
        cache.set_len(l, CTX_LEN)
    x = om.embed([17])
    build_s = time.perf_counter() - t_build

    def one_token(pos):
        t0 = time.perf_counter()
        h = om.forward_layers(x, 0, n_sample, pos, cache)
        t1 = time.perf_counter()
        lg = om.logits(h)
        O.argmax(lg)
        t2 = time.perf_counter()
        return t1 - t0, t2 - t1

    one_token(CTX_LEN)  # warm-up (page-in)
    tl, th = [], []
    spent = 0.0
    for i in range(n_tokens):
        a, b = one_token(CTX_LEN + 1 + i)
        tl.append(a)
        th.append(b)
        spent += a + b
        if spent > budget_s and i >= 1:
            break
    t_layers = statistics.mean(tl) * nl / n_sample
    t_tok = t_layers + statistics.mean(th)
    return {"value": 1.0 / t_tok, "unit": "tok/s", "cores": O.lib().ora_num_threads(), "kind": "port",
            "sample": f"{len(tl)} decode tokens at KV length {CTX_LEN}, {n_sample}/{nl} layers timed and scaled + full lm_head "
                      f"(oracle/cake_oracle.c, OpenMP; setup {build_s:.0f}s untimed)",
            "ms_per_token": t_tok * 1e3, "host": host_info()}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation cannot be built (Rust, no toolchain);
    the oracle port stands in, timed on this box's host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from cake_b200.config import llama3_8b
    cfg = llama3_8b(max_seq=4096)
    steps, warm = args.steps, args.warmup
    # keep the whole run within a few minutes: first probe 4 layers, then size the sample
    probe = cpu_reference_tok_s(cfg, 1, 5.0, layers_cap=4)
    est_full = probe["ms_per_token"] / 1e3
    budget = 150.0
    cap = cfg.num_hidden_layers if est_full * (steps + warm) <= budget else max(2, int(cfg.num_hidden_layers * budget / (est_full * (steps + warm))))
    res = cpu_reference_tok_s(cfg, steps, budget, layers_cap=cap)
    line = {
        "impl": "reference", "metric": METRIC, "value": res["value"], "unit": "tok/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warm, "ms_per_step": res["ms_per_token"], "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "Llama-3-8B bf16, 1 process, bs=1 decode, 2k context, CPU (reference path: oracle port)",
                   "kv_len": CTX_LEN, "reference_binary": "unbuildable here (Rust; no cargo/rustc) -> oracle port"},
        "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": res["value"], "unit": "tok/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "host": res["host"],
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
def run_cuda(args):
    import torch
    import torch.distributed as dist
    from cake_b200 import capi
    from cake_b200.capi import byref, c_uint32, check, int_array, lib, ptr_array
    from cake_b200.config import llama3_8b
    from cake_b200.model import B200Transformer, Context, TextModelBase
    from cake_b200.synth import make_head, make_layer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if args.gpus > 1 and world == 1:
            raise SystemExit("launch with: python -m torch.distributed.run --nproc-per-node N bench.py --gpus N ...")
    torch.cuda.set_device(local)
    if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
        os.environ["NCCL_DEBUG"] = "WARN"  # keep NCCL's version banner off stdout: rank 0 prints exactly one JSON line
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    cfg = llama3_8b(max_seq=4096)
    K, W = args.steps, max(args.warmup, 3)
    cache_cap = CTX_LEN + K + W + args.e2e_steps + E2E_WARM + 64

    # ---- model: synthetic random-init weights generated on the GPU, HF layout, loaded through the C ABI
    class LazySD(dict):
        """var_builder that materialises one layer at a time on the device (15 GB total)."""
        def __init__(self): super().__init__()
        def get(self, k, d=None):
            if k not in self and ".layers." in k:
                i = int(k.split(".layers.")[1].split(".")[0])
                self.clear_layers()
                self.update(make_layer(cfg, i, "bf16", seed=1234, device=f"cuda:{local}"))
            return dict.get(self, k, d)
        def clear_layers(self):
            for kk in [kk for kk in self if ".layers." in kk]:
                del self[kk]

    sd = LazySD()
    ctx = Context(cfg, sd, "bf16", device=local, max_seq=4096)
    ctx.cache = None
    from cake_b200.model import Cache
    ctx.cache = Cache(ctx, 1, cache_cap)
    if world == 1:
        sd.update(make_head(cfg, "bf16", seed=1234, device=f"cuda:{local}"))
        model = TextModelBase.load(ctx)
        sd.clear_layers()
        for k in list(sd):
            del sd[k]
        torch.cuda.empty_cache()
        blocks, idx = model.blocks, list(range(cfg.num_hidden_layers))
        ctx.cache.fill_synthetic(idx, CTX_LEN, 7)   # KV cache at 2k context (prefill is outside the metric)
        ctx.sync()
        model.index_pos = CTX_LEN
        model.decode_build()
        runner = model
        def decode(first, n): return model.decode_greedy(first, n)
    else:
        from cake_b200.parallel import ShardedMaster, Worker, init_comm
        init_comm(ctx, rank, world)
        if rank == 0:
            sd.update(make_head(cfg, "bf16", seed=1234, device=f"cuda:{local}"))
            master = ShardedMaster(ctx, world)
            sd.clear_layers()
            torch.cuda.empty_cache()
            model = master.model
        else:
            worker = Worker(ctx, rank, world)
            sd.clear_layers()
            torch.cuda.empty_cache()

    def sync_all():
        ctx.sync()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- ranks > 0: serve until shutdown -------------------------------------------------------
    if world > 1 and rank > 0:
        worker.serve_bench = True
        # the bench protocol mirrors rank 0's sequence below
        def timed_decode(n):
            blks, bidx = worker.block_list()
            if not getattr(worker, "_graph", False):
                check(lib().cake_b200_decode_build(ctx.h, ptr_array([b.h for b in blks]), int_array(bidx), len(blks), ctx.cache.h, rank, world))
                worker._graph = True
            pos = ctx.cache.len(bidx[0]) if bidx else 0
            check(lib().cake_b200_decode_begin(ctx.h, 0, pos))
            sync_all()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(ctx.torch_stream)
            check(lib().cake_b200_decode_run(ctx.h, n))
            e1.record(ctx.torch_stream)
            # sleep-poll instead of a spinning cudaStreamSynchronize: with N-1 worker processes spinning next to
            # rank 0's per-token sync the host-stepped leg showed 100 ms stalls at N=8 (CPU contention / quota)
            while not e1.query():
                time.sleep(0.0005)
            sync_all()
            return e0.elapsed_time(e1)
        _, bidx = worker.block_list()
        ctx.cache.fill_synthetic(bidx, CTX_LEN, 7)
        sync_all()
        timed_decode(W)
        ms = timed_decode(K)
        t = torch.tensor([ms], device=f"cuda:{local}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        # e2e leg: rank 0 drives one step at a time; workers pre-enqueue the same number of replays
        timed_decode(args.e2e_steps + E2E_WARM)
        dist.barrier()
        dist.destroy_process_group()
        return

    # ---- rank 0 -----------------------------------------------------------------------------------
    if world > 1:
        master.ctx.cache.fill_synthetic(master.local_idx, CTX_LEN, 7)
        model.index_pos = CTX_LEN
        master.decode_build()
        sync_all()

    def timed_decode(first, n):
        check(lib().cake_b200_decode_begin(ctx.h, first, model.index_pos))
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = ctx.launch_count()
        e0.record(ctx.torch_stream)
        check(lib().cake_b200_decode_run(ctx.h, n))
        e1.record(ctx.torch_stream)
        sync_all()
        model.index_pos += n
        out = (c_uint32 * n)()
        check(lib().cake_b200_decode_tokens(ctx.h, out, n))
        return e0.elapsed_time(e1), ctx.launch_count() - l0, int(out[n - 1])

    _, _, tok = timed_decode(17, W)                       # warm-up steps (untimed)
    sampler = ClockSampler(local)
    sampler.start()
    ms, launches, tok = timed_decode(tok, K)              # EXACTLY K timed steps
    clocks = sampler.stop()
    if world > 1:
        t = torch.tensor([ms], device=f"cuda:{local}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    tok_s = K / (ms / 1e3)

    # ---- e2e: one token per call through the public API with HOST buffers (token id H2D, token D2H, sync)
    n_e2e = args.e2e_steps
    nxt = c_uint32()
    check(lib().cake_b200_decode_begin(ctx.h, tok, model.index_pos))
    sync_all()
    cur = tok
    lat = []
    for _ in range(E2E_WARM):  # untimed warm-up of the host-stepped path (first replays after a (re)build are slow)
        check(lib().cake_b200_decode_step_host(ctx.h, cur, byref(nxt)))
        cur = nxt.value
    t0 = time.perf_counter()
    for _ in range(n_e2e):  # workers (N>1) have pre-enqueued the same number of replays
        t1 = time.perf_counter()
        check(lib().cake_b200_decode_step_host(ctx.h, cur, byref(nxt)))
        cur = nxt.value
        lat.append(time.perf_counter() - t1)
    e2e_s = time.perf_counter() - t0
    if world > 1:
        sync_all()
    model.index_pos += n_e2e + E2E_WARM
    e2e_tok_s = n_e2e / e2e_s
    lat_ms = sorted(x * 1e3 for x in lat)

    # ---- roofline.  The dominant kernel of the timed region is decode_mega_kernel (one launch per token per shard,
    # ~100% of the step): its achieved bandwidth = algorithmic bytes of that launch / its average duration over the K
    # timed steps (CUDA events on the library's stream, above).  The same GEMV / attention phases are additionally timed
    # live as stand-alone kernels (cake_b200_bench_kernel) -- an aid that shows what each phase reaches without the
    # phase boundaries, reported separately and never as the headline fraction.
    peak, peak_src = peaks()
    L_mean = CTX_LEN + W + K / 2.0
    bpt = bytes_per_token(cfg, L_mean)
    H, I = cfg.hidden_size, cfg.intermediate_size
    n_local = cfg.num_hidden_layers if world == 1 else len(master.local_idx)
    per_layer_bytes = 2 * (H * (cfg.num_attention_heads + 2 * cfg.num_key_value_heads) * cfg.hd + cfg.num_attention_heads * cfg.hd * H
                           + 3 * H * I + 2 * H) + 2 * 2 * cfg.num_key_value_heads * cfg.hd * (L_mean + 1)
    head_bytes = 2 * cfg.vocab_size * H + 2 * H + 2 * H
    launch_bytes = float(n_local * per_layer_bytes + head_bytes)   # rank 0's launch: its layers + ln_f/lm_head/embed row
    launch_ms = ms / K
    achieved = launch_bytes / (launch_ms * 1e-3) / 1e9
    roof = {"bound": "hbm", "kernel": "decode_mega_kernel<bf16,128,4> (all local layers + ln_f + lm_head + argmax in one launch)",
            "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
            # dram__bytes_read.sum + dram__bytes_write.sum of one launch, ncu --set full at N=1 (profiles/mega_r01_raw.csv)
            "traffic": (15280821000 + 7653120) if world == 1 else None,
            "peak_source": peak_src, "algorithmic_bytes_per_launch": launch_bytes, "launch_ms": launch_ms,
            "launches_per_step": 1, "share_of_step": 1.0}
    if world > 1:
        roof["note"] = (f"rank 0's launch holds {n_local}/{cfg.num_hidden_layers} layers + head and spans the whole step: it waits "
                        "inside the kernel for the ring to come back, so this fraction is bytes of one shard over the time of all shards")
    try:
        import ctypes
        local_blocks = model.blocks if world == 1 else master.local
        local_idx = list(range(len(local_blocks))) if world == 1 else master.local_idx
        hs, ix = ptr_array([b.h for b in local_blocks]), int_array(local_idx)
        per_kernel = {}
        kb = {0: 2 * (cfg.size_q + 2 * cfg.size_kv) * H, 1: 2 * H * cfg.size_q, 2: 2 * 2 * I * H, 3: 2 * H * I}
        names = {0: "qkv_gemv", 1: "o_gemv", 2: "gate_up_gemv", 3: "down_gemv", 4: "attn_decode"}
        for which in (2, 3, 0, 1, 4):
            msl = ctypes.c_float()
            check(lib().cake_b200_bench_kernel(ctx.h, hs, ix, len(local_blocks), ctx.cache.h, which, 20, byref(msl)))
            nbytes = kb.get(which, 2 * 2 * cfg.num_key_value_heads * cfg.hd * ctx.cache.len(local_idx[0]))
            per_kernel[names[which]] = {"ms": round(msl.value, 5), "GB/s": round(nbytes / (msl.value * 1e-3) / 1e9, 1),
                                        "frac": round(nbytes / (msl.value * 1e-3) / 1e9 / peak, 4)}
        roof["isolated_per_op_kernels"] = {"what": "the same phases as stand-alone kernels (gemv_kernel / attn_decode_kernel), 20 launches "
                                                   "each over distinct layers, timed live; not part of the timed decode step",
                                           "gate_up_traffic_ncu": 234938368 + 3269120, **per_kernel}
    except Exception as e:  # keep the headline line even if the aid fails
        roof["isolated_per_op_kernels"] = {"error": str(e)}

    line = {
        "metric": METRIC, "value": tok_s, "unit": "tok/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": f"Llama-3-8B bf16, bs=1 decode, 2k context, layers sharded over {world} B200 (contiguous ranges, one hidden-state hand-off per boundary)" if world > 1
                   else "Llama-3-8B bf16, 1xB200, bs=1 decode, 2k context (BASELINE.json configs[1])",
                   "kv_len_start": CTX_LEN + W, "kv_cache": "synthetic fill to 2048 positions (prefill is outside the metric, master.rs:131-134)",
                   "weights": "random-init N(0,0.02) bf16, HF layout, seed 1234", "greedy": True,
                   "parallelism": f"pp{world}" if world > 1 else "single",
                   "handoff": (os.environ.get("CAKE_B200_RING", "p2p") + (" (fused into the decode kernel over NVLink peer memory)" if os.environ.get("CAKE_B200_RING", "p2p") == "p2p" else " (ncclSend/ncclRecv graph nodes)")) if world > 1 else None,
                   "l2": "inputs (15 GB of weights per step) larger than L2; no flush needed"},
        "clocks": clocks,
        "e2e": {"value": e2e_tok_s, "unit": "tok/s", "h2d_bytes_per_step": 4, "d2h_bytes_per_step": 4, "steps": n_e2e, "warmup": E2E_WARM,
                "ms_per_step_p50": lat_ms[len(lat_ms) // 2], "ms_per_step_max": lat_ms[-1],
                "api": "cake_b200_decode_step_host (token id from host, sampled token back to host, sync per token)"},
        "gpu_launches": int(launches),
        "roofline": roof,
        "token_roofline": {"bytes_per_token": bpt, "achieved_GB_s": bpt / (ms / K * 1e-3) / 1e9,
                           "frac_of_one_gpu_hbm": bpt / (ms / K * 1e-3) / 1e9 / peak,
                           "frac_of_aggregate_hbm": bpt / (ms / K * 1e-3) / 1e9 / (peak * world),
                           "roofline_tok_s_one_gpu": peak * 1e9 / bpt,
                           "traffic_per_token_ncu": 15280821000 + 7653120,  # decode_mega_kernel, profiles/mega_r01_raw.csv (N=1)
                           "kernel": "decode_mega_kernel (one launch per token per shard)"},
    }
    if world == 1 and not args.no_cpu:
        try:
            line["cpu_baseline"] = {k: v for k, v in cpu_reference_tok_s(cfg, 3, 25.0, layers_cap=8).items()
                                    if k in ("value", "unit", "cores", "kind", "sample")}
        except Exception as e:
            line["cpu_baseline"] = {"error": str(e)}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    ctx.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=128)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=64)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_cuda(args)


if __name__ == "__main__":
    main()
