#!/usr/bin/env python
"""bench.py — hamming_top_n queries/s on B200 (BASELINE.json metric), one JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload c2|c4]
                    [--shard auto|rows|queries]

Workload (default, BASELINE config C2): GloVe-6B-shaped synthetic 400k x 300 fp32 vectors,
640-bit signatures (10 x uint64), 10 000 query vectors, top-10.  One STEP = one pass of the hot
path over the whole query batch: hash the 10k query vectors (K1T, tcgen05 TF32-split GEMM),
batched Hamming scan + top-10 over the signature table (K2T: tcgen05 kind::mxf4 phases + select
kernels; a gated XOR+POPC repair pass that is a no-op unless a candidate buffer overflowed)
[+ one NCCL all-gather when N > 1].  The signature table is built once by K1T before timing and
stays resident in HBM.  With N GPUs (strong scaling: total work fixed) the 32 MB table is
replicated and the query batch split N ways (`--shard queries`, the automatic choice for tables
<= 4 GiB); `--shard rows` splits the table by rows instead (all-gather of candidate keys + K4
merge), which is what the 12.8 GB C4 workload uses.

`value`  : queries/s with the query vectors already in HBM (device-timed, CUDA events, L2
           evicted between iterations).
`e2e`    : the same step through the public API with HOST (pinned) query vectors in and HOST
           ids out, copies inside the timed region, one blocking call per step;
           `e2e.pipelined`: the same batches streamed through np_sims.pipeline.QueryStream
           (two batches in flight, copies overlapped with the search).
`roofline`: the dominant launch (mma_scan_kernel, last phase) timed with CUDA events recorded by
           the library on the launching stream (nps_profile_*), against 4 x the measured bf16
           peak of MEASURED_PEAKS.json (kind::mxf4 : bf16 = 4 : 1); `roofline_stream`: the
           HBM-bound single-query scan (scan_topk_kernel) against the measured HBM peak.
`cpu_baseline` / --impl reference: the reference's own ufunc (oracle/_ref, compiled from the
           unmodified np_sims/ufunc.c) driven exactly like benchmark/glove.py:183-199 — a
           ThreadPoolExecutor over all host cores, one query per task.

bench.py is one of the three places allowed to touch oracle/ — only for the CPU legs and for the
parity spot-check of the timed step.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "np-sims_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

WORKLOADS = {
    # name: rows, dims, bits, queries, k
    "c2": dict(rows=400_000, dims=300, bits=640, queries=10_000, k=10,
               desc="C2: GloVe-6B-shaped synthetic 400k x 300 fp32, 640-bit signatures, 10k queries, top-10"),
    "c4": dict(rows=100_000_000, dims=0, bits=1024, queries=100_000, k=100,
               desc="C4: 100M x 1024-bit synthetic signature table, 100k query batch, top-100"),
}
L2_FLUSH_BYTES = 256 << 20


def ncu_traffic(kernel):
    """DRAM bytes per launch of the dominant kernel, from the committed ncu capture (profiles/)."""
    path = os.path.join(ROOT, "profiles", "r01_traffic.json")
    try:
        return json.load(open(path)).get(kernel)
    except (OSError, ValueError):
        return None


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)", float(d.get("sm_max_mhz", 1965.0))
    return 6650.0, "fallback (B200_PROFILING.md)", 1965.0


# ----------------------------------------------------------------------------- synthetic data
def make_vectors(n, dims, seed):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, dims), dtype=np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    return x


def make_projections(bits, dims):
    import np_sims.lsh as lsh
    np.random.seed(0)
    return lsh.create_projections(bits, dims)


# ----------------------------------------------------------------------------- clocks sampler
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.lines, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-i", str(self.index),
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smmax, reasons = [], [], set()
        for ts, line in self.lines:
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                clk, mx = float(parts[0]), float(parts[1])
            except ValueError:
                continue
            smmax.append(mx)
            if t0 <= ts <= t1 + 0.2:
                sm.append(clk)
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"),
                                     parts[3:7]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
        if not sm:
            sm = [float(l.split(",")[0]) for _, l in self.lines[-3:] if l.split(",")[0].strip().replace(".", "").isdigit()]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smmax) if smmax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------- CPU (reference) leg
def pick_reference_variant():
    from oracle import oracle
    for variant in ("native", "popcnt", ""):
        if oracle.ref_variant_runs(variant):
            return variant, oracle.ref_ufuncs(variant)
    return None, None


def cpu_reference_run(wl, sample_queries, steps, warmup, H=None, W=None, qvecs=None, time_budget_s=None):
    """The reference path on the host cores: per query  np.dot(W, x) -> packbits -> hamming_top_10
    (np_sims/lsh.py:70-74), one query per ThreadPoolExecutor task, GIL released inside the ufunc
    (benchmark/glove.py:183-199).  Returns (queries/s, ms_per_step, info)."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import oracle
    variant, uf = pick_reference_variant()
    kind = "reference"
    if uf is None:                       # oracle/_ref was not built: fall back to the C port
        kind, variant = "port", "oracle.c"
    cores = os.cpu_count() or 1
    k = wl["k"]

    def one(i):
        qh = np.packbits(np.dot(W, qvecs[i]) >= 0, bitorder="little").view(np.uint64)
        if uf is not None and k == 10:
            return uf.hamming_top_10(H, qh)
        if uf is not None and k == 1000:
            return uf.hamming_top_cand(H, qh)
        if uf is not None:               # no reference kernel for this k: distances + argpartition
            d = uf.hamming(H, qh)
            return np.argpartition(d, k)[:k]
        return oracle.top_k_queue(H, qh, k)[0]

    times = []
    with ThreadPoolExecutor(max_workers=cores) as ex:
        t_begin = time.perf_counter()
        for s in range(warmup + steps):
            idx = [(s * sample_queries + j) % len(qvecs) for j in range(sample_queries)]
            t0 = time.perf_counter()
            list(ex.map(one, idx))
            dt = time.perf_counter() - t0
            if s >= warmup:
                times.append(dt)
            if time_budget_s and time.perf_counter() - t_begin > time_budget_s and len(times) >= 1:
                break
    total = float(np.sum(times))
    qps = sample_queries * len(times) / total
    info = {"value": qps, "unit": "queries/s", "cores": cores, "kind": kind,
            "sample": f"{sample_queries} queries/step x {len(times)} steps over the full "
                      f"{H.shape[0]}x{H.shape[1] * 64}-bit table, ThreadPoolExecutor({cores}), build={variant or '-O3'}"}
    return qps, 1e3 * total / len(times), info


def cpu_signatures(wl, X, W):
    from oracle import oracle
    out = np.empty((len(X), wl["bits"] // 64), dtype=np.uint64)
    for i in range(0, len(X), 50_000):
        out[i:i + 50_000] = oracle.index_np(X[i:i + 50_000], W)
    return out


def run_reference(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    X = make_vectors(wl["rows"], wl["dims"], 0)
    qvecs = make_vectors(wl["queries"], wl["dims"], 1)
    np.random.seed(0)
    from oracle import oracle
    W = oracle.create_projections(wl["bits"], wl["dims"])
    H = cpu_signatures(wl, X, W)
    sample = args.ref_sample
    qps, ms, info = cpu_reference_run(wl, sample, args.steps, args.warmup, H=H, W=W, qvecs=qvecs)
    line = {
        "impl": "reference", "metric": "hamming_top_n queries/s", "value": qps, "unit": "queries/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": wl["desc"], "rows": wl["rows"], "bits": wl["bits"], "queries": wl["queries"],
                   "k": wl["k"], "step": f"{sample}-query sample of the workload per step (CPU arm)"},
        "cpu_baseline": info,
        "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------- B200 arm
def run_b200(args, wl):
    import torch
    import torch.distributed as dist
    import np_sims
    import np_sims.lsh as lsh
    from np_sims import _lib
    from np_sims.sharded import ReplicatedIndex, ShardedIndex, choose_sharding, query_bounds, shard_bounds

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    hbm_peak, peak_src, sm_max_mhz = peaks()
    k, nq, words = wl["k"], wl["queries"], wl["bits"] // 64
    shard = args.shard if args.shard != "auto" else choose_sharding(wl["rows"], words, nq, world)
    if world == 1:
        shard = "rows"
    by_query = shard == "queries"
    qb = query_bounds(nq, world) if by_query else [0] * rank + [0, nq]
    qlo, qhi = qb[rank], qb[rank + 1]        # the queries THIS rank hashes and answers

    # ---- inputs (synthetic), resident in HBM before timing
    b = [0] * (rank + 1) + [wl["rows"]] if by_query else shard_bounds(wl["rows"], world)
    if wl["dims"]:
        W = make_projections(wl["bits"], wl["dims"])
        qvecs_all = make_vectors(nq, wl["dims"], 1)
        qvecs_host = qvecs_all[qlo:qhi]
        X = make_vectors(wl["rows"], wl["dims"], 0)
        t0 = time.perf_counter()
        H_local = lsh.hash_vectors(torch.from_numpy(X[b[rank]:b[rank + 1]]).to(dev), W)   # K1: index build
        torch.cuda.synchronize()
        index_s = time.perf_counter() - t0
        qvecs_dev = torch.from_numpy(qvecs_host).to(dev)
        qvecs_pinned = torch.from_numpy(qvecs_host).pin_memory()
    else:   # signature-only workload (C4): i.i.d. bits drawn on the device, one generator stream per shard
        W = None
        g = torch.Generator(device=dev)
        g.manual_seed(3 + (0 if by_query else rank))
        H_local = torch.randint(-2**63, 2**63 - 1, (b[rank + 1] - b[rank], words), dtype=torch.int64, device=dev,
                                generator=g)
        g.manual_seed(4)
        qsig_dev = torch.randint(-2**63, 2**63 - 1, (nq, words), dtype=torch.int64, device=dev,
                                 generator=g)[qlo:qhi].contiguous()
        qsig_pinned = qsig_dev.cpu().pin_memory()
        index_s = None
    index = ReplicatedIndex(H_local) if by_query else ShardedIndex(H_local, b[rank])
    out_pinned = torch.empty((nq, k), dtype=torch.int64).pin_memory()
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=dev)

    def search(sigs):
        """-> (rows, dists) [nq, k] for the WHOLE batch, on every rank."""
        if by_query:      # ids only, like the reference's hamming_top_10: one all-gather
            return index.gather(index.query_local(sigs, k, return_distances=False), nq), None
        return index.query(sigs, k)

    def step_device():
        sigs = lsh.hash_vectors(qvecs_dev, W) if W is not None else qsig_dev
        return search(sigs)

    def step_e2e():
        if W is not None:
            sigs = lsh.hash_vectors(qvecs_pinned.to(dev, non_blocking=True), W)
        else:
            sigs = qsig_pinned.to(dev, non_blocking=True)
        rows, _ = search(sigs)
        out_pinned.copy_(rows.view(torch.int64))            # D2H of the step's result (blocking)
        return out_pinned

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        starts = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        stops = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        launches0 = _lib.raw("nps_launch_count")()
        w0 = time.perf_counter()
        for i in range(steps):
            flush.fill_(i & 0xFF)         # evict L2 between timed iterations (untimed)
            starts[i].record()
            fn()
            stops[i].record()
        torch.cuda.synchronize()
        w1 = time.perf_counter()
        launches = _lib.raw("nps_launch_count")() - launches0
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        total_ms = sum(s.elapsed_time(e) for s, e in zip(starts, stops))
        if world > 1:
            t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            total_ms = float(t.item())
        return total_ms, launches, (w0, w1)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    total_ms, launches, (w0, w1) = timed(step_device, args.steps, args.warmup)
    clocks = sampler.stop(w0, w1) if rank == 0 else None
    value = nq * args.steps / (total_ms * 1e-3)
    e2e_ms, _, _ = timed(step_e2e, args.steps, max(1, args.warmup // 2))
    e2e_value = nq * args.steps / (e2e_ms * 1e-3)

    # ---- the same end-to-end step as a STREAM of batches (np_sims.pipeline.QueryStream, 2 in flight): the
    # H2D copy of batch i+1 and the D2H of batch i-1 overlap the search of batch i.  Every batch's copies
    # are inside the timed region; L2 is evicted before every batch on the search stream (timed).
    pipelined = None
    if world == 1 and W is not None:
        from np_sims.pipeline import QueryStream
        qs = QueryStream(H_local, W, n=k, depth=2, flush_l2_bytes=L2_FLUSH_BYTES)
        for _ in qs.map(qvecs_pinned for _ in range(max(3, args.warmup))):
            pass
        torch.cuda.synchronize()
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(qs.s_in):
            p0.record()
        last = None
        for ids in qs.map(qvecs_pinned for _ in range(args.steps)):
            last = ids
        with torch.cuda.stream(qs.s_out):
            p1.record()
        torch.cuda.synchronize()
        pipe_ms = p0.elapsed_time(p1)
        pipelined = {"value": nq * args.steps / (pipe_ms * 1e-3), "unit": "queries/s", "ms_per_step": pipe_ms / args.steps,
                     "in_flight": 2, "api": "np_sims.pipeline.QueryStream", "last_ids": last.copy()}

    # ---- parity spot-check of what was just timed (rank 0, a few queries, against the oracle)
    rows, dists = step_device()
    torch.cuda.synchronize()

    # ---- roofline pass: the dominant kernel timed per launch on its stream
    _lib.raw("nps_profile_enable")(1)
    scan_ms, merge_ms, phase_runs = [], [], []
    for i in range(max(3, min(args.steps, 10))):
        flush.fill_(i)
        step_device()
        s_ms, m_ms = _lib.profile_last()
        scan_ms.append(s_ms)
        merge_ms.append(m_ms)
        if _lib.raw("nps_last_scan_path")() == _lib.SCAN_MMA:
            phase_runs.append(_lib.profile_last_phases())
    _lib.raw("nps_profile_enable")(0)
    scan_avg = float(np.median(scan_ms))
    n_local = b[rank + 1] - b[rank]
    nq_local = qhi - qlo
    sm_mhz = (clocks or {}).get("sm_mhz") or sm_max_mhz
    step_ms = total_ms / args.steps
    if phase_runs:
        # Tensor-core scan (mma_scan_kernel): bound by the tensor pipe.  Algorithmic work of a launch
        # = 2 * rows * queries * bits flop (one +-1 multiply-add per signature bit and (row, query)
        # pair; the folded-threshold slice is overhead, not counted).  Peak: kind::mxf4 runs at 4x
        # the bf16 rate (9 vs 2.25 PFLOP/s nominal); MEASURED_PEAKS.json holds the bf16 figure.
        plan = _lib.last_mma_plan()
        bf16_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops"] \
            if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 1590.0
        tensor_peak = 4.0 * bf16_peak
        nph = len(phase_runs[0])
        tiles = [phase_runs[0][j][0] for j in range(nph)]
        ph_ms = [float(np.median([r[j][1] for r in phase_runs])) for j in range(nph)]
        dom = int(np.argmax(ph_ms))
        rows_dom = min(tiles[dom] * 128, n_local)
        flop_dom = 2.0 * rows_dom * nq_local * wl["bits"]
        ach = flop_dom / (ph_ms[dom] * 1e-3) / 1e12
        flop_all = 2.0 * n_local * nq_local * wl["bits"]
        roofline = {
            "kernel": f"mma_scan_kernel<dense=false> (tcgen05 kind::mxf4, M128 x N{plan['n_tile']} x K64), phase {dom} of {nph}",
            "bound": "tensor", "achieved": ach, "peak": tensor_peak, "unit": "TFLOP/s", "frac": ach / tensor_peak,
            "traffic": ncu_traffic("mma_scan_kernel") if (world == 1 and wl is WORKLOADS["c2"]) else None,
            "peak_source": "4 x measured bf16 burst (MEASURED_PEAKS.json): mxf4 : bf16 = 9 : 2.25 PFLOP/s nominal; "
                           "scripts/ubench/umma_mxf4_probe measured 16359 MAC/clk/SM issue rate in-kernel",
            "frac_of_measured_mxf4_issue_rate": ach / (16359 * 2 * 148 * sm_mhz * 1e6 / 1e12),
            "measured_mxf4_issue_rate_note": "scripts/ubench/umma_mxf4_probe.cu: 16359 MAC/clk/SM back-to-back = "
                                             f"{16359 * 2 * 148 * sm_mhz * 1e6 / 1e12:.0f} TFLOP/s at {sm_mhz:.0f} MHz — a stricter "
                                             "denominator than 4 x the measured bf16 GEMM rate",
            "ms_per_launch": ph_ms[dom], "share_of_step": ph_ms[dom] / step_ms,
            "algorithmic": {"rows": rows_dom, "queries": nq_local, "bits": wl["bits"], "flop": flop_dom},
            "phases": [{"tiles": t, "scan_ms": m} for t, m in zip(tiles, ph_ms)],
            "all_phases": {"scan_ms": float(np.sum(ph_ms)), "scan_plus_select_ms": scan_avg,
                           "achieved_tflops": flop_all / (scan_avg * 1e-3) / 1e12,
                           "frac": flop_all / (scan_avg * 1e-3) / 1e12 / tensor_peak},
            "plan": plan,
            "popc_equivalent": {"gpopc32_s": 2.0 * n_local * nq_local * words / (scan_avg * 1e-3) / 1e9,
                                "popc_pipe_peak_gpopc32_s": 148 * 16 * sm_mhz * 1e6 / 1e9,
                                "note": "the XOR+POPC formulation of the same batch would need this POPC rate; "
                                        "the integer pipe peaks at 148 SM x 16 /clk"},
        }
    else:
        plan = _lib.last_scan_plan()
        # algorithmic bytes of one scan launch: signatures + queries + results
        alg_bytes = n_local * words * 8 + nq_local * words * 8 + nq_local * plan["num_chunks"] * k * 8
        achieved_gbs = alg_bytes / (scan_avg * 1e-3) / 1e9
        popc32 = 2.0 * n_local * nq_local * words
        popc_peak = 148 * 16 * sm_mhz * 1e6
        roofline = {
            "kernel": f"scan_topk_kernel<W={words},R={plan['rows_per_thread']}>",
            "bound": "hbm", "achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s",
            "frac": achieved_gbs / hbm_peak,
            "traffic": ncu_traffic("scan_topk_kernel") if (world == 1 and wl is WORKLOADS["c2"]) else None,
            "peak_source": peak_src, "ms_per_launch": scan_avg, "share_of_step": scan_avg / step_ms,
            "note": ("at this batch size the scan re-reads an L2-resident table once per query tile and is bound "
                     "by the integer POPC pipe, not HBM; see `popc` for that roofline and `roofline_stream` for "
                     "the HBM-bound single-query scan"),
            "popc": {"achieved_gpopc32_s": popc32 / (scan_avg * 1e-3) / 1e9,
                     "peak_gpopc32_s": popc_peak / 1e9, "frac": popc32 / (scan_avg * 1e-3) / popc_peak,
                     "peak_formula": f"148 SM x 16 POPC/clk x {sm_mhz:.0f} MHz (median clock under load)"},
            "plan": plan, "merge_ms_per_launch": float(np.mean(merge_ms)),
        }

    line = None
    if rank == 0:
        stream = stream_roofline(torch, np_sims, _lib, dev, hbm_peak, peak_src) if world == 1 and not args.no_stream else None
        cpu = None
        if world == 1 and not args.no_cpu_baseline and W is not None:
            from oracle import oracle
            Wn = np.asarray(W)
            H_host = H_local.cpu().numpy().view(np.uint64)
            # parity of the timed step against the oracle on a handful of queries
            sig_host = oracle.index_np(qvecs_all[:4], Wn)
            got = rows[:4].cpu().numpy().view(np.uint64)
            for i in range(4):
                want, _ = oracle.top_k_canonical(H_host, sig_host[i], k)
                assert np.array_equal(got[i], want), "timed step disagrees with the oracle"
                if pipelined is not None:
                    assert np.array_equal(pipelined["last_ids"][i], want), "streamed step disagrees with the oracle"
            _, _, cpu = cpu_reference_run(wl, args.cpu_sample, steps=64, warmup=1, H=H_host, W=Wn, qvecs=qvecs_all,
                                          time_budget_s=20.0)
        # summed over ranks: row sharding uploads the whole batch on every rank, query sharding one slice each
        # hashing parity of the timed step's K1T (north-star: "flipped-bit rate reported"): bare tensor-core
        # signs and final signatures of the query batch against the float64 kernel
        hash_report = None
        if W is not None:
            def bits_differing(a, b):
                x = (a ^ b).cpu().numpy().view(np.uint64)
                return int(np.unpackbits(x.view(np.uint8)).sum())
            s_exact = lsh.hash_vectors(qvecs_dev, W, exact_only=True)
            s_final = lsh.hash_vectors(qvecs_dev, W)
            st = lsh.hash_stats()
            s_raw = lsh.hash_vectors(qvecs_dev, W, tf32_raw=True)
            nbits = int(qvecs_dev.shape[0]) * wl["bits"]
            hash_report = {"path": st["path"], "bits": nbits,
                           "encoding": "bf16 x 2 (kind::f16), three products, separate correction accumulator",
                           "tolerance": f"(52 + dims/4) * 2^-22 * |x| * max|w| = {(52 + wl['dims'] / 4) * 2.0 ** -22:.3e} |x|",
                           "recomputed_in_float64": st.get("recomputed"), "recomputed_frac": (st.get("recomputed") or 0) / nbits,
                           "flipped_bits_before_fixup": bits_differing(s_raw, s_exact),
                           "flipped_bits_final": bits_differing(s_final, s_exact)}
            assert hash_report["flipped_bits_final"] == 0, "K1T signatures differ from the float64 kernel"
        h2d = (qvecs_all.nbytes if W is not None else nq * words * 8) * (1 if by_query else world)
        line = {
            "metric": "hamming_top_n queries/s", "value": value, "unit": "queries/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u64",
            "data": "synthetic",
            "config": {"workload": wl["desc"], "rows": wl["rows"], "bits": wl["bits"], "queries": nq, "k": k,
                       "parallelism": (f"table replicated, query batch sharded x{world} + NCCL all-gather of results"
                                       if by_query else
                                       f"row-sharded x{world}" + (" + NCCL all-gather of candidate keys" if world > 1 else "")),
                       "step": "hash query batch (K1) + batched scan/top-n (K2T tensor-core phases + select)" +
                               ((" + all-gather of results" if by_query else " + all-gather + K4 merge") if world > 1 else ""),
                       "l2": f"flushed between timed iterations ({L2_FLUSH_BYTES >> 20} MiB write)",
                       "index_build_s": index_s},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "queries/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(nq * k * 8) * world, "ms_per_step": e2e_ms / args.steps,
                    "mode": "one blocking call per step (copy in, search, copy out)",
                    "pipelined": ({kk: vv for kk, vv in pipelined.items() if kk != "last_ids"} if pipelined else None)},
            "gpu_launches": int(launches),
            "roofline": roofline,
            "roofline_stream": stream,
            "hash": hash_report,
            "cpu_baseline": cpu,
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return line


def stream_roofline(torch, np_sims, _lib, dev, hbm_peak, peak_src, rows=25_000_000, words=16, k=10, reps=10):
    """HBM-bound regime of the same kernel: ONE query against a 3.2 GB table (>> 126 MB L2), the
    configuration the north-star's '>= 70 % of HBM roofline' target is stated on."""
    H = torch.randint(-2**63, 2**63 - 1, (rows, words), dtype=torch.int64, device=dev)
    q = torch.randint(-2**63, 2**63 - 1, (1, words), dtype=torch.int64, device=dev)
    for _ in range(3):
        np_sims.hamming_top_n(H, q, k)
    _lib.raw("nps_profile_enable")(1)
    ms = []
    for _ in range(reps):
        np_sims.hamming_top_n(H, q, k)
        ms.append(_lib.profile_last()[0])
    _lib.raw("nps_profile_enable")(0)
    plan = _lib.last_scan_plan()
    avg = float(np.mean(ms))
    alg = rows * words * 8 + words * 8 + plan["num_chunks"] * k * 8
    gbs = alg / (avg * 1e-3) / 1e9
    del H
    return {"kernel": f"scan_topk_kernel<W={words},R={plan['rows_per_thread']}>", "workload":
            f"1 query x {rows} x {words * 64}-bit signatures ({rows * words * 8 / 1e9:.1f} GB, larger than L2), top-{k}",
            "bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak,
            "ms_per_launch": avg, "best_ms": float(np.min(ms)), "peak_source": peak_src, "plan": plan}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--shard", default="auto", choices=["auto", "rows", "queries"],
                    help="N > 1: split the table by rows, or replicate it and split the query batch (auto: by table size)")
    ap.add_argument("--ref-sample", type=int, default=512, help="queries per step of the reference arm")
    ap.add_argument("--cpu-sample", type=int, default=256, help="queries per batch of the cpu_baseline leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-stream", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, wl)
    else:
        run_b200(args, wl)


if __name__ == "__main__":
    main()
