#!/usr/bin/env python
"""bench.py — hamming_top_n queries/s on B200 (BASELINE.json metric), one JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
                    [--shard auto|rows|queries] [--skip c3,c4,single,cpu,stream,pipe]

Headline workload (BASELINE config C2): GloVe-6B-shaped synthetic 400k x 300 fp32 vectors,
640-bit signatures (10 x uint64), 10 000 query vectors, top-10.  One STEP = one pass of the hot
path over the whole query batch: hash the 10k query vectors (K1T, tcgen05 split-BF16 GEMM), batched
Hamming scan + top-10 over the signature table (K2T: tcgen05 kind::mxf4 phases + selects; a gated
XOR+POPC repair pass that is a no-op unless a candidate buffer overflowed) [+ one NCCL all-gather
when N > 1].  The signature table is built once by K1T before timing and stays resident in HBM.
With N GPUs (strong scaling: total work fixed) the 32 MB table is replicated and the query batch
split N ways (`--shard queries`, the automatic choice for tables <= 4 GiB).

`value`  : queries/s with the query vectors already in HBM (device-timed, CUDA events, L2 evicted
           between iterations).
`e2e`    : the same step through the public API with HOST (pinned) query vectors in and HOST ids
           out, copies inside the timed region, one blocking call per step; `e2e.pipelined`: the same
           batches streamed through np_sims.pipeline.QueryStream.
`roofline`: mma_scan_kernel over ALL its launches of a step (per-phase table kept), timed with CUDA
           events recorded by the library on the launching stream, against 4 x the measured bf16
           peak of MEASURED_PEAKS.json; `roofline_stream`: the HBM-bound single-query scan.
`cpu_baseline` / --impl reference: the reference's own ufunc (oracle/_ref, compiled from the
           unmodified np_sims/ufunc.c) driven exactly like benchmark/glove.py:183-199 — a
           ThreadPoolExecutor over all host cores, one query per task; two flag sets, all cores and
           one thread, plus the hashing baseline.
`single_query`: the reference's own calling pattern on the B200 — `lsh.query` one vector at a time
           from ThreadPoolExecutor(os.cpu_count()) host threads over the same table.
`c4`     : the north-star multi-GPU config at THIS run's N: 100M x 1024-bit signatures sharded by
           rows, 100k-query batch, top-100, NCCL all-gather of candidate keys + K4 merge, with the
           merged result checked against the oracle on planted neighbours and a row subsample.
`c3`     : projection only, 10M x 768 fp32 -> 1024-bit signatures (rows split over the N GPUs), a
           row subsample checked against the float64 oracle.

bench.py is one of the three places allowed to touch oracle/ — only for the CPU legs and for the
parity checks of what was timed.
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
}
C3 = dict(rows=10_000_000, dims=768, bits=1024)
C4 = dict(rows=100_000_000, bits=1024, queries=100_000, k=100)
L2_FLUSH_BYTES = 256 << 20
KEY_ROW_BITS = 40


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return {"hbm": float(d["hbm_gbs"]), "bf16": float(d["bf16_tflops"]), "src": "measured (MEASURED_PEAKS.json)",
                "sm_max_mhz": float(d.get("sm_max_mhz", 1965.0))}
    return {"hbm": 6650.0, "bf16": 1590.0, "src": "fallback (B200_PROFILING.md)", "sm_max_mhz": 1965.0}


def ncu_traffic(kernel):
    """DRAM bytes per launch of the dominant kernel: a COMMITTED constant from the ncu --set full
    capture under profiles/ (not a measurement of this run)."""
    for name in ("r02_traffic.json", "r01_traffic.json"):
        try:
            v = json.load(open(os.path.join(ROOT, "profiles", name))).get(kernel)
            if v is not None:
                return v, f"profiles/{name} (committed ncu capture, not measured in this run)"
        except (OSError, ValueError):
            pass
    return None, None


def make_config(wl, world, shard):
    """The `config` object, identical in both arms (--impl b200 / reference) of one run."""
    by_query = shard == "queries"
    return {"workload": wl["desc"], "rows": wl["rows"], "bits": wl["bits"], "queries": wl["queries"], "k": wl["k"],
            "parallelism": ("1 GPU" if world == 1 else
                            f"table replicated, query batch sharded x{world} + NCCL all-gather of results" if by_query else
                            f"row-sharded x{world} + NCCL all-gather of candidate keys + K4 merge"),
            "l2": f"flushed between timed iterations ({L2_FLUSH_BYTES >> 20} MiB write)"}


def resolve_shard(args, wl, world):
    if world == 1:
        return "rows"
    if args.shard != "auto":
        return args.shard
    from np_sims.sharded import choose_sharding
    return choose_sharding(wl["rows"], wl["bits"] // 64, wl["queries"], world)


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
                 "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
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


# ----------------------------------------------------------------------------- CPU (reference) legs
def reference_variants():
    """[(label, module)] of the compiled reference builds that run on this host: '-O3 -march=native'
    (fair: hardware popcnt) first, then '-O3' (what NumPy's default build gives the reference)."""
    from oracle import oracle
    out = []
    if oracle.ref_variant_runs("native"):
        out.append(("-O3 -march=native", oracle.ref_ufuncs("native")))
    elif oracle.ref_variant_runs("popcnt"):
        out.append(("-O3 -mpopcnt", oracle.ref_ufuncs("popcnt")))
    if oracle.ref_variant_runs(""):
        out.append(("-O3", oracle.ref_ufuncs("")))
    return out


def cpu_reference_run(wl, sample_queries, steps, warmup, H, W, qvecs, uf=None, build="", threads=None,
                      time_budget_s=None):
    """The reference path on the host cores: per query  np.dot(W, x) -> packbits -> hamming_top_10
    (np_sims/lsh.py:70-74), one query per ThreadPoolExecutor task, GIL released inside the ufunc
    (benchmark/glove.py:183-199).  Returns (queries/s, ms_per_step, info)."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import oracle
    kind = "reference" if uf is not None else "port"
    cores = threads or os.cpu_count() or 1
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
                      f"{H.shape[0]}x{H.shape[1] * 64}-bit table, ThreadPoolExecutor({cores}), "
                      f"build={build or ('oracle.c port' if uf is None else '-O3')}"}
    return qps, 1e3 * total / len(times), info


def cpu_baseline_full(wl, H, W, qvecs, X, sample):
    """the reference ufunc under both flag sets, all cores and one thread, plus the
    hashing baseline (lsh.index's per-row loop and the batched float64 GEMM + packbits)."""
    variants = reference_variants()
    cores = os.cpu_count() or 1
    table = {}
    head = None
    for label, uf in variants:
        _, _, a = cpu_reference_run(wl, sample, steps=64, warmup=1, H=H, W=W, qvecs=qvecs, uf=uf, build=label,
                                    time_budget_s=8.0 if head is None else 5.0)
        _, _, o = cpu_reference_run(wl, max(8, sample // 16), steps=16, warmup=1, H=H, W=W, qvecs=qvecs, uf=uf,
                                    build=label, threads=1, time_budget_s=3.0)
        table[label] = {"all_cores_qps": a["value"], "one_thread_qps": o["value"], "cores": cores,
                        "sample_all": a["sample"], "sample_one": o["sample"]}
        if head is None:
            head = a
    if head is None:       # oracle/_ref not built: the C port
        _, _, head = cpu_reference_run(wl, sample, steps=64, warmup=1, H=H, W=W, qvecs=qvecs, time_budget_s=10.0)
    # hashing: np_sims/lsh.py:36-44 (Python loop over rows, float64 dgemv + packbits per row)
    n_loop = min(20_000, len(X))
    t0 = time.perf_counter()
    for i in range(n_loop):
        np.packbits(np.dot(W, X[i]) >= 0, bitorder="little").view(np.uint64)
    loop_rps = n_loop / (time.perf_counter() - t0)
    n_gemm = min(100_000, len(X))
    t0 = time.perf_counter()
    np.packbits((X[:n_gemm].astype(np.float64) @ W.T) >= 0, axis=-1, bitorder="little")
    gemm_rps = n_gemm / (time.perf_counter() - t0)
    head = dict(head)
    head["variants"] = table
    head["hashing"] = {"lsh_index_rows_per_s": loop_rps, "lsh_index_sample": f"{n_loop} rows, per-row loop (lsh.py:36-44)",
                       "batched_f64_gemm_packbits_rows_per_s": gemm_rps, "batched_sample": f"{n_gemm} rows, NumPy BLAS"}
    try:
        model = [l.split(":", 1)[1].strip() for l in open("/proc/cpuinfo") if l.startswith("model name")]
        head["cpu_model"] = model[0] if model else None
    except OSError:
        head["cpu_model"] = None
    return head


def cpu_signatures(wl, X, W):
    from oracle import oracle
    out = np.empty((len(X), wl["bits"] // 64), dtype=np.uint64)
    for i in range(0, len(X), 50_000):
        out[i:i + 50_000] = oracle.index_np(X[i:i + 50_000], W)
    return out


def run_reference(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    X = make_vectors(wl["rows"], wl["dims"], 0)
    qvecs = make_vectors(wl["queries"], wl["dims"], 1)
    np.random.seed(0)
    from oracle import oracle
    W = oracle.create_projections(wl["bits"], wl["dims"])
    H = cpu_signatures(wl, X, W)
    sample = args.ref_sample
    variants = reference_variants()
    label, uf = variants[0] if variants else ("", None)
    qps, ms, info = cpu_reference_run(wl, sample, args.steps, args.warmup, H=H, W=W, qvecs=qvecs, uf=uf, build=label)
    info["step"] = f"{sample}-query sample of the workload per step (CPU arm)"
    line = {
        "impl": "reference", "metric": "hamming_top_n queries/s", "value": qps, "unit": "queries/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": make_config(wl, world, resolve_shard(args, wl, world)),
        "cpu_baseline": info,
        "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------- timing helper
class Timer:
    def __init__(self, torch, dist, dev, world, flush):
        self.torch, self.dist, self.dev, self.world, self.flush = torch, dist, dev, world, flush

    def run(self, fn, steps, warmup, flush=True):
        """W untimed warm-up steps, then exactly `steps` steps, each bracketed by CUDA events on the
        launching stream, barrier + synchronize on both sides, MAX over ranks.  -> (total_ms, launches, wall)."""
        torch, dist = self.torch, self.dist
        from np_sims import _lib
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        if self.world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        starts = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        stops = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        launches0 = _lib.raw("nps_launch_count")()
        w0 = time.perf_counter()
        for i in range(steps):
            if flush:
                self.flush.fill_(i & 0xFF)         # evict L2 between timed iterations (untimed)
            starts[i].record()
            fn()
            stops[i].record()
        torch.cuda.synchronize()
        w1 = time.perf_counter()
        launches = _lib.raw("nps_launch_count")() - launches0
        if self.world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        total_ms = sum(s.elapsed_time(e) for s, e in zip(starts, stops))
        if self.world > 1:
            t = torch.tensor([total_ms], dtype=torch.float64, device=self.dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            total_ms = float(t.item())
        return total_ms, launches, (w0, w1)


def keys_of(rows, dists):
    return (dists.astype(np.uint64) << np.uint64(KEY_ROW_BITS)) | rows.astype(np.uint64)


# ----------------------------------------------------------------------------- C4 sub-record
C4_BLOCK = 500_000
C4_SAMPLE = 64          # queries checked against the oracle
C4_REGIONS = 8          # planted neighbours / oracle subsample live at the start of each eighth of the table
C4_SUB = 65_536         # subsample rows per region


def c4_plants(rows, qsig_host, words):
    """Planted neighbours (SURVEY §8d C4 "planted neighbours" variant), independent of N: for each of
    the C4_SAMPLE checked queries and each table region, 4 rows = the query with a few bits flipped.
    -> (global row ids [P], signatures [P, words], query ids [C4_SAMPLE])."""
    nq = len(qsig_host)
    qids = np.arange(C4_SAMPLE, dtype=np.int64) * max(1, nq // C4_SAMPLE)
    qids = np.minimum(qids, nq - 1)
    rng = np.random.default_rng(44)
    ids, sigs = [], []
    region = rows // C4_REGIONS
    for j, qid in enumerate(qids):
        for r in range(C4_REGIONS):
            for t in range(4):
                row = r * region + 1000 + 4 * j + t
                if row >= min((r + 1) * region, rows):
                    continue
                s = qsig_host[qid].copy()
                for b in rng.choice(words * 64, size=3 + 5 * t + r, replace=False):
                    s[b // 64] ^= np.uint64(1) << np.uint64(b % 64)
                ids.append(row)
                sigs.append(s)
    return np.asarray(ids, dtype=np.int64), np.asarray(sigs, dtype=np.uint64).reshape(-1, words), qids


def c4_build_shard(torch, dev, lo, hi, words, plant_ids, plant_sigs):
    """Rows [lo, hi) of the synthetic table: i.i.d. bits drawn on the device in fixed 500k-row blocks, one
    generator seed per block, so that the table is the same for every N; then the planted rows."""
    H = torch.empty((hi - lo, words), dtype=torch.int64, device=dev)
    g = torch.Generator(device=dev)
    for blk in range(lo // C4_BLOCK, (hi + C4_BLOCK - 1) // C4_BLOCK):
        g.manual_seed(3_000_000 + blk)
        block = torch.randint(-2**63, 2**63 - 1, (C4_BLOCK, words), dtype=torch.int64, device=dev, generator=g)
        b0 = blk * C4_BLOCK
        s, e = max(lo, b0), min(hi, b0 + C4_BLOCK)
        H[s - lo:e - lo] = block[s - b0:e - b0]
        del block
    mine = (plant_ids >= lo) & (plant_ids < hi)
    if mine.any():
        H[torch.from_numpy(plant_ids[mine] - lo).to(dev)] = torch.from_numpy(plant_sigs[mine].view(np.int64)).to(dev)
    return H


def run_c4(args, torch, dist, dev, rank, world, timer, pk):
    from np_sims import _lib
    from np_sims.sharded import ShardedIndex, shard_bounds
    from np_sims.ufuncs import last_overflow_count
    rows = args.c4_rows or C4["rows"]
    nq = args.c4_queries or C4["queries"]
    k, words = C4["k"], C4["bits"] // 64
    b = shard_bounds(rows, world)
    lo, hi = b[rank], b[rank + 1]
    g = torch.Generator(device=dev)
    g.manual_seed(4)
    qsig_dev = torch.randint(-2**63, 2**63 - 1, (nq, words), dtype=torch.int64, device=dev, generator=g)
    qsig_host = qsig_dev.cpu().numpy().view(np.uint64)
    plant_ids, plant_sigs, qids = c4_plants(rows, qsig_host, words)
    t0 = time.perf_counter()
    H_local = c4_build_shard(torch, dev, lo, hi, words, plant_ids, plant_sigs)
    torch.cuda.synchronize()
    build_s = time.perf_counter() - t0
    index = ShardedIndex(H_local, lo)
    qsig_pinned = torch.from_numpy(qsig_host.view(np.int64)).pin_memory()
    out_pinned = torch.empty((nq, k), dtype=torch.int64).pin_memory()
    result = {}

    def step_device():
        result["rows"], result["dists"] = index.query(qsig_dev, k)

    def step_e2e():
        r, _ = index.query(qsig_pinned.to(dev, non_blocking=True), k)
        out_pinned.copy_(r.view(torch.int64))

    steps, warmup = args.c4_steps, 3
    total_ms, launches, _ = timer.run(step_device, steps, warmup, flush=False)
    overflow = last_overflow_count()
    scan_path = _lib.raw("nps_last_scan_path")()
    ms = total_ms / steps
    e2e_ms, _, _ = timer.run(step_e2e, max(1, steps - 1), 1, flush=False)
    e2e_ms /= max(1, steps - 1)

    # ---- parity of what was timed, at this N: every rank holds the merged result
    step_device()
    torch.cuda.synchronize()
    rows_t, dists_t = result["rows"].view(torch.int64), result["dists"].view(torch.int64)
    sel = torch.from_numpy(qids).to(dev)
    r_s, d_s = rows_t[sel], dists_t[sel]                              # [S, k]
    # identical on every rank
    chk = torch.stack([rows_t.sum(), dists_t.sum(), (rows_t[:, 0] * 31 + rows_t[:, -1]).sum()])
    same = True
    if world > 1:
        hi_c, lo_c = chk.clone(), chk.clone()
        dist.all_reduce(hi_c, op=dist.ReduceOp.MAX)
        dist.all_reduce(lo_c, op=dist.ReduceOp.MIN)
        same = bool(torch.equal(hi_c, lo_c))
    # signatures of the returned rows (each rank contributes the rows of its shard)
    own = (r_s >= lo) & (r_s < hi)
    G = H_local[(r_s - lo).clamp(0, hi - lo - 1)] * own.unsqueeze(-1).to(torch.int64)
    # the oracle's row subsample: the first C4_SUB rows of every region
    region = rows // C4_REGIONS
    sub = torch.zeros((C4_REGIONS, min(C4_SUB, region), words), dtype=torch.int64, device=dev)
    for r in range(C4_REGIONS):
        s0, s1 = r * region, r * region + sub.shape[1]
        a, e = max(s0, lo), min(s1, hi)
        if a < e:
            sub[r, a - s0:e - s0] = H_local[a - lo:e - lo]
    if world > 1:
        dist.all_reduce(G)
        dist.all_reduce(sub)
    rec = None
    if rank == 0:
        from oracle import oracle
        Gh = G.cpu().numpy().view(np.uint64)
        subh = sub.cpu().numpy().view(np.uint64).reshape(-1, words)
        sub_ids = (np.arange(C4_REGIONS, dtype=np.uint64)[:, None] * np.uint64(region)
                   + np.arange(sub.shape[1], dtype=np.uint64)[None, :]).reshape(-1)
        rs, ds = r_s.cpu().numpy().view(np.uint64), d_s.cpu().numpy().view(np.uint64)
        planted_found = planted_total = ties = 0
        for j, qid in enumerate(qids):
            q = qsig_host[qid]
            got = keys_of(rs[j], ds[j])
            assert np.all(got[1:] > got[:-1]), "c4: merged keys not strictly ascending"
            assert np.array_equal(oracle.hamming(Gh[j], q), ds[j]), "c4: distances of the returned rows differ from the oracle"
            d_sub = oracle.hamming(subh, q)
            keys_sub = keys_of(sub_ids, d_sub)
            want = np.sort(keys_sub[keys_sub <= got[-1]])
            have = got[np.isin(rs[j], sub_ids)]
            assert np.array_equal(want, have), "c4: result restricted to the oracle's row subsample differs"
            mine = plant_ids[np.arange(len(plant_ids)) // (4 * C4_REGIONS) == j] if len(plant_ids) == C4_SAMPLE * 4 * C4_REGIONS else None
            if mine is not None:
                planted_total += len(mine)
                planted_found += int(np.isin(mine.astype(np.uint64), rs[j]).sum())
            ties += int(ds[j][-1] == ds[j][-2])
        assert same, "c4: ranks disagree on the merged result"
        if planted_total:
            assert planted_found == planted_total, f"c4: {planted_total - planted_found} planted neighbours missing"
        n_local = hi - lo
        flop = 2.0 * n_local * nq * C4["bits"]
        tensor_peak = 4.0 * pk["bf16"]
        rec = {
            "workload": f"C4: {rows} x {C4['bits']}-bit synthetic signature table (i.i.d. bits + planted neighbours), "
                        f"{nq} query batch, top-{k}",
            "value": nq / (ms * 1e-3), "unit": "queries/s", "ms_per_step": ms, "steps": steps, "warmup": warmup,
            "n_gpus": world, "scaling": "strong",
            "parallelism": ("1 GPU" if world == 1 else f"row-sharded x{world} + NCCL all-gather of candidate keys + K4 merge"),
            "rows_per_gpu": int(n_local), "table_bytes_per_gpu": int(n_local * words * 8),
            "l2": "inputs larger than L2 (table shard >= 1.6 GB), no flush",
            "e2e": {"value": nq / (e2e_ms * 1e-3), "unit": "queries/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": int(nq * words * 8) * world, "d2h_bytes_per_step": int(nq * k * 8) * world},
            "gpu_launches_per_step": int(launches // steps),
            "scan_path": "mma (tcgen05 kind::mxf4)" if scan_path == _lib.SCAN_MMA else "popc",
            "roofline": {"bound": "tensor", "unit": "TFLOP/s", "achieved": flop / (ms * 1e-3) / 1e12, "peak": tensor_peak,
                         "frac": flop / (ms * 1e-3) / 1e12 / tensor_peak,
                         "note": "2 * rows_per_gpu * queries * bits flop over the WHOLE step of the slowest rank "
                                 "(scan phases + selects + all-gather + merge), peak = 4 x measured bf16 burst"},
            "overflowed_queries": int(overflow), "table_build_s": build_s,
            "parity": {"queries_checked": int(len(qids)), "ranks_identical": same,
                       "distances_of_returned_rows_vs_oracle": "equal",
                       "oracle_row_subsample": f"{len(sub_ids)} rows ({C4_REGIONS} regions x {sub.shape[1]}): result restricted "
                                               "to the subsample == oracle keys <= k-th key, in order",
                       "planted_neighbours_found": f"{planted_found}/{planted_total}",
                       "queries_with_boundary_tie": ties},
        }
        if world == 1 and "cpu" not in args.skip:
            rec["cpu_baseline"] = c4_cpu_baseline(H_local, qsig_host, k)
    del H_local, index, qsig_dev, G, sub
    torch.cuda.empty_cache()
    return rec


def c4_cpu_baseline(H_local, qsig_host, k, sub_rows=5_000_000, nq=32):
    """the reference ufunc on a stated row-subsample of the C4 table, extrapolated
    linearly to the full table (k = 100 has no reference kernel: `hamming` + argpartition, and the
    k = 1000 superset kernel `hamming_top_cand`)."""
    from concurrent.futures import ThreadPoolExecutor
    variants = reference_variants()
    if not variants:
        return None
    label, uf = variants[0]
    n = min(sub_rows, H_local.shape[0])
    H = H_local[:n].cpu().numpy().view(np.uint64)
    cores = os.cpu_count() or 1
    full = H_local.shape[0]
    out = {"kind": "reference", "cores": cores, "build": label, "unit": "queries/s",
           "sample": f"{nq} queries over the first {n} rows, ThreadPoolExecutor({cores}); q/s extrapolated linearly "
                     f"to {full} rows (x{n / full:.4f})"}
    for name, fn in (("hamming_argpartition", lambda q: np.argpartition(uf.hamming(H, q), k)[:k]),
                     ("hamming_top_cand_k1000", lambda q: uf.hamming_top_cand(H, q))):
        with ThreadPoolExecutor(max_workers=cores) as ex:
            list(ex.map(fn, qsig_host[:cores]))
            t0 = time.perf_counter()
            list(ex.map(fn, qsig_host[:nq]))
            dt = time.perf_counter() - t0
        out[name + "_qps_on_sample"] = nq / dt
        out[name + "_qps_extrapolated"] = nq / dt * n / full
    out["value"] = out["hamming_argpartition_qps_extrapolated"]
    return out


# ----------------------------------------------------------------------------- C3 sub-record
def run_c3(args, torch, dist, dev, rank, world, timer, pk):
    import np_sims.lsh as lsh
    from np_sims.sharded import shard_bounds
    rows = args.c3_rows or C3["rows"]
    dims, bits = C3["dims"], C3["bits"]
    b = shard_bounds(rows, world)
    n_local = b[rank + 1] - b[rank]
    g = torch.Generator(device=dev)
    g.manual_seed(2_000 + rank)
    X = torch.empty((n_local, dims), dtype=torch.float32, device=dev)
    for i in range(0, n_local, 1_000_000):           # generated on the device in chunks (30.7 GB at full size)
        c = X[i:i + 1_000_000]
        c.normal_(generator=g)
        c /= c.norm(dim=1, keepdim=True)
    np.random.seed(0)
    W = lsh.create_projections(bits, dims)
    out = {}

    def step():
        out["H"] = lsh.hash_vectors(X, W)

    total_ms, launches, _ = timer.run(step, args.c3_steps, 3, flush=False)
    ms = total_ms / args.c3_steps
    st = lsh.hash_stats()
    rec = None
    # parity: a row subsample of rank 0's shard against the float64 ORACLE (np_sims/lsh.py:31-44 arithmetic)
    if rank == 0:
        from oracle import oracle
        n_sub = min(args.c3_check_rows, n_local)
        sub = torch.randint(0, n_local, (n_sub,), device=dev, generator=g)
        Xs = X[sub].cpu().numpy()
        got = out["H"][sub].cpu().numpy().view(np.uint64)
        want = oracle.index_np(Xs, np.asarray(W))
        flipped = int(np.bitwise_count(got ^ want).sum())
        raw = lsh.hash_vectors(X[sub].contiguous(), W, tf32_raw=True).cpu().numpy().view(np.uint64)
        flipped_raw = int(np.bitwise_count(raw ^ want).sum())
        assert flipped == 0, f"c3: {flipped} signature bits differ from the float64 oracle"
        flop = 2.0 * rows * dims * bits
        t = ms * 1e-3
        rec = {
            "workload": f"C3: {rows} x {dims} fp32 -> {bits}-bit signatures (K1T: TMA-fed tcgen05 split-BF16 GEMM + "
                        "sign/bit-pack epilogue + float64 fix-up)",
            "ms_per_step": ms, "steps": args.c3_steps, "warmup": 3, "n_gpus": world, "rows_per_gpu": int(n_local),
            "value": rows / t, "unit": "rows/s", "scaling": "strong",
            "l2": "inputs larger than L2 (>= 3.8 GB of vectors per GPU), no flush",
            "roofline": {"bound": "tensor", "unit": "TFLOP/s", "peak": pk["bf16"] * world,
                         "achieved": flop / t / 1e12, "frac": flop / t / 1e12 / (pk["bf16"] * world),
                         "executed_flop_factor": 3, "frac_executed": 3 * flop / t / 1e12 / (pk["bf16"] * world),
                         "note": "algorithmic 2*N*D*B flop; the exactness scheme executes 3 BF16 products per flop "
                                 "(xh*wh + xh*wl + xl*wh); peak = measured bf16 burst x n_gpus",
                         "input_gbs_per_gpu": n_local * dims * 4 / t / 1e9, "hbm_peak_gbs": pk["hbm"]},
            "gpu_launches_per_step": int(launches // args.c3_steps),
            "hash": {"recomputed_in_float64": st.get("recomputed"), "by_fixer_warps_in_kernel": st.get("in_kernel"),
                     "by_fix_up_list": st.get("listed"), "pairs": st.get("pairs"), "overflowed": st.get("overflowed"),
                     "rows_checked_vs_float64_oracle": int(n_sub), "flipped_bits_before_fixup": flipped_raw,
                     "flipped_bit_rate_before_fixup": flipped_raw / (n_sub * bits), "flipped_bits_final": flipped},
        }
    del X, out
    torch.cuda.empty_cache()
    return rec


# ----------------------------------------------------------------------------- single-query record
def run_single_query(args, torch, lsh, X, W, H_local, H_host, qvecs_all, cpu_qps):
    """The reference's own calling pattern (benchmark/glove.py:183-199) through the drop-in API:
    `lsh.query(vector, hashes, projections)` one NumPy vector at a time, from a ThreadPoolExecutor of
    os.cpu_count() host threads, over the C2 table; plus the same calls from one thread."""
    from concurrent.futures import ThreadPoolExecutor
    from np_sims._device import attach
    from oracle import oracle
    H_np = attach(H_host, H_local)                 # what lsh.index returns: NumPy array + its HBM copy
    cores = os.cpu_count() or 1
    n = args.single_queries

    def one(i):
        return lsh.query(qvecs_all[i % len(qvecs_all)], H_np, W)[0]

    for i in range(64):
        one(i)
    t0 = time.perf_counter()
    seq = [one(i) for i in range(min(n, 2048))]
    seq_dt = time.perf_counter() - t0
    with ThreadPoolExecutor(max_workers=cores) as ex:
        list(ex.map(one, range(4 * cores)))
        best = None
        for _ in range(3):
            t0 = time.perf_counter()
            res = list(ex.map(one, range(n)))
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
    Wn = np.asarray(W)
    for i in (0, 1, n // 2, n - 1):
        sig = oracle.index_np(qvecs_all[i % len(qvecs_all)][None, :], Wn)[0]
        want = oracle.top_k_queue(H_host, sig, 10)[0]
        assert np.array_equal(res[i], want), "single_query: lsh.query differs from the reference queue order"
        if i < len(seq):
            assert np.array_equal(seq[i], want)
    qps = n / best
    return {"api": "lsh.query(vector, hashes, projections) -> (ids, None), reference queue order (hamming_top_10)",
            "value": qps, "unit": "queries/s", "threads": cores, "queries": n,
            "driver": f"ThreadPoolExecutor({cores}), one query per task (benchmark/glove.py:183-199), best of 3",
            "one_thread_qps": len(seq) / seq_dt, "one_thread_us_per_query": 1e6 * seq_dt / len(seq),
            "vs_cpu_baseline": (qps / cpu_qps) if cpu_qps else None,
            "parity": "4 results == oracle.top_k_queue (slot order)"}


# ----------------------------------------------------------------------------- B200 arm
def run_b200(args, wl):
    import torch
    import torch.distributed as dist
    import np_sims
    import np_sims.lsh as lsh
    from np_sims import _lib
    from np_sims.sharded import ReplicatedIndex, ShardedIndex, query_bounds, shard_bounds

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    pk = peaks()
    hbm_peak, peak_src, sm_max_mhz = pk["hbm"], pk["src"], pk["sm_max_mhz"]
    k, nq, words = wl["k"], wl["queries"], wl["bits"] // 64
    shard = resolve_shard(args, wl, world)
    by_query = shard == "queries"
    qb = query_bounds(nq, world) if by_query else [0] * rank + [0, nq]
    qlo, qhi = qb[rank], qb[rank + 1]        # the queries THIS rank hashes and answers

    # ---- inputs (synthetic), resident in HBM before timing
    b = [0] * (rank + 1) + [wl["rows"]] if by_query else shard_bounds(wl["rows"], world)
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
    index = ReplicatedIndex(H_local) if by_query else ShardedIndex(H_local, b[rank])
    out_pinned = torch.empty((nq, k), dtype=torch.int64).pin_memory()
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=dev)
    timer = Timer(torch, dist, dev, world, flush)

    def search(sigs):
        """-> (rows, dists) [nq, k] for the WHOLE batch, on every rank."""
        if by_query:      # ids only, like the reference's hamming_top_10: one all-gather
            return index.gather(index.query_local(sigs, k, return_distances=False), nq), None
        return index.query(sigs, k)

    def step_plain():
        return search(lsh.hash_vectors(qvecs_dev, W))

    # The step as ONE CUDA graph (np_sims.pipeline.GraphedQueryBatch: hash + scan phases + selects + gated
    # repairs captured once, replayed per step; the all-gather of N > 1 stays outside the graph).
    graphed = None
    if (world == 1 or by_query) and not args.no_graph:
        from np_sims.pipeline import GraphedQueryBatch
        graphed = GraphedQueryBatch(H_local, W, num_queries=qhi - qlo, dims=wl["dims"], n=k,
                                    return_distances=world == 1, vectors=qvecs_dev)

    def finish(out):
        if world == 1:
            return out                                      # (rows, dists)
        return index.gather(out, nq), None

    def step_device():
        return finish(graphed()) if graphed is not None else step_plain()

    def step_e2e():
        if graphed is not None:
            graphed.vectors.copy_(qvecs_pinned, non_blocking=True)      # H2D of the step's query vectors
            rows, _ = finish(graphed())
        else:
            rows, _ = search(lsh.hash_vectors(qvecs_pinned.to(dev, non_blocking=True), W))
        out_pinned.copy_(rows.view(torch.int64))            # D2H of the step's result (blocking)
        return out_pinned

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    total_ms, launches, (w0, w1) = timer.run(step_device, args.steps, args.warmup)
    clocks = sampler.stop(w0, w1) if rank == 0 else None
    value = nq * args.steps / (total_ms * 1e-3)
    plain_ms = None
    if graphed is not None:       # the same step launch by launch (what the graph replays), for comparison
        plain_ms, launches, _ = timer.run(step_plain, args.steps, 3)
        plain_ms /= args.steps
    e2e_ms, _, _ = timer.run(step_e2e, args.steps, max(3, args.warmup // 2))
    e2e_value = nq * args.steps / (e2e_ms * 1e-3)

    # ---- the same end-to-end step as a STREAM of batches (np_sims.pipeline.QueryStream, 2 in flight): the
    # H2D copy of batch i+1 and the D2H of batch i-1 overlap the search of batch i.  Every batch's copies
    # are inside the timed region; L2 is evicted before every batch on the search stream (timed).
    pipelined = None
    if world == 1 and "pipe" not in args.skip:
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

    # ---- what was just timed, once more, for the parity check below
    rows, dists = step_device()
    torch.cuda.synchronize()

    # ---- roofline pass: the dominant kernel timed per launch on its stream
    _lib.raw("nps_profile_enable")(1)
    scan_ms, merge_ms, phase_runs = [], [], []
    for i in range(max(3, min(args.steps, 10))):
        flush.fill_(i)
        step_plain()
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
    traffic, traffic_src = ncu_traffic("mma_scan_kernel" if phase_runs else "scan_topk_kernel")
    if world != 1:
        traffic, traffic_src = None, None
    if phase_runs:
        # Tensor-core scan (mma_scan_kernel): bound by the tensor pipe.  Algorithmic work of a launch
        # = 2 * rows * queries * bits flop (one +-1 multiply-add per signature bit and (row, query)
        # pair; the folded-threshold slice is overhead, not counted).  Peak: kind::mxf4 runs at 4x
        # the bf16 rate (9 vs 2.25 PFLOP/s nominal); MEASURED_PEAKS.json holds the bf16 figure.
        # `achieved` = flop of ALL launches of the kernel in a step / their summed durations
        # (= average flop per launch / average launch duration).
        plan = _lib.last_mma_plan()
        tensor_peak = 4.0 * pk["bf16"]
        nph = len(phase_runs[0])
        tiles = [phase_runs[0][j][0] for j in range(nph)]
        ph_ms = [float(np.median([r[j][1] for r in phase_runs])) for j in range(nph)]
        launched = [j for j in range(nph) if tiles[j]]
        flop_all = 2.0 * n_local * nq_local * wl["bits"]
        kern_ms = float(np.sum([ph_ms[j] for j in launched]))
        ach = flop_all / (kern_ms * 1e-3) / 1e12
        issue_peak = 16359 * 2 * 148 * sm_mhz * 1e6 / 1e12
        dom = int(np.argmax(ph_ms))
        rows_dom = min(tiles[dom] * 128, n_local)
        roofline = {
            "kernel": f"mma_scan_kernel (tcgen05 kind::mxf4, M128 x N{plan['n_tile']} x K64), all {len(launched)} launches of a step",
            "bound": "tensor", "achieved": ach, "peak": tensor_peak, "unit": "TFLOP/s", "frac": ach / tensor_peak,
            "traffic": traffic, "traffic_source": traffic_src,
            "peak_source": "4 x measured bf16 burst (MEASURED_PEAKS.json): mxf4 : bf16 = 9 : 2.25 PFLOP/s nominal",
            "frac_of_measured_mxf4_issue_rate": ach / issue_peak,
            "measured_mxf4_issue_rate_note": "scripts/ubench/umma_mxf4_probe.cu: 16359 MAC/clk/SM back-to-back = "
                                             f"{issue_peak:.0f} TFLOP/s at {sm_mhz:.0f} MHz — a stricter denominator",
            "launches_per_step": len(launched), "ms_per_launch": kern_ms / max(1, len(launched)),
            "kernel_ms_per_step": kern_ms, "share_of_step": kern_ms / step_ms,
            "algorithmic": {"rows": n_local, "queries": nq_local, "bits": wl["bits"], "flop_per_step": flop_all},
            "phases": [{"tiles": t, "scan_ms": m,
                        "tflops": (2.0 * min(t * 128, n_local) * nq_local * wl["bits"] / (m * 1e-3) / 1e12) if (t and m) else None}
                       for t, m in zip(tiles, ph_ms)],
            "dominant_launch": {"phase": dom, "ms": ph_ms[dom],
                                "frac": 2.0 * rows_dom * nq_local * wl["bits"] / (ph_ms[dom] * 1e-3) / 1e12 / tensor_peak},
            "scan_plus_select_ms": scan_avg, "frac_scan_plus_select": flop_all / (scan_avg * 1e-3) / 1e12 / tensor_peak,
            "frac_of_step": flop_all / (step_ms * 1e-3) / 1e12 / tensor_peak,
            "plan": plan,
        }
    else:
        plan = _lib.last_scan_plan()
        # algorithmic bytes of one scan launch: signatures + queries + results
        alg_bytes = n_local * words * 8 + nq_local * words * 8 + (
            nq_local * n_local * 2 if plan["cap"] == 0 else nq_local * plan["num_chunks"] * k * 8)
        achieved_gbs = alg_bytes / (scan_avg * 1e-3) / 1e9
        popc32 = 2.0 * n_local * nq_local * words
        popc_peak = 148 * 16 * sm_mhz * 1e6
        roofline = {
            "kernel": f"scan_topk_kernel<W={words},R={plan['rows_per_thread']}>",
            "bound": "hbm", "achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s",
            "frac": achieved_gbs / hbm_peak, "traffic": traffic, "traffic_source": traffic_src,
            "peak_source": peak_src, "ms_per_launch": scan_avg, "share_of_step": scan_avg / step_ms,
            "popc": {"achieved_gpopc32_s": popc32 / (scan_avg * 1e-3) / 1e9,
                     "peak_gpopc32_s": popc_peak / 1e9, "frac": popc32 / (scan_avg * 1e-3) / popc_peak},
            "plan": plan, "merge_ms_per_launch": float(np.mean(merge_ms)),
        }

    # ---- parity of the timed step at THIS N against the oracle (rank 0; the table is gathered when row-sharded)
    if by_query or world == 1:
        H_full = H_local
    else:
        H_full = torch.zeros((wl["rows"], words), dtype=torch.int64, device=dev)
        H_full[b[rank]:b[rank + 1]] = H_local
        dist.all_reduce(H_full)
    parity = None
    H_host = None
    if rank == 0:
        from oracle import oracle
        Wn = np.asarray(W)
        H_host = H_full.cpu().numpy().view(np.uint64)
        n_chk = args.parity_queries
        sel = np.unique(np.linspace(0, nq - 1, n_chk).astype(np.int64))
        sig_host = oracle.index_np(qvecs_all[sel], Wn)
        got = rows.view(torch.int64)[torch.from_numpy(sel).to(dev)].cpu().numpy().view(np.uint64)
        got_d = dists.view(torch.int64)[torch.from_numpy(sel).to(dev)].cpu().numpy().view(np.uint64) if dists is not None else None
        for i in range(len(sel)):
            want, want_d = oracle.top_k_canonical(H_host, sig_host[i], k)
            assert np.array_equal(got[i], want), f"timed step disagrees with the oracle (query {sel[i]})"
            if got_d is not None:
                assert np.array_equal(got_d[i], want_d), f"timed step distances disagree with the oracle (query {sel[i]})"
            if pipelined is not None:
                assert np.array_equal(pipelined["last_ids"][sel[i]], want), "streamed step disagrees with the oracle"
        t_rows = np.unique(np.linspace(0, wl["rows"] - 1, 20_000).astype(np.int64))
        t_flips = int(np.bitwise_count(H_host[t_rows] ^ oracle.index_np(X[t_rows], Wn)).sum())
        assert t_flips == 0, f"table signatures differ from the float64 oracle in {t_flips} bits"
        parity = {"queries_checked_vs_oracle": int(len(sel)), "spread": "evenly over the whole batch (every rank's slice)",
                  "table_rows_checked_vs_float64_oracle": int(len(t_rows)), "table_flipped_bits": t_flips,
                  "n_gpus": world, "result": "bit-identical ids" + (" and distances" if got_d is not None else "")}
    del H_full

    line = None
    stream = cpu = single = hash_report = None
    if rank == 0:
        if world == 1 and "stream" not in args.skip:
            stream = stream_roofline(torch, np_sims, _lib, dev, hbm_peak, peak_src)
        if world == 1 and "cpu" not in args.skip:
            cpu = cpu_baseline_full(wl, H_host, np.asarray(W), qvecs_all, X, args.cpu_sample)
        if world == 1 and "single" not in args.skip:
            single = run_single_query(args, torch, lsh, X, W, H_local, H_host, qvecs_all, cpu["value"] if cpu else None)
        # hashing parity of the timed step's K1T (north-star: "flipped-bit rate reported"): bare tensor-core
        # signs and final signatures of the query batch against the float64 kernel and the float64 oracle
        def bits_differing(a, b_):
            x = (a ^ b_).cpu().numpy().view(np.uint64)
            return int(np.bitwise_count(x).sum())
        s_exact = lsh.hash_vectors(qvecs_dev, W, exact_only=True)
        s_final = lsh.hash_vectors(qvecs_dev, W)
        st = lsh.hash_stats()
        s_raw = lsh.hash_vectors(qvecs_dev, W, tf32_raw=True)
        nbits = int(qvecs_dev.shape[0]) * wl["bits"]
        from oracle import oracle
        s_oracle = oracle.index_np(qvecs_host, np.asarray(W))
        hash_report = {"path": st["path"], "bits": nbits,
                       "encoding": "bf16 x 2 (kind::f16), three products, separate correction accumulator",
                       "tolerance": f"(52 + 0.006 dims) * 2^-22 |x| max|w| + 2.68 * 2^-22 * sum_i |x[0:16i)| max_j |w_j[0:16i)| "
                                    f"<= (52 + 0.17 dims) * 2^-22 |x| max|w| = {(52 + 0.17 * wl['dims']) * 2.0 ** -22:.3e} |x|",
                       "recomputed_in_float64": st.get("recomputed"), "recomputed_frac": (st.get("recomputed") or 0) / nbits,
                       "flipped_bits_before_fixup": bits_differing(s_raw, s_exact),
                       "flipped_bits_final": bits_differing(s_final, s_exact),
                       "flipped_bits_final_vs_float64_oracle": int(np.bitwise_count(
                           s_final.cpu().numpy().view(np.uint64) ^ s_oracle).sum())}
        assert hash_report["flipped_bits_final"] == 0, "K1T signatures differ from the float64 kernel"
        assert hash_report["flipped_bits_final_vs_float64_oracle"] == 0, "K1T signatures differ from the float64 oracle"
    # free the C2 working set before the large sub-records
    used_graph = graphed is not None
    graphed = None
    del index, H_local, qvecs_dev, flush
    timer.flush = None
    torch.cuda.empty_cache()

    c3 = run_c3(args, torch, dist, dev, rank, world, timer, pk) if "c3" not in args.skip else None
    c4 = run_c4(args, torch, dist, dev, rank, world, timer, pk) if "c4" not in args.skip else None

    if rank == 0:
        h2d = qvecs_all.nbytes * (1 if by_query else world)
        line = {
            "metric": "hamming_top_n queries/s", "value": value, "unit": "queries/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u64",
            "data": "synthetic",
            "config": make_config(wl, world, shard),
            "step": "hash query batch (K1T) + batched scan/top-n (K2T tensor-core phases + select)" +
                    (", replayed as one CUDA graph (np_sims.pipeline.GraphedQueryBatch)" if used_graph else "") +
                    ((" + all-gather of results" if by_query else " + all-gather + K4 merge") if world > 1 else ""),
            "ms_per_step_launch_by_launch": plain_ms,
            "index_build_s": index_s,
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "queries/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(nq * k * 8) * world, "ms_per_step": e2e_ms / args.steps,
                    "mode": "one blocking call per step (copy in, search, copy out)",
                    "pipelined": ({kk: vv for kk, vv in pipelined.items() if kk != "last_ids"} if pipelined else None)},
            "gpu_launches": int(launches),
            "roofline": roofline,
            "roofline_stream": stream,
            "parity": parity,
            "hash": hash_report,
            "cpu_baseline": cpu,
            "single_query": single,
            "c3": c3,
            "c4": c4,
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
    ms, sel_ms = [], []
    for _ in range(reps):
        np_sims.hamming_top_n(H, q, k)
        a, b_ = _lib.profile_last()
        ms.append(a)
        sel_ms.append(b_)
    _lib.raw("nps_profile_enable")(0)
    plan = _lib.last_scan_plan()
    avg = float(np.mean(ms))
    hist = plan["cap"] == 0       # K2H (hist_select.cuh): distances out as uint16, selection by counting
    alg = rows * words * 8 + words * 8 + (rows * 2 if hist else plan["num_chunks"] * k * 8)
    gbs = alg / (avg * 1e-3) / 1e9
    out = {"kernel": (f"hist_scan_kernel<W={words},R={plan['rows_per_thread']}> (H1 of K2H: table stream + uint16 distances "
                      "+ histogram)" if hist else f"scan_topk_kernel<W={words},R={plan['rows_per_thread']}>"),
           "algorithmic_bytes": alg, "workload":
           f"1 query x {rows} x {words * 64}-bit signatures ({rows * words * 8 / 1e9:.1f} GB, larger than L2), top-{k}",
           "bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak,
           "frac_of_nominal_8TBs": gbs / 8000.0,
           "ms_per_launch": avg, "best_ms": float(np.min(ms)), "select_ms": float(np.mean(sel_ms)),
           "frac_whole_call": rows * words * 8 / ((avg + float(np.mean(sel_ms))) * 1e-3) / 1e9 / hbm_peak,
           "peak_source": peak_src, "plan": plan}
    # the reference's own API on the same table: hamming_top_10 (queue order), whole call device-timed
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ref_ms = []
    for i in range(6):
        e0.record()
        np_sims.hamming_top_10(H, q[0])
        e1.record()
        torch.cuda.synchronize()
        if i >= 2:
            ref_ms.append(e0.elapsed_time(e1))
    r_avg = float(np.mean(ref_ms))
    out["hamming_top_10"] = {"api": "np_sims.hamming_top_10 (reference queue order), whole call, device-timed",
                             "ms_per_call": r_avg, "achieved": rows * words * 8 / (r_avg * 1e-3) / 1e9, "unit": "GB/s",
                             "frac": rows * words * 8 / (r_avg * 1e-3) / 1e9 / hbm_peak}
    del H
    torch.cuda.empty_cache()
    return out


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
    ap.add_argument("--parity-queries", type=int, default=256, help="queries of the timed C2 step checked against the oracle")
    ap.add_argument("--single-queries", type=int, default=4096, help="queries of the single_query record")
    ap.add_argument("--skip", default="", help="comma list of sub-records to skip: c3,c4,single,cpu,stream,pipe")
    ap.add_argument("--c4-rows", type=int, default=0, help="development: smaller C4 table")
    ap.add_argument("--c4-queries", type=int, default=0)
    ap.add_argument("--c4-steps", type=int, default=3)
    ap.add_argument("--c3-rows", type=int, default=0, help="development: smaller C3 input")
    ap.add_argument("--c3-steps", type=int, default=5)
    ap.add_argument("--c3-check-rows", type=int, default=20_000)
    ap.add_argument("--no-graph", action="store_true", help="time the step launch by launch instead of as a CUDA graph")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-stream", action="store_true")
    args = ap.parse_args()
    args.skip = set(x for x in args.skip.split(",") if x)
    if args.no_cpu_baseline:
        args.skip.add("cpu")
    if args.no_stream:
        args.skip.add("stream")
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, wl)
    else:
        run_b200(args, wl)


if __name__ == "__main__":
    main()
