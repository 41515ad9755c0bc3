"""BASELINE config C5: latency sweep of hamming_top_n on one B200 — batch 1..4096, 64..2048-bit
signatures, k = 1..1000 — showing the memory-bound / compute-bound crossover and which kernel
family the library picks (XOR+POPC stream vs tcgen05 phases).  Writes one JSON document.

    python scripts/sweep_c5.py [--rows 16000000] [--out profiles/r01_c5_sweep.json]
"""
import argparse, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "np-sims_b200")]
import torch
import np_sims
from np_sims import _lib


def time_call(fn, reps):
    fn(); torch.cuda.synchronize()
    ms = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    return float(np.median(ms)), float(np.min(ms))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=16_000_000)
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r02_c5_sweep.json"))
    ap.add_argument("--table", default=None, help="also write the latency table (text) here")
    ap.add_argument("--force", default="auto", choices=["auto", "popc", "mma", "both"])
    args = ap.parse_args()
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev); g.manual_seed(5)
    out = {"rows": args.rows, "note": "i.i.d. random signatures (seed 5); median of 5 after 1 warm-up; table > L2 for every width",
           "hbm_peak_gbs": peaks["hbm_gbs"], "points": []}
    for bits in (64, 128, 256, 512, 640, 1024, 2048):
        w = bits // 64
        H = torch.randint(-2**63, 2**63 - 1, (args.rows, w), dtype=torch.int64, device=dev, generator=g)
        for k in (1, 10, 100, 1000):
            for batch in (1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096):
                Q = torch.randint(-2**63, 2**63 - 1, (batch, w), dtype=torch.int64, device=dev, generator=g)
                paths = {"auto": [_lib.SCAN_AUTO], "popc": [_lib.SCAN_POPC], "mma": [_lib.SCAN_MMA],
                         "both": [_lib.SCAN_POPC, _lib.SCAN_MMA]}[args.force]
                for path in paths:
                    if path == _lib.SCAN_MMA and (w > 23 or batch < 16):
                        continue
                    prev = _lib.set_scan_path(path)
                    try:
                        med, best = time_call(lambda: np_sims.hamming_top_n(H, Q, k), 5 if batch * args.rows * w < 4e12 else 2)
                        used = "mma" if _lib.raw("nps_last_scan_path")() == _lib.SCAN_MMA else (
                            "hist" if _lib.last_scan_plan()["cap"] == 0 else "popc")
                    finally:
                        _lib.set_scan_path(prev)
                    table_gb = args.rows * w * 8 / 1e9
                    pt = {"bits": bits, "k": k, "batch": batch, "path": used, "ms": med, "best_ms": best,
                          "queries_per_s": batch / (med * 1e-3), "table_gbs": table_gb / (med * 1e-3),
                          "hbm_frac": table_gb / (med * 1e-3) / peaks["hbm_gbs"],
                          "pair_tflops": 2.0 * args.rows * batch * bits / (med * 1e-3) / 1e12}
                    out["points"].append(pt)
                    print(json.dumps(pt), flush=True)
        del H
    json.dump(out, open(args.out, "w"), indent=0)
    if args.table:
        cols = (1, 4, 16, 64, 256, 1024, 4096)
        with open(args.table, "w") as f:
            f.write(f"# C5 sweep, {args.rows // 1_000_000}M-row table, one B200, hamming_top_n call latency in ms (median of 5), "
                    "h = K2H table stream + counting select, p = K2 XOR+POPC lists, m = tcgen05 phases\n")
            f.write("bits    k | " + " | ".join(f"{c:7d}" for c in cols) + "\n")
            for bits in (64, 128, 256, 512, 640, 1024, 2048):
                for k in (1, 10, 100, 1000):
                    cells = []
                    for c in cols:
                        pt = [p for p in out["points"] if p["bits"] == bits and p["k"] == k and p["batch"] == c]
                        cells.append(f"{pt[0]['ms']:6.2f}{pt[0]['path'][0]}" if pt else "      -")
                    f.write(f"{bits:5d} {k:4d} | " + " | ".join(cells) + "\n")


if __name__ == "__main__":
    main()
