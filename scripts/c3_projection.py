"""BASELINE config C3: projection-only, 10M x 768 fp32 -> 1024-bit signatures on one B200 (K1T), with a
parity check of a 100k-row subsample against the float64 kernel.  Prints one JSON line.

    python scripts/c3_projection.py [--rows 10000000]
"""
import argparse, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "np-sims_b200")]
import torch
import np_sims.lsh as lsh

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=10_000_000)
ap.add_argument("--dims", type=int, default=768)
ap.add_argument("--bits", type=int, default=1024)
args = ap.parse_args()
g = torch.Generator(device="cuda"); g.manual_seed(2)
X = torch.empty((args.rows, args.dims), dtype=torch.float32, device="cuda")
for i in range(0, args.rows, 1_000_000):           # generated on the device in chunks (30.7 GB at full size)
    c = X[i:i + 1_000_000]
    c.normal_(generator=g)
    c /= c.norm(dim=1, keepdim=True)
np.random.seed(0)
W = lsh.create_projections(args.bits, args.dims)
lsh.hash_vectors(X[:100_000], W); torch.cuda.synchronize()
ms = []
for _ in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); H = lsh.hash_vectors(X, W); e1.record(); torch.cuda.synchronize()
    ms.append(e0.elapsed_time(e1))
st = lsh.hash_stats()
sub = torch.randint(0, args.rows, (100_000,), device="cuda", generator=g)
ref = lsh.hash_vectors(X[sub].contiguous(), W, exact_only=True)
same = bool(torch.equal(ref, H[sub]))
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"bf16_tflops": 1590.0, "hbm_gbs": 6650.0}
t = float(np.median(ms)) * 1e-3
flop = 2.0 * args.rows * args.dims * args.bits
print(json.dumps({"workload": f"C3: {args.rows} x {args.dims} fp32 -> {args.bits}-bit signatures (K1T)", "ms": t * 1e3,
                  "rows_per_s": args.rows / t, "algorithmic_tflops": flop / t / 1e12,
                  "tf32_peak_tflops": peaks["bf16_tflops"] / 2, "frac_of_tf32_peak_algorithmic": flop / t / 1e12 / (peaks["bf16_tflops"] / 2),
                  "executed_flop_factor": 3, "input_gbs": args.rows * args.dims * 4 / t / 1e9, "hbm_peak_gbs": peaks["hbm_gbs"],
                  "recomputed_in_float64": st["recomputed"], "pairs": st["pairs"], "overflowed": st["overflowed"],
                  "subsample_100k_equals_float64_kernel": same}))
