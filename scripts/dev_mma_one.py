import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "np-sims_b200"), os.path.join(ROOT, "scripts")]
from dev_mma_check import run
n, w, q, k = (int(x) for x in sys.argv[1:5])
run(n, w, q, k)
