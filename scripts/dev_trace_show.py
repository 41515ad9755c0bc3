import sys, numpy as np
a = np.fromfile(sys.argv[1], dtype=np.uint64).reshape(4, 4096, 2)
ev = []
names = {0: {1: "mma0: acc free", 2: "mma0: A full", 3: "mma0: issued+committed"},
         3: {1: "mma1: acc free", 2: "mma1: A full", 3: "mma1: issued+committed"},
         1: {1: "exp: stage free", 2: "exp: stage written", 3: "exp: packed landed", 4: "exp: expanded", 5: "exp: fenced"},
         2: {1: "epi: acc ready", 2: "epi: acc released", 3: "epi: flush done"}}
for r in range(4):
    for t, e in a[r]:
        if t > 10**9 and int(e) in names[r]: ev.append((int(t), r, int(e)))
ev.sort()
t0 = ev[0][0]
lo, hi = int(sys.argv[2]), int(sys.argv[3])
for t, r, e in ev:
    if lo <= t - t0 <= hi:
        print(f"{t - t0:9d}  {'  ' * (r % 3) * 10}{names[r][e]}")
