"""SASS mnemonic counts per kernel of the built library (cuobjdump -sass): the evidence that the tcgen05 / TMA / TMEM
paths are what the shipped binary contains.  No GPU needed.

    python scripts/sass_mnemonics.py > profiles/r02_sass_mnemonics.txt
"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "np-sims_b200", "lib", "libnpsims_b200.so")
KEYS = ["UTCOMMA", "UTCHMMA", "LDTM", "STTM", "UTMALDG", "UBLKCP", "UTCBAR", "UTCATOMSWS", "SYNCS", "REDUX", "POPC", "DFMA",
        "ATOMS", "ATOMG", "RED", "MEMBAR", "ERRBAR", "BAR.SYNC", "UCGABAR", "ELECT"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    archs = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
    print(f"# SASS mnemonic counts per kernel of np-sims_b200/lib/libnpsims_b200.so (cuobjdump -sass, sm_100a only), round 2 final build")
    print("# UTCOMMA = tcgen05.mma kind::mxf4 block-scaled, UTCHMMA = tcgen05.mma kind::f16/tf32, LDTM/STTM = tcgen05.ld/st, UTMALDG = TMA tensor load,")
    print("# UBLKCP = TMA 1-D bulk copy, UTCBAR = tcgen05.commit, SYNCS = mbarrier ops, REDUX = redux.sync, UCGABAR = cluster barrier")
    print(f"# architectures in the fat binary: {archs}")
    names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
    blocks = re.split(r"Function : \S+", sass)[1:]
    for name, body in zip(names, blocks):
        ops = collections.Counter()
        n = 0
        for m in re.finditer(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", body):
            n += 1
            op = m.group(1)
            for k in KEYS:
                if op == k or op.startswith(k + ".") or (k == "RED" and op.startswith("REDG")):
                    ops[k] += 1
                    break
        short = re.sub(r"\(.*", "", name)
        print(f"{short:90s} instr={n:6d}  " + " ".join(f"{k}={ops[k]}" for k in KEYS if ops[k]))


if __name__ == "__main__":
    main()
