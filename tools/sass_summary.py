"""Per-kernel SASS mnemonic counts of libvft_b200.so (cuobjdump -sass): what proves the Blackwell-native paths.
tcgen05.mma -> UTC*MMA, tcgen05.ld/st -> LDTM/STTM, TMA -> UTMALDG/UTMASTG/UBLKCP, tcgen05.commit -> UTCBAR,
legacy mma.sync -> HMMA.  Usage: python tools/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "vision-ft_b200", "vft_b200", "libvft_b200.so")
PAT = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTMAPF", "HMMA", "STSM", "LDSM",
       "SYNCS", "REDG", "ATOMG", "STL", "LDL"]

def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
    kernels, cur = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.search(r"/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1)
            kernels[cur]["_total"] += 1
            for p in PAT:
                if op.startswith(p):
                    kernels[cur][p] += 1
    print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)} (sm_100a): instruction counts per kernel")
    print("# " + " ".join(f"{p:>8}" for p in ["total"] + PAT) + "  kernel")
    for name, c in kernels.items():
        d = demangle(name)
        d = re.sub(r"vft::\(anonymous namespace\)::", "", d)
        d = re.sub(r"\(.*$", "", d)
        print("  " + " ".join(f"{c.get(p if p != 'total' else '_total', 0):>8}" for p in ["total"] + PAT) + "  " + d)

main()
