"""Per-kernel census of the SASS mnemonics that prove a Blackwell-native kernel (B200_PROFILING.md):
UTC*MMA (tcgen05.mma), LDTM/STTM (tcgen05.ld/st), UTMALDG/UTMASTG/UBLKCP (TMA / bulk copies), MUFU.*, plus the legacy
HMMA (mma.sync) that must NOT appear.  `python tools/sass_census.py [lib.so] > profiles/r2_sass_census.txt`
(__graft_entry__.build() regenerates the file whenever it rebuilds the library)."""
import collections
import hashlib
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PATTERNS = [("UTC*MMA", r"\bUTC\w*MMA\b"), ("UTCHMMA.2CTA", r"\bUTC\w*MMA\.2CTA"), ("LDTM", r"\bLDTM\b"), ("STTM", r"\bSTTM\b"),
            ("UTMALDG", r"\bUTMALDG\b"), ("UTMASTG", r"\bUTMASTG\b"), ("UBLKCP", r"\bUBLKCP\b"), ("UBLKPF", r"\bUBLKPF\b"),
            ("MUFU.SIN", r"\bMUFU\.SIN\b"), ("MUFU.COS", r"\bMUFU\.COS\b"), ("MUFU.other", r"\bMUFU\.(?!SIN|COS)\w+"),
            ("SYNCS", r"\bSYNCS\b"), ("REDG/RED", r"\bRED(G)?\b"), ("MULTIMEM", r"\bMULTIMEM|\.MMEM|LDGMC|STGMC|REDGMC"),
            ("HMMA(legacy)", r"\bHMMA\b")]


def census(lib):
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    counts = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None or "/*" not in line:
            continue
        for name, pat in PATTERNS:
            if re.search(pat, line):
                counts[cur][name] += 1
    return counts


def demangle(names):
    try:
        out = subprocess.run(["c++filt"] + list(names), capture_output=True, text=True, check=True).stdout.splitlines()
        return dict(zip(names, out))
    except Exception:
        return {n: n for n in names}


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "reni_b200", "lib", "libreni_b200.so")
    counts = census(lib)
    names = demangle(list(counts))
    # (keyed by the hash of the kernel SOURCES, as profiles/ncu_traffic.json is: the built .so differs from build to build)
    h = hashlib.sha256()
    csrc = os.path.join(ROOT, "reni_b200", "csrc")
    for f in sorted(os.listdir(csrc)) + [os.path.join("..", "..", "include", "reni_b200.h")]:
        h.update(open(os.path.join(csrc, f), "rb").read())
    print(f"# SASS census of {os.path.relpath(lib, ROOT)} (kernel sources sha256[:16] {h.hexdigest()[:16]}); "
          "cuobjdump -sass, counts of instructions per kernel")
    cols = [n for n, _ in PATTERNS]
    print("| kernel | " + " | ".join(cols) + " |")
    print("|---|" + "---|" * len(cols))
    for fn, c in counts.items():
        if not any(c.values()):
            continue
        short = re.sub(r"\(.*", "", names[fn]).replace("void ", "").replace("reni::", "")
        print(f"| {short} | " + " | ".join(str(c.get(n, 0)) for n in cols) + " |")


if __name__ == "__main__":
    main()
