"""SASS opcode evidence for the built libraries (no GPU needed): per kernel, how many tensor-core / TMA / TMEM instructions it contains.
    python tools/sass_histogram.py [lib ...] > profiles/r02_sass_opcodes.txt
UTCHMMA = tcgen05.mma, UTMALDG / UTMASTG = TMA tensor load / store, LDTM = tcgen05.ld (TMEM -> registers), UTCBAR = tcgen05.commit,
HMMA = legacy mma.sync, FFMA2 / FADD2 / FMUL2 = packed fp32 pairs (fma / add / mul .f32x2), LDSM = ldmatrix, SYNCS = mbarrier ops, UBLKPF / CCTL.. = prefetches."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
libs = sys.argv[1:] or [os.path.join(ROOT, "yolo-sod_b200", n) for n in ("libysod.so", "libysod_f16.so")]
KEYS = ["UTCHMMA", "UTMALDG", "UTMASTG", "UTMAPF", "LDTM", "UTCBAR", "UTCATOMSWS", "HMMA", "LDSM", "SYNCS", "MUFU", "LDG", "STG", "LDS", "STS", "SHFL", "ELECT", "ACQBULK", "UTMACMDFLUSH", "FFMA2", "FADD2", "FMUL2"]
for lib in libs:
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    per = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = name.replace("(anonymous namespace)::", "").replace("void ", "")
            name = re.sub(r"\(.*", "", name)
            cur = per.setdefault(name, collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)((?:\.[A-Z0-9_]+)*)", line)
        if m and cur is not None:
            cur[m.group(1)] += 1
            cur["_total"] += 1
            if m.group(1) == "HMMA":
                cur["HMMA" + m.group(2)] += 1
    print(f"== {os.path.basename(lib)}: {len(per)} kernels")
    tot = collections.Counter()
    for c in per.values():
        tot.update(c)
    print("   library totals: " + "  ".join(f"{k}={tot[k]}" for k in KEYS if tot[k]) + "  " + "  ".join(f"{k}={v}" for k, v in sorted(tot.items()) if k.startswith("HMMA.")))
    print(f"   {'kernel':72s} {'instr':>6s}  tensor / TMA / TMEM opcodes")
    for name, c in per.items():
        hot = "  ".join(f"{k}={c[k]}" for k in KEYS[:10] if c[k])
        print(f"   {name[:72]:72s} {c['_total']:6d}  {hot}")
