"""Count the Blackwell-specific SASS mnemonics in every native library (evidence for profiles/).
Usage: python scripts/sass_summary.py > profiles/sass_summary.txt"""
import collections
import pathlib
import re
import subprocess

BUILD = pathlib.Path(__file__).resolve().parent.parent / "colossalai_b200" / "kernel" / "_build"
KEYS = ["UTCHMMA", "UTCQMMA", "UTCMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "SYNCS", "HMMA",
        "LDGSTS", "LDGMC", "LDG.E.128.STRONG.SYS", "STG.E.STRONG.SYS", "LDG.E.STRONG.SYS", "RED", "ATOM", "CCTL", "FENCE", "MEMBAR"]
for so in sorted(BUILD.glob("libcb200_*.so")):
    try:
        sass = subprocess.run(["cuobjdump", "-sass", str(so)], capture_output=True, text=True, timeout=600).stdout
    except Exception as e:
        print(f"{so.name}: cuobjdump failed: {e}")
        continue
    if "Function" not in sass:
        continue
    print(f"== {so.name}")
    cur, counts = None, collections.OrderedDict()
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        for k in KEYS:
            if re.search(r"\b" + re.escape(k), line):
                counts[cur][k] += 1
    for fn, c in counts.items():
        hot = {k: v for k, v in c.items() if k not in ("ATOM", "RED", "MEMBAR", "CCTL") or v}
        if any(k in c for k in ("UTCHMMA", "UTCQMMA", "LDTM", "UTMALDG", "LDGMC", "LDG.E.128.STRONG.SYS", "HMMA")):
            short = subprocess.run(["c++filt", fn], capture_output=True, text=True).stdout.strip()[:110]
            print(f"  {short}: " + ", ".join(f"{k}={v}" for k, v in hot.items()))
