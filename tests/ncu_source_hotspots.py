"""Source-line attribution of an `ncu --set full --import-source on` capture of the ring kernel (development aid).

    python tests/ncu_source_hotspots.py gpurun_out/prof.ncu-rep [whisper_medusa_b200/_lib/obj/decode.o] [--top 40]

`ncu -i <rep> --page source --csv --print-source sass` gives warp-state samples per SASS instruction; the line
table of the object file (nvdisasm -g, the build uses -lineinfo) maps instruction offsets to source lines.  Prints
the stall mix of the kernel, the source lines ranked by non-barrier samples (a warp waiting at a CTA barrier is
waiting for some other warp: the time is spent where THAT warp stalls) and the hottest single instructions.
Runs on the CPU box (no GPU needed)."""
from __future__ import annotations

import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

KERNEL = "ring_kernelILi1280ELb0"


def main():
    rep = sys.argv[1]
    obj = sys.argv[2] if len(sys.argv) > 2 and not sys.argv[2].startswith("--") else os.path.join(
        os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "whisper_medusa_b200", "_lib", "obj", "decode.o")
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]
    idx = {h: i for i, h in enumerate(hdr)}
    data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
    base = int(data[0][idx["Address"]], 16)
    with tempfile.TemporaryDirectory() as d:
        subprocess.run(f"cd {d} && cuobjdump -xelf all {os.path.abspath(obj)} > /dev/null && nvdisasm -g *.cubin > dis.txt 2>&1", shell=True, check=True)
        fn = None; file = None; line = 0; amap = {}
        for l in open(os.path.join(d, "dis.txt"), errors="ignore"):
            if l.startswith(".text."):
                fn = l
            elif "//## File" in l:
                m = re.search(r'File "([^"]*)", line (\d+)', l)
                if m:
                    file = os.path.basename(m.group(1)); line = int(m.group(2))
            elif fn and KERNEL in fn:
                m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(\S.*?);", l)
                if m:
                    amap[int(m.group(1), 16)] = (file, line)
    reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(int(r[idx["# Samples"]] or 0) for r in data)
    print(f"kernel samples {tot} over {len(data)} SASS instructions")
    agg = {h: sum(int(r[idx[h]] or 0) for r in data) for h in reasons}
    print("stall mix: " + ", ".join(f"{k[6:]} {100 * v / tot:.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v * 200 > tot))
    byline = collections.defaultdict(collections.Counter)
    for r in data:
        off = int(r[idx["Address"]], 16) - base
        f, ln = amap.get(off, ("?", 0))
        c = byline[(f, ln)]
        c["n"] += int(r[idx["# Samples"]] or 0)
        for h in reasons:
            c[h] += int(r[idx[h]] or 0)
    nb = sorted(((c["n"] - c["stall_barrier"], f, ln, c) for (f, ln), c in byline.items()), reverse=True)
    tn = sum(x[0] for x in nb)
    print(f"\nsource lines by non-barrier samples (total {tn}, {100 * tn / tot:.0f}% of all):")
    for n, f, ln, c in nb[:top]:
        rs = ", ".join(f"{k[6:]}={v}" for k, v in c.most_common(6) if k not in ("n", "stall_barrier") and v > 0)
        print(f"  {f}:{ln:<5d} {n:6d} {100 * n / tn:5.1f}%  {rs}")
    inst = sorted(((int(r[idx["# Samples"]] or 0) - int(r[idx["stall_barrier"]] or 0), i) for i, r in enumerate(data)), reverse=True)
    print("\nhottest instructions (non-barrier samples):")
    for n, i in inst[: top // 2]:
        r = data[i]
        off = int(r[idx["Address"]], 16) - base
        t3 = sorted(((int(r[idx[h]] or 0), h[6:]) for h in reasons), reverse=True)[:2]
        print(f"  +{off:05x} {n:6d}  {r[idx['Source']].strip()[:56]:56s} {amap.get(off)}  {t3}")


if __name__ == "__main__":
    main()
