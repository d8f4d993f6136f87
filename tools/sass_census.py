"""Per-kernel census of the tensor-core / TMA SASS in libccdm_b200.so (cuobjdump -sass) plus the ptxas resource lines of the
build (csrc/build/*.ptxas.log).  Writes profiles/<name> (default r2_sass_census.txt).  Runs on the build container (no GPU).

    python tools/sass_census.py [out_name]
"""
import collections
import glob
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MNEMONICS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "UTCCP", "SYNCS", "UCGABAR", "HMMA", "MUFU.TANH", "MUFU.EX2", "FFMA2", "RED", "ATOMG"]


def main():
    out_name = sys.argv[1] if len(sys.argv) > 1 else "r2_sass_census.txt"
    so = os.path.join(ROOT, "ccdm_b200", "libccdm_b200.so")
    sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
    counts, order, cur = {}, [], None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
            cur = re.sub(r"\(.*$", "", cur)
            counts[cur] = collections.Counter()
            order.append(cur)
            continue
        if cur is None:
            continue
        m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1)
            counts[cur]["_total"] += 1
            for mn in MNEMONICS:
                if op.startswith(mn):
                    counts[cur][mn] += 1
    res = {}
    for log in sorted(glob.glob(os.path.join(ROOT, "ccdm_b200", "csrc", "build", "*.ptxas.log"))):
        txt = open(log).read()
        for m in re.finditer(r"Compiling entry function '(\S+)' for 'sm_100a'\n(?:ptxas info\s+: Function properties for \S+\n)?\s*(?:ptxas info\s+: )?(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\nptxas info\s+: Used (\d+) registers", txt):
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            res[re.sub(r"\(.*$", "", name)] = (int(m.group(5)), int(m.group(2)), int(m.group(3)), int(m.group(4)))
    lines = ["# SASS census of ccdm_b200/libccdm_b200.so (cuobjdump -sass) + ptxas -v resources; sm_100a",
             "# columns: kernel | SASS instructions | regs | stack B | spill st/ld B | " + " ".join(MNEMONICS), ""]
    tens = [k for k in order if counts[k]["UTCHMMA"] or counts[k]["UTMALDG"] or counts[k]["LDTM"]]
    for title, ks in (("tensor-core / TMA kernels", tens), ("CUDA-core kernels", [k for k in order if k not in tens])):
        lines.append(f"## {title}")
        for k in ks:
            c = counts[k]
            r = res.get(k, ("?", "?", "?", "?"))
            mn = " ".join(f"{m}={c[m]}" for m in MNEMONICS if c[m])
            lines.append(f"{k} | {c['_total']} | regs {r[0]} | stack {r[1]} | spill {r[2]}/{r[3]} | {mn}")
        lines.append("")
    path = os.path.join(ROOT, "profiles", out_name)
    open(path, "w").write("\n".join(lines))
    print(path, len(order), "kernels,", len(tens), "with tcgen05 / TMA")


if __name__ == "__main__":
    main()
