"""Join an ncu launch list (gpu__time_duration.sum, --csv) of tools/prof_forward.py with the record names it wrote
(gpurun_out/forward_recs.json): the LAST len(recs) launches are the second forward, one launch per record.

    python tools/ncu_layers.py gpurun_out/launches.csv gpurun_out/forward_recs.json profiles/rN_forward_layers_ncu.json
"""
import csv
import json
import sys


def main():
    src, recs_path, dst = sys.argv[1:4]
    rows = list(csv.reader(open(src)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    launches = []
    for r in data:
        if len(r) <= mv:
            continue
        try:
            v = float(r[mv].replace(",", ""))
        except ValueError:
            continue
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[mu].strip(), 1e-3)
        launches.append((r[kn].split("(")[0].replace("void ", ""), v * scale))
    recs = json.load(open(recs_path))
    tail = launches[-len(recs):]
    assert len(tail) == len(recs), (len(launches), len(recs))
    out, tot = [], 0.0
    for rec, (kname, us) in zip(recs, tail):
        row = dict(rec, kernel=kname, us=round(us, 2))
        if rec.get("flops"):
            row["tflops"] = round(rec["flops"] / us / 1e6, 1)
        out.append(row)
        tot += us
    json.dump({"total_us": tot, "rows": out}, open(dst, "w"), indent=1)
    print(f"{len(out)} launches, {tot:.1f} us")
    for row in out:
        print(f"{row['name']:36s} {row['kernel'][:44]:44s} {row['us']:9.1f} us {row.get('tflops', '')}")


if __name__ == "__main__":
    main()
