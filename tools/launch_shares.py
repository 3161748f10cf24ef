"""Condense an ncu launch list (--csv --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum]) into
per-kernel totals: launches, serialised time, share of the step, DRAM bytes and achieved DRAM GB/s.

    python tools/launch_shares.py gpurun_out/r2_launches.csv profiles/r2_shares_ddim_step.json
"""
import collections
import csv
import json
import sys

SCALE_T = {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "s": 1e6, "second": 1e6}
SCALE_B = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def short(name):
    n = name.split("(")[0].replace("void ", "").replace("ctrl::", "")
    return n.strip()


def main():
    src = sys.argv[1]
    out = sys.argv[2] if len(sys.argv) > 2 else None
    rows = list(csv.reader(open(src)))
    hdr = next(r for r in rows if "Kernel Name" in r)
    ix = {k: hdr.index(k) for k in ("ID", "Kernel Name", "Metric Name", "Metric Unit", "Metric Value")}
    per = collections.OrderedDict()  # launch id -> {name, us, rd, wr}
    for r in rows:
        if len(r) != len(hdr) or r is hdr or not r[ix["ID"]].isdigit():
            continue
        rec = per.setdefault(r[ix["ID"]], {"name": short(r[ix["Kernel Name"]]), "us": 0.0, "rd": 0.0, "wr": 0.0})
        try:
            v = float(r[ix["Metric Value"]].replace(",", ""))
        except ValueError:
            continue
        m, u = r[ix["Metric Name"]], r[ix["Metric Unit"]]
        if m.startswith("gpu__time_duration"):
            rec["us"] = v * SCALE_T.get(u, 1.0)
        elif m.startswith("dram__bytes_read"):
            rec["rd"] = v * SCALE_B.get(u, 1.0)
        elif m.startswith("dram__bytes_write"):
            rec["wr"] = v * SCALE_B.get(u, 1.0)
    agg = collections.OrderedDict()
    for rec in per.values():
        a = agg.setdefault(rec["name"], {"launches": 0, "us": 0.0, "dram_read_bytes": 0.0, "dram_write_bytes": 0.0})
        a["launches"] += 1
        a["us"] += rec["us"]
        a["dram_read_bytes"] += rec["rd"]
        a["dram_write_bytes"] += rec["wr"]
    total = sum(a["us"] for a in agg.values())
    table = []
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
        gbs = (a["dram_read_bytes"] + a["dram_write_bytes"]) / (a["us"] * 1e-6) / 1e9 if a["us"] else 0.0
        table.append(dict(kernel=k, share=a["us"] / total if total else 0.0, dram_gbs=gbs, **a))
        print(f"{k[:52]:52s} n={a['launches']:4d} {a['us']:9.1f} us {100 * a['us'] / total:5.1f}%  "
              f"dram {(a['dram_read_bytes'] + a['dram_write_bytes']) / 1e6:9.1f} MB  {gbs:7.1f} GB/s")
    print(f"total {total:.1f} us over {len(per)} launches")
    fam = lambda pred: {"launches": sum(t["launches"] for t in table if pred(t["kernel"])),
                        "us": sum(t["us"] for t in table if pred(t["kernel"])),
                        "dram_bytes": sum(t["dram_read_bytes"] + t["dram_write_bytes"] for t in table if pred(t["kernel"]))}
    res = {"source": src, "total_us_serialised": total, "launches": len(per), "kernels": table,
           "families": {"gemm_tcgen05": fam(lambda k: k.startswith("gemm_tcgen05")), "groupnorm": fam(lambda k: k.startswith("gn_")),
                        "attention": fam(lambda k: k.startswith("attention")), "layernorm": fam(lambda k: k.startswith("layernorm"))}}
    if out:
        json.dump(res, open(out, "w"), indent=1)


if __name__ == "__main__":
    main()
