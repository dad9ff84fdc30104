"""Summarise an ncu --csv launch list (gpu__time_duration, dram bytes, lts bytes, lanes per instruction) per kernel name."""
import csv, sys, collections

def main(path):
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    h = rows[0]
    iK, iM, iV, iID = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("ID")
    iU = h.index("Metric Unit")
    per = collections.OrderedDict()
    for r in rows[1:]:
        k = (r[iID], r[iK])
        v = float(r[iV].replace(",", "")) if r[iV] else 0.0
        u = r[iU]
        if r[iM].startswith("gpu__time") and u == "ns": v /= 1e6
        elif r[iM].startswith("gpu__time") and u in ("us", "usecond"): v /= 1e3
        if "bytes" in r[iM]:
            v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        per.setdefault(k, {})[r[iM]] = v
    agg = collections.OrderedDict()
    for (i, k), m in per.items():
        name = k.split("(")[0]
        a = agg.setdefault(name, [0, 0.0, 0.0, 0.0, 0.0])
        a[0] += 1; a[1] += m.get("gpu__time_duration.sum", 0); a[2] += m.get("dram__bytes_read.sum", 0) + m.get("dram__bytes_write.sum", 0)
        a[3] += m.get("lts__t_bytes.sum", 0); a[4] += m.get("smsp__thread_inst_executed_per_inst_executed.ratio", 0) * m.get("gpu__time_duration.sum", 0)
    tot = sum(a[1] for a in agg.values())
    print(f"{'kernel':58s} {'n':>3s} {'ms':>9s} {'%':>5s} {'DRAM MB':>9s} {'DRAM GB/s':>9s} {'L2 GB/s':>8s} {'lanes':>5s}")
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        ms = a[1]
        print(f"{name[:58]:58s} {a[0]:3d} {ms:9.3f} {100 * ms / tot:5.1f} {a[2] / 1e6:9.1f} {a[2] / 1e9 / (ms * 1e-3) if ms else 0:9.1f} {a[3] / 1e9 / (ms * 1e-3) if ms else 0:8.1f} {a[4] / ms if ms else 0:5.1f}")
    print(f"{'total (serialised, cold cache)':58s} {sum(a[0] for a in agg.values()):3d} {tot:9.3f}")

if __name__ == "__main__":
    main(sys.argv[1])
