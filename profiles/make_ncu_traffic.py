#!/usr/bin/env python
"""profiles/ncu_traffic.json from `ncu --set full` captures: DRAM bytes per launch of the headline kernel and of the sampler
on the GDELT shapes (captured by the SAMPLER_NCU section of scratch/final_evidence.sh: launches in the order recent L0, recent L1, uniform L0, uniform L1),
joined with the algorithmic bytes the same run printed (bench_configs.py --config hbm_bound).

  python profiles/make_ncu_traffic.py <tag>     # reads gpurun_out/<tag>_headline.ncu-rep, <tag>_hbm_<shape>.ncu-rep / .log
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"]
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1e-3, "ms": 1.0, "ns": 1e-6, "s": 1e3, "%": 1.0}


def launches(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    names, units, data = rows[0], rows[1], rows[2:]
    col = {n: i for i, n in enumerate(names)}
    res = []
    for r in data:
        rec = {"kernel": r[col["Kernel Name"]]}
        for m in WANT:
            rec[m] = float(r[col[m]].replace(",", "")) * SCALE.get(units[col[m]], 1.0)
        res.append(rec)
    return res


def main():
    tag = sys.argv[1]
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    doc = json.load(open(path))
    h = launches(os.path.join(ROOT, "gpurun_out", tag + "_headline.ncu-rep"))[0]
    doc["sample_persistent_kernel<0, 4, 0>"] = {
        "dram_bytes": int(h["dram__bytes_read.sum"] + h["dram__bytes_write.sum"]), "read": int(h["dram__bytes_read.sum"]),
        "write": int(h["dram__bytes_write.sum"]), "ms_under_ncu": h["gpu__time_duration.sum"],
        "source": "profiles/%s_ncu_headline.txt" % tag}
    for shape in ("GDELT-16.7K", "GDELT-16.7M"):
        rep = os.path.join(ROOT, "gpurun_out", "%s_hbm_%s.ncu-rep" % (tag, shape))
        log = os.path.join(ROOT, "gpurun_out", "%s_hbm_%s.log" % (tag, shape))
        line = [l for l in open(log) if l.startswith("{")][-1]
        recs = [r for r in json.loads(line)["hbm_bound"]["launches"] if r["layer"] != "chain"]
        ls = launches(rep)
        assert len(ls) == len(recs) == 4, (len(ls), len(recs))
        out = []
        for r, l in zip(recs, ls):
            rd, wr = l["dram__bytes_read.sum"], l["dram__bytes_write.sum"]
            out.append({"strategy": r["strategy"], "layer": r["layer"], "targets": r["targets"], "neighbors": r["neighbors"],
                        "algorithmic_bytes": r["algorithmic_bytes"], "dram_read": rd, "dram_write": wr,
                        "dram_over_algorithmic": (rd + wr) / r["algorithmic_bytes"],
                        "dram_read_per_neighbor": rd / max(1, r["neighbors"]),
                        "ms_under_ncu": l["gpu__time_duration.sum"],
                        "dram_pct_of_peak": l["gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"],
                        "l1tex_pct_of_peak": l["l1tex__throughput.avg.pct_of_peak_sustained_elapsed"]})
        doc["hbm_bound:" + shape] = {"scale": json.loads(line)["hbm_bound"]["scale"], "launches": out,
                                     "source": "profiles/%s_ncu_hbm_*.txt" % tag}
    json.dump(doc, open(path, "w"), indent=1)
    for k in ("hbm_bound:GDELT-16.7K", "hbm_bound:GDELT-16.7M"):
        for l in doc[k]["launches"]:
            print(k, l["strategy"], l["layer"], "dram/alg %.2f" % l["dram_over_algorithmic"], "dram %.0f%%" % l["dram_pct_of_peak"],
                  "l1 %.0f%%" % l["l1tex_pct_of_peak"], "read/nbr %.0f B" % l["dram_read_per_neighbor"], "%.3f ms" % l["ms_under_ncu"])


if __name__ == "__main__":
    main()
