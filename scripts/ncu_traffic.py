"""profiles/traffic.json from the committed ncu summaries: per kernel class the DRAM bytes (dram__bytes_read.sum +
dram__bytes_write.sum) of ONE launch, parsed from `ncu --set full` captures (scripts/gpu_r2_profile.sh -> scripts/ncu_summary.py).
bench.py reads the JSON at run time for roofline.traffic; re-run this after every new capture.

    python scripts/ncu_traffic.py            # rewrites profiles/traffic.json
"""
import csv, json, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# kernel class (bench.py's names) -> (summary csv, substring of the kernel name, note)
SOURCES = {
    "decoder_chain": ("profiles/r2/r2_prof_chain_summary.csv", "decoder_chain_kernel", "one full launch (out_proj..next in_proj), cfg3, 4096 users"),
    "attention": ("profiles/r2/r2_prof_attn_summary.csv", "pim_attn_persistent_kernel", "one full-window launch, cfg3, 4096 users x 4 heads"),
    "scorer": ("profiles/r2/r2_prof_scorer_summary.csv", "score_tc_kernel<0>", "score_tc_kernel<0> only (the re-score kernel adds profiles/r2/r2_prof_rescore_summary.csv)"),
    "gather": ("profiles/r2/r2_prof_gather_summary.csv", "embed_gather", "cfg3, 4096 users x 201 positions"),
    "topk": ("profiles/r2/r2_prof_topk_select_summary.csv", "topk_select_kernel", "topk_select_kernel only, cfg5 Caser shape (8192 x 1M x 256, k=50)"),
}
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
out = {}
for cls, (path, needle, note) in SOURCES.items():
    p = os.path.join(ROOT, path)
    if not os.path.exists(p):
        continue
    rows = list(csv.reader(open(p)))
    if len(rows) < 3:
        continue
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        if needle not in r[hdr.index("Kernel Name")]:
            continue
        tot = 0.0
        for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            i = hdr.index(k)
            tot += float(r[i].replace(",", "")) * UNIT[units[i]]
        t = hdr.index("gpu__time_duration.sum")
        out[cls] = {"bytes_per_launch": tot, "source": path, "kernel": r[hdr.index("Kernel Name")][:60], "note": note,
                    "workload": "cfg5" if cls == "topk" else "cfg3",
                    "ncu_duration": f"{r[t]} {units[t]}"}
        break
json.dump(out, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
