"""profiles/r<N>_traffic.json from an `ncu --set full` capture of one dense-mode expansion:
    ncu --set full --clock-control none --import-source on -k regex:"k_numeric|k_symbolic" -s <skip> -c 6 \\
        -o gpurun_out/expand python bench.py --steps 4 --warmup 3 --dense --no-cpu-baseline --no-extras
    ncu -i gpurun_out/expand.ncu-rep --page raw --csv > profiles/<name>.csv
    python scripts/ncu_traffic.py profiles/<name>.csv <batches per step> > profiles/r2_traffic.json
One expansion = k_symbolic + k_numeric per trie depth; bench.py reports dram bytes per rl_expand_level call."""
import csv, json, sys

path, batches = sys.argv[1], int(sys.argv[2])
rows = list(csv.reader(open(path)))
hdr = rows[0]
col = {h: i for i, h in enumerate(hdr)}
pick = {"duration_us": "gpu__time_duration.sum", "dram_read_bytes": "dram__bytes_read.sum",
        "dram_write_bytes": "dram__bytes_write.sum", "dram_throughput_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "issue_active_pct": "sm__inst_issued.avg.pct_of_peak_sustained_active" if "sm__inst_issued.avg.pct_of_peak_sustained_active" in col
        else "smsp__issue_active.avg.pct", "warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
        "registers": "launch__registers_per_thread", "grid": "launch__grid_size"}
units = rows[1]
scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6}
launches = []
for r in rows[2:]:
    if len(r) != len(hdr):
        continue
    d = {"kernel": r[col["Kernel Name"]].split("(")[0]}
    for k, m in pick.items():
        if m in col:
            v = float(r[col[m]].replace(",", "") or 0)
            d[k] = v * scale.get(units[col[m]], 1.0) if ("bytes" in k or k == "duration_us") else v
    launches.append(d)
tot = sum(l.get("dram_read_bytes", 0) + l.get("dram_write_bytes", 0) for l in launches)
n_levels = sum(1 for l in launches if "k_numeric" in l["kernel"])
print(json.dumps({"source": "%s (ncu --set full --clock-control none, bench.py --dense --batches %d, one expansion = %d launches)"
                  % (path, batches, len(launches)), "batches_per_step": batches, "launches": launches,
                  "dram_bytes_per_expansion": tot, "dram_bytes_per_launch": tot / max(1, n_levels),
                  "launch_definition": "one rl_expand_level call = k_symbolic + k_numeric of one trie depth"}, indent=1))
