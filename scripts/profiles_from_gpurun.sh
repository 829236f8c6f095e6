#!/bin/bash
# turn the raw outputs of scripts/profile_round.sh (gpurun_out/) into the tracked summaries under profiles/ for round $1
R=${1:-r2}
set -e
grep -v "^==PROF==" gpurun_out/launches.csv > profiles/${R}_launches.csv
grep -v "^==PROF==" gpurun_out/traffic_full.csv > profiles/${R}_k_shade_wf_gi1_traffic_fullsize.csv
python scripts/ncu_summary.py gpurun_out/k_shade_wf_full.ncu-rep > profiles/${R}_k_shade_wf_gi1_ncu_full_256x256x256.txt
python scripts/ncu_phases.py gpurun_out/k_shade_wf_full.ncu-rep k_shade_wf > profiles/${R}_k_shade_wf_gi1_phases_256x256x256.txt
python scripts/ncu_phases.py gpurun_out/k_shade_wf_full.ncu-rep k_shade_wf --warpsync | awk 'NR<=2 || $5+0 >= 0.3' > profiles/${R}_k_shade_wf_gi1_phases_warpsync_256x256x256.txt
[ -f gpurun_out/k_prim_shade_full.ncu-rep ] && python scripts/ncu_summary.py gpurun_out/k_prim_shade_full.ncu-rep > profiles/${R}_k_prim_shade_ncu_full_512x512.txt || true
python - <<PY
import csv, json
rows = list(csv.DictReader(l for l in open("profiles/${R}_k_shade_wf_gi1_traffic_fullsize.csv") if l.startswith('"')))
m = {r["Metric Name"]: float(r["Metric Value"]) for r in rows}
t = json.load(open("profiles/traffic.json"))
t["k_shade_wf_gi1_dram_bytes_per_launch"] = m["dram__bytes_read.sum"] + m["dram__bytes_write.sum"]
t["k_shade_wf_gi1_l1_sector_bytes_per_launch"] = m["l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum"] * 32
t["gi1"] = {"kernel": rows[0]["Kernel Name"][:24], "dram_read": m["dram__bytes_read.sum"], "dram_write": m["dram__bytes_write.sum"],
            "lts_sectors": m["lts__t_sectors.sum"], "l1tex_global_load_sectors": m["l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum"],
            "gpu_time_duration_ms_under_ncu": m["gpu__time_duration.sum"] / 1e6,
            "source": "profiles/${R}_k_shade_wf_gi1_traffic_fullsize.csv: one k_shade_wf<1,0> launch, 512x512 / 1024 spp / global_illumination=true, first frame of the bench order (AIST frame 3)"}
json.dump(t, open("profiles/traffic.json", "w"), indent=1)
print(json.dumps(t["gi1"]))
PY
# launch-list shares
python - <<PY
import csv, collections
rows = list(csv.DictReader(l for l in open("profiles/${R}_launches.csv") if l.startswith('"')))
tot = collections.Counter(); cnt = collections.Counter()
for r in rows:
    if r["Metric Name"] == "gpu__time_duration.sum":
        k = r["Kernel Name"].split("(")[0][:40]; tot[k] += float(r["Metric Value"]); cnt[k] += 1
s = sum(tot.values())
for k, v in tot.most_common(8): print(f"{k:42s} {cnt[k]:5d} launches {v/1e6:9.2f} ms {100*v/s:5.1f} %")
PY
