#!/bin/bash
# GPU box, one GPU: DRAM bytes per launch of the dominant traversal kernel for every named scene (roofline.traffic).
# Writes gpurun_out/roofline_traffic_r02.json; copy it over profiles/roofline_traffic.json.
out=gpurun_out; mkdir -p $out
for spec in bunny:k_extend:7 spheres:k_extend:14 grid:k_extend:16 field:k_extend:16 bunny_ao:k_ao:1; do
  scene=${spec%%:*}; rest=${spec#*:}; kern=${rest%%:*}; cnt=${rest#*:}
  KSEL="regex:$kern"; [ "$kern" = "k_ao" ] && KSEL="k_ao"   # exact base name: the regex would also match k_ao_frames
  B="python bench.py --scene $scene --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-stats --no-fast-tree --tune 20,6,4,10,0,1,0"
  # skip the warm-up steps' launches (3 warm-up + 2 exclusive-pass warm-ups come before; take launches from the timed step on)
  timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k $KSEL -s $((cnt*3)) -c $cnt --csv --log-file $out/traffic_$scene.csv $B > $out/traffic_$scene.log 2>&1
done
python - <<'PY'
import csv, json, collections
res = {"note": "dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant traversal kernel (k_extend; k_ao for bunny_ao), mean over "
               "the launches of one step, ncu --clock-control none, round 2 build (pair walk, 256-bit loads), one wave lane, shadow kernel in line"}
for scene in ("bunny", "spheres", "grid", "field", "bunny_ao"):
    try:
        rows = [r for r in csv.reader(open(f"gpurun_out/traffic_{scene}.csv")) if len(r) > 10]
        hdr = rows[0]; ix = {h: i for i, h in enumerate(hdr)}
        per = collections.defaultdict(dict)
        for r in rows[1:]:
            v = float(r[ix["Metric Value"]].replace(",", ""))
            unit = r[ix["Metric Unit"]]
            scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "msecond": 1.0, "usecond": 1e-3, "nsecond": 1e-6}.get(unit, 1)
            per[r[ix["ID"]]][r[ix["Metric Name"]]] = v * scale
        tr = [p["dram__bytes_read.sum"] + p["dram__bytes_write.sum"] for p in per.values()]
        ms = [p["gpu__time_duration.sum"] for p in per.values()]
        res[scene] = int(sum(tr) / len(tr))
        res[scene + "_detail"] = {"launches": len(tr), "mean_launch_ms_under_ncu": sum(ms) / len(ms), "dram_gb_per_s": sum(tr) / (sum(ms) * 1e-3) * 1e-9}
    except Exception as e:
        res[scene + "_error"] = str(e)
json.dump(res, open("gpurun_out/roofline_traffic_r02.json", "w"), indent=1)
print(json.dumps(res)[:1500])
PY
