#!/bin/bash
# usage: tools/final_round.sh <tag>  (GPU box, one GPU): smoke, the two bench arms, the per-scene table, the launch list
tag=$1; out=gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --impl reference --steps 2 --warmup 1 > $out/bench_ref_$tag.json 2> $out/bench_ref_$tag.err; tail -c 600 $out/bench_ref_$tag.json; echo
python bench.py --steps 10 --warmup 3 > $out/bench_$tag.json 2> $out/bench_$tag.err || exit 1
for scene in bunny_ao spheres grid field; do
  python bench.py --scene $scene --steps 3 --warmup 3 --ref-budget 6 > $out/suite_${tag}_$scene.json 2> $out/suite_${tag}_$scene.err
done
python - "$tag" <<'PY'
import json, sys
tag = sys.argv[1]
for name, f in [("bunny", f"gpurun_out/bench_{tag}.json")] + [(s, f"gpurun_out/suite_{tag}_{s}.json") for s in ("bunny_ao", "spheres", "grid", "field")]:
    try:
        d = json.load(open(f))
    except Exception as e:
        print(name, "FAILED", e); continue
    r = d["roofline"]; c = d.get("cpu_baseline", {})
    print(name, "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "Mrays/s", round(d["mrays_per_s"], 1), "rays/sample", round(d["rays_per_sample"], 2),
          "B/ray", round(r["algorithmic_bytes_per_ray"]), "GB/s", round(r["achieved"] or 0), "frac", round(r["frac"] or 0, 3), "fast", round((d.get("fast_tree") or {}).get("value") or 0, 1),
          "cpu", c.get("value"), c.get("cores"), "ms", {k: round(v, 2) for k, v in r["kernel_ms_per_step"].items()})
PY
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-stats"
$B > $out/plain_$tag.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 111 -c 40 --csv --log-file $out/launches_$tag.csv $B > $out/ncu_l_$tag.log 2>&1
tail -n 1 $out/ncu_l_$tag.log
