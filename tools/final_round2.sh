#!/bin/bash
# usage: tools/final_round2.sh <tag>  (GPU box, one GPU): the whole -m gpu suite, smoke, both bench arms, the per-scene table,
# the ncu launch list of one S1 step.  Everything for the record goes to gpurun_out/ (copy to profiles/r02/).
tag=$1; out=gpurun_out; mkdir -p $out; rm -f $out/parity_vs_reference.jsonl
( time timeout 2400 python -m pytest tests -m gpu -q -rs 2>&1 | tail -12 ) > $out/pytest_gpu_$tag.log 2>&1; tail -6 $out/pytest_gpu_$tag.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --impl reference --steps 3 --warmup 1 > $out/bench_ref_$tag.json 2> $out/bench_ref_$tag.err
python bench.py --steps 20 --warmup 3 > $out/bench_$tag.json 2> $out/bench_$tag.err || exit 1
for scene in bunny_ao spheres grid field; do
  timeout 1500 python bench.py --scene $scene --steps 4 --warmup 3 --ref-budget 6 > $out/suite_${tag}_$scene.json 2> $out/suite_${tag}_$scene.err
done
python - "$tag" <<'PY' | tee gpurun_out/final_table_$1.md
import json, sys
tag = sys.argv[1]
ref = json.loads(open(f"gpurun_out/bench_ref_{tag}.json").read().strip().splitlines()[-1])
print("| scene | CPU g_ray Msamples/s (cores) | 1xB200 Msamples/s | ms/step | e2e pipelined | e2e serial | Mrays/s | B/ray | algorithmic GB/s (frac of peak) | DRAM frac | SAH tree | kernel ms/step (each by itself) |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|")
for name, f in [("bunny", f"gpurun_out/bench_{tag}.json")] + [(s, f"gpurun_out/suite_{tag}_{s}.json") for s in ("bunny_ao", "spheres", "grid", "field")]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print("|", name, "| FAILED", e, "|"); continue
    r = d["roofline"]; c = d.get("cpu_baseline", {}); e = d["e2e"]
    k = ", ".join(f"{a} {b:.2f}" for a, b in r["kernel_ms_per_step"].items())
    fast = (d.get("fast_tree") or {}).get("value")
    print(f"| {name} | {c.get('value', 0):.2f} ({c.get('cores')}) | {d['value']:.1f} | {d['ms_per_step']:.2f} | {e['value']:.1f} | {e['serial_value']:.1f} | {d['mrays_per_s']:.0f} | "
          f"{r['algorithmic_bytes_per_ray']:.0f} | {r['achieved'] or 0:.0f} ({r['frac'] or 0:.3f}) | {r.get('dram_frac') or 0:.3f} | {fast and round(fast, 1)} | {k} |")
print()
print("reference arm:", round(ref["value"], 3), "Msamples/s,", ref["reference_step"], "; same config:", ref["config"] == json.loads(open(f"gpurun_out/bench_{tag}.json").read().strip().splitlines()[-1])["config"])
PY
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-stats --no-fast-tree"
$B > $out/plain_$tag.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 187 -c 61 --csv --log-file $out/launches_$tag.csv $B > $out/ncu_l_$tag.log 2>&1
tail -n 1 $out/ncu_l_$tag.log | cut -c1-200
