#!/bin/bash
# GPU box, N GPUs (gpurun --gpus N): usage tools/r2_multi_gpu.sh N.  The N-GPU bench line with its strong-scaling records for the
# headline scene and for BASELINE.json's configs 4 (10 M-triangle grid) and 5 (instanced field, 3840x2160).
n=$1; out=gpurun_out; mkdir -p $out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
run() { timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $n "${@:3}" > $out/$2.json 2> $out/$2.err; tail -1 $out/$2.err | cut -c1-200; }
show() { python - "$1" <<'PY'
import json, sys
d = json.loads(open(f"gpurun_out/{sys.argv[1]}.json").read().strip().splitlines()[-1])
print(sys.argv[1], ': N', d['n_gpus'], 'value', round(d['value'], 1), 'ms', round(d['ms_per_step'], 2), 'e2e', round(d['e2e']['value'], 1), 'allreduce_check', d.get('allreduce_check'))
for k, v in (d.get('strong') or {}).items():
    if v: print('  strong', k, {a: (round(b, 3) if isinstance(b, float) else b) for a, b in v.items() if a not in ('limiter', 'workload', 'kernel_ms_per_frame_rank0', 'spp_per_gpu')})
PY
}
run 29521 bench_n${n}_r2 --steps 10 --warmup 3; show bench_n${n}_r2
run 29522 bench_n${n}_field_r2 --scene field --steps 3 --warmup 3 --no-fast-tree; show bench_n${n}_field_r2
