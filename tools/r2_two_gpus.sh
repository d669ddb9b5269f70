#!/bin/bash
# GPU box, TWO GPUs (gpurun --gpus 2): the multi-GPU tests that a one-GPU box skips, g_ray --gpus 2, and the N = 2 bench line
# with its strong-scaling records and all-reduce check.
out=gpurun_out; mkdir -p $out
nvidia-smi --query-gpu=index,name --format=csv,noheader
( timeout 900 python -m pytest tests/test_distributed.py tests/test_cli.py -m gpu -x -q -rs 2>&1 | tail -12 ) > $out/pytest_2gpu_r2.log 2>&1; cat $out/pytest_2gpu_r2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > $out/bench_n2_r2.json 2> $out/bench_n2_r2.err
tail -2 $out/bench_n2_r2.err | cut -c1-300
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n2_r2.json').read().strip().splitlines()[-1])
print('N=2: value', round(d['value'],1), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), 'allreduce_check', d.get('allreduce_check'))
for k,v in (d.get('strong') or {}).items():
    print(' strong', k, {a:(round(b,3) if isinstance(b,float) else b) for a,b in v.items() if a not in ('limiter','workload','kernel_ms_per_frame_rank0')})
PY
