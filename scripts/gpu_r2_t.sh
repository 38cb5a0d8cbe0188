#!/bin/bash
# round 2, GPU call T (8 GPUs): head step with the zero-copy exchange; NCCL CTA cap variants at N = 8; N = 4; OLAT N = 8
set -u
OUT=gpurun_out/r2t
mkdir -p $OUT
run() {  # n name env... -- args...
  n=$1; name=$2; shift 2
  envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) \
    bench.py --gpus $n "$@" > $OUT/bench_${name}.json 2> $OUT/bench_${name}.err
  echo "$name rc=$?" >> $OUT/status.log
}
run 8 head_n8 A=1 -- --steps 100 --warmup 10
run 8 head_n8_ctas16 NCCL_MAX_CTAS=16 -- --steps 100 --warmup 10
run 8 head_n8_ctas4 NCCL_MAX_CTAS=4 -- --steps 100 --warmup 10
run 4 head_n4 A=1 -- --steps 100 --warmup 10
run 8 olat_n8 A=1 -- --config olat --steps 5 --warmup 3
run 8 hand_mvp_n8 A=1 -- --config hand_mvp --steps 5 --warmup 3
cat $OUT/status.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2t/bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], "ms/step %.4f"%d["ms_per_step"], "value %.3f"%d["value"], "e2e %.3f"%d["e2e"]["value"], (d.get("per_rank") or {}).get("collectives"))
    except Exception as e:
        print(f, "ERR", e); print(open(f.replace('.json','.err')).read()[-1200:])
PY
