#!/bin/bash
# round 2, GPU call H (8 GPUs): the four bench configurations under torchrun, N = 8 (and N = 2, 4 for the head)
set -u
OUT=gpurun_out/r2h
mkdir -p $OUT
run() {  # n name args...
  n=$1; name=$2; shift 2
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) \
    bench.py --gpus $n "$@" > $OUT/bench_${name}_n$n.json 2> $OUT/bench_${name}_n$n.err
  echo "$name n=$n rc=$?" >> $OUT/status.log
}
nvidia-smi -L > $OUT/gpus.txt 2>&1
run 8 head --steps 100 --warmup 10
run 8 olat --config olat --steps 5 --warmup 3
run 8 hand_mvp --config hand_mvp --steps 5 --warmup 3
run 8 mvp_full --config mvp_full --steps 1 --warmup 3
run 4 head --steps 100 --warmup 10
run 2 head --steps 100 --warmup 10
cat $OUT/status.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2h/bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], "ms/step %.4f"%d["ms_per_step"], "value %.3f"%d["value"], "e2e %.3f"%d["e2e"]["value"])
        pr = d.get("per_rank")
        if pr: print("   ", {k:(v if not isinstance(v,list) else [round(x,3) if isinstance(x,float) else x for x in v]) for k,v in pr.items()})
    except Exception as e:
        print(f, "ERR", e); print(open(f.replace('.json','.err')).read()[-1500:])
PY
