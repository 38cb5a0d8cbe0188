#!/bin/bash
# round 2, GPU call S (2 GPUs): zero-copy exchange + owner's host table on its own stream + NCCL CTA cap, against the round-1 protocol
set -u
OUT=gpurun_out/r2s
mkdir -p $OUT
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) \
    bench.py --gpus 2 --steps 100 --warmup 10 > $OUT/bench_$name.json 2> $OUT/bench_$name.err
  echo "$name rc=$?" >> $OUT/status.log
}
run inplace A=1
run copy GOLIATH_B200_EXCHANGE=copy
run inplace_ctas32 NCCL_MAX_CTAS=32
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29911 bench.py --gpus 2 --config olat --steps 5 --warmup 3 > $OUT/bench_olat_n2.json 2> $OUT/bench_olat_n2.err
cat $OUT/status.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2s/bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], "ms/step %.4f"%d["ms_per_step"], "value %.3f"%d["value"], "e2e %.3f"%d["e2e"]["value"], d["per_rank"]["collectives"])
    except Exception as e:
        print(f, "ERR", e); print(open(f.replace('.json','.err')).read()[-1200:])
PY
