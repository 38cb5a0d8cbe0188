#!/bin/bash
# round 2, GPU call P (4 GPUs): what costs the head step 0.10 ms under the exchange?  cooperative rank sort vs NCCL's CTAs
set -u
OUT=gpurun_out/r2p
mkdir -p $OUT
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) \
    bench.py --gpus 4 --steps 100 --warmup 10 > $OUT/bench_$name.json 2> $OUT/bench_$name.err
  echo "$name rc=$?" >> $OUT/status.log
}
run base A=1
run passes GOLIATH_B200_RANKSORT=passes
run ctas8 NCCL_MAX_CTAS=8
run ctas8_passes NCCL_MAX_CTAS=8 GOLIATH_B200_RANKSORT=passes
cat $OUT/status.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2p/bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], "ms/step %.4f"%d["ms_per_step"], "value %.3f"%d["value"], "e2e %.3f"%d["e2e"]["value"], d["per_rank"]["collectives"])
    except Exception as e:
        print(f, "ERR", e); print(open(f.replace('.json','.err')).read()[-800:])
PY
