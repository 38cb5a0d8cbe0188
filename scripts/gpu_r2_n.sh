#!/bin/bash
# round 2, GPU call N: ranked staging with rank prefetch; ncu source-level capture of the ranked forward
set -u
OUT=gpurun_out/r2n
mkdir -p $OUT
timeout 900 python -m pytest tests/test_splat_gpu.py -q -k "ranked" > $OUT/pytest.log 2>&1
echo "rc=$?" >> $OUT/pytest.log
B="--steps 200 --warmup 10 --no-cpu-baseline --no-decoder"
timeout 600 python bench.py $B > $OUT/bench_head.json 2> $OUT/bench_head.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"blend_(bwd_mom|fwd_ilp)" -s 6 -c 4 -o $OUT/blend_ranked python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-decoder --no-graph > $OUT/ncu_full.log 2>&1
tail -4 $OUT/pytest.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2n/bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], "ms/step %.4f"%d["ms_per_step"], "value %.3f"%d["value"], "e2e %.3f"%d["e2e"]["value"])
        if d.get("roofline") and "kernels" in d["roofline"]: print("   ", {k[:30]:round(v["ms"]*1000,1) for k,v in d["roofline"]["kernels"].items()})
    except Exception as e:
        print(f, "ERR", e); print(open(f.replace('.json','.err')).read()[-1500:])
PY
