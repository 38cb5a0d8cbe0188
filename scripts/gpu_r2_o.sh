#!/bin/bash
# round 2, GPU call O: forward probe loops parked with nanosleep; ranked vs packed records again
set -u
OUT=gpurun_out/r2o
mkdir -p $OUT
timeout 1200 python -m pytest tests/test_splat_gpu.py tests/test_olat_gpu.py tests/test_fullpath_gpu.py -q > $OUT/pytest.log 2>&1
echo "rc=$?" >> $OUT/pytest.log
B="--steps 200 --warmup 10 --no-cpu-baseline --no-decoder"
timeout 600 python bench.py $B > $OUT/bench_head.json 2> $OUT/bench_head.err
GOLIATH_B200_RECORDS=packed timeout 600 python bench.py $B > $OUT/bench_head_packed.json 2> $OUT/bench_head_packed.err
timeout 900 python bench.py --config olat --steps 5 --warmup 3 > $OUT/bench_olat.json 2> $OUT/bench_olat.err
tail -4 $OUT/pytest.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2o/bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], "ms/step %.4f"%d["ms_per_step"], "value %.3f"%d["value"], "e2e %.3f"%d["e2e"]["value"])
        if d.get("roofline") and "kernels" in d["roofline"]: print("   ", {k[:30]:round(v["ms"]*1000,1) for k,v in d["roofline"]["kernels"].items()})
    except Exception as e:
        print(f, "ERR", e); print(open(f.replace('.json','.err')).read()[-1500:])
PY
