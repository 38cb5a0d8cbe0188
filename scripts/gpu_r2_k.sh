#!/bin/bash
# round 2, GPU call K: regrouped SG arithmetic in the fused shade; render as two autograd nodes with the shade on a side stream
set -u
OUT=gpurun_out/r2k
mkdir -p $OUT
timeout 1200 python -m pytest tests/test_sg_gpu.py tests/test_fullpath_gpu.py tests/test_splat_gpu.py tests/test_olat_gpu.py tests/test_rgca_extra_gpu.py tests/test_heads_gpu.py -q > $OUT/pytest.log 2>&1
echo "rc=$?" >> $OUT/pytest.log
timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-decoder > $OUT/bench_head.json 2> $OUT/bench_head.err
GOLIATH_B200_SHADE_STREAM=0 timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-decoder > $OUT/bench_head_nostream.json 2> $OUT/bench_head_nostream.err
GOLIATH_B200_SHADE_STREAM=0 GOLIATH_B200_RENDER_SPLIT=0 timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-decoder > $OUT/bench_head_single.json 2> $OUT/bench_head_single.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 330 --csv --log-file $OUT/launches_head.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-decoder --no-graph > $OUT/ncu_head.log 2>&1
tail -8 $OUT/pytest.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2k/bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], "ms/step %.4f"%d["ms_per_step"], "value %.3f"%d["value"], "e2e %.3f"%d["e2e"]["value"])
        if d.get("roofline") and "kernels" in d["roofline"]: print("   ", {k[:26]:round(v["ms"]*1000,1) for k,v in d["roofline"]["kernels"].items()})
    except Exception as e:
        print(f, "ERR", e); print(open(f.replace('.json','.err')).read()[-1500:])
PY
