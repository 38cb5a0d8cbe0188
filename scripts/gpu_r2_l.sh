#!/bin/bash
# round 2, GPU call L: late colours through a fill kernel, shade on a lowest-priority stream under a high-priority capture
set -u
OUT=gpurun_out/r2l
mkdir -p $OUT
timeout 1200 python -m pytest tests/test_fullpath_gpu.py tests/test_splat_gpu.py tests/test_olat_gpu.py -q > $OUT/pytest.log 2>&1
echo "rc=$?" >> $OUT/pytest.log
B="--steps 200 --warmup 10 --no-cpu-baseline --no-decoder"
timeout 600 python bench.py $B > $OUT/bench_head.json 2> $OUT/bench_head.err
GOLIATH_B200_RENDER_PRIO=0 timeout 600 python bench.py $B > $OUT/bench_head_noprio.json 2> $OUT/bench_head_noprio.err
GOLIATH_B200_SHADE_STREAM=0 timeout 600 python bench.py $B > $OUT/bench_head_nostream.json 2> $OUT/bench_head_nostream.err
GOLIATH_B200_SHADE_STREAM=0 GOLIATH_B200_RENDER_SPLIT=0 timeout 600 python bench.py $B > $OUT/bench_head_single.json 2> $OUT/bench_head_single.err
GOLIATH_B200_SHADE_STREAM=0 GOLIATH_B200_RENDER_SPLIT=0 GOLIATH_B200_RENDER_PRIO=0 timeout 600 python bench.py $B > $OUT/bench_head_single_noprio.json 2> $OUT/bench_head_single_noprio.err
tail -4 $OUT/pytest.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2l/bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], "ms/step %.4f"%d["ms_per_step"], "value %.3f"%d["value"], "e2e %.3f"%d["e2e"]["value"])
    except Exception as e:
        print(f, "ERR", e); print(open(f.replace('.json','.err')).read()[-1500:])
PY
