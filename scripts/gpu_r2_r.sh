#!/bin/bash
# round 2, GPU call R: wide weight-gradient kernel with conflict-free plane strides and two full waves
set -u
OUT=gpurun_out/r2r
mkdir -p $OUT
timeout 900 python -m pytest tests/test_nn_gpu.py tests/test_hand_mvp_gpu.py -q > $OUT/pytest.log 2>&1
echo "rc=$?" >> $OUT/pytest.log
timeout 300 python scripts/profile_tower_bwd.py > $OUT/tower_wide.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_tower_bwd.csv python scripts/profile_tower_bwd.py > $OUT/ncu_tower.log 2>&1
timeout 900 python bench.py --config hand_mvp --steps 5 --warmup 3 > $OUT/bench_hand_mvp.json 2> $OUT/bench_hand_mvp.err
tail -3 $OUT/pytest.log; cat $OUT/tower_wide.txt
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r2r/bench_hand_mvp.json").read().strip().splitlines()[-1])
    print("hand_mvp ms/step %.3f value %.2f" % (d["ms_per_step"], d["value"]))
except Exception as e:
    print("hand_mvp ERR", e); print(open("gpurun_out/r2r/bench_hand_mvp.err").read()[-800:])
PY
