#!/bin/bash
# round 2, GPU call G: view streams, select-free per-view slices, raymarch queue backward A/B
set -u
OUT=gpurun_out/r2g
mkdir -p $OUT
echo "== pytest" > $OUT/pytest.log
timeout 1500 python -m pytest tests -q -m gpu >> $OUT/pytest.log 2>&1
echo "rc=$?" >> $OUT/pytest.log
timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-decoder > $OUT/bench_head.json 2> $OUT/bench_head.err
echo "head rc=$?" >> $OUT/pytest.log
timeout 900 python bench.py --config olat --steps 3 --warmup 3 > $OUT/bench_olat_streams.json 2> $OUT/bench_olat_streams.err
echo "olat streams rc=$?" >> $OUT/pytest.log
GOLIATH_B200_VIEW_STREAMS=1 timeout 900 python bench.py --config olat --steps 3 --warmup 3 > $OUT/bench_olat_nostreams.json 2> $OUT/bench_olat_nostreams.err
echo "olat nostreams rc=$?" >> $OUT/pytest.log
GOLIATH_B200_RAYMARCH=queue-bwd timeout 900 python bench.py --config hand_mvp --steps 5 --warmup 3 > $OUT/bench_hand_mvp_queue_bwd.json 2> $OUT/bench_hand_mvp_queue_bwd.err
echo "hand_mvp queue-bwd rc=$?" >> $OUT/pytest.log
timeout 900 python bench.py --config hand_mvp --steps 5 --warmup 3 > $OUT/bench_hand_mvp.json 2> $OUT/bench_hand_mvp.err
echo "hand_mvp rc=$?" >> $OUT/pytest.log
tail -14 $OUT/pytest.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2g/bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], "ms/step %.4f"%d["ms_per_step"], "value %.1f"%d["value"], "e2e %.1f"%d["e2e"]["value"])
        if d.get("reference_extension"): print("   ", {k:round(v,3) for k,v in d["reference_extension"].items()})
    except Exception as e:
        print(f, "ERR", e); print(open(f.replace('.json','.err')).read()[-1200:])
PY
