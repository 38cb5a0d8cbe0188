#!/bin/bash
# round 2, GPU call F: coop rank sort v2 (no global atomics), raymarch queue forward, OLAT multi vs single, re-tests
set -u
OUT=gpurun_out/r2f
mkdir -p $OUT
echo "== pytest" > $OUT/pytest.log
timeout 1500 python -m pytest tests -q -m gpu >> $OUT/pytest.log 2>&1
echo "rc=$?" >> $OUT/pytest.log
timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-decoder > $OUT/bench_head.json 2> $OUT/bench_head.err
echo "head rc=$?" >> $OUT/pytest.log
GOLIATH_B200_RANKSORT=passes GOLIATH_B200_TILESORT=fused timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-decoder > $OUT/bench_head_r1binning.json 2> $OUT/bench_head_r1binning.err
echo "head r1 binning rc=$?" >> $OUT/pytest.log
timeout 900 python bench.py --config olat --steps 3 --warmup 3 > $OUT/bench_olat_multi.json 2> $OUT/bench_olat_multi.err
echo "olat multi rc=$?" >> $OUT/pytest.log
GOLIATH_B200_OLAT=single timeout 900 python bench.py --config olat --steps 3 --warmup 3 > $OUT/bench_olat_single.json 2> $OUT/bench_olat_single.err
echo "olat single rc=$?" >> $OUT/pytest.log
timeout 900 python bench.py --config hand_mvp --steps 5 --warmup 3 > $OUT/bench_hand_mvp.json 2> $OUT/bench_hand_mvp.err
echo "hand_mvp rc=$?" >> $OUT/pytest.log
GOLIATH_B200_RAYMARCH=legacy timeout 900 python bench.py --config hand_mvp --steps 5 --warmup 3 > $OUT/bench_hand_mvp_legacy.json 2> $OUT/bench_hand_mvp_legacy.err
echo "hand_mvp legacy rc=$?" >> $OUT/pytest.log
timeout 1200 python bench.py --config mvp_full --steps 1 --warmup 3 > $OUT/bench_mvp_full.json 2> $OUT/bench_mvp_full.err
echo "mvp_full rc=$?" >> $OUT/pytest.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/launches_head.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-decoder --no-graph > $OUT/ncu_head.log 2>&1
tail -22 $OUT/pytest.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2f/bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], "ms/step %.4f"%d["ms_per_step"], "value %.1f"%d["value"], "e2e %.1f"%d["e2e"]["value"])
        if d.get("roofline") and "kernels" in d["roofline"]: print("   ", {k[:26]:round(v["ms"]*1000,1) for k,v in d["roofline"]["kernels"].items()})
        if d.get("reference_extension"): print("   ", {k:round(v,3) for k,v in d["reference_extension"].items()})
    except Exception as e:
        print(f, "ERR", e); print(open(f.replace('.json','.err')).read()[-1200:])
PY
