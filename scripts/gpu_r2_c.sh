#!/bin/bash
# round 2, GPU call C: coop rank sort + OLAT path + new bench configs
set -u
OUT=gpurun_out/r2c
mkdir -p $OUT
echo "== pytest new" > $OUT/pytest.log
timeout 1200 python -m pytest tests/test_olat_gpu.py tests/test_splat_gpu.py -q -m gpu -x -k "olat or bin_tiles or shared" >> $OUT/pytest.log 2>&1
echo "rc=$?" >> $OUT/pytest.log
echo "== pytest all" >> $OUT/pytest.log
timeout 1500 python -m pytest tests -q -m gpu >> $OUT/pytest.log 2>&1
echo "rc=$?" >> $OUT/pytest.log
for rs in coop passes; do
  GOLIATH_B200_RANKSORT=$rs timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-decoder > $OUT/bench_rs_$rs.json 2> $OUT/bench_rs_$rs.err
  echo "bench ranksort $rs rc=$?" >> $OUT/pytest.log
done
( time timeout 900 python bench.py ) > $OUT/bench_default.json 2> $OUT/bench_default.err
echo "bench default rc=$?" >> $OUT/pytest.log
timeout 900 python bench.py --config olat --steps 3 --warmup 3 > $OUT/bench_olat.json 2> $OUT/bench_olat.err
echo "bench olat rc=$?" >> $OUT/pytest.log
timeout 900 python bench.py --config hand_mvp --steps 3 --warmup 3 > $OUT/bench_hand_mvp.json 2> $OUT/bench_hand_mvp.err
echo "bench hand_mvp rc=$?" >> $OUT/pytest.log
timeout 1200 python bench.py --config mvp_full --steps 1 --warmup 3 > $OUT/bench_mvp_full.json 2> $OUT/bench_mvp_full.err
echo "bench mvp_full rc=$?" >> $OUT/pytest.log
tail -25 $OUT/pytest.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2c/bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], "ms/step %.4f"%d["ms_per_step"], "value %.1f"%d["value"], "e2e %.1f"%d["e2e"]["value"], "launches", d.get("gpu_launches"))
        if d.get("roofline") and "kernels" in d["roofline"]: print("   ", {k[:26]:round(v["ms"]*1000,1) for k,v in d["roofline"]["kernels"].items()})
        if d.get("reference_extension"): print("   ", d["reference_extension"])
        if d.get("decoder"): print("   decoder", d["decoder"])
    except Exception as e:
        print(f, "ERR", e); print(open(f.replace('.json','.err')).read()[-1500:])
PY
