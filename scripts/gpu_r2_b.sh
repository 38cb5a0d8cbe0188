#!/bin/bash
# round 2, GPU call B: full gpu test-suite, blend A/B (pipe / mom / mom-affine), ncu of the mom-affine kernels
set -u
OUT=gpurun_out/r2b
mkdir -p $OUT
echo "== pytest all gpu" > $OUT/pytest.log
timeout 1500 python -m pytest tests -q -m gpu -x >> $OUT/pytest.log 2>&1
echo "rc=$?" >> $OUT/pytest.log
for mode in pipe mom mom-affine; do
  GOLIATH_B200_BLEND=$mode timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > $OUT/bench_$mode.json 2> $OUT/bench_$mode.err
  echo "bench $mode rc=$?" >> $OUT/pytest.log
done
GOLIATH_B200_BLEND=mom-affine timeout 900 ncu --set full --clock-control none --import-source on -k regex:"blend_(bwd_mom|fwd_ilp)" -s 4 -c 2 -o $OUT/blend_mom_affine python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > $OUT/ncu_full.log 2>&1
tail -15 $OUT/pytest.log
python - <<'PY'
import json
for m in ("pipe","mom","mom-affine"):
    try:
        d=json.loads(open("gpurun_out/r2b/bench_%s.json"%m).read().strip().splitlines()[-1])
        ks=d["roofline"]["kernels"]
        print(m, "ms/step %.4f"%d["ms_per_step"], "value %.1f"%d["value"], "e2e %.1f"%d["e2e"]["value"], {k[:24]:round(v["ms"]*1000,1) for k,v in ks.items()})
    except Exception as e:
        print(m, "ERR", e)
PY
