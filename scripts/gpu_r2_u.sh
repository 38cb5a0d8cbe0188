#!/bin/bash
# round 2, GPU call U (final, 1 GPU): full GPU suite, smoke, the default bench line, the other configs, final ncu captures
set -u
OUT=gpurun_out/r2u
mkdir -p $OUT
timeout 1500 python -m pytest tests -q -m gpu > $OUT/pytest.log 2>&1
echo "rc=$?" >> $OUT/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" >> $OUT/pytest.log 2>&1
echo "smoke rc=$?" >> $OUT/pytest.log
( time timeout 900 python bench.py ) > $OUT/bench_default.json 2> $OUT/bench_default.err
timeout 600 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-decoder > $OUT/bench_head.json 2> $OUT/bench_head.err
timeout 900 python bench.py --config olat --steps 5 --warmup 3 > $OUT/bench_olat.json 2> $OUT/bench_olat.err
timeout 1200 python bench.py --config mvp_full --steps 1 --warmup 3 > $OUT/bench_mvp_full.json 2> $OUT/bench_mvp_full.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 330 --csv --log-file $OUT/launches_head.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-decoder --no-graph > $OUT/ncu_head.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"blend_(bwd_mom|fwd_ilp)|tile_sort_kernel|gather_records|tile_scatter|rank_sort_coop|depth_keys|sg_(fwd|bwd)_kernel|project_(fwd|bwd)_kernel" -s 36 -c 12 -o $OUT/head_kernels python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-decoder --no-graph > $OUT/ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:"blend_(bwd|fwd)_multi" -s 4 -c 2 -o $OUT/olat_kernels python bench.py --config olat --steps 1 --warmup 3 --no-graph > $OUT/ncu_olat.log 2>&1
tail -6 $OUT/pytest.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2u/bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], "ms/step %.4f"%d["ms_per_step"], "value %.3f"%d["value"], "e2e %.3f"%d["e2e"]["value"])
        if d.get("roofline") and "kernels" in d["roofline"]: print("   ", {k[:30]:round(v["ms"]*1000,1) for k,v in d["roofline"]["kernels"].items()}, d["roofline"].get("issue_slot"))
    except Exception as e:
        print(f, "ERR", e); print(open(f.replace('.json','.err')).read()[-1200:])
PY
