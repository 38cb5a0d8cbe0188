#!/bin/bash
# round 2, GPU call V (last minutes of the budget): ncu --set full of the shipped blend kernels, the full GPU suite, a launch list
set -u
OUT=gpurun_out/r2v
mkdir -p $OUT
timeout 110 ncu --set full --clock-control none -k regex:"blend_(bwd_mom|fwd_ilp)" -s 6 -c 2 -o $OUT/blend_final python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-decoder --no-graph > $OUT/ncu_full.log 2>&1
timeout 130 python -m pytest tests -q -x -m gpu > $OUT/pytest.log 2>&1
echo "rc=$?" >> $OUT/pytest.log
timeout 70 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches_head.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-decoder --no-graph > $OUT/ncu_head.log 2>&1
timeout 60 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-decoder > $OUT/bench_head.json 2> $OUT/bench_head.err
tail -3 $OUT/pytest.log; tail -c 600 $OUT/bench_head.json
