#!/bin/bash
# round 2, GPU call A: new blend kernels (mode 3) — sanitizer on a small case, parity tests, A/B bench, ncu.
set -u
mkdir -p gpurun_out
OUT=gpurun_out/r2a
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $OUT/smi.txt 2>&1
echo "== sanitizer memcheck" > $OUT/sanitizer.log
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_splat_gpu.py -q -x -k "packed and mom and (dense96 or ragged)" >> $OUT/sanitizer.log 2>&1
echo "rc=$?" >> $OUT/sanitizer.log
echo "== sanitizer racecheck" >> $OUT/sanitizer.log
timeout 600 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_splat_gpu.py -q -x -k "packed and mom and dense96 and 4" >> $OUT/sanitizer.log 2>&1
echo "rc=$?" >> $OUT/sanitizer.log
echo "== pytest splat" > $OUT/pytest.log
timeout 900 python -m pytest tests/test_splat_gpu.py -q -m gpu -k "packed or sync_free or render" >> $OUT/pytest.log 2>&1
echo "rc=$?" >> $OUT/pytest.log
echo "== pytest fullpath" >> $OUT/pytest.log
timeout 900 python -m pytest tests/test_fullpath_gpu.py -q -m gpu >> $OUT/pytest.log 2>&1
echo "rc=$?" >> $OUT/pytest.log
for mode in pipe mom; do
  GOLIATH_B200_BLEND=$mode timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > $OUT/bench_$mode.json 2> $OUT/bench_$mode.err
  echo "bench $mode rc=$?" >> $OUT/pytest.log
done
# launch list of one step (shares) and a full capture of the two blend kernels
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > $OUT/ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"blend_(bwd_mom|fwd_ilp)" -s 4 -c 2 -o $OUT/blend_mom python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > $OUT/ncu_full.log 2>&1
echo done >> $OUT/pytest.log
tail -5 $OUT/sanitizer.log; tail -30 $OUT/pytest.log; cat $OUT/bench_pipe.json | head -c 600; echo; cat $OUT/bench_mom.json | head -c 600
