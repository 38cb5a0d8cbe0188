#!/bin/bash
# round 2, GPU call D: photo-loss tests, launch lists + ncu captures of the round-2 kernels
set -u
OUT=gpurun_out/r2d
mkdir -p $OUT
echo "== pytest photo loss" > $OUT/pytest.log
timeout 600 python -m pytest tests/test_photo_loss_gpu.py -q -m gpu >> $OUT/pytest.log 2>&1
echo "rc=$?" >> $OUT/pytest.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/launches_head.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-decoder --no-graph > $OUT/ncu_head.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"blend_(bwd_mom|fwd_ilp)|rank_sort_coop|tile_sort_pack|tile_scatter|depth_keys" -s 12 -c 6 -o $OUT/head_kernels python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-decoder --no-graph > $OUT/ncu_full.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 1500 --csv --log-file $OUT/launches_olat.csv python bench.py --config olat --steps 1 --warmup 3 --no-graph > $OUT/ncu_olat.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_hand_mvp.csv python bench.py --config hand_mvp --steps 1 --warmup 3 > $OUT/ncu_hand.log 2>&1
tail -8 $OUT/pytest.log
