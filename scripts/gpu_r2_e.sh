#!/bin/bash
# round 2, GPU call E (2 GPUs): new unit tests + the N=2 paths of every bench config
set -u
OUT=gpurun_out/r2e
mkdir -p $OUT
echo "== pytest new" > $OUT/pytest.log
timeout 900 python -m pytest tests/test_photo_loss_gpu.py tests/test_optim_gpu.py tests/test_geom_gpu.py tests/test_olat_gpu.py tests/test_splat_gpu.py tests/test_fullpath_gpu.py -q -m gpu >> $OUT/pytest.log 2>&1
echo "rc=$?" >> $OUT/pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus 2 --steps 100 --warmup 10 > $OUT/bench_head_n2.json 2> $OUT/bench_head_n2.err
echo "head n2 rc=$?" >> $OUT/pytest.log
timeout 600 $TR bench.py --gpus 2 --impl reference --steps 2 --warmup 1 > $OUT/bench_ref_n2.json 2> $OUT/bench_ref_n2.err
echo "ref n2 rc=$?" >> $OUT/pytest.log
timeout 900 $TR bench.py --gpus 2 --config olat --steps 3 --warmup 3 > $OUT/bench_olat_n2.json 2> $OUT/bench_olat_n2.err
echo "olat n2 rc=$?" >> $OUT/pytest.log
timeout 900 $TR bench.py --gpus 2 --config hand_mvp --steps 3 --warmup 3 > $OUT/bench_hand_mvp_n2.json 2> $OUT/bench_hand_mvp_n2.err
echo "hand_mvp n2 rc=$?" >> $OUT/pytest.log
timeout 1200 $TR bench.py --gpus 2 --config mvp_full --steps 1 --warmup 3 > $OUT/bench_mvp_full_n2.json 2> $OUT/bench_mvp_full_n2.err
echo "mvp_full n2 rc=$?" >> $OUT/pytest.log
tail -12 $OUT/pytest.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2e/bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], "n", d.get("n_gpus"), "ms/step %.4f"%d["ms_per_step"], "value %.1f"%d["value"], "e2e %.1f"%d["e2e"]["value"], d.get("per_rank"))
    except Exception as e:
        print(f, "ERR", e); print(open(f.replace('.json','.err')).read()[-1200:])
PY
