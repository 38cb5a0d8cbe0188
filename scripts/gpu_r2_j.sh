#!/bin/bash
# round 2, GPU call J: forward hit list per stage + scalar colour buffer in the backward; A/B against the rounds variant
set -u
OUT=gpurun_out/r2j
mkdir -p $OUT
timeout 900 python -m pytest tests/test_splat_gpu.py tests/test_fullpath_gpu.py tests/test_olat_gpu.py -q -x > $OUT/pytest.log 2>&1
echo "rc=$?" >> $OUT/pytest.log
timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-decoder > $OUT/bench_head_list.json 2> $OUT/bench_head_list.err
GOLIATH_B200_BLEND_FWD=rounds timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-decoder > $OUT/bench_head_rounds.json 2> $OUT/bench_head_rounds.err
timeout 900 python bench.py --config olat --steps 5 --warmup 3 > $OUT/bench_olat.json 2> $OUT/bench_olat.err
timeout 900 python bench.py --decoder-library > $OUT/decoder_library.json 2> $OUT/decoder_library.err
tail -5 $OUT/pytest.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2j/bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], "ms/step %.4f"%d["ms_per_step"], "value %.3f"%d["value"], "e2e %.3f"%d["e2e"]["value"])
        if d.get("roofline") and "kernels" in d["roofline"]: print("   ", {k[:26]:round(v["ms"]*1000,1) for k,v in d["roofline"]["kernels"].items()})
    except Exception as e:
        print(f, "ERR", e); print(open(f.replace('.json','.err')).read()[-1200:])
PY
cat $OUT/decoder_library.json | tail -2
