"""Consistency of the committed measurement artefacts with what bench.py reads from them (CPU)."""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_traffic_table_has_the_shipped_blend_kernels():
    """bench.py: kernel_roofline takes roofline.traffic / roofline.issue_slot of the dominant blend kernel from
    profiles/r02_traffic.json by kernel name; a renamed template parameter would silently turn both into null."""
    k = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))["kernels"]
    for name in ("blend_fwd_ilp_kernel<4, 1, 0>", "blend_bwd_mom_kernel<4, 0>"):
        assert name in k, name
        assert k[name]["warp_inst"] > 1e6 and k[name]["dram_bytes"] > 1e6 and 0 < k[name]["issue_active_pct"] <= 100
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert '"blend_fwd_ilp_kernel<4, %d, 0>"' in src and '"blend_bwd_mom_kernel<4, 0>"' in src


def test_bench_lines_in_profiles_are_valid_json_with_the_contract_keys():
    need = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "dtype",
            "data", "config", "e2e", "gpu_launches", "clocks"}
    seen = 0
    for f in sorted(os.listdir(os.path.join(ROOT, "profiles"))):
        if not (f.startswith("r02_bench_") and f.endswith(".json")):
            continue
        line = open(os.path.join(ROOT, "profiles", f)).read().strip().splitlines()[-1]
        d = json.loads(line)
        if d.get("impl") == "reference":
            continue
        assert need <= set(d), (f, sorted(need - set(d)))
        assert d["value"] > 0 and d["e2e"]["value"] > 0 and "workload" in d["config"], f
        seen += 1
    assert seen >= 10
