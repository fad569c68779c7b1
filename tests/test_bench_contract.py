"""CPU test of bench.py's reference arm: one JSON line with the keys the driver reads (the GPU arm's line is
checked by running it on the GPU box; its key set is asserted here against the source)."""
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    env = dict(os.environ, NK_REF_SAMPLE_BASES="2000000")
    out = subprocess.check_output([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                                   "--warmup", "0"], text=True, env=env, cwd=ROOT)
    lines = [l for l in out.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "kmers/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("canonical k-mers/sec") and d["n_gpus"] == 1 and d["steps"] == 1
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert set(d["cpu_baseline"]) >= {"value", "unit", "cores", "kind", "sample"} and d["cpu_baseline"]["kind"] == "port"
    assert d["e2e"] == {"value": d["value"], "unit": "kmers/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_gpu_arm_line_has_the_contract_keys():
    src = open(os.path.join(ROOT, "bench.py")).read()
    body = src[src.index("        line = {"):src.index("        sys.stdout.flush()")]
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "roofline", "cpu_baseline", "clocks"):
        assert re.search(r'"%s":' % key, body), key
    roof = src[src.index("        roofline = {"):src.index("        ph = {")]
    for key in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert re.search(r'"%s":' % key, roof), key
