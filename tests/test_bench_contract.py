"""bench.py's JSON contract, checked on the CPU through the reference arm (the only arm that
runs without a GPU): one line, the required keys, the reference-arm extras."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line(tmp_path):
    env = dict(os.environ, ESKF_BENCH_LEAD_IN="2", ESKF_BENCH_CACHE=str(tmp_path), ESKF_BENCH_DENSE_SRC="60000",
               ESKF_BENCH_DENSE_MAP="200000", ESKF_BENCH_CPU_SAMPLE="20000")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference",
                          "--steps", "2", "--warmup", "3"], capture_output=True, text=True, timeout=600,
                         env=env, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
              "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["metric"] == "mpts_per_s_per_gn_iteration" and d["unit"] == "Mpts/s"
    assert d["higher_is_better"] is True and d["scaling"] == "strong"
    assert d["vs_baseline"] is None and d["data"] == "synthetic"
    assert "workload" in d["config"] and "model" not in d["config"]
    assert d["steps"] == 2 and d["value"] > 0 and d["value"] == d["cpu_baseline"]["value"] == d["e2e"]["value"]
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    # non-zero ranks of a torchrun launch stay silent and exit 0
    out2 = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference"],
                          capture_output=True, text=True, timeout=60, env=dict(env, RANK="1", WORLD_SIZE="2"),
                          cwd=ROOT)
    assert out2.returncode == 0 and out2.stdout.strip() == ""
