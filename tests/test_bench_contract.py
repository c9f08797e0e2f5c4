"""The driver's contract for bench.py that can be checked without a GPU: the reference arm's JSON line
(the oracle port timed on the host cores) and the helpers that turn committed ncu counts into rooflines."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_reference_arm_prints_one_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines                    # ONE JSON line on stdout, nothing else
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "impl", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["metric"] == "scenario_cut_evals_per_sec" and d["unit"] == "evals/s"
    assert d["steps"] == 1 and d["warmup"] == 0 and d["n_gpus"] == 1 and d["higher_is_better"] is True
    assert d["vs_baseline"] is None and d["data"] == "synthetic" and "workload" in d["config"]
    assert d["value"] > 0 and d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["kind"] in ("port", "reference")
    assert d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0


def test_roofline_helpers_read_the_committed_counts():
    import bench
    for wl in ("c2", "c4"):
        t = bench.measured_traffic(wl)
        assert isinstance(t, int) and t > 0
        r = bench.issue_roofline(wl, 10.0, 1965.0, 148)
        assert r["bound"] == "issue" and abs(r["peak"] - 148 * 4 * 1.965) < 1e-9
        assert abs(r["achieved"] - r["warp_instructions_per_launch"] / 10e-3 / 1e9) < 1e-6 and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
    assert bench.issue_roofline("c5", 10.0, 1965.0, 148) is None        # no committed count: no claim
    assert bench.issue_roofline("c2", 10.0, None, 148) is None           # no clock sample: no claim
