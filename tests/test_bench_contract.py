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
    # the headline workload and the config keys the GPU arm prints (the driver compares the two config objects)
    assert d["scaling"] == "strong" and d["config"]["scenarios_total"] == 100000 and d["config"]["candidates_per_step"] == 8
    for k in ("n", "m", "L", "T", "scenarios_per_gpu", "candidate_paths", "l2", "timing"):
        assert k in d["config"], k


def test_strong_scaling_data_do_not_depend_on_the_number_of_gpus():
    """The C4 / C5 instance is drawn in 8 blocks: the shards of a 2- or 4-GPU run are slices of the 1-GPU instance."""
    import numpy as np
    import bench
    from sgufp_solver_b200.distributed import shard_bounds
    S = bench.totals("c4", 1)
    full = bench.scenario_range("c4", 0, S)
    assert full.S == S and full.upper.shape == (full.m, S)
    for world in (2, 4):
        for rank in (0, world - 1):
            lo, hi = shard_bounds(S, world, rank)
            part = bench.scenario_range("c4", lo, hi)
            assert (part.upper == full.upper[:, lo:hi]).all() and (part.lower == full.lower[:, lo:hi]).all()
            assert (part.tail == full.tail).all() and (part.reward[:, 0] == full.reward[:, 0]).all()
    odd = bench.scenario_range("c4", 1000, 2777)          # a range that starts and ends inside blocks
    assert (odd.upper == full.upper[:, 1000:2777]).all()
    assert bench.totals("c2", 4) == 4000 and bench.totals("c5", 8) == 100000


def test_roofline_helpers_read_the_committed_counts():
    import bench
    for wl in ("c2", "c4", "c5"):
        t = bench.measured_traffic(wl)
        assert isinstance(t, int) and t > 0
        r = bench.issue_roofline(wl, 10.0, 1965.0, 148)
        assert r["bound"] == "issue" and abs(r["peak"] - 148 * 4 * 1.965) < 1e-9
        assert abs(r["achieved"] - r["warp_instructions_per_launch"] / 10e-3 / 1e9) < 1e-6 and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
    assert bench.issue_roofline("c9", 10.0, 1965.0, 148) is None        # no committed count: no claim
    assert bench.issue_roofline("c2", 10.0, None, 148) is None           # no clock sample: no claim
