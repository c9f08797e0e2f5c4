"""Developer tool (GPU box): load time of an instance by the three routes:
  text   the reference's format (Network.cpp:18-51) parsed by the reference's own parser (oracle/_ref, when present) and by
         instances.read_text, then sgufp_create from the arc-major arrays;
  arrays sgufp_create from arc-major int32 arrays already in memory (upload + re-layout kernel);
  cache  sgufp_create_from_cache: scenario-major fp64 file -> pinned chunks -> HBM (page cache warm: the file was just written).
Prints one JSON line per workload.   python tools/time_cache_load.py c4 c5"""
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import sgufp_solver_b200 as sg  # noqa: E402
from sgufp_solver_b200 import instances as I  # noqa: E402


def t(fn):
    t0 = time.perf_counter()
    r = fn()
    return r, time.perf_counter() - t0


for wl in [a for a in sys.argv[1:] if a in ("c2", "c4", "c5")]:
    S = {"c2": 1000, "c4": 10000, "c5": 100000}[wl]
    inst = I.config2(S) if wl == "c2" else I.config4(S)
    line = {"workload": wl, "n": inst.n, "m": inst.m, "S": S, "tokens_in_text_form": int(inst.m * (2 + 3 * S))}
    with tempfile.TemporaryDirectory() as d:
        cpath = os.path.join(d, "inst.sgufpc")
        _, line["cache_write_s"] = t(lambda: I.save_cache(inst, cpath))
        line["cache_bytes"] = os.path.getsize(cpath)
        g, line["create_from_arrays_s"] = t(lambda: sg.GuroSolver(inst))
        paths = I.random_paths(g, 2, 5, 0.1)
        want = g.solve_paths(paths, want_obj=False, want_status=False)
        g.close()
        for rep in range(2):
            h, dt = t(lambda: sg.GuroSolver.from_cache(cpath))
            line["create_from_cache_s"] = min(dt, line.get("create_from_cache_s", 1e9))
            got = h.solve_paths(paths, want_obj=False, want_status=False)
            assert (got.rhs == want.rhs).all() and (got.coef_dense == want.coef_dense).all()
            h.close()
        line["cache_GBps"] = line["cache_bytes"] / line["create_from_cache_s"] / 1e9
        h, line["create_one_eighth_block_from_cache_s"] = t(lambda: sg.GuroSolver.from_cache(cpath, scenario_offset=S // 8 * 3, S_local=S // 8))
        h.close()
        if S <= 10000:            # the text form of C5 is 3e8 tokens (> 1 GB): measured at C4 size, linear in S
            tpath = os.path.join(d, "inst.txt")
            _, line["text_write_s"] = t(lambda: inst.write_text(tpath))
            line["text_bytes"] = os.path.getsize(tpath)
            _, line["text_parse_python_s"] = t(lambda: I.read_text(tpath))
            try:
                from oracle import ref_dd
                if ref_dd.available():
                    L = ref_dd.lib()
                    hh, line["text_parse_reference_parser_s"] = t(lambda: L.ref_net_load(tpath.encode()))
                    L.ref_net_free(hh)
            except Exception as e:  # noqa: BLE001
                line["text_parse_reference_parser_s"] = None
    print(json.dumps(line), flush=True)
