import sys, time, numpy as np
sys.path.insert(0, '.')
import sgufp_solver_b200 as sg
from sgufp_solver_b200 import instances as I
from oracle.oracle import OracleNet

def run(inst, K, seed, unm):
    onet = OracleNet(inst)
    gs = sg.GuroSolver(inst)
    assert (gs.layer_arc == onet.layer_arc).all() and gs.T == onet.T
    paths = I.random_paths(onet, K, seed, unm)
    t = time.time(); res = gs.solve_paths(paths); dt = time.time() - t
    bad = 0
    for k in range(K):
        oc = onet.solve_path(paths[k])
        ok = (oc.cut_type == res.cut_type[k])
        if oc.cut_type == 0:
            ok &= bool((oc.status == res.status[k]).all()) and bool((oc.obj == res.obj[k]).all())
        ok &= oc.first_infeasible == res.first_infeasible[k]
        div = 1.0 if oc.cut_type else inst.S
        ok &= abs(oc.isum[0]/div - res.rhs[k]) <= 1e-12*max(1,abs(res.rhs[k]))
        ok &= bool(np.allclose(oc.isum[1:]/div, res.coef_dense[k], rtol=1e-12, atol=0))
        ok &= bool(np.allclose(oc.coef_dense, res.coef_dense[k], rtol=1e-9, atol=1e-9*max(1,np.abs(oc.coef_dense).max())))
        if not ok:
            bad += 1
            print('MISMATCH', inst.name, k, oc.cut_type, res.cut_type[k], oc.rhs, res.rhs[k], oc.first_infeasible, res.first_infeasible[k],
                  np.abs(oc.coef_dense-res.coef_dense[k]).max(), (oc.status != res.status[k]).sum(), (oc.obj != res.obj[k]).sum() if oc.cut_type==0 else -1)
    print(inst.name, 'S', inst.S, 'K', K, 'bad', bad, 'gpu call s', round(dt,4), 'stats', gs.last_stats())
    return bad

bad = 0
bad += run(I.config1(S=50), 8, 1, 0.15)
bad += run(I.config1(S=50, lower_prob=0.3), 8, 2, 0.3)
bad += run(I.config2(S=200), 6, 3, 0.1)
bad += run(I.config2(S=200, lower_prob=0.05), 6, 4, 0.3)
bad += run(I.config4(S=64), 3, 5, 0.1)
bad += run(I.config4(S=64, lower_prob=0.02), 3, 6, 0.3)
print('TOTAL BAD', bad)
sys.exit(1 if bad else 0)
