"""N-1 contingency sweep (BASELINE config 4 pattern): every single-line outage of every base state in one
launch == the same outages solved one by one by the oracle (reference pattern:
examples/backend_dependant_code/_obs_with_n1.py:111-125: _disconnect_line(i); runpf(); get_relative_flow())."""
import numpy as np
import pytest

from conftest import env_grid

from grid2op_b200.gridmodel import GridModel
from oracle.c_oracle import COracle
from test_c_oracle import random_cases

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,nbase", [("l2rpn_case14_sandbox", 48), ("rte_case5_example", 32), ("l2rpn_neurips_2020_track1", 6),
                                        ("l2rpn_wcci_2022_dev", 3)])
@pytest.mark.parametrize("policy", [1, 2])
def test_n1_sweep_matches_oracle(cuda_required, name, nbase, policy):
    from grid2op_b200.engine import PowerFlowEngine
    path = env_grid(name)
    if path is None:
        pytest.skip("reference grid files not available")
    gm = GridModel(path)
    topo, inj = random_cases(gm, nbase, seed=5, p_disc=0.0)
    eng = PowerFlowEngine(gm, max_batch=nbase * gm.n_line)
    eng.set_kernel_policy(policy)
    rho, status = eng.n1_sweep(topo, inj)
    assert (eng.plan_stats()["last_kernel"].startswith("planned")) == (policy == 2)
    # oracle: explicit records with line i out of service
    t2 = np.repeat(topo, gm.n_line, axis=0)
    i2 = np.repeat(inj, gm.n_line, axis=0)
    for b in range(nbase):
        for l in range(gm.n_line):
            t2[b * gm.n_line + l, gm.line_or_pos[l]] = -1
            t2[b * gm.n_line + l, gm.line_ex_pos[l]] = -1
    ref, rs, _, _ = COracle(gm).run(t2, i2)
    rs = rs.reshape(nbase, gm.n_line)
    assert np.array_equal(status == 0, rs == 0)
    a_or = ref[:, 3 * gm.n_line:4 * gm.n_line].reshape(nbase, gm.n_line, gm.n_line)
    want = a_or / gm.thermal_limit_a[None, None, :]
    ok = status == 0
    assert ok.sum() > ok.size // 2
    assert np.isnan(rho[~ok]).all()
    assert np.max(np.abs(rho[ok] - want[ok])) <= 2e-5 * max(1.0, float(np.max(want[ok])))
    # the outaged line carries nothing
    for l in range(gm.n_line):
        sel = ok[:, l]
        assert np.all(rho[sel, l, l] == 0.0)
    eng.close()
