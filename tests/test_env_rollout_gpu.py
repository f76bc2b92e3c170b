"""Drop-in check at the env.step() level: unmodified grid2op Environment + B200Backend on the GPU replays the
reference's own recorded DoNothing rollouts (grid2op/data/rte_case5_example/_statistics, produced with
PandaPowerBackend, AC, protections ON: cascading-failure logic runs through Backend.next_grid_state).
Float32 observations must match the recording (integers bit-exact)."""
import json
import os
import warnings

import numpy as np
import pytest

from conftest import grid2op_root

pytestmark = pytest.mark.gpu

KEYS = ["p_or", "q_or", "v_or", "a_or", "p_ex", "q_ex", "v_ex", "a_ex", "prod_p", "prod_q", "prod_v", "load_p", "load_q",
        "load_v", "rho", "topo_vect", "line_status"]


@pytest.mark.parametrize("scenarios", [(0, 1, 2, 4)])
def test_case5_recorded_rollouts(cuda_required, scenarios):
    root = grid2op_root()
    stat = os.path.join(root, "data", "rte_case5_example", "_statistics") if root else None
    if stat is None or not os.path.isdir(stat):
        pytest.skip("reference rollouts not available")
    from grid2op_b200.backend import B200Backend
    import grid2op
    meta = json.load(open(os.path.join(stat, "metadata.json")))
    gold = {k: np.load(os.path.join(stat, f"obs_{k}.npz"))["data"] for k in KEYS}
    sid = np.load(os.path.join(stat, "scenario_ids.npz"))["data"].ravel().astype(int)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        env = grid2op.make("rte_case5_example", test=True, backend=B200Backend(), _add_to_name="b200stat")
    worst = {k: 0.0 for k in KEYS}
    n_cmp = 0
    for s in scenarios:
        rows = np.flatnonzero(sid == s)
        env.set_id(s)
        obs = env.reset()
        seq = [obs]
        done = False
        while not done:
            obs, _, done, _ = env.step(env.action_space())
            seq.append(obs)
        assert len(seq) == meta[str(s)]["nb_step"] == len(rows)        # game over at the very same step
        n = len(rows) - 1                                              # last recorded row repeats the previous one
        for k in KEYS:
            mine = np.stack([getattr(o, k) for o in seq[:n]])
            ref = gold[k][rows[:n]]
            if ref.dtype.kind in "bi":
                assert np.array_equal(mine, ref), k
            else:
                worst[k] = max(worst[k], float(np.max(np.abs(mine.astype(np.float64) - ref))))
        n_cmp += n
    assert n_cmp > 300
    # sn_mva = 1 here: 1e-4 p.u. = 1e-4 MW; amps / kV: float32 resolution of the recording
    for k in ("p_or", "q_or", "p_ex", "q_ex", "prod_p", "prod_q", "load_p", "load_q"):
        assert worst[k] <= 1e-4, (k, worst[k])
    for k in ("v_or", "v_ex", "prod_v", "load_v"):
        assert worst[k] <= 2e-5, (k, worst[k])
    for k in ("a_or", "a_ex"):
        assert worst[k] <= 2e-3, (k, worst[k])
    assert worst["rho"] <= 1e-6
