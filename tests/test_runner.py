"""``grid2op.Runner.Runner`` (reference Runner/runner.py: builds its own environments from ``env.get_params_for_runner()``, a new
backend object per episode batch through ``backendClass``) and ``EpisodeData`` (Episode/EpisodeData.py: the stored episode is
re-read and its observations compared) with ``B200Backend`` — the callers SURVEY.md 8(f).4 names for the wire formats.

CPU: backend host logic over the oracle adapter (test infrastructure); GPU: the CUDA engine.  Both are compared with the same
run on the oracle's restatement of PandaPowerBackend."""
import tempfile
import warnings

import numpy as np
import pytest

from conftest import env_grid

ENV = "l2rpn_case14_sandbox"
MAX_ITER = 12


def _run(backend_cls, tag, path=None):
    import grid2op
    from grid2op.Agent import RecoPowerlineAgent
    from grid2op.Runner import Runner
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        env = grid2op.make(ENV, test=True, backend=backend_cls(), _add_to_name=f"runner_{tag}")
        runner = Runner(**env.get_params_for_runner(), agentClass=RecoPowerlineAgent)
        res = runner.run(nb_episode=2, max_iter=MAX_ITER, path_save=path, env_seeds=[0, 1], agent_seeds=[0, 1], episode_id=[0, 1])
    env.close()
    return res


def _compare(backend_cls, tag):
    from oracle.ppbackend_ref import PandaPowerBackendRef as OraclePandaPowerBackend
    with tempfile.TemporaryDirectory() as d1:
        ours = _run(backend_cls, tag, path=d1)
        ref = _run(OraclePandaPowerBackend, tag + "ref")
        assert len(ours) == len(ref) == 2
        for a, b in zip(ours, ref):
            # (chronics folder, name, cumulative reward, steps survived, max steps)
            assert a[1] == b[1] and a[3] == b[3] == MAX_ITER and a[4] == b[4]
            assert abs(a[2] - b[2]) <= 1e-3 * max(1.0, abs(b[2])), (a[2], b[2])
        # the stored episode reads back (EpisodeData.from_disk) with the observations the backend produced
        from grid2op.Episode import EpisodeData
        ep = EpisodeData.from_disk(d1, ours[0][1])
        assert len(ep.observations) == MAX_ITER + 1
        rho = np.array([o.rho for o in ep.observations])
        assert np.isfinite(rho).all() and (rho > 0).any()
        assert np.isclose(float(np.sum(ep.rewards[:MAX_ITER])), ours[0][2], rtol=1e-5)


def test_runner_host_logic():
    if env_grid(ENV) is None:
        pytest.skip("reference data not available")
    import grid2op_b200.backend as bk
    from oracle_engine import OracleEngine

    class HostLogicBackend(bk.B200Backend):
        def _make_engine(self, gm):
            return OracleEngine(gm)

    _compare(HostLogicBackend, "cpu")


@pytest.mark.gpu
def test_runner_gpu(cuda_required):
    if env_grid(ENV) is None:
        pytest.skip("reference data not available")
    from grid2op_b200.backend import B200Backend
    _compare(B200Backend, "gpu")
