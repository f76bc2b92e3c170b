"""GPU variant of tests/test_env_random_agent_cpu.py: the same sampled-action runs with ``B200Backend`` THROUGH CUDA next to the oracle's
restatement of ``PandaPowerBackend``.

Written after this round's GPU budget was spent: the function has never run on hardware (its CPU counterpart, which shares the whole
body, is green).  It is therefore marked ``xfail(strict=False)`` — a pass shows up as XPASS in the driver's round-end GPU run, a
failure cannot break the suite; drop the marker once it has passed on a B200.  (The file name sorts LAST on purpose: these are the only
GPU tests that never met hardware, nothing runs after them.)"""
import pytest

from conftest import env_grid

pytestmark = pytest.mark.gpu


@pytest.mark.xfail(strict=False, reason="added after the round's GPU budget was spent: not yet run on hardware")
@pytest.mark.parametrize("name,n_steps,sn_mva,dc,with_simulate", [("l2rpn_case14_sandbox", 80, 100.0, False, True),
                                                                  ("educ_case14_storage", 60, 100.0, False, False),
                                                                  ("l2rpn_case14_sandbox", 40, 100.0, True, False),
                                                                  ("l2rpn_neurips_2020_track1", 30, 1.0, False, False)])
def test_random_agent_side_by_side_through_cuda(cuda_required, name, n_steps, sn_mva, dc, with_simulate):
    if env_grid(name) is None:
        pytest.skip("reference data not available")
    from grid2op_b200.backend import B200Backend
    from test_env_random_agent_cpu import run_side_by_side
    run_side_by_side(B200Backend, name, n_steps, sn_mva, dc, 2, with_simulate, tag="fuzzgpu")


@pytest.mark.xfail(strict=False, reason="added after the round's GPU budget was spent: not yet run on hardware")
@pytest.mark.parametrize("seed,protections", [(0, False), (2, True)])
def test_random_mixed_actions_through_cuda(cuda_required, seed, protections):
    """tests/test_batched_env.py::test_random_mixed_actions_host_logic with the CUDA engine on both sides (BatchedEnv and the unmodified
    environments on B200Backend)"""
    if env_grid("l2rpn_neurips_2020_track1") is None:
        pytest.skip("reference data not available")
    from grid2op_b200.backend import B200Backend
    from test_batched_env import run_random_mixed_actions
    run_random_mixed_actions(seed, protections, B200Backend, lambda gm: None)


@pytest.mark.xfail(strict=False, reason="added after the round's GPU budget was spent: not yet run on hardware")
def test_side_by_side_environment_with_angles_through_cuda(cuda_required, monkeypatch):
    """tests/test_env_vs_oracle_env_gpu.py with the voltage angles compared too (open lines: the angle of the bus the end was last
    attached to, pPB:1163-1187) — green on the CPU with the host logic (tests/test_env_hostlogic_cpu.py)"""
    if env_grid("l2rpn_case14_sandbox") is None:
        pytest.skip("reference data not available")
    import test_env_vs_oracle_env_gpu as T
    monkeypatch.setattr(T, "WITH_THETA", True)
    T.test_case14_sandbox_env_side_by_side(None)
