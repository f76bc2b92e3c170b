"""CPU counterpart of tests/test_env_vs_oracle_env_gpu.py: the same side-by-side environment run with the
backend's HOST logic only (engine replaced by the oracle adapter, test infrastructure) against the oracle's
restatement of PandaPowerBackend.  Catches bookkeeping errors in apply_action / copy / reset / simulate without
a GPU."""
import pytest

from conftest import env_grid


def test_side_by_side_host_logic(monkeypatch):
    if env_grid("l2rpn_case14_sandbox") is None:
        pytest.skip("reference data not available")
    import grid2op_b200.backend as bk
    from oracle_engine import OracleEngine
    import test_env_vs_oracle_env_gpu as T

    class HostLogicBackend(bk.B200Backend):
        def _make_engine(self, gm):
            return OracleEngine(gm)

    monkeypatch.setattr(bk, "B200Backend", HostLogicBackend)
    monkeypatch.setattr(T, "WITH_THETA", True)          # angles too, incl. the reference's value for open lines (pPB:1163-1187)
    T.test_case14_sandbox_env_side_by_side(None)
