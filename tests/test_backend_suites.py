"""The reference's own backend test-suites (grid2op._create_test_suite.create_test_suite: the 41
AAATestBackendAPI tests + the BaseBackendTest mixins with their golden vectors) applied to
``B200Backend``.

* CPU (``not gpu``): the backend's HOST logic is exercised with the engine replaced by the fp64 oracle
  adapter (tests/oracle_engine.py) - test infrastructure only, monkeypatched from here.
* GPU (``-m gpu``): the same suites through the real CUDA engine.
The suites live in the reference tree (``grid2op/tests``); they are skipped when it is not available.
"""
import os
import sys
import unittest
import warnings

import pytest

from conftest import data_test_dir, have_cuda

from grid2op_b200._bootstrap import ensure_grid2op

_ok = ensure_grid2op()
_have_suites = False
if _ok:
    try:
        import grid2op
        _tests_dir = os.path.join(os.path.dirname(grid2op.__file__), "tests")
        _have_suites = os.path.isdir(_tests_dir) and os.path.exists(os.path.join(_tests_dir, "aaa_test_backend_interface.py"))
    except Exception:  # pragma: no cover
        _have_suites = False

if _have_suites:
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        from grid2op._create_test_suite import create_test_suite
        from grid2op_b200.backend import B200Backend
        from oracle_engine import OracleEngine

    class _HostLogicBackend(B200Backend):
        """B200Backend with the oracle adapter as engine (CPU check of the host logic only)."""

        def _make_engine(self, gm):
            return OracleEngine(gm)

    def _mk_cpu(self, detailed_infos_for_cascading_failures=False):
        return _HostLogicBackend(detailed_infos_for_cascading_failures=detailed_infos_for_cascading_failures)

    def _mk_gpu(self, detailed_infos_for_cascading_failures=False):
        return B200Backend(detailed_infos_for_cascading_failures=detailed_infos_for_cascading_failures)

    _extended = data_test_dir() is not None and os.path.isdir(os.path.join(os.path.dirname(grid2op.__file__), "data_test"))
    _skip_cls = ("TestLoadingBackendPandaPower",)   # hard-codes PandaPowerBackend() in its setUp

    def _build(make, tag, marks):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            classes = create_test_suite(make_backend_fun=make, add_name_cls=tag, add_to_module=None,
                                        extended_test=_extended)
        if classes is None:   # extended_test=False returns None: rebuild only the API suite
            from grid2op.tests.aaa_test_backend_interface import AAATestBackendAPI

            def mk(self, detailed_infos_for_cascading_failures=False):
                return make(self, detailed_infos_for_cascading_failures)

            classes = [type(f"AAATestBackendAPI_{tag}", (AAATestBackendAPI, unittest.TestCase), {"make_backend": mk})]
        for cls in classes:
            if any(cls.__name__.startswith(s) for s in _skip_cls):
                continue
            for m in marks:
                cls = m(cls)
            globals()[cls.__name__] = cls

    _build(_mk_cpu, "hostlogic", [pytest.mark.filterwarnings("ignore")])
    _build(_mk_gpu, "b200", [pytest.mark.gpu, pytest.mark.filterwarnings("ignore")])
else:
    def test_reference_suites_unavailable():
        pytest.skip("reference test-suites (grid2op/tests) not available")
