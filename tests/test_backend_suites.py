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

from conftest import have_cuda  # noqa: F401

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

    def _data_test_fixture_dir():
        """The located grid2op has no ``data_test`` next to it (pip-installed reference, e.g. on the GPU box): unpack the
        committed fixtures (tests/golden/data_test.tar.gz, the files the extended suites open, made by make_golden.py)
        and point the reference's ``helper_path_test`` constants at them BEFORE the suites import them."""
        import tarfile
        import tempfile
        gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "data_test.tar.gz")
        if not os.path.exists(gold):
            return None
        dst = os.path.join(tempfile.gettempdir(), f"grid2op_b200_data_test_{os.getuid()}")
        if not os.path.exists(os.path.join(dst, "data_test", "test_PandaPower", "test_case14.json")):
            with tarfile.open(gold) as tar:
                tar.extractall(dst, filter="data")
        return os.path.join(dst, "data_test")

    _extended = os.path.isdir(os.path.join(os.path.dirname(grid2op.__file__), "data_test"))
    if not _extended:
        _fix = _data_test_fixture_dir()
        if _fix is not None:
            # `import grid2op` has already imported grid2op.tests.* (grid2op/__init__.py pulls in _create_test_suite), so the
            # path constants are re-pointed in every module that holds a copy
            _old = os.path.join(os.path.dirname(os.path.abspath(grid2op.__file__)), "data_test")
            for _name, _mod in list(sys.modules.items()):
                if not _name.startswith("grid2op.tests") or _mod is None:
                    continue
                for _k, _v in list(vars(_mod).items()):
                    if isinstance(_v, str) and _v.startswith(_old):
                        setattr(_mod, _k, _fix + _v[len(_old):])
            _extended = True
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

    def _mk_ref(self, detailed_infos_for_cascading_failures=False):
        # the oracle's restatement of PandaPowerBackend itself (oracle/ppbackend_ref.py): pins the ORACLE to the reference's
        # golden vectors, independent of B200Backend's host logic
        from oracle.ppbackend_ref import PandaPowerBackendRef
        return PandaPowerBackendRef(detailed_infos_for_cascading_failures=detailed_infos_for_cascading_failures)

    _build(_mk_ref, "oracleref", [pytest.mark.filterwarnings("ignore")])
    _build(_mk_cpu, "hostlogic", [pytest.mark.filterwarnings("ignore")])
    _build(_mk_gpu, "b200", [pytest.mark.gpu, pytest.mark.filterwarnings("ignore")])
else:
    def test_reference_suites_unavailable():
        pytest.skip("reference test-suites (grid2op/tests) not available")
