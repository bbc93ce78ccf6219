"""The C-ABI library: builds, loads, exports every symbol include/boundmpc_b200.h declares, and
fails loudly (no fallback) where no CUDA device exists."""
import os
import re
import ctypes
import pytest
import torch
from boundmpc_b200 import _cabi, build as bld

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    src = open(os.path.join(ROOT, "include", "boundmpc_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bmpc_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_header_symbols():
    path = bld.build()
    assert os.path.exists(path)
    L = ctypes.CDLL(path)
    names = _declared_functions()
    assert len(names) >= 9
    for nm in names:
        assert hasattr(L, nm), f"{nm} declared in include/boundmpc_b200.h but not exported"
    assert sorted(_cabi.EXPORTS) == names


def test_sm100a_code_is_in_the_library():
    out = os.popen(f"cuobjdump -lelf {bld.build()} 2>/dev/null").read()
    assert "sm_100a" in out


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_fails_loudly_without_gpu():
    from boundmpc_b200.ocp import default_solver
    with pytest.raises(_cabi.BmpcError):
        default_solver()


def test_tolerance_option_resolution():
    """`ipopt.tol` is honoured as the reference's Ipopt honours it (BoundMPC.py:121), `b200.tol` overrides, default 1e-9;
    a controller that builds its own solver passes the reference's option dictionary."""
    from boundmpc_b200.ocp import resolve_tol
    assert resolve_tol(None) == 1e-9 and resolve_tol({}) == 1e-9
    assert resolve_tol({"ipopt": {"tol": 10e-6, "max_iter": 500}}) == 1e-5
    assert resolve_tol({"ipopt": {"tol": 10e-6}, "b200": {"tol": 1e-8}}) == 1e-8
    assert resolve_tol({"b200": {"threads": 384}}) == 1e-9
    import inspect
    from boundmpc_b200 import bound_mpc
    assert "'tol': 10e-6" in inspect.getsource(bound_mpc.BoundMPC.__init__)


def test_initial_multipliers_are_reported_as_ignored():
    """`solver(..., lam_g0=, lam_x0=)` keeps CasADi's signature; non-zero initial multipliers are not used and say so once."""
    import warnings
    import numpy as np
    from boundmpc_b200 import ocp
    ocp._DUAL_NOTE[0] = False
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        assert not ocp.note_dual_start(None, None) and not ocp.note_dual_start(0, np.zeros(4))     # the reference's initial values (BoundMPC.py:114-116)
        assert ocp.note_dual_start(np.ones(3), None) and ocp.note_dual_start(np.ones(3), None)
    assert len([x for x in w if "lam_g0" in str(x.message)]) == 1
