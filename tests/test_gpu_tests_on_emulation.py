"""The GPU cases that were written after round 1's GPU budget was spent, executed here on the CPU: the very same test
functions, with `cake_b200.model.Context` replaced by the CPU stand-in and the library by the oracle-backed emulation
(tests/fake_b200).  Numerics are trivially exact this way; the point is that the tests' own code — shapes, call
sequences, expectations on errors, process handling — is known to be right before it meets hardware."""
import os

import pytest

from tests.cpu_ctx import CpuContext, use_emulation


@pytest.fixture
def emulated(monkeypatch, tmp_path):
    import cake_b200.model as M
    use_emulation(monkeypatch, tmp_path / "emu_lib")
    monkeypatch.setattr(M, "Context", CpuContext)
    # child processes (cake_worker) resolve libcake_b200.so through the loader path: point it at the emulation
    monkeypatch.setenv("LD_LIBRARY_PATH", str(tmp_path / "emu_lib") + os.pathsep + os.environ.get("LD_LIBRARY_PATH", ""))
    return tmp_path


def test_loader_gpu_case(emulated):
    from tests import test_loader
    d = emulated / "case"
    d.mkdir()
    test_loader.test_gpu_master_loads_from_disk(d)


def test_python_wire_worker_gpu_case(emulated):
    from tests import test_wire
    test_wire.test_gpu_worker_behind_the_wire_equals_local_blocks()


def test_master_with_tcp_worker_gpu_case(emulated):
    from tests import test_wire
    d = emulated / "case"
    d.mkdir()
    test_wire.test_gpu_master_with_tcp_worker_generates_the_same_tokens(d)


def test_cpp_wire_worker_gpu_case(emulated):
    from tests import test_wire_cpp
    d = emulated / "case"
    d.mkdir()
    test_wire_cpp.test_gpu_cpp_worker_equals_local_blocks(d)


def test_cpp_host_gpu_cases(emulated):
    from tests import test_cpp_host
    for sharded in (False, True):
        d = emulated / f"case{int(sharded)}"
        d.mkdir()
        test_cpp_host.test_cake_run_tokens_equal_python_master(d, sharded)


def _expand(fn):
    """The parameter sets of a pytest-parametrized test function, as a list of kwargs."""
    import inspect
    combos = [dict()]
    for m in [m for m in getattr(fn, "pytestmark", []) if m.name == "parametrize"]:
        names = [n.strip() for n in m.args[0].split(",")]
        combos = [dict(c, **dict(zip(names, v if len(names) > 1 else [v]))) for c in combos for v in m.args[1]]
    return combos, "monkeypatch" in inspect.signature(fn).parameters


def test_the_hardware_validated_parity_suite_still_runs(emulated):
    """tests/test_gpu_parity.py (run on a B200 in round 1) executed over the emulation: a regression net for the host
    code those tests go through (loading, forward_batch plumbing, caches, the decode loop API, repeat penalty, error
    paths) whenever it is refactored without a GPU at hand.  Numerics are exact here by construction."""
    import inspect
    from tests import test_gpu_parity as T
    ran = 0
    for name, fn in inspect.getmembers(T, inspect.isfunction):
        if not name.startswith("test_"):
            continue
        combos, wants_mp = _expand(fn)
        for kw in combos:
            mp = pytest.MonkeyPatch() if wants_mp else None
            try:
                fn(**(dict(kw, monkeypatch=mp) if wants_mp else kw))
            finally:
                if mp is not None:
                    mp.undo()
            ran += 1
    assert ran >= 20
