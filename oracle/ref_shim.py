"""Import shim for the REAL reference (yoraish/mmd) -- test infrastructure only.

Only usable where /root/reference exists (the build container).  Nothing on the
GPU box may import this module; it is used by oracle/gen_golden.py to produce
the fixtures under tests/golden/ and by `-m "not gpu"` tests (skipped when the
reference tree is absent) to pin oracle/port.py against the reference itself.

Shims: matplotlib is stubbed (hot-path modules import it
at top level), sys.path gets the four un-installed package roots, cwd must be a
git work tree because mmd/datasets/trajectories.py:19 opens git.Repo('.').
"""
import os
import sys
import types

REF_ROOT = os.environ.get("MMD_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "mmd"))


class _Stub(types.ModuleType):
    def __getattr__(self, k):
        if k.startswith("__"):
            raise AttributeError(k)
        m = _Stub(self.__name__ + "." + k)
        setattr(self, k, m)
        return m

    def __call__(self, *a, **k):
        return _Stub("call")


_installed = False


def install():
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    try:
        import matplotlib  # noqa: F401
    except Exception:
        for n in ["matplotlib", "matplotlib.pyplot", "matplotlib.collections", "matplotlib.patches",
                  "matplotlib.transforms", "matplotlib.animation", "matplotlib.cm", "matplotlib.colors"]:
            sys.modules[n] = _Stub(n)
    for p in ["", "/deps/torch_robotics", "/deps/experiment_launcher", "/deps/motion_planning_baselines"]:
        if REF_ROOT + p not in sys.path:
            sys.path.insert(0, REF_ROOT + p)
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    os.chdir(repo)
    _installed = True
