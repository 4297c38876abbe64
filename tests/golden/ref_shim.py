"""Import the UNMODIFIED reference (read-only, /root/reference) in this container.

Used only by tests/golden/gen_golden.py to produce committed fixtures; nothing on the GPU
box imports this (the reference does not travel).  Two shims:
  * package alias ``bsr`` -> /root/reference/codes (bsr_class.py:10-12 uses absolute imports)
  * a stub ``matplotlib.pyplot`` (bsr_class.py:22 imports it, never uses it)
plus a value-level RNG recorder: every draw the reference makes is appended, in call order,
to a tape of doubles (uniform -> u, randint -> int, choice -> index, norm.rvs -> value,
invgamma.rvs -> value).
"""
import importlib.util
import os
import sys
import types

import numpy as np

REF_ROOT = os.environ.get("BSR_REFERENCE_ROOT", "/root/reference")


def load_reference():
    """Return (funcs_module, bsr_class_module) of the unmodified reference."""
    if "bsr.funcs" in sys.modules:
        return sys.modules["bsr.funcs"], sys.modules["bsr.bsr_class"]
    codes = os.path.join(REF_ROOT, "codes")
    if not os.path.isdir(codes):
        raise RuntimeError("reference not present at %s" % codes)
    if "matplotlib" not in sys.modules:
        try:
            import matplotlib.pyplot  # noqa: F401
        except Exception:
            mpl = types.ModuleType("matplotlib")
            plt = types.ModuleType("matplotlib.pyplot")
            mpl.pyplot = plt
            sys.modules["matplotlib"] = mpl
            sys.modules["matplotlib.pyplot"] = plt
    pkg = types.ModuleType("bsr")
    pkg.__path__ = [codes]
    sys.modules["bsr"] = pkg
    mods = []
    for name in ("funcs", "bsr_class"):
        spec = importlib.util.spec_from_file_location("bsr." + name, os.path.join(codes, name + ".py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules["bsr." + name] = mod
        spec.loader.exec_module(mod)
        setattr(pkg, name, mod)
        mods.append(mod)
    return mods[0], mods[1]


class _DistProxy:
    """Wraps scipy.stats norm / invgamma: same calls, records what .rvs returns."""

    def __init__(self, dist, tape):
        self._dist = dist
        self._tape = tape

    def rvs(self, *a, **k):
        v = self._dist.rvs(*a, **k)
        self._tape.append(float(v))
        return v

    def __getattr__(self, name):
        return getattr(self._dist, name)


class TapeRecorder:
    """Context manager: while active, all reference draws are appended to ``self.tape``."""

    def __init__(self):
        self.tape = []

    def __enter__(self):
        funcs, cls = load_reference()
        self._funcs, self._cls = funcs, cls
        self._saved = dict(
            uniform=np.random.uniform, randint=np.random.randint, choice=np.random.choice,
            f_norm=funcs.norm, f_ig=funcs.invgamma, c_norm=cls.norm, c_ig=cls.invgamma)
        tape = self.tape
        o_uniform, o_randint, o_choice = np.random.uniform, np.random.randint, np.random.choice

        def uniform(*a, **k):
            v = o_uniform(*a, **k)
            tape.extend(np.atleast_1d(v).astype(float).tolist())
            return v

        def randint(*a, **k):
            v = o_randint(*a, **k)
            tape.extend(np.atleast_1d(v).astype(float).tolist())
            return v

        def choice(*a, **k):
            v = o_choice(*a, **k)
            tape.append(float(v))
            return v

        np.random.uniform, np.random.randint, np.random.choice = uniform, randint, choice
        funcs.norm = _DistProxy(self._saved["f_norm"], tape)
        funcs.invgamma = _DistProxy(self._saved["f_ig"], tape)
        cls.norm = _DistProxy(self._saved["c_norm"], tape)
        cls.invgamma = _DistProxy(self._saved["c_ig"], tape)
        return self

    def __exit__(self, *exc):
        s = self._saved
        np.random.uniform, np.random.randint, np.random.choice = s["uniform"], s["randint"], s["choice"]
        self._funcs.norm, self._funcs.invgamma = s["f_norm"], s["f_ig"]
        self._cls.norm, self._cls.invgamma = s["c_norm"], s["c_ig"]
        return False

    def mark(self):
        return len(self.tape)
