"""TEST INFRASTRUCTURE ONLY - pre-flight of the `-m gpu` parity suite against the CPU logic-check build: imports
tests/test_gpu_parity.py, substitutes a numpy-backed stand-in for the handful of torch calls the tests and the host
wrapper make ("device" memory is host memory in the emulation), expands the parametrisations and calls every test.
It answers "will the GPU suite's logic pass with the current sources?" before GPU minutes are spent; the parity claims
themselves come only from the real `pytest -m gpu` run on the B200.

    FDFD_B200_LIB=build/emu/libfdfd_emu.so python tests/emu/run_gpu_suite_emulated.py [-k substring]
"""
import itertools
import os
import sys
import time
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]


class FakeTensor:
    """numpy array posing as a CUDA tensor (only what operator.py and the tests touch)"""
    is_cuda = True
    device = "emu"

    def __init__(self, a):
        self.a = np.ascontiguousarray(a)

    dtype = property(lambda self: fake_torch.complex128 if self.a.dtype == np.complex128 else self.a.dtype)

    def data_ptr(self):
        return self.a.ctypes.data

    def numel(self):
        return self.a.size

    def is_contiguous(self):
        return True

    def cuda(self):
        return self

    def cpu(self):
        return self

    def numpy(self):
        return self.a

    def clone(self):
        return FakeTensor(self.a.copy())


fake_torch = types.ModuleType("torch")
fake_torch.complex128 = "complex128"
fake_torch.from_numpy = lambda a: FakeTensor(a)
fake_torch.empty_like = lambda t: FakeTensor(np.full_like(t.a, np.nan))
fake_torch.zeros_like = lambda t: FakeTensor(np.zeros_like(t.a))
fake_torch.cuda = types.SimpleNamespace(is_available=lambda: True, set_device=lambda d: None,
                                        current_stream=lambda d=None: types.SimpleNamespace(synchronize=lambda: None))
sys.modules["torch"] = fake_torch

import maxwellfdm_jl_b200 as fb                     # noqa: E402

assert "EMULATED" in fb._lib.lib().fdfd_version().decode(), "point FDFD_B200_LIB at build/emu/libfdfd_emu.so"
import test_gpu_parity as T                          # noqa: E402

SKIP = {"test_large_grid_properties": "draws its inputs with torch.randn on the device",
        "test_reduced_operator_on_device_tensors": "indexes torch tensors on the device (the same code runs on CPU tensors "
                                                   "against the emulation build: run_emu_cases.py reduced)"}
SLOW = {"test_tfsf_rhs_reproduces_the_incident_wave": "a 40^3 solve to 1e-10: half an hour of emulation (run with --slow)"}


def cases(fn):
    marks = [m for m in getattr(fn, "pytestmark", []) if m.name == "parametrize"]
    if not marks:
        return [{}]
    axes = []
    for m in marks:
        names = [n.strip() for n in m.args[0].split(",")] if isinstance(m.args[0], str) else list(m.args[0])
        vals = [v if len(names) > 1 else (v,) for v in m.args[1]]
        axes.append([dict(zip(names, v)) for v in vals])
    out = []
    for combo in itertools.product(*axes):
        d = {}
        for c in combo:
            d.update(c)
        out.append(d)
    return out


def main():
    key = sys.argv[sys.argv.index("-k") + 1] if "-k" in sys.argv else ""
    n = 0
    t00 = time.time()
    for name in sorted(dir(T)):
        fn = getattr(T, name)
        if not name.startswith("test_") or not callable(fn) or key not in name:
            continue
        if name in SKIP or (name in SLOW and "--slow" not in sys.argv):
            print(f"SKIP {name}: {SKIP.get(name) or SLOW[name]}", flush=True)
            continue
        for kw in cases(fn):
            t0 = time.time()
            fn(**kw)
            n += 1
            print(f"ok   {name}{kw if kw else ''} ({time.time() - t0:.1f}s)", flush=True)
    print(f"emulated GPU suite: {n} test cases ok in {time.time() - t00:.0f}s")


if __name__ == "__main__":
    main()
