"""CPU-side checks of the boundary: the shared library loads, exports every symbol the header
declares, the ctypes mirrors match the C structs, and the product path fails loudly (no CPU
fallback, no oracle import)."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADERS = [os.path.join(ROOT, "include", f) for f in sorted(os.listdir(os.path.join(ROOT, "include"))) if f.endswith(".h")]


def _declared_functions():
    names = []
    for h in HEADERS:
        txt = re.sub(r"/\*.*?\*/", "", open(h).read(), flags=re.S)
        names += re.findall(r"\b(dronecu_[a-z0-9_]+)\s*\(", txt)
    return sorted(set(names))


@pytest.fixture(scope="module")
def lib():
    from drone_rl_b200 import build, _lib
    build.build()
    return _lib.load()


def test_every_declared_symbol_is_exported(lib):
    from drone_rl_b200 import _lib
    declared = _declared_functions()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/*.h but not exported"
    # ... and the Python binding types every one of them
    assert sorted(_lib.declared_symbols()) == declared


def test_struct_mirrors_match_header_sizes(lib, tmp_path):
    """Compile a tiny C program against include/dronecu.h and compare sizeof() with ctypes."""
    from drone_rl_b200 import _lib
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "dronecu.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu\\n",'
                   "sizeof(dronecu_config),sizeof(dronecu_rollout_out),sizeof(dronecu_step_out),"
                   "sizeof(dronecu_state_view),sizeof(dronecu_stats),sizeof(dronecu_policy_out),sizeof(dronecu_ppo_config));return 0;}\n")
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    sizes = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    assert sizes[:7] == [C.sizeof(_lib.Config), C.sizeof(_lib.RolloutOut), C.sizeof(_lib.StepOut),
                         C.sizeof(_lib.StateView), C.sizeof(_lib.Stats), C.sizeof(_lib.PolicyOut), C.sizeof(_lib.PPOConfig)]


def test_reference_defaults(lib):
    from drone_rl_b200 import _lib, EnvConfig
    c = _lib.Config()
    lib.dronecu_config_single(C.byref(c))
    assert (c.dt, c.mass, c.gravity, c.max_steps, c.obs_dim, c.bonus_radius) == (0.02, 1.0, 9.81, 200, 15, 0.05)
    assert list(c.inertia) == [0.005, 0.005, 0.01] and c.curriculum_period == 2000 and c.flags == 3
    py = EnvConfig.single().to_c()
    assert bytes(py) == bytes(c)
    lib.dronecu_config_vector(C.byref(c))
    assert (c.max_steps, c.obs_dim, c.bonus_radius, c.flags) == (1000, 12, 1.0, 0)
    assert list(c.fixed_target) == [0.0, 0.0, 10.0] and list(c.fixed_start) == [0.1, 0.1, 0.1]
    assert bytes(EnvConfig.vector().to_c()) == bytes(c)


def test_no_cpu_fallback_and_no_oracle_in_product():
    import torch
    import drone_rl_b200
    if not torch.cuda.is_available():
        with pytest.raises(drone_rl_b200.DronecuError):
            drone_rl_b200.DroneBatch(4)
        # the C ABI itself refuses as well
        from drone_rl_b200 import _lib
        lib = _lib.load()
        cfg = drone_rl_b200.EnvConfig.single().to_c()
        h = C.c_void_p()
        assert lib.dronecu_create(C.byref(cfg), 0, 4, 0, 0, C.byref(h)) == -2
        assert b"no CUDA device" in lib.dronecu_last_error()
    pkg = os.path.join(ROOT, "drone_rl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f"{f} imports the oracle"
