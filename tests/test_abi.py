"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/cmpc.h declares, validates constructor arguments like the reference's asserts, and
FAILS LOUDLY without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "cmpc.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cmpc_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_expected_entry_points():
    syms = declared_symbols()
    for s in ("cmpc_create", "cmpc_setup", "cmpc_update_weights", "cmpc_solve_batch", "cmpc_solve_batch_device",
              "cmpc_build_batch", "cmpc_rollout", "cmpc_destroy", "cmpc_last_error"):
        assert s in syms


def test_library_exports_every_declared_symbol(pkg):
    lib = pkg.load_library()
    for s in declared_symbols():
        assert hasattr(lib, s), f"libcmpc_b200.so does not export {s}"
    out = subprocess.run(["nm", "-D", "--defined-only", pkg.lib_path()], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (cmpc_[a-z0-9_]+)", out))
    assert set(declared_symbols()) <= exported
    assert lib.cmpc_version().decode().startswith("cmpc_b200")


def test_config_struct_layout_matches_header(pkg):
    # 8+4+4+8+32+360+4+4+8+4+4 (polish, presolve) +4+4 (qp_backend, reserved)
    assert C.sizeof(pkg.CmpcConfig) == 448
    assert pkg.CmpcConfig.presolve.offset == 436 and pkg.CmpcConfig.qp_backend.offset == 440
    assert pkg.CmpcConfig.weights.offset == 56 and pkg.CmpcConfig.ipm_tol.offset == 424
    lib = pkg.load_library()
    cfg = pkg.CmpcConfig()
    w = (C.c_double * 45)(*pkg.workloads.F1_WEIGHTS)
    mu = (C.c_double * 4)(0.8, 0.8, 0.8, 0.8)
    assert lib.cmpc_config_init(C.byref(cfg), 8.0, 4, 6, 0.01, w, mu) == 0
    ref = pkg.make_config(pkg.workloads.default_config(6))
    assert bytes(cfg) == bytes(ref)


@pytest.mark.parametrize("bad", [dict(mass=0.0), dict(num_legs=0), dict(num_legs=5), dict(horizon=0), dict(horizon=33),
                                 dict(dt=-1.0), dict(mu=[0.8, 0.8, 0.0, 0.8])])
def test_create_rejects_what_the_reference_asserts(pkg, bad):
    """CentroidalMPC.cpp:24-25: assert(mass > 0 && num_legs > 0 && predict_horizon > 0), mu.size()==num_legs."""
    cfg = pkg.workloads.default_config(6)
    cfg.update(bad)
    if "num_legs" in bad and bad["num_legs"] > 4:
        cfg["mu"] = [0.8] * 5
        cfg["weights"] = list(cfg["weights"]) + [0.1] * 9
        c = pkg.CmpcConfig(); c.mass = 8; c.num_legs = 5; c.horizon = 6; c.dt = 0.01
    elif "num_legs" in bad:
        c = pkg.make_config(pkg.workloads.default_config(6)); c.num_legs = 0
    else:
        c = pkg.make_config(cfg)
    h = C.c_void_p()
    assert pkg.load_library().cmpc_create(C.byref(c), C.byref(h)) == -1
    assert not h.value


def test_create_destroy_and_call_order_without_gpu(pkg):
    lib = pkg.load_library()
    c = pkg.make_config(pkg.workloads.default_config(6))
    h = C.c_void_p()
    assert lib.cmpc_create(C.byref(c), C.byref(h)) == 0 and h.value
    # solve before setup -> CMPC_ERR_STATE with a message, no crash
    z = np.zeros(8)
    rc = lib.cmpc_solve_batch(h, 1, z.ctypes.data, z.ctypes.data, z.ctypes.data, z.ctypes.data, z.ctypes.data,
                              None, None, None, None, None)
    assert rc == -3 and b"cmpc_setup" in lib.cmpc_last_error(h)
    assert lib.cmpc_update_weights(h, (C.c_double * 44)(), 44) == -1
    lib.cmpc_destroy(h)


def test_no_cpu_fallback(pkg):
    """Without a CUDA device the product must fail loudly (CMPC_ERR_NO_DEVICE), never solve on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    m = pkg.CentroidalMPC.from_dict(pkg.workloads.default_config(6))
    with pytest.raises(pkg.CmpcError, match="no CUDA device"):
        m.SetupMPC(4)
    m.close()


def test_product_does_not_reference_the_oracle():
    """oracle/ is test infrastructure: nothing under the package or include/ may name it."""
    for base in ("cheeta-mpc_b200", "include"):
        for dp, _, fns in os.walk(os.path.join(ROOT, base)):
            for fn in fns:
                if fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp", "Makefile")):
                    txt = open(os.path.join(dp, fn), errors="ignore").read()
                    assert "cmpc_oracle" not in txt and "oracle_py" not in txt and "numpy_mirror" not in txt, (dp, fn)
    out = subprocess.run(["ldd", os.path.join(ROOT, "cheeta-mpc_b200", "csrc", "libcmpc_b200.so")],
                         capture_output=True, text=True).stdout
    assert "oracle" not in out


def test_tree_is_free_of_the_cuda_batch_copy_entry_points():
    """The driver refuses GPU runs whose tree names one of four CUDA batch-copy entry points (a statically linked
    cudart carries them in its symbol table: a stray binary once did).  Every tracked file is checked, binaries
    included; the names are assembled here so that this file does not contain them."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    stems = [p + "Memcpy" + d + "Batch" + "Async" for p in ("cuda", "cu") for d in ("", "3D")]
    try:
        files = subprocess.run(["git", "ls-files"], cwd=root, capture_output=True, text=True, check=True).stdout.split("\n")
    except Exception:
        files = []
    if not [f for f in files if f]:
        files = [os.path.relpath(os.path.join(d, f), root) for d, _, fs in os.walk(root) if ".git" not in d.split(os.sep)
                 and "gpurun_out" not in d for f in fs]
    bad = []
    for f in files:
        path = os.path.join(root, f)
        if not f or not os.path.isfile(path) or os.path.getsize(path) > 64 << 20:
            continue
        data = open(path, "rb").read()
        if any(s.encode() in data for s in stems):
            bad.append(f)
    assert not bad, bad
