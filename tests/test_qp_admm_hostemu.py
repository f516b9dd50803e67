"""CPU check of the ALGORITHM of the device QP solver for nv > 4 (asif_b200/csrc/qp_admm.cuh).

The solver is written against a Team concept; tests/hostemu/qp_admm_hostemu.cpp instantiates the same text with a
one-thread host team and exports the two symbols libasif_ref_b200.so imports.  Loaded ahead of that library (in a
subprocess, so that symbol resolution order is under control), it runs the reference's own ASIFrealizable / ASIFrobust
classes on that algorithm, against the OSQP stand-in build.  TEST INFRASTRUCTURE: the product library has no CPU path
(test_capi.py) and this emulation is never loaded by it.  Parallel execution is tests/test_gpu_qp_admm.py's job."""
import os
import subprocess
import sys

import pytest

import conftest as cf

SCRIPT = r"""
import ctypes as C, sys, numpy as np
sys.path.insert(0, %(root)r); sys.path.insert(0, %(tests)r)
C.CDLL(%(emu)r, mode=C.RTLD_GLOBAL)
import conftest as cf
from oracle import pyref
L = pyref.RefLib(pyref.REF_B200_SO)
L.set_qp_mode()
for cfg, opts, gen, n in ((5, cf.C4_OPTS, cf.c4_inputs, 120), (4, cf.C3B_OPTS, cf.c3b_inputs, 3)):
    x, ud = gen(n, seed=cf.SEED + 1900 + cfg)
    L.select_backend(0); f0 = L.create(cfg, opts)
    L.select_backend(1); f1 = L.create(cfg, opts)
    u0, r0, rc0, _, st0, it0 = f0.filter_batch_ex(x, ud)
    u1, r1, rc1 = f1.filter_batch(x, ud)
    r0, r1 = np.asarray(r0).reshape(n, -1)[:, -1:], np.asarray(r1).reshape(n, -1)[:, -1:]  # realizable: the second relax slot is the QP's (as the golden tests)
    D = cf.disagree((u1, r1, rc1), (u0, r0, rc0))
    print("cfg", cfg, "n", n, "disagree", int(D.sum()), "max|du|", float(np.abs(u1 - u0).max()), "rc", sorted(set(rc1.tolist())))
    assert D.sum() == 0
print("HOSTEMU_OK")
"""


def test_device_qp_algorithm_on_host_team(tmp_path):
    from oracle import pyref
    if not os.path.exists(pyref.REF_B200_SO):
        pytest.skip("oracle/_ref/libasif_ref_b200.so not built (needs /root/reference at build time)")
    src = os.path.join(cf.ROOT, "tests", "hostemu", "qp_admm_hostemu.cpp")
    emu = str(tmp_path / "libqp_admm_hostemu.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-o", emu, src])
    code = SCRIPT % dict(root=cf.ROOT, tests=os.path.join(cf.ROOT, "tests"), emu=emu)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=900)
    print(r.stdout[-2000:], r.stderr[-2000:])
    assert r.returncode == 0 and "HOSTEMU_OK" in r.stdout
