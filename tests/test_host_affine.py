"""Host-side affine evaluator + kernel CSV loader (asif_b200/host/affine.hpp, realizable_host.hpp): the facet table
they produce must equal the one the reference build computed with libaffa.  CPU only."""
import os
import subprocess

import numpy as np

import conftest as cf


def write_kernel_csv(path, k):
    """The file format of examples/InvertedPendulum_RealizableSampled.cpp::loadRealizableKernel."""
    with open(path, "w") as f:
        f.write("%d,\n%d,\n%d,\n%d,\n" % (len(k["vertices"]), len(k["normals"]), int(k["max_critical_facets"]),
                                            int(k["max_active_constraints"])))
        for v in k["vertices"]:
            f.write("%r,%r\n" % (float(v[0]), float(v[1])))
        for i in range(len(k["normals"])):
            f.write("%d,%d\n" % tuple(int(t) for t in k["facet_vertices"][i]))
            f.write("%r,%r\n" % tuple(float(t) for t in k["normals"][i]))
            f.write(",".join(str(int(t)) for t in k["facet_active"][i] if t >= 0) + "\n")


def test_facet_table_matches_libaffa(tmp_path):
    host = os.path.join(cf.ROOT, "asif_b200", "host")
    subprocess.check_call(["make", "-C", host, "-s", "facet_table_check"])
    k = cf.realizable_kernel()
    csv = str(tmp_path / "kernel.csv")
    write_kernel_csv(csv, k)
    out = subprocess.run([os.path.join(host, "facet_table_check"), csv], capture_output=True, text=True, check=True).stdout
    lines = out.strip().split("\n")
    assert [int(t) for t in lines[0].split()] == [50, 50, 3, 3]
    table = np.array([[float(t) for t in ln.split()] for ln in lines[1:]]).reshape(50, 3, 4)
    ref = k["facet_lie"]
    assert np.abs(table - ref).max() <= 1e-15, np.abs(table - ref).max()
    # the intervals are not degenerate: the least-squares sine and the gain interval both contribute
    assert np.median(ref[:, :, 1] - ref[:, :, 0]) > 1e-2 and np.median(ref[:, :, 3] - ref[:, :, 2]) > 1e-2


def test_host_copier_threads_copy_every_byte():
    """csrc/host_copier.hpp (the pool that moves pageable caller arrays through pinned staging): 1200 copy() calls of
    random shapes with 1 / 2 / 5 / 8 threads, compared byte for byte (asif_b200/host/host_copier_check.cpp)."""
    host = os.path.join(cf.ROOT, "asif_b200", "host")
    subprocess.check_call(["make", "-C", host, "-s", "host_copier_check"])
    r = subprocess.run([os.path.join(host, "host_copier_check")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "host_copier ok" in r.stdout, r.stdout + r.stderr
