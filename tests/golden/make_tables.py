#!/usr/bin/env python
"""Extract the numeric tables of the named configs from the reference headers into .npy fixtures
(and oracle/oracle_tables.h).  Tables are inputs of configs 3b / 4, not code.  Dev container only."""
import os
import re

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference/include"


def halfplanes(name):
    src = open(os.path.join(REF, name)).read()
    rows = re.findall(r"\{\s*([-+0-9.eE]+)\s*,\s*([-+0-9.eE]+)\s*\}", src)
    return rows


def main():
    rows = halfplanes("KernelData_70-135kg.h")
    np.save(os.path.join(HERE, "halfplanes_70_135.npy"), np.array(rows, dtype=np.float64))
    out = ["/* oracle_tables.h -- numeric tables of the named configs (TEST INFRASTRUCTURE ONLY).",
           " * Values of include/KernelData_70-135kg.h:5-104 (100 half-planes {a0,a1}, h = 1 - a.x), extracted by",
           " * tests/golden/make_tables.py; the table is an input of config 3b, not code. */",
           "#ifndef ORACLE_TABLES_H", "#define ORACLE_TABLES_H", "#define ORACLE_N_HALFPLANES %d" % len(rows),
           "static const double oracle_halfplanes_70_135[2 * ORACLE_N_HALFPLANES] = {"]
    out += ["\t%s, %s," % (a, b) for a, b in rows] + ["};", "#endif", ""]
    open(os.path.join(ROOT, "oracle", "oracle_tables.h"), "w").write("\n".join(out))
    print("half-planes:", len(rows))


if __name__ == "__main__":
    main()


def realizable_kernel():
    """Geometry + x-independent facet interval table of config 4, exported by the reference build itself
    (libaffa evaluates the interval dynamics over each facet; src/asif_realizable.cpp:137-157,470-500)."""
    import sys
    sys.path.insert(0, ROOT)
    from oracle import pyref
    L = pyref.RefLib()
    f = L.create(pyref.CFG_IP_REALIZABLE)
    k = pyref.realizable_export(f)
    np.savez(os.path.join(HERE, "realizable_kernel_100hz_50pt.npz"), **k)
    nF, mA = k["facet_active"].shape
    out = ["", "/* include/RealizableKernelData_100Hz_50pt.h (50 vertices / 50 facets, maxCriticalFacets 3, maxActiveConstraints 3)",
           " * + [LfLo, LfHi, LgLo, LgHi] per (facet, active constraint) as the reference build computes them */",
           "#define ORACLE_RZ_NV %d" % len(k["vertices"]), "#define ORACLE_RZ_NF %d" % nF,
           "#define ORACLE_RZ_MAXCRIT %d" % k["max_critical_facets"], "#define ORACLE_RZ_MAXACT %d" % mA,
           "static const double oracle_rz_vertices[2 * ORACLE_RZ_NV] = {"]
    out += ["\t%r, %r," % (float(a), float(b)) for a, b in k["vertices"]] + ["};",
           "static const double oracle_rz_normals[2 * ORACLE_RZ_NF] = {"]
    out += ["\t%r, %r," % (float(a), float(b)) for a, b in k["normals"]] + ["};",
           "static const int oracle_rz_facet_vertices[2 * ORACLE_RZ_NF] = {"]
    out += ["\t%d, %d," % (a, b) for a, b in k["facet_vertices"]] + ["};",
           "static const int oracle_rz_facet_active[ORACLE_RZ_MAXACT * ORACLE_RZ_NF] = {"]
    out += ["\t" + ", ".join(str(int(v)) for v in row) + "," for row in k["facet_active"]] + ["};",
           "static const double oracle_rz_facet_lie[4 * ORACLE_RZ_MAXACT * ORACLE_RZ_NF] = {"]
    out += ["\t" + ", ".join(repr(float(v)) for v in row) + "," for row in k["facet_lie"].reshape(-1, 4)] + ["};", ""]
    p = os.path.join(ROOT, "oracle", "oracle_tables.h")
    s = open(p).read().replace("#endif\n", "")
    open(p, "w").write(s + "\n".join(out) + "#endif\n")
    print("realizable kernel: %d vertices, %d facets" % (len(k["vertices"]), nF))


if __name__ == "__main__":
    realizable_kernel()
