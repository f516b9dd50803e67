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
