#!/usr/bin/env python
"""Run the reference's own applications (unmodified sources, compiled against the stand-ins of oracle/compat with the
reference's base::solver::Eigen3: oracle/_ref/apps/*) on small inputs and store what they print under
tests/golden/refrun_apps/.  The tests then require the SAME applications compiled against the B200 binding
(include/insilico_b200_reference.hpp, oracle/_ref/apps_b200/*) to print the same.

    make -C oracle ref && python tools/make_ref_app_goldens.py
"""
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import ref_apps_cases as C  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "refrun_apps")


def main():
    os.makedirs(OUT, exist_ok=True)
    for name in C.CASES:
        with tempfile.TemporaryDirectory() as wd:
            exe, args = C.prepare(name, wd)
            out = subprocess.run([os.path.join(ROOT, "oracle", "_ref", "apps", exe)] + args, cwd=wd, check=True,
                                 capture_output=True, text=True, timeout=900).stdout
            if C.output_file(name):
                out += open(os.path.join(wd, C.output_file(name))).read()
        with open(os.path.join(OUT, name + ".out"), "w") as f:
            f.write(out)
        print(name, "->", len(out.splitlines()), "lines")


if __name__ == "__main__":
    main()
