"""Legacy VTK writer (base/io/vtk/LegacyWriter.hpp) against the reference's own golden file reference/05-mixedPoisson/ref.vtk
(written by the unmodified mixedPoissonWithDriver; the application test of the binding compares the same file)."""
import io
import os

import numpy as np
import pytest

from insilico_b200 import engine as E
from insilico_b200 import smf, vtk

REFERENCE = "/root/reference"


def test_writer_layout():
    coords = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]], dtype=float)
    out = io.StringIO()
    w = vtk.LegacyWriter(out)
    w.write_unstructured_grid(E.TET, coords, np.array([[0, 1, 2, 3]]))
    w.write_point_data(np.array([[1.0, 2.0, 3.0]] * 4), "u")
    w.write_point_data(np.arange(4.0), "p")
    w.write_cell_data(np.array([[0.5]]), "c")
    text = out.getvalue().splitlines()
    assert text[3:6] == ["DATASET UNSTRUCTURED_GRID", "POINTS 4 float", "0 0 0 "]
    assert "CELLS 1 5" in text and "4 0 1 2 3 " in text and "CELL_TYPES 1" in text and "10" in text
    assert text.count("POINT_DATA 4") == 1 and "VECTORS u float " in text and "1 2 3 " in text
    assert "SCALARS p float " in text and "CELL_DATA 1" in text and "0.5 " in text


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="needs /root/reference")
def test_reproduces_the_reference_golden_file():
    d = os.path.join(REFERENCE, "reference", "05-mixedPoisson")
    gold = open(os.path.join(d, "ref.vtk")).read()
    lines = gold.splitlines()
    shape, deg, coords, conn = smf.read(os.path.join(d, "square_020.smf"))
    i_pd = lines.index("POINT_DATA 441")
    i_cd = lines.index("CELL_DATA 400")
    temperature = np.array([float(x) for x in lines[i_pd + 3:i_cd]])[:, None]          # base::Vector<1> values
    flux = np.array([[float(t) for t in x.split()[:2]] for x in lines[i_cd + 2:i_cd + 2 + 400]])
    out = io.StringIO()
    w = vtk.LegacyWriter(out)
    w.write_unstructured_grid(shape, coords, conn)
    w.write_point_data(temperature, "temperature")
    w.write_cell_data(flux, "flux")
    assert out.getvalue() == gold
