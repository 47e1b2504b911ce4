"""Simulation-cube loaders (turbulence_tracing_b200/io.py): VTK XML ImageData without the vtk package.  CPU only."""
import itertools
import os

import numpy as np
import pytest

from turbulence_tracing_b200 import io as tio


def _cube(shape, dtype, seed=0):
    rng = np.random.RandomState(seed)
    return (rng.rand(*shape) * 1e3).astype(dtype)


@pytest.mark.parametrize("fmt,encoding,compress", [("ascii", "raw", False), ("binary", "raw", False), ("binary", "raw", True),
                                                   ("appended", "raw", False), ("appended", "raw", True),
                                                   ("appended", "base64", False), ("appended", "base64", True)])
@pytest.mark.parametrize("header_type,byte_order", [("UInt64", "LittleEndian"), ("UInt32", "BigEndian")])
def test_vti_round_trip(tmp_path, fmt, encoding, compress, header_type, byte_order):
    for shape, dtype in (((5, 4, 3), np.float64), ((4, 3, 6, 3), np.float32), ((7, 1, 2), np.int32)):
        img = _cube(shape, dtype, seed=len(shape))
        fn = tmp_path / "c.vti"
        tio.write_vti(fn, img, spacing=(1e-4, 2e-4, 3e-4), origin=(1.0, 2.0, 3.0), name="rnec", fmt=fmt, encoding=encoding,
                      compress=compress, header_type=header_type, byte_order=byte_order)
        got, origin, spacing, name = tio.read_vti(fn)
        assert got.dtype == img.dtype and got.shape == img.shape and name == "rnec"
        np.testing.assert_array_equal(got, img)
        np.testing.assert_array_equal(spacing, [1e-4, 2e-4, 3e-4])
        np.testing.assert_array_equal(origin, [1.0, 2.0, 3.0])


def test_vti_layout_is_x_fastest(tmp_path):
    """VTK stores x fastest: the flat ascii payload of img[ix, iy, iz] = ix + 10 iy + 100 iz reads 0 1 2 10 11 12 ..."""
    ix, iy, iz = np.meshgrid(np.arange(3), np.arange(2), np.arange(2), indexing="ij")
    img = (ix + 10 * iy + 100 * iz).astype(np.int32)
    fn = tmp_path / "l.vti"
    tio.write_vti(fn, img, fmt="ascii")
    txt = open(fn).read()
    payload = txt[txt.index('format="ascii">') + 15: txt.index("</DataArray>")].split()
    assert [int(float(v)) for v in payload] == [0, 1, 2, 10, 11, 12, 100, 101, 102, 110, 111, 112]
    np.testing.assert_array_equal(tio.read_vti(fn)[0], img)
    # compressed payload spanning several zlib blocks
    big = _cube((40, 30, 20), np.float64)
    tio.write_vti(fn, big, fmt="appended", compress=True)
    np.testing.assert_array_equal(tio.read_vti(fn)[0], big)


@pytest.mark.parametrize("pieces", [(2, 1, 1), (2, 3, 2), (1, 1, 1)])
def test_pvti_readin_matches_the_examples_helper(tmp_path, pieces):
    """(img, dim, spacing) as example_kitchensink.py:7-36 returns them; vector arrays keep the component axis last."""
    rnec = _cube((12, 9, 10), np.float64, 1)
    bvec = _cube((12, 9, 10, 3), np.float32, 2)
    tio.write_pvti(tmp_path / "x08_rnec-400.pvti", rnec, spacing=(5e-5, 5e-5, 1e-4), name="rnec", pieces=pieces, compress=True)
    tio.write_pvti(tmp_path / "x08_Bvec-400.pvti", bvec, spacing=(5e-5, 5e-5, 1e-4), name="Bvec", pieces=pieces, fmt="binary")
    img, dim, spacing = tio.pvti_readin(str(tmp_path / "x08_rnec-400.pvti"))
    np.testing.assert_array_equal(img, rnec)
    assert tuple(dim) == (12, 9, 10)
    np.testing.assert_array_equal(spacing, [5e-5, 5e-5, 1e-4])
    B, dimB, _ = tio.pvti_readin(str(tmp_path / "x08_Bvec-400.pvti"))
    np.testing.assert_array_equal(B, bvec)
    assert tuple(dimB) == (12, 9, 10, 3)
    # the examples' axes (example_kitchensink.py:49-57): every other cell, symmetric linspace
    x, y, z = tio.centred_axes(dim, spacing, stride=2)
    M = dim[1] // 2
    ext = 2 * spacing[1] * ((M - 1) / 2)
    np.testing.assert_array_equal(y, np.linspace(-ext, ext, M))
    assert len(x) == 6 and len(z) == 5 and z[-1] == pytest.approx(2 * 1e-4 * 2)
    np.testing.assert_array_equal(tio.load_cube(tmp_path / "x08_rnec-400.pvti"), rnec)
    np.testing.assert_array_equal(tio.load_cube(tmp_path / "x08_rnec-400" / "x08_rnec-400_0.vti"),
                                  rnec[: 12 // pieces[0], : 9 // pieces[1], : 10 // pieces[2]])


def test_load_cube_npy_npz_and_errors(tmp_path):
    a = _cube((4, 5, 6), np.float32)
    np.save(tmp_path / "a.npy", a)
    np.savez(tmp_path / "a.npz", ne=a)
    np.savez(tmp_path / "two.npz", ne=a, te=a)
    np.testing.assert_array_equal(tio.load_cube(tmp_path / "a.npy"), a)
    np.testing.assert_array_equal(tio.load_cube(tmp_path / "a.npz"), a)
    np.testing.assert_array_equal(tio.load_cube(tmp_path / "two.npz", key="te"), a)
    with pytest.raises(ValueError):
        tio.load_cube(tmp_path / "two.npz")
    with pytest.raises(ValueError):
        tio.load_cube(tmp_path / "a.h5")
    (tmp_path / "bad.vti").write_text('<?xml version="1.0"?><VTKFile type="PolyData"></VTKFile>')
    with pytest.raises(tio.VTKFormatError):
        tio.read_vti(tmp_path / "bad.vti")
    tio.write_vti(tmp_path / "ok.vti", a)
    with pytest.raises(tio.VTKFormatError):
        tio.read_vti(tmp_path / "ok.vti", array="missing")


def test_pvti_readin_of_a_hand_written_vtk_file():
    """tests/golden/vtk_handwritten/: a .pvti with two .vti pieces written by hand to the VTK XML format (legacy-style
    headers: version 0.1, UInt32 block header; one piece ascii, one inline base64), NOT by this package's writers.
    Cell data 4 x 3 x 2, value(i, j, k) = 100 i + 10 j + k + 0.5, x fastest -- what vtkXMLPImageDataReader +
    vtk_to_numpy(...).reshape(order="F") of the reference's helper returns (example_kitchensink.py:7-36)."""
    from turbulence_tracing_b200 import io as tio
    path = os.path.join(os.path.dirname(__file__), "golden", "vtk_handwritten", "cube.pvti")
    img, dim, spacing = tio.pvti_readin(path)
    assert tuple(dim) == (4, 3, 2) and img.shape == (4, 3, 2)
    i, j, k = np.meshgrid(np.arange(4), np.arange(3), np.arange(2), indexing="ij")
    np.testing.assert_array_equal(img, (100.0 * i + 10.0 * j + k + 0.5).astype(np.float32))
    np.testing.assert_allclose(spacing, [5e-5, 5e-5, 1e-4])
