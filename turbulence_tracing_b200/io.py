"""I/O adapters for simulation cubes: the loaders the reference's example scripts
rely on, without the ``vtk`` dependency.

* :func:`pvti_readin` -- same name, arguments and return value ``(img, dim, spacing)`` as the helper the
  examples define around ``vtk.vtkXMLPImageDataReader`` (example_kitchensink.py:7-36, example_MPI.py): the
  first CellData array of a ``.pvti`` (parallel ImageData, pieces in ``.vti`` files) or a single ``.vti``,
  as an array indexed ``[ix, iy, iz(, component)]``.
* :func:`load_cube` -- ``.npy`` / ``.npz`` / ``.vti`` / ``.pvti`` by extension.
* :func:`centred_axes` -- the symmetric ``linspace`` axes the examples build from ``dim`` and ``spacing``
  (example_kitchensink.py:49-57).
* :func:`write_vti` / :func:`write_pvti` -- writers (ascii, base64 "binary", raw or base64 "appended", zlib) so
  that cubes produced here (e.g. ``turboGen.gaussian3D_FFT``) open in ParaView, and for the round-trip tests.

Host-side file parsing only; arrays come back as numpy and go to the device through
``ElectronCube.external_ne`` / ``external_B`` as usual.  The VTK XML layout handled: ``<DataArray>`` formats
``ascii`` | ``binary`` | ``appended`` (encoding ``raw`` | ``base64``), ``header_type`` UInt32 | UInt64, both byte
orders, ``vtkZLibDataCompressor`` blocks.
"""
from __future__ import annotations

import base64
import os
import re
import struct
import xml.etree.ElementTree as ET
import zlib

import numpy as np

_VTK_TYPES = {
    "Float32": "f4", "Float64": "f8", "Int8": "i1", "UInt8": "u1", "Int16": "i2", "UInt16": "u2",
    "Int32": "i4", "UInt32": "u4", "Int64": "i8", "UInt64": "u8",
}
_NP_TO_VTK = {np.dtype(v).str[1:]: k for k, v in _VTK_TYPES.items()}


class VTKFormatError(ValueError):
    pass


def _split_appended(raw):
    """(xml bytes without the appended payload, payload bytes after the '_' marker or None, encoding)."""
    m = re.search(rb"<AppendedData[^>]*>", raw)
    if not m:
        return raw, None, None
    enc = re.search(rb'encoding\s*=\s*"([^"]+)"', m.group(0))
    start = raw.index(b"_", m.end()) + 1
    end = raw.rindex(b"</AppendedData>")
    head = raw[:m.start()] + b"</VTKFile>"
    return head, raw[start:end], (enc.group(1).decode() if enc else "base64")


class _Blob:
    """Decoder of one binary data block (inline base64, appended base64 or appended raw)."""

    def __init__(self, byte_order, header_type, compressed):
        self.bo = "<" if byte_order == "LittleEndian" else ">"
        self.hfmt = {"UInt32": "I", "UInt64": "Q"}[header_type]
        self.hsize = struct.calcsize(self.hfmt)
        self.compressed = compressed

    def _ints(self, b, n):
        return struct.unpack(self.bo + self.hfmt * n, b[: n * self.hsize])

    def from_raw(self, buf):
        if not self.compressed:
            (nbytes,) = self._ints(buf, 1)
            return bytes(buf[self.hsize: self.hsize + nbytes])
        nblocks, _, _ = self._ints(buf, 3)
        sizes = self._ints(buf[3 * self.hsize:], nblocks)
        pos = (3 + nblocks) * self.hsize
        out = []
        for s in sizes:
            out.append(zlib.decompress(bytes(buf[pos: pos + s])))
            pos += s
        return b"".join(out)

    def from_base64(self, txt):
        """txt: base64 characters starting at the block (may run on into later blocks)."""
        txt = bytes(txt).strip()

        def dec(chars):
            return base64.b64decode(chars + b"=" * (-len(chars) % 4))

        def take(nbytes, at):
            """decode the base64 group(s) covering nbytes starting at char `at`: (bytes, chars consumed)"""
            nchar = -(-nbytes // 3) * 4
            return dec(txt[at: at + nchar])[:nbytes], nchar

        if not self.compressed:
            # length header + data in ONE stream (VTK >= 5); older writers pad the header separately
            h, nchar = take(self.hsize, 0)
            (nbytes,) = self._ints(h, 1)
            if txt[nchar - 1: nchar] == b"=":            # separately encoded header
                return take(nbytes, nchar)[0]
            return take(self.hsize + nbytes, 0)[0][self.hsize:]
        h, n1 = take(3 * self.hsize, 0)
        nblocks, _, _ = self._ints(h, 3)
        h, nchar = take((3 + nblocks) * self.hsize, 0)      # the header is its own base64 stream
        sizes = self._ints(h[3 * self.hsize:], nblocks)
        data = take(sum(sizes), nchar)[0]
        out, pos = [], 0
        for s in sizes:
            out.append(zlib.decompress(data[pos: pos + s]))
            pos += s
        return b"".join(out)


def _read_data_array(el, blob, bo, appended, app_enc):
    try:
        dt = np.dtype(_VTK_TYPES[el.get("type")])
    except KeyError:
        raise VTKFormatError(f"unsupported DataArray type {el.get('type')!r}") from None
    ncomp = int(el.get("NumberOfComponents", "1"))
    fmt = el.get("format", "ascii")
    if fmt == "ascii":
        a = np.array((el.text or "").split(), dtype=np.float64).astype(dt)
    else:
        if fmt == "binary":
            b = blob.from_base64((el.text or "").encode())
        elif fmt == "appended":
            if appended is None:
                raise VTKFormatError("DataArray format='appended' but the file has no <AppendedData>")
            off = int(el.get("offset", "0"))
            b = blob.from_raw(memoryview(appended)[off:]) if app_enc == "raw" else blob.from_base64(appended[off:])
        else:
            raise VTKFormatError(f"unknown DataArray format {fmt!r}")
        a = np.frombuffer(b, dtype=dt.newbyteorder(bo))
        a = a.astype(dt)              # native byte order, owns its memory
    return a, ncomp


def _extent(s):
    e = [int(v) for v in s.split()]
    if len(e) != 6:
        raise VTKFormatError(f"bad extent {s!r}")
    return e


def _read_vti_pieces(path, array=0):
    """[(extent, flat data, ncomp)], whole extent, origin, spacing, array name of one .vti file."""
    with open(path, "rb") as f:
        raw = f.read()
    head, appended, app_enc = _split_appended(raw)
    root = ET.fromstring(head)
    if root.tag != "VTKFile" or root.get("type") != "ImageData":
        raise VTKFormatError(f"{path}: not a VTK XML ImageData file (type={root.get('type')!r})")
    bo = "<" if root.get("byte_order", "LittleEndian") == "LittleEndian" else ">"
    comp = root.get("compressor")
    if comp not in (None, "", "vtkZLibDataCompressor"):
        raise VTKFormatError(f"{path}: unsupported compressor {comp}")
    blob = _Blob(root.get("byte_order", "LittleEndian"), root.get("header_type", "UInt32"), bool(comp))
    img = root.find("ImageData")
    whole = _extent(img.get("WholeExtent"))
    origin = np.array([float(v) for v in img.get("Origin", "0 0 0").split()])
    spacing = np.array([float(v) for v in img.get("Spacing", "1 1 1").split()])
    pieces, name = [], None
    for piece in img.findall("Piece"):
        cd = piece.find("CellData")
        arrays = [] if cd is None else cd.findall("DataArray")
        if isinstance(array, str):
            arrays = [a for a in arrays if a.get("Name") == array]
            idx = 0
        else:
            idx = array
        if idx >= len(arrays):
            raise VTKFormatError(f"{path}: CellData array {array!r} not found")
        el = arrays[idx]
        name = el.get("Name")
        flat, ncomp = _read_data_array(el, blob, bo, appended, app_enc)
        pieces.append((_extent(piece.get("Extent")), flat, ncomp))
    return pieces, whole, origin, spacing, name


def _assemble(pieces, whole):
    ncomp = pieces[0][2]
    dims = [whole[1] - whole[0], whole[3] - whole[2], whole[5] - whole[4]]          # cells per axis
    shape = dims + ([ncomp] if ncomp > 1 else [])
    out = np.empty(shape, dtype=pieces[0][1].dtype)
    for ext, flat, nc in pieces:
        if nc != ncomp:
            raise VTKFormatError("pieces disagree on NumberOfComponents")
        n = [ext[1] - ext[0], ext[3] - ext[2], ext[5] - ext[4]]
        if flat.size != n[0] * n[1] * n[2] * nc:
            raise VTKFormatError(f"piece with extent {ext} holds {flat.size} values, expected {n[0] * n[1] * n[2] * nc}")
        # VTK order: x fastest, components innermost  ->  [ix, iy, iz(, c)]
        blk = flat.reshape((n[2], n[1], n[0]) + ((nc,) if nc > 1 else ())).transpose((2, 1, 0) + ((3,) if nc > 1 else ()))
        lo = (ext[0] - whole[0], ext[2] - whole[2], ext[4] - whole[4])
        out[lo[0]: lo[0] + n[0], lo[1]: lo[1] + n[1], lo[2]: lo[2] + n[2]] = blk
    return out


def read_vti(filename, array=0):
    """First (or named) CellData array of a .vti file -> (img[ix, iy, iz(, c)], origin, spacing, name)."""
    pieces, whole, origin, spacing, name = _read_vti_pieces(filename, array)
    if not pieces:
        raise VTKFormatError(f"{filename}: no <Piece>")
    # a piece file of a .pvti carries the WHOLE extent of the data set: return what the file holds
    ext = np.array([p[0] for p in pieces])
    box = [ext[:, 0].min(), ext[:, 1].max(), ext[:, 2].min(), ext[:, 3].max(), ext[:, 4].min(), ext[:, 5].max()]
    return _assemble(pieces, [int(v) for v in box]), origin, spacing, name


def read_pvti(filename, array=0):
    """First (or named) cell array of a .pvti file, its pieces read from the .vti files it lists."""
    root = ET.parse(filename).getroot()
    if root.tag != "VTKFile" or root.get("type") != "PImageData":
        raise VTKFormatError(f"{filename}: not a VTK XML PImageData file")
    pimg = root.find("PImageData")
    whole = _extent(pimg.get("WholeExtent"))
    origin = np.array([float(v) for v in pimg.get("Origin", "0 0 0").split()])
    spacing = np.array([float(v) for v in pimg.get("Spacing", "1 1 1").split()])
    base = os.path.dirname(os.path.abspath(filename))
    if isinstance(array, int):
        names = [a.get("Name") for a in (pimg.find("PCellData").findall("PDataArray") if pimg.find("PCellData") is not None else [])]
        if array < len(names) and names[array]:
            array = names[array]             # pieces are looked up by name
    pieces, name = [], None
    for p in pimg.findall("Piece"):
        sub, _, _, _, name = _read_vti_pieces(os.path.join(base, p.get("Source")), array)
        ext = _extent(p.get("Extent"))
        for e, flat, nc in sub:
            pieces.append((e if len(sub) > 1 else ext, flat, nc))
    if not pieces:
        raise VTKFormatError(f"{filename}: no <Piece>")
    return _assemble(pieces, whole), origin, spacing, name


def pvti_readin(filename):
    """Drop-in for the helper of the reference's examples (example_kitchensink.py:7-36): returns
    ``(img, dim, spacing)`` -- the first cell array as ``img[ix, iy, iz(, component)]`` (what
    ``vtk_to_numpy(...).reshape(vec, order="F")`` yields there), ``dim = img.shape`` and the grid spacing."""
    img, _, spacing, _ = (read_pvti if str(filename).lower().endswith(".pvti") else read_vti)(filename)
    return img, img.shape, spacing


def load_cube(filename, key=None):
    """Cube from ``.npy`` | ``.npz`` (array ``key`` or the only one) | ``.vti`` | ``.pvti``."""
    ext = os.path.splitext(str(filename))[1].lower()
    if ext == ".npy":
        return np.load(filename)
    if ext == ".npz":
        with np.load(filename) as z:
            if key is None:
                if len(z.files) != 1:
                    raise ValueError(f"{filename} holds {z.files}; pass key=")
                key = z.files[0]
            return z[key]
    if ext in (".vti", ".pvti"):
        return (read_pvti if ext == ".pvti" else read_vti)(filename, 0 if key is None else key)[0]
    raise ValueError(f"unknown cube format {ext!r}")


def centred_axes(dim, spacing, stride=1):
    """Symmetric axes for a cube of ``dim`` cells with cell size ``spacing``, sub-sampled by ``stride``:
    ``linspace(-e, e, M)`` with ``M = dim // stride`` and ``e = stride * spacing * (M - 1) / 2``, as the
    examples build them (example_kitchensink.py:49-57)."""
    axes = []
    for n, h in zip(dim[:3], np.broadcast_to(spacing, (3,))):
        m = int(n) // int(stride)
        e = stride * float(h) * ((m - 1) / 2)
        axes.append(np.linspace(-e, e, m))
    return tuple(axes)


# ---------------------------------------------------------------------------------------------- writers
def _encode_block(data, bo, header_type, compress, block=1 << 15):
    hf = bo + {"UInt32": "I", "UInt64": "Q"}[header_type]
    if not compress:
        return struct.pack(hf, len(data)), data
    chunks = [data[i: i + block] for i in range(0, len(data), block)] or [b""]
    comp = [zlib.compress(c) for c in chunks]
    head = struct.pack(hf[0] + hf[1] * (3 + len(comp)), len(comp), block, len(chunks[-1]) % block if len(chunks[-1]) != block else 0,
                       *[len(c) for c in comp])
    return head, b"".join(comp)


def write_vti(filename, img, spacing=(1.0, 1.0, 1.0), origin=(0.0, 0.0, 0.0), name="data", fmt="appended",
              encoding="raw", compress=False, header_type="UInt64", byte_order="LittleEndian",
              extent=None, whole_extent=None):
    """Write ``img[ix, iy, iz(, c)]`` as the CellData array ``name`` of a VTK XML ImageData file.
    fmt: 'ascii' | 'binary' | 'appended' (encoding 'raw' | 'base64')."""
    img = np.asarray(img)
    nc = img.shape[3] if img.ndim == 4 else 1
    n = img.shape[:3]
    if extent is None:
        extent = [0, n[0], 0, n[1], 0, n[2]]
    if whole_extent is None:
        whole_extent = extent
    bo = "<" if byte_order == "LittleEndian" else ">"
    vt = _NP_TO_VTK[img.dtype.str[1:]]
    flat = np.ascontiguousarray(img.transpose((2, 1, 0) + ((3,) if img.ndim == 4 else ()))).reshape(-1)
    ext_s = " ".join(str(v) for v in extent)
    attrs = f'type="{vt}" Name="{name}" NumberOfComponents="{nc}"'
    comp_attr = ' compressor="vtkZLibDataCompressor"' if compress and fmt != "ascii" else ""
    head = (f'<?xml version="1.0"?>\n<VTKFile type="ImageData" version="1.0" byte_order="{byte_order}" '
            f'header_type="{header_type}"{comp_attr}>\n'
            f'  <ImageData WholeExtent="{" ".join(str(v) for v in whole_extent)}" '
            f'Origin="{" ".join(repr(float(v)) for v in origin)}" Spacing="{" ".join(repr(float(v)) for v in spacing)}">\n'
            f'    <Piece Extent="{ext_s}">\n      <CellData Scalars="{name}">\n').encode()
    tail_piece = b"      </CellData>\n    </Piece>\n  </ImageData>\n"
    with open(filename, "wb") as f:
        f.write(head)
        if fmt == "ascii":
            f.write(f"        <DataArray {attrs} format=\"ascii\">\n".encode())
            f.write(" ".join(repr(v) for v in flat.tolist()).encode())
            f.write(b"\n        </DataArray>\n" + tail_piece)
        else:
            h, d = _encode_block(flat.astype(flat.dtype.newbyteorder(bo)).tobytes(), bo, header_type, compress)
            # uncompressed: header + data form one base64 stream; compressed: the header is its own stream
            b64 = (base64.b64encode(h) + base64.b64encode(d)) if compress else base64.b64encode(h + d)
            if fmt == "binary":
                f.write(f"        <DataArray {attrs} format=\"binary\">\n".encode() + b64 + b"\n        </DataArray>\n" + tail_piece)
            elif fmt == "appended":
                f.write(f"        <DataArray {attrs} format=\"appended\" offset=\"0\"/>\n".encode() + tail_piece)
                f.write(f'  <AppendedData encoding="{encoding}">\n   _'.encode())
                f.write(h + d if encoding == "raw" else b64)
                f.write(b"\n  </AppendedData>\n")
            else:
                raise ValueError(fmt)
        f.write(b"</VTKFile>\n")


def write_pvti(filename, img, spacing=(1.0, 1.0, 1.0), origin=(0.0, 0.0, 0.0), name="data", pieces=(2, 1, 1), **kw):
    """Write ``img`` as a .pvti whose pieces (``pieces`` per axis) go to ``<stem>/<stem>_<i>.vti``."""
    img = np.asarray(img)
    stem = os.path.splitext(os.path.basename(filename))[0]
    d = os.path.join(os.path.dirname(os.path.abspath(filename)), stem)
    os.makedirs(d, exist_ok=True)
    n = img.shape[:3]
    nc = img.shape[3] if img.ndim == 4 else 1
    whole = [0, n[0], 0, n[1], 0, n[2]]
    cuts = [np.linspace(0, n[a], pieces[a] + 1).astype(int) for a in range(3)]
    lines, i = [], 0
    for a in range(pieces[0]):
        for b in range(pieces[1]):
            for c in range(pieces[2]):
                ext = [cuts[0][a], cuts[0][a + 1], cuts[1][b], cuts[1][b + 1], cuts[2][c], cuts[2][c + 1]]
                ext = [int(v) for v in ext]
                src = f"{stem}/{stem}_{i}.vti"
                write_vti(os.path.join(d, f"{stem}_{i}.vti"), img[ext[0]:ext[1], ext[2]:ext[3], ext[4]:ext[5]], spacing, origin,
                          name, extent=ext, whole_extent=whole, **kw)
                lines.append(f'    <Piece Extent="{" ".join(str(v) for v in ext)}" Source="{src}"/>')
                i += 1
    vt = _NP_TO_VTK[img.dtype.str[1:]]
    with open(filename, "w") as f:
        f.write(f'<?xml version="1.0"?>\n<VTKFile type="PImageData" version="1.0" byte_order="LittleEndian" header_type="UInt64">\n'
                f'  <PImageData WholeExtent="{" ".join(str(v) for v in whole)}" GhostLevel="0" '
                f'Origin="{" ".join(repr(float(v)) for v in origin)}" Spacing="{" ".join(repr(float(v)) for v in spacing)}">\n'
                f'    <PCellData Scalars="{name}">\n      <PDataArray type="{vt}" Name="{name}" NumberOfComponents="{nc}"/>\n'
                f'    </PCellData>\n' + "\n".join(lines) + "\n  </PImageData>\n</VTKFile>\n")
