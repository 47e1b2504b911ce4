"""Drop-in mirror of the reference's ``ray_transfer_matrix`` module: optical element functions,
``Rays`` and the Shadowgraphy / Schlieren_DF / Schlieren_LF / AFR detectors.

Reference: particle_tracking/ray_transfer_matrix.py (elements :37-154, Rays :156-206, detectors
:208-299).  Every element chain and the histogram run in ONE fused CUDA kernel
(csrc/optics_hist.cu) instead of one numpy temporary per element; a detector's ``solve()`` only
records its element program, the kernel runs when ``rf`` or ``histogram()`` is first needed.

Rays are 4 x N arrays (x, theta, y, phi); numpy arrays, torch tensors and DeviceArray are accepted.
Rejected rays are NaN in all four rows, as in the reference.
"""
from __future__ import annotations

import numpy as np

from . import _lib
from ._lib import DeviceArray

__all__ = ["OpticsProgram", "m_to_mm", "lens", "sym_lens", "distance", "circular_aperture", "circular_stop", "annular_stop",
           "angular_filter", "rect_aperture", "knife_edge", "plot_afr", "Rays", "Shadowgraphy", "Schlieren_DF",
           "Schlieren_LF", "AFR", "Refractometer", "ShadowgraphyRays", "SchlierenRays", "BurdiscopeRays"]


# ---- element programs ---------------------------------------------------------------------------
def _op(code, a=0.0, b=0.0):
    return (int(code), float(a), float(b))


def _op_distance(d):
    return _op(_lib.OP_DISTANCE, d)


def _op_lens(f1, f2):
    return _op(_lib.OP_LENS, f1, f2)


def _op_knife(offset, axis, direction):
    if axis not in ("x", "y"):
        raise ValueError("axis must be 'x' or 'y'")
    if direction == 0:
        raise ValueError("Direction must be <0 or >0")
    return _op(_lib.OP_KNIFE_EDGE, offset, (1.0 if axis == "x" else 2.0) * (1.0 if direction > 0 else -1.0))


def _ops_angular_filter(Rs):
    Rs = np.asarray(Rs, dtype=np.float64)
    return [_op(_lib.OP_ANNULAR_STOP, Rs[2 * i], Rs[2 * i + 1]) for i in range(len(Rs) // 2)]


_kernel_events = None      # measurement aid (bench.py): a list collects the (start, end) CUDA events of every launch below


def _run(r_dev, program, pos_scale=1.0, perm=None, hist=None, want_rf=True, weights=None, Hw=None):
    """Launch the fused kernel.  hist = (xedges_dev, yedges_dev, H_dev) or None; weights/Hw: optional
    per-ray weights and the FP64 image they are summed into."""
    torch = _lib.torch_cuda()
    lib = _lib.load()
    if r_dev.dim() != 2 or r_dev.shape[0] != 4:
        raise NotImplementedError("ray arrays must have shape (4, N): x, theta, y, phi")
    if len(program) > _lib.MAX_OPTICS:
        raise ValueError(f"element program longer than {_lib.MAX_OPTICS}")
    n = r_dev.shape[1]
    prog = (_lib.Optic * max(1, len(program)))()
    for k, (code, a, b) in enumerate(program):
        prog[k].op, prog[k].a, prog[k].b = code, a, b
    out = torch.empty_like(r_dev) if want_rf else None
    xe = ye = H = None
    nbx = nby = 0
    if hist is not None:
        xe, ye, H = hist
        nbx, nby = xe.numel() - 1, ye.numel() - 1
    ev = _kernel_events
    if ev is not None:
        e0 = torch.cuda.Event(enable_timing=True)
        e0.record()
    _lib.check(lib.tt_optics_hist_weighted(_lib.ptr(r_dev), n, _lib.ptr(perm), float(pos_scale), prog, len(program),
                                           _lib.ptr(xe), nbx, _lib.ptr(ye), nby, _lib.ptr(H), _lib.ptr(weights),
                                           _lib.ptr(Hw), _lib.ptr(out), _lib.stream_ptr()), "tt_optics_hist")
    if ev is not None:
        e1 = torch.cuda.Event(enable_timing=True)
        e1.record()
        ev.append((e0, e1))
    return out


def _apply(r, program, inplace=False):
    """Apply an element program to a user array, returning the same kind of array."""
    torch = _lib.torch_cuda()
    dev = _lib.to_device(r, torch.float64)
    out = _run(dev, program)
    if isinstance(r, DeviceArray):
        if inplace:
            r.torch.copy_(out)
            r._np = None
            return r
        return DeviceArray(out)
    if isinstance(r, torch.Tensor):
        if inplace:
            r.copy_(out.to(r.device, r.dtype))
            return r
        return out
    res = out.cpu().numpy()
    if inplace:
        r[...] = res
        return r
    return res


def m_to_mm(r):
    """Copy with positions (rows 0, 2) scaled m -> mm (:37-40)."""
    if isinstance(r, DeviceArray) or not isinstance(r, np.ndarray):
        torch = _lib.torch_cuda()
        t = _lib.to_device(r, torch.float64).clone()
        t[0::2] *= 1e3
        return DeviceArray(t)
    rr = np.ndarray.copy(r)
    rr[0::2, :] *= 1e3
    return rr


def lens(r, f1, f2):
    """Thin lens, focal lengths f1 and f2 in the two orthogonal axes (:42-54)."""
    return _apply(r, [_op_lens(f1, f2)])


def sym_lens(r, f):
    """Axisymmetric lens (:56-60)."""
    return lens(r, f, f)


def distance(r, d):
    """Free-space propagation over d (:62-71)."""
    return _apply(r, [_op_distance(d)])


def circular_aperture(r, R):
    """Rejects rays outside radius R, in place (:73-79)."""
    return _apply(r, [_op(_lib.OP_CIRC_APERTURE, R)], inplace=True)


def circular_stop(r, R):
    """Rejects rays inside radius R, in place (:81-87)."""
    return _apply(r, [_op(_lib.OP_CIRC_STOP, R)], inplace=True)


def annular_stop(r, R1, R2):
    """Boolean mask of the rays that fall between R1 and R2 (:89-97); r is not modified."""
    torch = _lib.torch_cuda()
    dev = _lib.to_device(r, torch.float64)
    out = _run(dev, [_op(_lib.OP_ANNULAR_STOP, R1, R2)])
    mask = torch.isnan(out[0]) & ~torch.isnan(dev[0])
    return mask if isinstance(r, (torch.Tensor, DeviceArray)) else mask.cpu().numpy()


def angular_filter(r, Rs):
    """Rejects rays inside the annuli (Rs[0],Rs[1]), (Rs[2],Rs[3]), ..., in place (:99-111)."""
    return _apply(r, _ops_angular_filter(Rs), inplace=True)


def rect_aperture(r, Lx, Ly):
    """Rejects rays outside a 2Lx x 2Ly rectangle in BOTH axes, in place (:128-136)."""
    return _apply(r, [_op(_lib.OP_RECT_APERTURE, Lx, Ly)], inplace=True)


def knife_edge(r, offset, axis, direction):
    """Knife edge in 'x' or 'y'; direction > 0 rejects above the offset, < 0 below (:138-154)."""
    return _apply(r, [_op_knife(offset, axis, direction)], inplace=True)


def plot_afr(Rs):
    """Plot the angular filter (:113-126).  matplotlib is imported lazily (plotting is outside the
    CUDA path)."""
    import matplotlib as mpl
    import matplotlib.pyplot as plt
    fig, ax = plt.subplots(figsize=(4, 4), dpi=200)
    for i in range(0, len(Rs) // 2):
        R1, R2 = Rs[2 * i], Rs[2 * i + 1]
        dR = R2 - R1
        an = mpl.patches.Circle((0, 0), R2) if dR >= R2 else mpl.patches.Annulus((0, 0), R2, dR)
        ax.add_patch(an)
    ax.set_xlim([-Rs.max(), Rs.max()])
    ax.set_ylim([-Rs.max(), Rs.max()])
    return fig, ax


_EDGE_CACHE = {}


def _edges_on_device(lo, hi, n):
    """numpy.histogramdd builds its bin edges as linspace(range_min, range_max, bins + 1); they are
    computed on the host by numpy itself and cached on the device (no per-call upload / host sync)."""
    torch = _lib.torch_cuda()
    key = (float(lo), float(hi), int(n), torch.cuda.current_device())
    hit = _EDGE_CACHE.get(key)
    if hit is None:
        e = np.linspace(lo, hi, n + 1)
        if len(_EDGE_CACHE) > 64:
            _EDGE_CACHE.clear()
        hit = _EDGE_CACHE[key] = (e, torch.from_numpy(e).cuda())
    return hit


class OpticsProgram:
    """A user-composed chain of the reference's optical elements that runs as ONE fused kernel pass
    (instead of one array pass per element function):

        prog = OpticsProgram().distance(400).circular_aperture(25).lens(400, 200).knife_edge(0.5, 'y', 1)
        det = Rays(rf); det.solve_program(prog.distance(400)); det.histogram()

    Each method appends the element with the semantics of the function of the same name
    (ray_transfer_matrix.py:42-154)."""

    def __init__(self, ops=None):
        self.ops = list(ops or [])

    def _add(self, *ops):
        return OpticsProgram(self.ops + list(ops))

    def distance(self, d):
        return self._add(_op_distance(d))

    def lens(self, f1, f2):
        return self._add(_op_lens(f1, f2))

    def sym_lens(self, f):
        return self._add(_op_lens(f, f))

    def circular_aperture(self, R):
        return self._add(_op(_lib.OP_CIRC_APERTURE, R))

    def circular_stop(self, R):
        return self._add(_op(_lib.OP_CIRC_STOP, R))

    def angular_filter(self, Rs):
        return self._add(*_ops_angular_filter(Rs))

    def rect_aperture(self, Lx, Ly):
        return self._add(_op(_lib.OP_RECT_APERTURE, Lx, Ly))

    def knife_edge(self, offset, axis, direction):
        return self._add(_op_knife(offset, axis, direction))

    def __len__(self):
        return len(self.ops)


# ---- detectors --------------------------------------------------------------------------------------
class Rays:
    """Inheritable class for ray diagnostics (:156-206)."""

    def __init__(self, r0, focal_plane=0, L=400, R=25, Lx=18, Ly=13.5):
        """r0: 4 x N rays [x, theta, y, phi] in m / rad; L: length scale (first lens at L), R: lens
        radius, Lx, Ly: detector size in mm (:160-172)."""
        self.focal_plane, self.L, self.R, self.Lx, self.Ly = focal_plane, L, R, Lx, Ly
        torch = _lib.torch_cuda()
        # Rays are visited in their stored order (coalesced 32 B/ray).  Visiting them in the Morton order
        # of the trace (r0.perm) would keep a CTA's rays in one shared-memory histogram window, but the
        # permuted gather costs a 128-byte DRAM transaction per 8-byte element: measured on B200 with
        # 1e8 rays, 7.8 ms against 2.3 ms for the streaming order.  Opt in with ``use_ray_order(perm)``.
        self._perm = None
        # the reference copies at construction (self.r0 = m_to_mm(r0), :172): a device input is cloned so that later
        # in-place edits of the caller's array (e.g. cube.rf) cannot change this detector's image (32 B/ray, ~1 ms
        # for 1e8 rays); host inputs are copied by the upload anyway
        t = _lib.to_device(r0, torch.float64)
        if isinstance(r0, (DeviceArray, torch.Tensor)) and t.data_ptr() == (r0.torch if isinstance(r0, DeviceArray) else r0).data_ptr():
            t = t.clone()
        self._r0_m = t                                      # metres; m_to_mm happens in the kernel
        self._r0_mm = None
        self._program = None
        self._rf = None
        self.H_dev = None

    def _perm_dev(self):
        return None if self._perm is None else _lib.to_device(self._perm)

    def use_ray_order(self, perm):
        """Visit the rays in the order perm[0..N) (int32 device tensor, e.g. ``cube.rf.perm``)."""
        self._perm = perm

    # r0 in mm, as the reference stores it (:172)
    @property
    def r0(self):
        if self._r0_m is None:
            return None
        if self._r0_mm is None:
            t = _lib.to_device(self._r0_m).clone()
            t[0::2] *= 1e3
            self._r0_mm = DeviceArray(t)
        return self._r0_mm

    @r0.setter
    def r0(self, v):
        if v is None:
            self._r0_m = self._r0_mm = None
            return
        torch = _lib.torch_cuda()
        t = _lib.to_device(v, torch.float64).clone()
        t[0::2] *= 1e-3
        self._r0_m, self._r0_mm = t, None

    def _set_program(self, program):
        self._program = list(program)
        self._rf = None

    def solve_program(self, program):
        """Custom detector: propagate r0 (mm) through a user-composed :class:`OpticsProgram`."""
        self._set_program(program.ops if isinstance(program, OpticsProgram) else program)

    @property
    def rf(self):
        if self._rf is None and self._program is not None and self._r0_m is not None:
            self._rf = DeviceArray(_run(_lib.to_device(self._r0_m), self._program, pos_scale=1e3, perm=self._perm_dev()))
        return self._rf

    @rf.setter
    def rf(self, v):
        self._rf = v if (v is None or isinstance(v, DeviceArray)) else DeviceArray(_lib.to_device(v))
        if v is None:
            self._program = None

    def histogram(self, bin_scale=10, pix_x=3448, pix_y=2574, clear_mem=False, to_host=True, weights=None):
        """Bin detector-plane rays; defaults are for a KAF-8300 (:173-195).  Sets ``H``
        (pix_y//bin_scale, pix_x//bin_scale) float64 like numpy.histogram2d(...).T, ``xedges``,
        ``yedges``; ``H_dev`` keeps the integer counts on the device (for NCCL all-reduce).
        ``to_host=False`` skips the device->host copy of H (multi-GPU drivers reduce H_dev first).
        ``weights`` (N,): additionally sums the weights of the binned rays into ``Hw`` / ``Hw_dev`` (FP64),
        like ``numpy.histogram2d(..., weights=)`` in example_kitchensink.py:108-129 (amplitude- or
        polarisation-weighted images; the weights refer to the rays in their original order)."""
        torch = _lib.torch_cuda()
        nbx, nby = pix_x // bin_scale, pix_y // bin_scale
        self.xedges, xe = _edges_on_device(-self.Lx / 2, self.Lx / 2, nbx)
        self.yedges, ye = _edges_on_device(-self.Ly / 2, self.Ly / 2, nby)
        H = torch.zeros((nby, nbx), dtype=torch.int64, device="cuda")
        w = Hw = None
        if weights is not None:
            w = _lib.to_device(weights, torch.float64).reshape(-1)
            Hw = torch.zeros((nby, nbx), dtype=torch.float64, device="cuda")
        if self._rf is not None:                      # rays already at the detector plane
            rfd = _lib.to_device(self._rf, torch.float64)
            if w is not None and w.numel() != rfd.shape[1]:
                raise ValueError("weights must have one entry per ray")
            _run(rfd, [], perm=self._perm_dev(), hist=(xe, ye, H), want_rf=False, weights=w, Hw=Hw)
        elif self._program is not None and self._r0_m is not None:   # fused: optics + binning, one pass
            r0d = _lib.to_device(self._r0_m, torch.float64)
            if w is not None and w.numel() != r0d.shape[1]:
                raise ValueError("weights must have one entry per ray")
            _run(r0d, self._program, pos_scale=1e3, perm=self._perm_dev(), hist=(xe, ye, H), want_rf=False,
                 weights=w, Hw=Hw)
        else:
            raise AttributeError("no rays to bin: call solve() first")
        self.H_dev = H
        self.H = H.double().cpu().numpy() if to_host else None
        self.Hw_dev = Hw
        self.Hw = Hw.cpu().numpy() if (Hw is not None and to_host) else None
        if clear_mem:
            self.clear_rays()

    def plot(self, ax, clim=None, cmap=None):
        ax.imshow(self.H, interpolation="nearest", origin="lower", clim=clim, cmap=cmap,
                  extent=[self.xedges[0], self.xedges[-1], self.yedges[0], self.yedges[-1]])

    def clear_rays(self):
        """Clears the r0 and rf variables to save memory (:201-206)."""
        self._r0_m = self._r0_mm = None
        self._rf = None
        self._program = None
        self._perm = None

    def __getstate__(self):          # detectors stay picklable (example_MPI.py:152-166); device data -> numpy
        d = dict(self.__dict__)
        for k in ("_r0_m", "_r0_mm", "_rf", "_perm", "H_dev", "Hw_dev"):
            v = d.get(k)
            if v is not None:
                d[k] = np.asarray(v.detach().cpu() if hasattr(v, "detach") else v)
        return d

    def __setstate__(self, d):
        # host arrays are kept as they are (a pickle can be opened without a GPU, e.g. for plotting H);
        # they are uploaded again on first use
        self.__dict__.update(d)


class Shadowgraphy(Rays):
    """Two-lens M = 1 telescope, apertures of radius R at both lenses (:208-227)."""

    def solve(self, displacement=None):
        """``displacement`` is the keyword of the reference's stale example scripts
        (``sh.solve(displacement=0)``, example_MPI.py:71-72, example_multiprocess.py:69): the object plane's offset
        from the focal plane of the first lens in mm, i.e. what the checked-in class takes as ``focal_plane`` at
        construction (:160-172).  None keeps the constructor's value."""
        L, R = self.L, self.R
        if displacement is not None:
            self.focal_plane = displacement
        self._set_program([
            _op_distance(L - self.focal_plane), _op(_lib.OP_CIRC_APERTURE, R), _op_lens(L, L),
            _op_distance(L * 2),
            _op(_lib.OP_CIRC_APERTURE, R), _op_lens(L, L),
            _op_distance(L)])


class Schlieren_DF(Rays):
    """Dark-field schlieren: circular stop of radius R [mm] at the focal plane (:229-250)."""

    def solve(self, R=1):
        L, Ra = self.L, self.R
        self._set_program([
            _op_distance(L - self.focal_plane), _op(_lib.OP_CIRC_APERTURE, Ra), _op_lens(L, L),
            _op_distance(L), _op(_lib.OP_CIRC_STOP, R),
            _op_distance(L), _op(_lib.OP_CIRC_APERTURE, Ra), _op_lens(L, L),
            _op_distance(L)])


class Schlieren_LF(Rays):
    """Light-field schlieren: circular aperture of radius R [mm] at the focal plane (:252-273)."""

    def solve(self, R=1):
        L, Ra = self.L, self.R
        self._set_program([
            _op_distance(L - self.focal_plane), _op(_lib.OP_CIRC_APERTURE, Ra), _op_lens(L, L),
            _op_distance(L), _op(_lib.OP_CIRC_APERTURE, R),
            _op_distance(L), _op(_lib.OP_CIRC_APERTURE, Ra), _op_lens(L, L),
            _op_distance(L)])


class AFR(Rays):
    """Angular filter refractometer: annular stops Rs at the focal plane (:275-299)."""

    def solve(self, Rs):
        L, Ra = self.L, self.R
        self._set_program([
            _op_distance(L / 2 - self.focal_plane), _op(_lib.OP_CIRC_APERTURE, Ra), _op_lens(L / 2, L / 2),
            _op_distance(L / 4), *_ops_angular_filter(Rs),
            _op_distance(L / 4), _op(_lib.OP_CIRC_APERTURE, Ra), _op_lens(L / 2, L / 2),
            _op_distance(L / 2)])


class Refractometer(Rays):
    """Imaging refractometer ("Burdiscope"): the detector shows POSITION along x and ray ANGLE along y.

    PARITY UNPINNED: the reference checkout only has call sites (``rtm.BurdiscopeRays(rf)``,
    example_MPI.py:67-69, example_multiprocess.py:65) and the building block ``lens(r, f1, f2)`` with
    independent focal lengths (ray_transfer_matrix.py:42-54); this program is composed here from those
    elements.  A 4f relay (as Shadowgraphy) forms an image at L behind lens 2; a hybrid (cylindrical)
    lens L/2 further on with f_x = L/3, f_y = L images that plane onto the detector in x
    (magnification -2: x_det = 2 x0 - ...) and puts the detector in its focal plane in y
    (y_det = -L * phi0), L behind it.  Apertures of radius R at the relay lenses, a rectangular
    aperture (R, R) at the hybrid lens."""

    def solve(self):
        L, R = self.L, self.R
        self._set_program([
            _op_distance(L - self.focal_plane), _op(_lib.OP_CIRC_APERTURE, R), _op_lens(L, L),
            _op_distance(L * 2),
            _op(_lib.OP_CIRC_APERTURE, R), _op_lens(L, L),
            _op_distance(L + L / 2), _op(_lib.OP_RECT_APERTURE, R, R), _op_lens(L / 3, L),
            _op_distance(L)])


# names used by the reference's (stale) example scripts: example_MPI.py:67-77, example_multiprocess.py:65-67
ShadowgraphyRays = Shadowgraphy
SchlierenRays = Schlieren_DF
BurdiscopeRays = Refractometer
