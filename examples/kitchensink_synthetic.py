"""The reference's example_kitchensink.py flow (particle_tracking/example_kitchensink.py:38-134) on the
B200 backend, with synthetic data instead of the (missing) simulation files: ne + B cubes are written as
.pvti files and read back with ``pvti_readin`` like the example does (:7-36, no vtk needed), probing along
y, phase + Faraday rotation, amplitude- and polarisation-weighted detector images.

    python examples/kitchensink_synthetic.py [Np]
"""
import os
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from turbulence_tracing_b200 import particle_tracker as pt      # noqa: E402
from turbulence_tracing_b200 import ray_transfer_matrix as rtm  # noqa: E402
from turbulence_tracing_b200 import turboGen as tg              # noqa: E402
from turbulence_tracing_b200.io import pvti_readin, write_pvti, centred_axes   # noqa: E402

Np = int(float(sys.argv[1])) if len(sys.argv) > 1 else int(1e5)
N = 32                                        # cube of (2N+1)^3 = 65^3 points
M = 2 * N + 1
ne_extent = 5e-3

f = tg.gaussian3D_FFT(N, lambda k: k ** (-11.0 / 3.0), seed=1)
rnec = 1e25 * np.clip(1 + 0.3 * f / f.std(), 0, None)
Bvec = np.zeros((M, M, M, 3))
Bvec[..., 1] = 10.0                           # 10 T along the probing direction
Bvec[..., 0] = 2.0 * tg.gaussian3D_FFT(N, lambda k: k ** (-11.0 / 3.0), seed=2)


# "simulation output" on disk, then the example's loading step (example_kitchensink.py:43-62)
with tempfile.TemporaryDirectory() as d:
    h = 2 * ne_extent / (M - 1)
    write_pvti(os.path.join(d, "x08_rnec-400.pvti"), rnec, spacing=(h, h, h), name="rnec", pieces=(2, 2, 1), compress=True)
    write_pvti(os.path.join(d, "x08_Bvec-400.pvti"), Bvec, spacing=(h, h, h), name="Bvec", pieces=(2, 1, 1))
    rnec, dim, spacing = pvti_readin(os.path.join(d, "x08_rnec-400.pvti"))
    Bvec, dim, spacing = pvti_readin(os.path.join(d, "x08_Bvec-400.pvti"))
ne_x, ne_y, ne_z = centred_axes(dim, spacing)
assert ne_y[-1] == ne_extent or abs(ne_y[-1] - ne_extent) < 1e-15

test = pt.ElectronCube(ne_x, ne_y, ne_z, ne_extent, B_on=True, inv_brems=False, phaseshift=True,
                       probing_direction="y")                 # the call of example_kitchensink.py:72, verbatim
test.external_ne(rnec)
test.external_B(Bvec)
test.calc_dndr()
test.set_up_interps()
test.clear_memory()

beam_size, divergence = 4e-3, 0.05e-3
np.random.seed(0)
ss = pt.init_beam(Np=Np, beam_size=beam_size, divergence=divergence, ne_extent=ne_extent, probing_direction="y")
rf = test.solve(ss)                           # (x, theta, y, phi) at the exit plane
Jf = test.Jf

amp = np.sqrt(np.abs(Jf[0, :] ** 2 + Jf[1, :] ** 2))
aEy = np.arctan(np.real(Jf[0, :] / Jf[1, :]))

det = rtm.ShadowgraphyRays(rf, L=400, R=25, Lx=12, Ly=12)
det.solve()
det.histogram(bin_scale=25, weights=amp)
amp_hist = det.Hw
det.histogram(bin_scale=25, weights=aEy * amp)
rot_hist = np.divide(det.Hw, amp_hist, out=np.zeros_like(amp_hist), where=amp_hist > 0)
print(f"{Np} rays, {test.ray_steps} ray-steps; accepted {int(det.H.sum())}; "
      f"mean Faraday rotation {np.mean(aEy) * 1e3:.3f} mrad; image rotation range "
      f"[{rot_hist[amp_hist > 0].min() * 1e3:.3f}, {rot_hist[amp_hist > 0].max() * 1e3:.3f}] mrad")
expected = pt.VERDET * 1053e-9**2 * 1e25 * 10.0 * 2 * ne_extent
print(f"uniform-plasma estimate V*ne*B*L = {expected * 1e3:.3f} mrad")
