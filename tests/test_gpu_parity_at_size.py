"""GPU parity at BASELINE-config sizes (run with -m gpu on the B200 box).

tests/test_gpu_parity.py compares the device with the numpy oracle on hierarchies of <= 33^3 points, where AUTO
never picks the production kernels (SELL-C-sigma in persistent CTAs needs >= 200k rows per part).  Here the PRODUCT
host setup builds the hierarchy at 64^3 ... 128^3 (config 2 of BASELINE.json), the device runs it with AUTO formats,
and the C oracle (oracle/pamg_oracle.c, the OpenMP restatement of amg_oracle.py) runs the SAME operators, copied out
through the C ABI queries.  Bars (north star): V-cycle <= 1e-12 relative, identical PCG iteration counts to 1e-8,
residual histories to 1e-7.  The setup itself is checked bit-exact against the numpy oracle in test_setup_parity.py."""
import numpy as np
import pytest

import c_oracle
from parallel_amg_b200 import _lib as L
from util import det_vector, rel_err

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600, method="thread")]

TOL_VCYCLE = 1e-12   # north_star: per V-cycle, fp64


def _context(kind, dims, pp, opts):
    nparts = int(np.prod(pp))
    c = L.Context(nparts)
    if kind == "poisson":
        c.gallery_poisson(dims, pp)
    elif kind == "jump":
        c.gallery_diffusion_jump(dims, pp, blocks=8, kmax=1.0e4, eps_z=1.0e-3)
    else:
        c.gallery_elasticity(dims, pp)
    c.setup(c.default_options(**opts))
    c.device_init()
    return c, nparts


# kind, dims, parts, options, format expected for A on level 0 (None: do not care)
CASES = [
    ("poisson", (64, 64, 64), (1, 1, 1), {}, L.FORMAT_SELL),        # 262k rows: the persistent stride loop makes > 1 pass
    ("poisson", (128, 128, 128), (1, 1, 1), {}, L.FORMAT_SELL),     # BASELINE config 2
    ("poisson", (128, 128, 128), (2, 1, 1), {}, L.FORMAT_SELL),     # two 1M-row parts: SELL + halo roles + replicated tail
    ("poisson", (96, 80, 72), (2, 2, 1), {}, None),                 # ragged sizes, 4 parts, CSR-stream / SELL mix
    ("elasticity", (24, 24, 24), (2, 1, 1), {}, None),              # config 4 operator, 41k DOFs, 2 parts
    ("elasticity", (42, 42, 42), (1, 1, 1), {}, L.FORMAT_SELL),     # 222k DOFs x 79 nnz/row: long-row SELL
    ("jump", (48, 48, 48), (2, 2, 2), dict(eps_strength=0.08), None),   # config 5 operator, 8 parts, filtered strength
    ("jump", (64, 64, 64), (1, 1, 1), dict(eps_strength=0.08), L.FORMAT_SELL),
]


@pytest.mark.parametrize("kind,dims,pp,opts,fmt", CASES, ids=[f"{k}-{'x'.join(map(str, d))}-p{int(np.prod(p))}" for k, d, p, o, f in CASES])
def test_auto_production_path_matches_c_oracle(kind, dims, pp, opts, fmt):
    c, nparts = _context(kind, dims, pp, opts)
    st = c.stats()
    if fmt is not None:
        assert st.format[0] == fmt, "AUTO did not select the production kernel family this test is about"
    co = c_oracle.COracle.from_product_context(c, nparts)
    n, _ = c.global_size()
    own = [c.index_maps(0, p)[0] for p in range(nparts)]
    # V-cycle (the preconditioner apply) on two right-hand sides
    for seed in (71, 72):
        b = det_vector(n, seed)
        z_ref = co.vcycle([b[o] for o in own])
        z = c.vcycle([b[o] for o in own])
        assert rel_err(z, z_ref) <= TOL_VCYCLE
    # AMG-PCG: identical iteration count, same residual history, same solution
    rhs = c.host_matvec_global(det_vector(n, 1))
    x_ref, it_ref, hist_ref = co.pcg([rhs[o] for o in own], 1e-8, 400, True)
    for rep in range(2):   # the second solve replays the captured graph
        x, it, hist, ok = c.pcg([rhs[o] for o in own], rtol=1e-8, maxiter=400)
        assert ok and it == it_ref, (it, it_ref)
        assert np.allclose(hist, hist_ref, rtol=1e-7)
        assert rel_err(x, x_ref) <= 1e-9
    xg = np.zeros(n)
    for o, xp in zip(own, x):
        xg[o] = xp
    assert np.linalg.norm(c.host_matvec_global(xg) - rhs) <= 1.001e-8 * np.linalg.norm(rhs)
    co.close()
    c.close()


def test_l1_jacobi_and_two_sweeps_at_size():
    """Non-default smoothing options through the same production kernels (64^3, one part)."""
    for opts, kw in ((dict(smoother=L.SMOOTHER_L1JACOBI), dict(smoother="l1jacobi")), (dict(nu_pre=2, nu_post=2), dict(nu_pre=2, nu_post=2))):
        c, nparts = _context("poisson", (64, 64, 64), (2, 1, 1), opts)
        co = c_oracle.COracle.from_product_context(c, nparts, **kw)
        n, _ = c.global_size()
        own = [c.index_maps(0, p)[0] for p in range(nparts)]
        b = det_vector(n, 5)
        assert rel_err(c.vcycle([b[o] for o in own]), co.vcycle([b[o] for o in own])) <= TOL_VCYCLE
        x_ref, it_ref, hist_ref = co.pcg([b[o] for o in own], 1e-8, 200, True)
        x, it, hist, ok = c.pcg([b[o] for o in own])
        assert ok and it == it_ref and np.allclose(hist, hist_ref, rtol=1e-7)
        co.close()
        c.close()


@pytest.mark.parametrize("kind,dims,pp,opts", [("poisson", (96, 96, 96), (2, 1, 1), {}), ("jump", (64, 64, 64), (1, 1, 1), dict(eps_strength=0.08))],
                         ids=["poisson-96-p2", "jump-64-p1"])
def test_fgmres_at_size_matches_c_oracle(kind, dims, pp, opts):
    """pamg_fgmres through the production kernels (SELL, persistent CTAs) against orc_fgmres on the same operators: identical
    inner iteration counts with and without a restart inside the solve, estimates to 1e-7, and the true residual."""
    c, nparts = _context(kind, dims, pp, opts)
    co = c_oracle.COracle.from_product_context(c, nparts)
    n, _ = c.global_size()
    own = [c.index_maps(0, p)[0] for p in range(nparts)]
    rhs = c.host_matvec_global(det_vector(n, 1))
    for restart in (30, 6):
        x_ref, it_ref, hist_ref = co.fgmres([rhs[o] for o in own], 1e-8, 400, restart, True)
        x, it, hist, ok = c.fgmres([rhs[o] for o in own], rtol=1e-8, maxiter=400, restart=restart)
        assert ok and it == it_ref, (it, it_ref)
        assert np.allclose(hist, hist_ref, rtol=1e-7)
        assert rel_err(x, x_ref) <= 1e-9
        xg = np.zeros(n)
        for o, xp in zip(own, x):
            xg[o] = xp
        assert np.linalg.norm(c.host_matvec_global(xg) - rhs) <= 1.01e-8 * np.linalg.norm(rhs)
    co.close()
    c.close()
