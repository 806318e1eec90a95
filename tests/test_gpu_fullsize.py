"""Full-size checks at BASELINE.json's C3 graph (N = 500k, average degree 32, K = 300), where the oracle itself would
take minutes: size-independent properties instead.

* the three SpMM gather engines accumulate a row in CSR order, so their results agree BIT FOR BIT at full size
  (the L2-panel engine, the bulk-copy engine and the register-gather engine are independent kernels);
* a sample of rows equals the float64 SciPy product to fp32 rounding;
* linearity: A.(2 H1 + H2) = 2 A.H1 + A.H2 to fp32 rounding;
* A_hat.1 equals the row sums of A_hat.
"""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_c3_spmm_engines_agree_bitwise_and_match_scipy_on_a_sample():
    import torch
    from geographconv_b200 import synth
    from geographconv_b200.engine import DeviceCsr, HostCsr
    from geographconv_b200.layers import get_dev
    from geographconv_b200.partition import ld_of
    n, K = 500_000, 300
    ld = ld_of(K)
    A = synth.synthetic_graph(n, 32, 77)
    d = get_dev()
    csr = DeviceCsr(d, HostCsr(A, 1024), 0)
    d.ensure_ws(max(d.ctx.lib.gcnb_spmm_workspace_bytes(C.byref(csr.struct), K), 1 << 20))
    g = torch.Generator(device="cuda").manual_seed(1)
    H1 = torch.randn(n, ld, device=d.dev, generator=g)
    H2 = torch.randn(n, ld, device=d.dev, generator=g)
    H1[:, K:] = 0
    H2[:, K:] = 0
    ones = torch.zeros(n, ld, device=d.dev)
    ones[:, :K] = 1
    out = {}

    def run(engine, B, name):
        csr.struct.engine = engine
        Cb = torch.zeros(n, ld, device=d.dev)
        d.fence()
        d.ctx.call("gcnb_spmm_csr_f32", C.byref(csr.struct), C.c_void_p(B.data_ptr()), ld, C.c_void_p(Cb.data_ptr()), ld, K,
                   None)
        d.ctx.sync()
        out[name] = Cb
        return Cb

    assert d.ctx.lib.gcnb_spmm_engine_for(d.ctx.h, C.byref(csr.struct), ld, K) in (0, 1, 2)
    csr.struct.engine = -2
    assert d.ctx.lib.gcnb_spmm_engine_for(d.ctx.h, C.byref(csr.struct), ld, K) == 2   # the per-call choice at C3
    p2, p1, p0 = run(2, H1, "panel"), run(1, H1, "bulk"), run(0, H1, "ldg")
    assert torch.equal(p2, p1) and torch.equal(p2, p0)
    assert not bool(p2[:, K:].any())                                                  # padding columns stay zero
    rows = np.random.RandomState(0).choice(n, size=2000, replace=False)
    want = (A[rows].astype(np.float64) @ H1[:, :K].double().cpu().numpy())
    got = p2[torch.from_numpy(rows).to(d.dev)][:, :K].cpu().numpy()
    np.testing.assert_allclose(got, want, rtol=1e-3, atol=1e-5 * float(np.abs(want).max()))
    lin = run(2, 2.0 * H1 + H2, "lin")
    rhs = 2.0 * p2 + run(2, H2, "h2")
    assert float((lin - rhs).abs().max()) <= 1e-5 * float(rhs.abs().max())
    rs = run(2, ones, "ones")[:, 0].cpu().numpy()
    np.testing.assert_allclose(rs, np.asarray(A.sum(axis=1)).reshape(-1), rtol=1e-5)
