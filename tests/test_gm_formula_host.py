"""CPU check of the arithmetic of the mixture-density kernel (elfi_b200/csrc/smc.cu, gm_pdf_kernel):
the centred / expanded squared distance with the folded log-weight and the degree-6 minimax
2^f are restated in NumPy term by term and compared with a float128 evaluation of
GMDistribution.logpdf (elfi/methods/utils.py:174-197).  This pins the accuracy the header claims
(< 2e-9 relative per term) independently of the device; tests/test_smc_gpu.py checks the CUDA
build against the reference's goldens."""
import numpy as np
import pytest

COEF = [1.5345812158740182e-04, 1.3399931209474140e-03, 9.6184889565227916e-03,
        5.5503287769976638e-02, 2.4022646890639572e-01, 6.9314720573725268e-01,
        1.0000000005541663e+00]
SCALE = 0.8493218002880191   # sqrt(log2(e) / 2)


def exp2_neg(nt):
    with np.errstate(invalid='ignore'):      # nt = +inf for zero-weight components
        k = np.rint(-nt)
        f = -nt - k
    pz = np.full_like(nt, COEF[0])
    for c in COEF[1:]:
        pz = pz * f + c
    with np.errstate(over='ignore', invalid='ignore'):
        return np.where(nt <= 1020.0, np.ldexp(pz, np.maximum(k, -1100).astype(np.int64)), 0.0)


def kernel_logpdf(x, means, cov, w):
    L = np.linalg.cholesky(cov)
    Linv = np.linalg.inv(L)
    p = x.shape[1]
    centre = means[0]
    y = (x - centre) @ Linv.T * SCALE
    m = (means - centre) @ Linv.T * SCALE
    with np.errstate(divide='ignore'):
        cj = np.sum(m * m, axis=1) - np.log2(w / w.sum())
    g = cj[None, :] + y @ (-2.0 * m).T
    nt = g + np.sum(y * y, axis=1)[:, None]
    acc = exp2_neg(nt).sum(axis=1)
    lognorm = -0.5 * (p * np.log(2 * np.pi) + 2 * np.sum(np.log(np.diag(L))))
    return np.log(acc) + lognorm


def exact_logpdf(x, means, cov, w):
    ld = np.longdouble
    prec = np.linalg.inv(cov).astype(ld)
    d = x[:, None, :].astype(ld) - means[None, :, :].astype(ld)
    maha = np.einsum('ijk,kl,ijl->ij', d, prec, d)
    wn = (w / w.sum()).astype(ld)
    p = x.shape[1]
    dens = (wn[None, :] * np.exp(-maha / 2)).sum(axis=1)
    return (np.log(dens) - ld(0.5) * (p * np.log(ld(2) * np.pi) + np.log(ld(np.linalg.det(cov))))
            ).astype(np.float64)


@pytest.mark.parametrize('p,loc,sd', [(2, 0.5, 0.2), (2, 0.6, 0.004), (1, -3.0, 1.5), (4, 5.0, 0.05),
                                      (3, 1e3, 1e-2), (2, 0.0, 30.0)])
def test_kernel_arithmetic_matches_float128(p, loc, sd):
    rs = np.random.RandomState(p * 7 + int(sd * 1000) % 97)
    M, N = 3000, 400
    means = loc + sd * rs.randn(M, p) * rs.uniform(0.5, 2.0, p)
    w = rs.rand(M) ** 3
    w[::17] = 0.0                                    # zero weights never contribute
    A = rs.randn(p, p) * 0.3 + np.eye(p)
    cov = 2 * sd ** 2 * (A @ A.T)
    x = np.vstack([means[rs.choice(M, N - 40)] + np.sqrt(2) * sd * rs.randn(N - 40, p),
                   loc + 8 * sd * rs.randn(40, p)])  # incl. points far in the tails
    got = kernel_logpdf(x, means, cov, w)
    want = exact_logpdf(x, means, cov, w)
    ok = np.isfinite(want) & (want > -600)
    assert ok.sum() > N // 2
    np.testing.assert_allclose(np.exp(got[ok] - want[ok]), 1.0, rtol=5e-9)
