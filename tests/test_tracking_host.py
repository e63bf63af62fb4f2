"""Host logic of the trackingCT twin (gnssacq/tracking.py): C/N0 estimator and bit-edge index, CPU only."""
import numpy as np

from gnssacq.tracking import bit_edge_index, cn0_estimates


def test_cn0_follows_the_reference_formula():
    rng = np.random.default_rng(3)
    a, sigma, n = 2000.0, 300.0, 60                         # prompt amplitude and per-component noise of the sums
    p_i = a + sigma * rng.standard_normal(n)
    p_q = sigma * rng.standard_normal(n)
    got = cn0_estimates(p_i, p_q)
    assert got.shape == (3,)
    z = (p_i ** 2 + p_q ** 2)[:20]                          # trackingCT.m:121-133 written out for the first block
    m, v = z.mean(), z.var(ddof=1)
    na2 = np.sqrt(m * m - v)
    want = abs(10 * np.log10(1 / 1e-3 * na2 / (2 * (0.5 * (m - na2)))))
    assert got[0] == want
    truth = 10 * np.log10(a * a / (2 * sigma * sigma) / 1e-3)            # C/N0 = (A^2 / 2 sigma^2) / T
    assert abs(np.median(got) - truth) < 2.0


def test_bit_edge_index():
    p = np.ones(1000)
    assert bit_edge_index(p) == 0                           # no transition: countinx stays 0 (trackingCT.m:20)
    for edge in (601, 615, 640, 777):                       # 1-based index of the first period with the new sign
        q = np.ones(1000)
        q[edge - 1:] = -1.0
        assert bit_edge_index(q) == edge % 20 - 1           # :207
    q = np.ones(1000)
    q[299:] = -1.0                                          # before i = 600: ignored (:205)
    q[659:] = 1.0
    assert bit_edge_index(q) == 660 % 20 - 1
    q = np.ones(1000)
    q[619] = -1.0                                           # a single outlier is not an edge
    assert bit_edge_index(q) == 0
    r = np.ones(1000)
    r[629:] = -1.0
    r[633] = 1.0                                            # the 17 successors must all agree
    assert bit_edge_index(r) == 0
