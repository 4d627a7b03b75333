"""include/mkf_expf.h -- the single-precision exp of the legacy filter's GMM prior (src/pf2D.cpp:105-109) -- against
the host libm's expf.  The header restates glibc's algorithm (FMA build) so that the oracle and the device share one
definition; these tests pin it to the libm this image ships (glibc 2.39): a dense sample by default, every float bit
pattern with MKF_EXPF_EXHAUSTIVE=1 (tools/check_expf.c is the standalone version; profiles/r02_expf_exhaustive.json
holds its result)."""
import os

import numpy as np
import pytest

import mkf_oracle as orc


def test_expf_known_values():
    assert orc.expf(0.0) == 1.0
    assert orc.expf(-np.inf) == 0.0
    assert orc.expf(np.inf) == np.inf
    assert np.isnan(orc.expf(np.nan))
    assert orc.expf(-104.0) == 0.0 and orc.expf(89.0) == np.inf
    assert orc.expf(1.0) == float(np.float32(np.e))
    # the two arguments on which glibc's FMA and non-FMA builds differ (the header follows the FMA build)
    for x in (float.fromhex("-0x1.f8cbb2p+5"), float.fromhex("0x1.04845ep+5")):
        assert orc.expf(x) == orc.libm_expf(x)
    # subnormal results
    assert orc.expf(-100.0) == orc.libm_expf(-100.0) and 0.0 < orc.expf(-100.0) < 1.2e-38


def test_expf_matches_libm_on_dense_sample():
    rng = np.random.default_rng(2026)
    # the range src/pf2D.cpp:108 can produce: -q/2 with q >= 0; below -104 the result is 0
    neg = np.arange(np.float32(-0.0).view(np.uint32), np.float32(-104.5).view(np.uint32), 257, dtype=np.uint64)
    pos = np.arange(0, np.float32(89.5).view(np.uint32), 1031, dtype=np.uint64)
    rnd = rng.integers(0, 1 << 32, 2_000_000, dtype=np.uint64)
    bits = np.concatenate([neg, pos, rnd]).astype(np.uint32)
    assert orc.expf_compare(bits) == 0


@pytest.mark.skipif(os.environ.get("MKF_EXPF_EXHAUSTIVE") != "1", reason="set MKF_EXPF_EXHAUSTIVE=1 (about a minute)")
def test_expf_matches_libm_exhaustively():
    bad = 0
    for hi in range(256):
        bits = (np.arange(1 << 24, dtype=np.uint64) + (hi << 24)).astype(np.uint32)
        bad += orc.expf_compare(bits)
    assert bad == 0
