import sys
sys.path[:0] = ["/root/repo", "/root/repo/assignment-for-aae6102_gnss-sdr_b200", "/root/repo/tests"]
import test_gpu_property as T
from hypothesis import settings, HealthCheck, seed
fn = T.test_random_satellite_matches_oracle
inner = fn.hypothesis.inner_test
from hypothesis import given, strategies as st
N = T.N
for sd in range(6):
    g = seed(sd)(settings(max_examples=150, deadline=None, suppress_health_check=list(HealthCheck), database=None)(
        given(prn=st.integers(1, 32), delay=st.integers(0, N - 1), dopp=st.floats(-9900.0, 9900.0),
              amp=st.floats(0.5, 6.0), seed=st.integers(0, 2 ** 20), data_type=st.sampled_from([1, 2]),
              skip_ms=st.integers(0, 40))(inner)))
    try:
        g(); print("seed", sd, "ok", flush=True)
    except BaseException as e:
        print("seed", sd, "FAILED:", repr(e)[:1500], flush=True)
