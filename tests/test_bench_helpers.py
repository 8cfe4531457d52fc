"""CPU test of bench.py's workload builders (base selection, candidate list through the oracle stages)
so that the benchmark code paths stay importable and sane without a GPU."""
import numpy as np

import bench


def test_candidate_list_from_oracle_stages():
    raw, P, Q, cp, cq = bench.build_workload(20000)
    T, mix = bench.make_candidates(128, P, Q, cp, cq, 7, bench.OracleStages())
    assert T.shape == (128, 16) and T.dtype == np.float32
    assert mix["near_gt"] == bench.N_NEAR and mix["near_gt"] + mix["quad_derived"] + mix["random"] == 128
    M = T.reshape(-1, 4, 4).transpose(0, 2, 1)
    R = M[:, :3, :3].astype(np.float64)
    assert np.abs(np.einsum("kij,klj->kil", R, R) - np.eye(3)).max() < 1e-4     # rigid motions
    assert np.allclose(M[:, 3], [0, 0, 0, 1])


def test_select_base_is_wide_and_deterministic():
    raw, P, Q, cp, cq = bench.build_workload(5000)
    d = float(np.linalg.norm(P.max(0) - P.min(0)))
    a = bench.select_base(P, np.random.RandomState(3), d)
    b = bench.select_base(P, np.random.RandomState(3), d)
    assert np.array_equal(a[0], b[0]) and a[1] == b[1] and a[2] == b[2]
    assert len(set(a[0].tolist())) == 4 and 0.0 <= a[1] <= 1.0 and 0.0 <= a[2] <= 1.0
