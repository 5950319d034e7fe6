"""oracle/dqn.py pinned to the UNMODIFIED reference agents' learn() (tests/golden/*_small.npz etc.)."""
import numpy as np
import pytest

import gen_inputs as G
from helpers import load_golden, run_q_oracle


@pytest.mark.parametrize("name", list(G.Q_CASES.keys()))
def test_q_oracle_matches_reference(name):
    case = G.Q_CASES[name]
    out, inp = run_q_oracle(case)
    gold = load_golden(name)
    for k, v in gold.items():
        if k.startswith("param."):
            np.testing.assert_allclose(G.subsample(out["params"][k[6:]].numpy()), v, rtol=1e-6, atol=1e-7, err_msg=k)
    np.testing.assert_allclose(out["loss"], float(gold["result.loss"]), rtol=1e-6)
    np.testing.assert_allclose(out["max_Q"], float(gold["result.max_Q"]), rtol=1e-6)
    if "prio.p" in gold:
        np.testing.assert_allclose(out["priority"].numpy(), gold["prio.p"], rtol=1e-6, atol=1e-9)
        assert np.array_equal(inp["indices"], gold["prio.idx"])
    if "result.max_logit" in gold:
        np.testing.assert_allclose(out["max_logit"], float(gold["result.max_logit"]), rtol=1e-6)
        np.testing.assert_allclose(out["min_logit"], float(gold["result.min_logit"]), rtol=1e-6)
