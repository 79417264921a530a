"""CPU: the Prodigy restatement (oracle/prodigy_oracle.py) against the golden produced by the unmodified reference
optimizer (oracle/make_golden_prodigy.py -> tests/golden/prodigy.pt)."""
import os

import pytest
import torch


@pytest.mark.parametrize("case", ["default", "decay_biascorr", "coupled_decay_growth"])
def test_prodigy_oracle_matches_reference(case):
    from oracle.prodigy_oracle import ProdigyOracle
    g = torch.load(os.path.join(os.path.dirname(__file__), "golden", "prodigy.pt"))[case]
    params = [t.clone() for t in g["init"]]
    opt = ProdigyOracle(params, lr=1.0, **g["kw"])
    for step, ns in enumerate(g["noises"]):
        opt.step([(p - t) + n for p, t, n in zip(params, g["targets"], ns)])
        assert abs(opt.d - g["d"][step]) <= 1e-6 * abs(g["d"][step])
    for p, ref in zip(params, g["final"]):
        assert torch.allclose(p, ref, rtol=1e-5, atol=1e-7)
