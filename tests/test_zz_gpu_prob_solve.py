"""GPU tier of tests/test_prob_solve_end_to_end.py: ``prob.solve(nlp=True, solver=cp.IPOPT)`` through the unmodified
reference (oracle/_ref on the GPU box) with the cyipopt protocol stand-in, once on the reference's ``Oracles`` and once
on ``GpuOracles`` on the real device - same status, iteration count, callback counts, optimum within 1e-8.  Sorted
last with the other tests that joined the GPU tier after the round's GPU budget was spent."""
import pytest

from test_prob_solve_end_to_end import CASES, README_OPTIMUM, _both_arms, cp  # noqa: F401  (cp is a fixture)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_gpu_prob_solve_is_unchanged_by_the_oracle_swap(name, cp, monkeypatch):
    """The same on the real device: the CUDA path behind the reference's own solver interface."""
    ref, ours = _both_arms(cp, CASES[name], monkeypatch, standin_device=False)
    assert ours.nlps[0].obj.kernel_launches() > 0
    if name == "readme_toy":
        assert abs(ours.value - README_OPTIMUM) < 1e-7
