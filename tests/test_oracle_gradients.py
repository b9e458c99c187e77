"""CPU: autograd through the oracle port reproduces the gradients the UNMODIFIED reference produced
(fixtures with ``grad/*`` entries, oracle/make_goldens.py::reference_gradients)."""
import pytest

from tests import grad_util as gu
from tests import parity_util as pu
from tests.golden_io import golden_names, load_golden

GRAD_FIXTURES = [n for n in golden_names() if load_golden(n).grad]


def test_there_are_gradient_fixtures():
    assert len(GRAD_FIXTURES) >= 2


@pytest.mark.parametrize("name", GRAD_FIXTURES)
def test_oracle_autograd_matches_reference_gradients(name):
    g = load_golden(name)
    wl = pu.workload_from_golden(g)
    got = gu.oracle_gradients(wl, g.sd)
    for k, ref in g.grad.items():
        assert got[k].shape == ref.shape, k
        assert gu.rel_err(got[k], ref) <= 1e-5, (k, gu.rel_err(got[k], ref))
