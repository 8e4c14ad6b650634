"""A small slice of the randomised parity sweep (tests/gpu_fuzz.py) as a regular GPU test."""
import numpy as np
import pytest

from oracle import ts_oracle as O
from tests.gpu_fuzz import random_case
from tests.helpers import compare_runs

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("sweep_seed", [11, 12])
def test_random_configurations_identical_to_oracle(sweep_seed):
    rng = np.random.default_rng(sweep_seed)
    for i in range(8):
        case = random_case(rng, i).build()
        if rng.random() < 0.25:  # alpha as a full cost channel
            for e in case.examples:
                e[..., 3] = rng.integers(0, 256, e.shape[:2], dtype=np.uint8)
            case.pyramids = [O.pyramid_build(e, max(1, case.stages)) for e in case.examples]
            if case.inpaint:
                case.inpaint_color = case.examples[0].copy()
        st = compare_runs(case.run_oracle(), case.run_gpu())
        assert st["color_mismatch"] == 0 and st["coord_mismatch"] == 0 and st["id_mismatch"] == 0, (case.name, st)
        assert st["order_equal"] and st.get("score_bit_mismatch", 0) == 0, (case.name, st)
