"""Plan identity on the GPU: the product's BatchedAStarPlanner (path_planner_b200/harness/, expansion
batched into the CUDA engine through the C ABI) against the reference's own AStarPlanner on the
same world, start state, seed and VIRTUAL clock.  BASELINE.json north_star: "The final plan must be
identical to the reference's for the same seed and inputs" -- the chosen Dubins words, radii, plan
depth and the search counters (Samples / Generated / Expanded / Iterations / now() calls) must be
equal; continuous fields within 1e-9 relative.

Both planners live in oracle/_ref/libppe_harness.so (built in the container that holds
/root/reference; it travels to the GPU box as a prebuilt file and links path_planner_b200/libppe.so)."""
import numpy as np
import pytest

from tests import common, plan_cases

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not common.have_harness(), reason="oracle/_ref/libppe_harness.so not built (needs /root/reference)")]


@pytest.fixture(scope="module")
def lib():
    return common.load_harness()


@pytest.mark.parametrize("case", plan_cases.CASES, ids=plan_cases.CASE_IDS)
def test_plan_is_identical_to_the_reference(lib, case):
    plan_cases.compare(lib, case, exact=False)


def test_knn_chunk_does_not_change_the_plan(lib):
    case = plan_cases.CASES[0][:4] + (2e-3, 100)
    a, plan_a = plan_cases.compare(lib, case, exact=False, knn_chunk=16)
    b, plan_b = plan_cases.compare(lib, case, exact=False, knn_chunk=512)
    assert np.array_equal(plan_a, plan_b)
    assert a["dubins_solves"] < b["dubins_solves"]
