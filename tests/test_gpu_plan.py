"""Plan identity on the GPU: the product's BatchedAStarPlanner (path_planner_b200/harness/, expansion
batched into the CUDA engine through the C ABI) against the reference's own AStarPlanner on the
same world, start state, seed and VIRTUAL clock.  BASELINE.json north_star: "The final plan must be
identical to the reference's for the same seed and inputs" -- the chosen Dubins words, radii, plan
depth and the search counters (Samples / Generated / Expanded / Iterations / now() calls) must be
equal; continuous fields within 1e-9 relative.

Both planners live in oracle/_ref/libplan_compare.so (built in the container that holds
/root/reference; it travels to the GPU box as a prebuilt file and links path_planner_b200/libppe.so)."""
import numpy as np
import pytest

from tests import common, plan_cases

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not common.have_harness(), reason="oracle/_ref/libplan_compare.so not built (needs /root/reference)")]


@pytest.fixture(scope="module")
def lib():
    return common.load_harness()


@pytest.mark.parametrize("case", plan_cases.CASES, ids=plan_cases.CASE_IDS)
def test_plan_is_identical_to_the_reference(lib, case):
    plan_cases.compare(lib, case, exact=False)


@pytest.mark.parametrize("case", plan_cases.FOLLOWUP_CASES, ids=plan_cases.FOLLOWUP_IDS)
def test_followup_cycle_is_identical_to_the_reference(lib, case):
    """Second planning cycle: non-empty previousPlan (re-validated through ppe_true_cost_batch, AStarPlanner.cpp:46-59)
    and / or useBrownPaths (AStarPlanner.cpp:150-162)."""
    plan_cases.compare_followup(lib, case, exact=False)


@pytest.mark.parametrize("frontier", [0, 1, 16])
def test_frontier_width_does_not_change_the_plan(lib, frontier):
    """0 = exact host replay per vertex (K1 chunks + one K2 launch per vertex), 1 = one vertex per ppe_expand_batch."""
    for i in (0, 5, 6):
        got, _ = plan_cases.compare(lib, plan_cases.CASES[i], exact=False, frontier=frontier)
        if frontier == 0:
            assert got["frontier_vertices"] == 0
        else:
            assert got["exact_expansions"] == 0 and got["frontier_vertices"] >= got["expanded"]


def test_knn_chunk_does_not_change_the_plan(lib):
    case = plan_cases.CASES[0][:4] + (2e-3, 100)
    a, plan_a = plan_cases.compare(lib, case, exact=False, knn_chunk=16, frontier=0)
    b, plan_b = plan_cases.compare(lib, case, exact=False, knn_chunk=512, frontier=0)
    assert np.array_equal(plan_a, plan_b)
    assert a["dubins_solves"] < b["dubins_solves"]


def test_expand_test_1_ribbons_on_the_engine(lib):
    """ExpandTest1Ribbons (test_planner.cpp:1061-1082) with the CUDA engine behind the adapter: exactly 40 vertices,
    non-decreasing f, the reference's f-values within 1e-9."""
    import ctypes as C
    from path_planner_b200 import synth
    D = C.POINTER(C.c_double)
    lib.lib.ref_expand_once.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, D, C.c_int]
    lib.lib.harness_expand_once.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, D, C.c_int]
    world = synth.world_c1()
    sid = world.upload_ref(lib)
    f_ref, f_har = np.zeros(64), np.zeros(64)
    n_ref = lib.lib.ref_expand_once(lib.ctx, sid, 1000, 9, f_ref.ctypes.data_as(D), 64)
    n_har = lib.lib.harness_expand_once(lib.ctx, 0, sid, 1000, 9, f_har.ctypes.data_as(D), 64)
    assert n_ref == 40 and n_har == 40
    assert np.all(np.diff(f_har[:40]) >= 0)
    assert np.allclose(f_ref[:40], f_har[:40], rtol=common.RTOL, atol=common.ATOL)
