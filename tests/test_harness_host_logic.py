"""Host logic of the C++ adapter (path_planner_b200/harness/BatchedAStarPlanner.cpp) without a GPU.

oracle/_ref/libplan_compare_cpu.so links the adapter against a TEST DOUBLE of the C ABI
(oracle/ppe_on_oracle.c: ppe_* forwarded to the CPU oracle, which is bit-identical to the compiled
reference).  With a bit-identical evaluator behind it, everything the adapter does on the host --
batch assembly for the three call sites of SamplingBasedPlanner::expand (:76, :119, :145), the
replay of the Euclid / Dubins heaps, ribbon-set interning, handing results back through the
reference's Edge / Vertex members in push order -- must reproduce the reference's plan BIT FOR BIT,
together with its Samples / Generated / Expanded / Iterations counters."""
import numpy as np
import pytest

from tests import common, plan_cases

CPU_SO = common.HARNESS_SO.replace("libplan_compare.so", "libplan_compare_cpu.so")
pytestmark = pytest.mark.skipif(not __import__("os").path.exists(CPU_SO),
                                reason="oracle/_ref/libplan_compare_cpu.so not built (needs /root/reference)")


@pytest.fixture(scope="module")
def lib():
    return common.load_harness(CPU_SO)


@pytest.mark.parametrize("case", plan_cases.CASES, ids=plan_cases.CASE_IDS)
def test_batched_planner_reproduces_the_reference_plan_bit_for_bit(lib, case):
    got, plan = plan_cases.compare(lib, case, exact=True)
    assert got["expanded"] >= 2


def test_default_path_is_the_frontier_path(lib):
    """Default width: vertices travel to ppe_expand_batch in groups, most expansions are served from a cached group
    result, and no expansion needs the exact host replay (no two samples at exactly equal distance)."""
    got, _ = plan_cases.compare(lib, plan_cases.CASES[0], exact=True)
    assert got["frontier_vertices"] >= got["expanded"] and got["frontier_hits"] > 0.5 * got["expanded"]
    assert got["exact_expansions"] == 0
    assert got["batches"] < 0.5 * got["expanded"]


@pytest.mark.parametrize("frontier", [0, 1, 7])
@pytest.mark.parametrize("i", [0, 6, 8])
def test_frontier_width_does_not_change_the_plan(lib, i, frontier):
    """Width 0 = the exact host replay for every vertex (k-nearest heaps on the host, K1 per chunk, K2 per vertex);
    width 1 = one vertex per device call; any width must give the reference's plan bit for bit."""
    got, _ = plan_cases.compare(lib, plan_cases.CASES[i], exact=True, frontier=frontier)
    if frontier == 0:
        assert got["exact_expansions"] == got["expanded"] and got["frontier_vertices"] == 0


@pytest.mark.parametrize("case", plan_cases.FOLLOWUP_CASES, ids=plan_cases.FOLLOWUP_IDS)
def test_previous_plan_and_brown_paths_through_the_engine(lib, case):
    """The other callers of the path: previous-plan re-validation (AStarPlanner.cpp:46-59) and the Brown-path
    expansion (:150-162) are evaluated by the engine in BatchedAStarPlanner::plan; the second planning cycle must be the
    reference's bit for bit."""
    got, plan = plan_cases.compare_followup(lib, case, exact=True)
    assert len(plan) >= 1


def test_knn_chunk_only_changes_the_launch_count(lib):
    case = plan_cases.CASES[0][:4] + (2e-3, 100)
    a, plan_a = plan_cases.compare(lib, case, exact=True, knn_chunk=16, frontier=0)
    b, plan_b = plan_cases.compare(lib, case, exact=True, knn_chunk=512, frontier=0)
    assert np.array_equal(plan_a, plan_b)
    assert a["dubins_solves"] < b["dubins_solves"] and a["batches"] > b["batches"]


def test_mispredicted_chunks_do_not_change_the_plan():
    """The K1 chunk is a PREDICTION of the samples the replayed loop will pop.  With the prediction scrambled on purpose
    (reversed, a third dropped) every pop goes through the look-up or the single-solve fallback -- the plan must still be
    the reference's bit for bit.  Runs in a fresh process: the hook is read once per process."""
    import subprocess
    import sys
    code = (
        "import os, sys\n"
        "os.environ['PPE_HARNESS_TEST_MISPREDICT'] = '1'\n"
        "sys.path.insert(0, %r)\n"
        "from tests import common, plan_cases\n"
        "lib = common.load_harness(%r)\n"
        "for i in (0, 2, 5):\n"
        "    got, plan = plan_cases.compare(lib, plan_cases.CASES[i], exact=True, knn_chunk=16, frontier=0)\n"
        "print('ok')\n" % (common.ROOT, CPU_SO))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0 and out.stdout.strip().endswith("ok"), out.stderr[-2000:]


def test_mixing_device_and_exact_expansions_does_not_change_the_plan():
    """An expansion the device flags (two samples at exactly equal distance) is replayed on the host, whose sample heap must
    then be in the arrangement the reference's m_Samples has at that moment: the logged device expansions are replayed into
    it first.  PPE_HARNESS_TEST_FORCE_EXACT sends every third expansion down that path."""
    import subprocess
    import sys
    code = (
        "import os, sys\n"
        "os.environ['PPE_HARNESS_TEST_FORCE_EXACT'] = '1'\n"
        "sys.path.insert(0, %r)\n"
        "from tests import common, plan_cases\n"
        "lib = common.load_harness(%r)\n"
        "for i in (0, 6, 8):\n"
        "    got, plan = plan_cases.compare(lib, plan_cases.CASES[i], exact=True)\n"
        "    assert got['exact_expansions'] > 0 and got['frontier_vertices'] > 0, got\n"
        "print('ok')\n" % (common.ROOT, CPU_SO))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0 and out.stdout.strip().endswith("ok"), out.stderr[-2000:]


def test_keyed_heap_matches_libstdcxx(host_helpers):
    """path_planner_b200/harness/KeyedHeap.h restates libstdc++'s heap algorithms on (key, index) arrays; the arrangement
    after make_heap and after every pop_heap must be the one std::make_heap / std::pop_heap produce -- with heavy ties."""
    import ctypes as C
    hh = host_helpers
    hh.hh_keyed_heap_check.argtypes = [C.POINTER(C.c_double), C.c_int, C.c_int]
    rng = np.random.default_rng(12)
    for n in (1, 2, 3, 4, 5, 7, 8, 16, 17, 100, 101, 1000, 4097):
        for mode in ("distinct", "ties", "all-equal", "sorted", "reversed"):
            if mode == "distinct":
                k = rng.uniform(0, 100, n)
            elif mode == "ties":
                k = rng.integers(0, max(2, n // 8), n).astype(np.float64)
            elif mode == "all-equal":
                k = np.full(n, 3.5)
            elif mode == "sorted":
                k = np.sort(rng.uniform(0, 100, n))
            else:
                k = np.sort(rng.uniform(0, 100, n))[::-1].copy()
            k = np.ascontiguousarray(k, dtype=np.float64)
            rc = hh.hh_keyed_heap_check(k.ctypes.data_as(C.POINTER(C.c_double)), n, min(n, 300))
            assert rc == 0, (n, mode, rc)


def test_expand_test_1_ribbons(lib):
    """ExpandTest1Ribbons, test_planner.cpp:1061-1082: one expansion over 1000 samples leaves exactly 40 vertices on
    the open list (2 speeds x 2 radii to the ribbon end point + 2 radii x k = 9 winners x 2 speeds), popped in
    non-decreasing f.  The adapter's expansion must push the same 40 f-values in the same order as the reference's."""
    import ctypes as C
    from path_planner_b200 import synth
    D = C.POINTER(C.c_double)
    lib.lib.ref_expand_once.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, D, C.c_int]
    lib.lib.harness_expand_once.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, D, C.c_int]
    world = synth.world_c1()          # RibbonManager().add(0, 10, 0, 30), Map base, no obstacles
    sid = world.upload_ref(lib)
    f_ref, f_har = np.zeros(64), np.zeros(64)
    n_ref = lib.lib.ref_expand_once(lib.ctx, sid, 1000, 9, f_ref.ctypes.data_as(D), 64)
    n_har = lib.lib.harness_expand_once(lib.ctx, 0, sid, 1000, 9, f_har.ctypes.data_as(D), 64)
    assert n_ref == 40 and n_har == 40
    assert np.all(np.diff(f_ref[:40]) >= 0)
    assert np.array_equal(f_ref[:40], f_har[:40])
