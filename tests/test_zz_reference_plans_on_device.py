"""The reference optimizer's own plans, lowered by the in-tree binding, EXECUTED ON THE DEVICE through the C ABI.

Same interpreter as tests/test_reference_plans.py / test_reference_plans_more.py (DAG order, every operator fed exactly the
lowered objects its serialized work order names), but every operator runs on the GPU: each call stages its input
relations, runs the C-ABI entry (qsgpu_select / qsgpu_build_lip_filter / qsgpu_join_build + probe / qsgpu_agg_create + run +
finalize / qsgpu_topk) and reads the output relation back.  The answers must be the ones the unmodified engine printed.

STATUS (profiles/r4e_reference_plans_on_device.log, one B200, the round's last seconds of GPU budget): Q6, Q1, Q3 and Q17
ran and gave the engine's answers on the device (XPASS while they were still marked xfail) and are plain `-m gpu` tests
now; Q19 ran and failed -- as lower.cu reads, on `p_brand = 'Brand#12'` inside the join's residual predicate: a CHAR
comparison on a BUILD-side attribute is refused with QSGPU_ERR_UNSUPPORTED ("CHAR comparison other than attribute-vs-literal"),
and its 19 string literals exceed the VM's 96-byte string pool; Q4, Q5 and Q21 have NOT run on hardware (Q4 and Q21 sort on CHAR(15) / CHAR(25) keys, which
qsgpu_topk refuses above 8 bytes).  Those four stay `xfail(strict=False)`: an XPASS in a later log is their first hardware
evidence.  The file sorts last so that nothing runs after it.
"""
import pytest

import tpch_data as D
from backends import GpuBackend

import test_reference_plans as P
import test_reference_plans_more as M

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]
NOT_YET = pytest.mark.xfail(strict=False, reason="not yet green on hardware: see the module docstring (Q19 failed in r4e; Q4 / Q5 / Q21 have not run)")


@pytest.fixture()
def device(engine):
    b = P.StagingBackend(GpuBackend(engine))
    yield b
    try:                # a plan the device refuses must show up as that test's xfail, not as a teardown error
        b.close()
        engine.synchronize()
    except Exception as ex:          # noqa: BLE001
        print("teardown after a failed plan:", repr(ex))


@pytest.fixture(scope="module")
def hot_tables(oracle):
    return D.golden_tables()


@pytest.fixture(scope="module")
def full_tables(oracle):
    return M.load_full_tables()


@pytest.mark.parametrize("query,check", [("q6", P.check_q6), ("q1", P.check_q1), ("q3", P.check_q3)])
def test_hot_path_plans_on_the_device(device, hot_tables, query, check):
    # the device top-k takes LIMIT <= 1024: the query's own LIMIT (= the rows the engine printed) instead of a full sort
    check(P.Interpreter(P.PLANS[query], hot_tables, device, limit=len(P.ENGINE["sf0.01"][query]["rows"])).run())


@pytest.mark.parametrize("query", ["q17"] + [pytest.param(q, marks=NOT_YET) for q in ("q4", "q5", "q19", "q21")])
def test_next_plans_on_the_device(device, full_tables, query):
    it, out = M.run(query, full_tables, device, limit=max(1, len(M.ENGINE[query]["rows"])))
    M.check(query, it, out)
