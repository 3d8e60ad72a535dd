"""The roofline numerator is counted, not estimated: the constants frozen in
csrc/mppi_b200.cu must equal what oracle/flop_count.py counts over the restated maths."""
import pytest


@pytest.fixture(scope="module")
def table(oracle):
    from oracle import flop_count
    return flop_count.table()


def test_library_constants_equal_the_static_counter(table):
    from quadrotor_manipulator_mppi_b200 import _native
    ids = {"drone3": _native.MODEL_DRONE3, "quad4": _native.MODEL_QUAD4, "arm7": _native.MODEL_ARM7, "wb11": _native.MODEL_WB11}
    for name, mid in ids.items():
        assert _native.algorithmic_flops(mid) == table[name]["dense"]["flop"], name
        assert _native.structural_flops(mid) == table[name]["structural"]["flop"], name


def test_counter_sanity(table):
    # transcendental evaluations are listed separately and are not in the FLOP figure
    assert table["arm7"]["dense"]["sincos"] == 7 and table["arm7"]["dense"]["atan2"] == 2 and table["arm7"]["dense"]["asin"] == 1
    assert table["wb11"]["dense"]["sincos"] == 10 and table["wb11"]["dense"]["rcp"] == 1
    # the whole body is the quadrotor step + the arm step + the per-step moving base
    wb, arm, quad = (table[m]["dense"]["flop"] for m in ("wb11", "arm7", "quad4"))
    assert arm + quad < wb < arm + quad + 120
    # skipping structural zeros can only remove work
    for m in table:
        assert table[m]["structural"]["flop"] <= table[m]["dense"]["flop"]
    # arm FK dominates: 7 joints x (column rotation 18 + constant product 45 + translation 18)
    assert table["arm7"]["dense"]["flop"] == 7 * 81 + 7 * 11 + 42
