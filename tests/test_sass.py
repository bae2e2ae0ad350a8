"""The shipped library's machine code holds what DESIGN.md says it does (cuobjdump on the in-tree .so; no GPU needed)."""
import pathlib
import re
import shutil
import subprocess

import pytest

LIB = pathlib.Path(__file__).resolve().parent.parent / "gecco_b200" / "libgecco_crf_b200.so"


@pytest.fixture(scope="module")
def sass():
    if shutil.which("cuobjdump") is None or not LIB.exists():
        pytest.skip("cuobjdump or the built library is missing")
    text = subprocess.run(["cuobjdump", "-sass", str(LIB)], check=True, capture_output=True, text=True).stdout
    functions, name = {}, None
    for line in text.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            functions[name] = []
        elif name is not None and re.match(r"\s+/\*[0-9a-f]{4}\*/", line):
            functions[name].append(line.split("*/", 1)[1].strip())
    archs = set(re.findall(r"arch = (\S+)", text))
    return functions, archs


def find(functions, *parts):
    hits = [body for name, body in functions.items() if all(p in name for p in parts)]
    assert hits, parts
    return hits


def count(body, mnemonic):
    return sum(1 for ins in body if re.search(rf"(^|\s){re.escape(mnemonic)}", ins))


def test_built_for_sm_100a_only(sass):
    _, archs = sass
    assert archs == {"sm_100a"}


def test_streaming_kernel_uses_bulk_copies_mbarriers_and_packed_math(sass):
    """ids and the delta table arrive by cp.async.bulk (UBLKCP) on mbarriers (SYNCS), the DP runs on packed
    f32x2 instructions with MUFU reciprocals / exponentials, outputs leave as 64-bit stores."""
    functions, _ = sass
    for body in find(functions, "stream_kernel", "ILi20E"):
        assert count(body, "UBLKCP") >= 2 and count(body, "SYNCS.PHASECHK") >= 1
        assert count(body, "FFMA2") >= 40 and count(body, "FMUL2") >= 40
        assert count(body, "MUFU.RCP") >= 40 and count(body, "MUFU.EX2") >= 1
        assert count(body, "STG.E.64") >= 1
        assert not any("LDL" in ins or "STL" in ins for ins in body), "no register spills in the hot kernel"


def test_feature_kernel_is_branch_light_and_works_in_shared_memory(sass):
    """uniform loop bounds through REDUX (no BRA.DIV in front of the shuffles), shared-memory atomics for
    the bitmaps, the staged uint16 table, streaming loads / stores, 16-byte bitmap wipes."""
    functions, _ = sass
    staged = find(functions, "features_kernel", "Li8ELi8ELi512ELb1E")
    for body in staged:
        assert count(body, "BRA.DIV") == 0
        assert count(body, "CREDUX") + count(body, "REDUX") >= 3
        assert count(body, "ATOMS.OR") >= 8
        assert count(body, "LDS.U16") >= 8
        assert count(body, "STS.128") >= 1
        assert any(ins.startswith("@") and "STG.E.EF" in ins for ins in body), "predicated streaming stores"
        assert any("LDG.E.EF" in ins for ins in body)
        assert not any("LDL" in ins or "STL" in ins for ins in body)
    for body in find(functions, "features_kernel", "Li32ELi2ELi256ELb0E"):  # sparse tables: hashed bitmaps
        assert count(body, "BRA.DIV") == 0 and count(body, "ATOMS.OR") >= 2


def test_streaming_kernel_variants_exist(sass):
    """Nine window sizes, the half / quarter-tile variants of W = 5 and 20, and the peer-store instantiation."""
    functions, _ = sass
    for w in (5, 10, 15, 20, 25, 30, 40, 50, 64):  # from W = 25 on three CTAs fit an SM (two at 64): their share of the registers
        assert find(functions, "stream_kernel", f"ILi{w}ELi128ELi{4 if w <= 20 else 3 if w <= 50 else 2}EiLi256ELb0E")
    for w in (5, 20):
        for slots in (128, 64):
            assert find(functions, "stream_kernel", f"ILi{w}ELi128ELi4EiLi{slots}ELb0E")
        assert find(functions, "stream_kernel", f"ILi{w}ELi128ELi4EiLi256ELb1E")


def test_reference_order_f64_kernel(sass):
    """separate double multiplies and adds (the recursion is not contracted into FMAs: the only DFMAs are
    those of the IEEE division sequence and of the double-double exponential), 64-bit atomic max for the pool."""
    functions, _ = sass
    for body in find(functions, "exact_window_kernel", "ILi20E"):
        assert count(body, "DMUL") >= 200 and count(body, "DADD") >= 60
        assert count(body, "MUFU.RCP64H") >= 20, "IEEE divisions (reciprocal seed + Newton steps)"
        assert any("ATOMG" in ins and "MAX" in ins and "64" in ins for ins in body) or any("RED" in ins and "MAX" in ins for ins in body)
        # four blocks per SM = 128 registers: a few doubles of the stored forward half spill (0.41 ms against 0.47 ms
        # at three blocks of 152 registers), nothing like the whole half
        assert sum("LDL" in ins or "STL" in ins for ins in body) <= 48
    for body in find(functions, "exact_unary_kernel"):
        assert count(body, "DFMA") >= 20 and count(body, "DADD") >= 10  # the double-double exponential


def test_segment_kernels_are_ballot_and_popcount_code(sass):
    """gene masks by warp votes, bit-parallel replay, release / acquire flags instead of grid barriers."""
    functions, _ = sass
    for body in find(functions, "scan_kernel", "SegmentsArgs"):
        assert count(body, "VOTE") >= 5 and count(body, "POPC") >= 10 and count(body, "FLO") >= 1
        assert any("LDG" in ins and "STRONG.GPU" in ins for ins in body), "acquire loads of the published summaries"
        assert not any("LDL" in ins or "STL" in ins for ins in body)
