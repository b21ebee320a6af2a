import numpy as np

from portablert_b200 import hitreg


def test_layouts_match_reference_types(reference):
    """hitreg.layout() == sizeof/offsetof of the reference's own HitReg<Tags...> for all 31 combos."""
    assert len(hitreg.TAG_COMBOS) == 31
    assert sorted(hitreg.mask_of(c) for c in hitreg.TAG_COMBOS) == list(range(1, 32))
    for combo in hitreg.TAG_COMBOS:
        m = hitreg.mask_of(combo)
        assert reference.layout(m) == hitreg.layout_tuple(m), combo


def test_layout_table_of_survey_8b():
    L = hitreg.layout
    assert L(hitreg.VALID) == (8, dict(u=-1, v=-1, t=-1, primitive_id=-1, valid=4, px=-1, py=-1, pz=-1))
    assert L(hitreg.T)[0] == 16 and L(hitreg.T)[1]["t"] == 4
    assert L(hitreg.T | hitreg.VALID)[1]["valid"] == 9
    s, o = L(hitreg.T | hitreg.PID)
    assert (s, o["t"], o["primitive_id"]) == (16, 4, 8)
    s, o = L(hitreg.ALL)
    assert s == 32 and (o["u"], o["v"], o["t"], o["primitive_id"], o["valid"], o["px"], o["py"],
                        o["pz"]) == (0, 4, 8, 12, 16, 20, 24, 28)
    assert L(hitreg.P)[0] == 20 and L(hitreg.T | hitreg.P)[0] == 24
    assert L(hitreg.T | hitreg.PID | hitreg.P)[0] == 28


def test_dtype_and_masks():
    assert hitreg.dtype(hitreg.ALL).itemsize == 32
    assert hitreg.dtype(hitreg.VALID).names == ("valid",)
    assert hitreg.mask_of(()) == hitreg.ALL
    assert hitreg.mask_of(("valid", "t")) == hitreg.mask_of(("t", "valid")) == 18
    assert hitreg.hitreg_name(("uv", "t")) == "uv_t"
    a = np.zeros(3, hitreg.dtype(hitreg.T | hitreg.VALID))
    assert a.strides == (16,)
