"""Hit-record layouts of the reference, restated for the Python host mirror.

The reference's ``HitReg<Tags...>`` (include/portableRT/hitreg.hpp:29-59) is a struct whose eight
members u, v, t, primitive_id, valid, px, py, pz are each either the real field or a 1-byte
``Empty`` placeholder (hitreg.hpp:16-23), so absent fields still occupy one byte and the record
size depends on the tag set.  The C++ header of this backend (include/portableRT/intersect_cuda.hpp)
derives size/offsets with sizeof/offsetof at compile time; this module computes the same table
with the Itanium C++ ABI rules so that numpy can view the AoS records the C-ABI fills.
``tests/test_layout.py`` pins it against the table the reference's own types produce
(oracle/_ref, ``ref_layout``) and against SURVEY.md section 8b.
"""
from __future__ import annotations

import numpy as np

# canonical tag order of the reference (hitreg.hpp:24-26) == bit order of the C-ABI tag mask
TAGS = ("uv", "t", "primitive_id", "p", "valid")
UV, T, PID, P, VALID = 1, 2, 4, 8, 16
ALL = UV | T | PID | P | VALID
_BIT = dict(zip(TAGS, (UV, T, PID, P, VALID)))

# the reference's TAG_COMBOS list, in its order (hitreg.hpp:146-177)
TAG_COMBOS = [
    ("uv",), ("t",), ("primitive_id",), ("p",), ("valid",),
    ("uv", "t"), ("uv", "primitive_id"), ("uv", "p"), ("uv", "valid"), ("t", "primitive_id"),
    ("t", "p"), ("t", "valid"), ("primitive_id", "p"), ("primitive_id", "valid"), ("p", "valid"),
    ("uv", "t", "primitive_id"), ("uv", "t", "p"), ("uv", "t", "valid"),
    ("uv", "primitive_id", "p"), ("uv", "primitive_id", "valid"), ("uv", "p", "valid"),
    ("t", "primitive_id", "p"), ("t", "primitive_id", "valid"), ("t", "p", "valid"),
    ("primitive_id", "p", "valid"),
    ("uv", "t", "primitive_id", "p"), ("uv", "t", "primitive_id", "valid"),
    ("uv", "t", "p", "valid"), ("uv", "primitive_id", "p", "valid"),
    ("t", "primitive_id", "p", "valid"),
    ("uv", "t", "primitive_id", "p", "valid"),
]


def mask_of(tags) -> int:
    """Tag names (any order, like ``HitReg<>``'s order-insensitive has_tag fold,
    hitreg.hpp:46,56-59) -> 5-bit mask.  An empty list means all tags, like the zero-tag
    ``nearest_hits(rays)`` overload (backend.hpp:77-79)."""
    if isinstance(tags, int):
        if not 0 < tags <= ALL:
            raise ValueError(f"tag mask out of range: {tags}")
        return tags
    if isinstance(tags, str):
        tags = (tags,)
    tags = tuple(tags)
    if not tags:
        return ALL
    m = 0
    for t in tags:
        if t not in _BIT:
            raise ValueError(f"unknown filter tag {t!r}; expected one of {TAGS}")
        m |= _BIT[t]
    return m


def tags_of(mask: int):
    return tuple(t for t in TAGS if mask & _BIT[t])


def hitreg_name(tags) -> str:
    """core.hpp:75-81 (order-sensitive join with '_')."""
    return "_".join(tags)


# member order of HitRegImpl (hitreg.hpp:36-43): (name, owning tag bit, size, alignment)
_MEMBERS = (
    ("u", UV, 4, 4), ("v", UV, 4, 4), ("t", T, 4, 4), ("primitive_id", PID, 4, 4),
    ("valid", VALID, 1, 1), ("px", P, 4, 4), ("py", P, 4, 4), ("pz", P, 4, 4),
)
_NP = {"u": "<f4", "v": "<f4", "t": "<f4", "primitive_id": "<u4", "valid": "?",
       "px": "<f4", "py": "<f4", "pz": "<f4"}


def layout(mask: int):
    """-> (stride, {member: offset or -1}) for HitReg of this tag mask."""
    mask = mask_of(mask)
    off = 0
    align = 1
    offs = {}
    for name, bit, size, al in _MEMBERS:
        if mask & bit:
            off = (off + al - 1) // al * al
            offs[name] = off
            off += size
            align = max(align, al)
        else:  # Empty: size 1, alignment 1
            offs[name] = -1
            off += 1
    stride = (off + align - 1) // align * align
    return stride, offs


def layout_tuple(mask: int):
    """(stride, off_u, off_v, off_t, off_pid, off_valid, off_px, off_py, off_pz) -- the field order
    of ``prt_hit_layout`` in include/prt_b200.h."""
    stride, o = layout(mask)
    return (stride, o["u"], o["v"], o["t"], o["primitive_id"], o["valid"], o["px"], o["py"], o["pz"])


def dtype(mask: int) -> np.dtype:
    """numpy structured dtype with exactly the reference's offsets and itemsize."""
    stride, o = layout(mask)
    names = [n for n, *_ in _MEMBERS if o[n] >= 0]
    return np.dtype({"names": names, "formats": [_NP[n] for n in names],
                     "offsets": [o[n] for n in names], "itemsize": stride})
