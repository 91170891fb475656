"""Host-side helpers for balanced presentations + the single-call reduction functions.

Mirrors the names and behaviour of the reference's ``ac_solver/envs/utils.py`` (paths below
are relative to the reference root).  The predicates and layout converters are plain host
logic; ``simplify_relator`` / ``simplify_presentation`` run on the GPU through the generic
byte kernel of libacsolver_b200 (any int8 alphabet) -- there is no CPU implementation.
"""

from __future__ import annotations

import numpy as np

from .. import _lib
from .._host import generic_call


def is_array_valid_presentation(array):
    """envs/utils.py:13-54 -- both halves non-empty and zero-padded on the right."""
    assert isinstance(array, (list, np.ndarray)), f"array must be a list or a numpy array, got {type(array)}"
    a = np.asarray(array)
    if a.ndim != 1 or a.size % 2 != 0:
        return False
    m = a.size // 2
    ok = True
    for half in (a[:m], a[m:]):
        n = int(np.count_nonzero(half))
        ok = ok and n > 0 and not half[n:].any()
    return bool(ok)


def is_presentation_trivial(presentation):
    """envs/utils.py:57-87 -- each relator is a single letter and both generators occur."""
    if not is_array_valid_presentation(presentation):
        return False
    a = np.asarray(presentation)
    m = a.size // 2
    if np.count_nonzero(a[:m]) != 1 or np.count_nonzero(a[m:]) != 1:
        return False
    return sorted(abs(int(v)) for v in a[a != 0]) == [1, 2]


def generate_trivial_states(max_relator_length):
    """envs/utils.py:91-114 -- the eight trivial states as an (8, 2*mrl) array."""
    out = np.zeros((8, 2 * max_relator_length), dtype=np.int64)
    k = 0
    for g in (1, 2):
        for s1 in (-1, 1):
            for s2 in (-1, 1):
                out[k, 0] = s1 * g
                out[k, max_relator_length] = s2 * (3 - g)
                k += 1
    return out


def convert_relators_to_presentation(relator1, relator2, max_relator_length):
    """envs/utils.py:117-145 -- two letter lists -> zero-padded int8 presentation."""
    assert 0 not in relator1 and 0 not in relator2, "relator1 and relator2 must not be padded with zeros."
    assert max_relator_length >= max(len(relator1), len(relator2)), (
        "max_relator_length must be greater than or equal to the lengths of relator1 and rel2."
    )
    assert isinstance(relator1, list) and isinstance(relator2, list), (
        f"got types {type(relator1)} for relator1 and {type(relator2)} for relator2"
    )
    out = np.zeros(2 * max_relator_length, dtype=np.int8)
    out[: len(relator1)] = relator1
    out[max_relator_length : max_relator_length + len(relator2)] = relator2
    return out


def change_max_relator_length_of_presentation(presentation, new_max_length):
    """envs/utils.py:148-172 -- re-pad a presentation to another max_relator_length."""
    p = np.asarray(presentation)
    m = p.size // 2
    n0 = int(np.count_nonzero(p[:m]))
    n1 = int(np.count_nonzero(p[m:]))
    # the reference slices the first n letters of each half and requires list inputs
    return convert_relators_to_presentation(
        relator1=p[:n0].tolist(), relator2=p[m : m + n1].tolist(), max_relator_length=new_max_length
    )


def simplify_relator(relator, max_relator_length, cyclical=False, padded=True):
    """envs/utils.py:175-240 -- free (and optionally cyclic) reduction of one word."""
    assert isinstance(relator, np.ndarray), "expect relator to be a numpy array"
    width = max(int(relator.size), 1)
    row = np.zeros(width, dtype=np.int8)
    row[: relator.size] = relator
    out, aux, status = generic_call(_lib.OP_SIMPLIFY_RELATOR, row[None, :], cyclical=cyclical)
    assert status[0] == 0, "expect all zeros to be at the right end"
    n = int(aux[0])
    res = out[0, :n].astype(relator.dtype)
    if padded:
        res = np.pad(res, (0, max_relator_length - len(res)))
    assert max_relator_length >= n, "Increase max length! Length of simplified word is bigger than maximum allowed length."
    return res, n


def simplify_presentation(presentation, max_relator_length, lengths_of_words, cyclical=True):
    """envs/utils.py:243-280 -- validate, then reduce both relators."""
    p = np.array(presentation)
    assert p.size == 2 * max_relator_length
    out, aux, status = generic_call(_lib.OP_SIMPLIFY_PRESENTATION, p.astype(np.int8)[None, :], cyclical=cyclical)
    assert status[0] == 0, f"{p} is not a valid presentation. Expect all zeros to be padded to the right."
    return out[0].astype(p.dtype), [int(aux[0, 0]), int(aux[0, 1])]
