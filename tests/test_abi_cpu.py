"""CPU: the C-ABI library loads without a GPU and exports every symbol include/indm_b200.h declares; the ctypes binding
(indm_b200/_lib.py) declares a signature for each of them and mirrors the igemm descriptor field for field."""
import ctypes
import os
import re

import pytest

from indm_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, 'include', 'indm_b200.h')


def _declared():
    src = open(HEADER).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(?:int|const char\*)\s+(indm_[a-z0-9_]+)\s*\(', src)))


def test_header_declares_what_the_binding_binds():
    names = _declared()
    assert 'indm_igemm' in names and 'indm_upfirdn2d_f32' in names and len(names) >= 40
    missing = [n for n in names if n not in _lib.EXPORTS]
    extra = [n for n in _lib.EXPORTS if n not in names]
    assert not missing, f'declared in the header but not bound in _lib.py: {missing}'
    assert not extra, f'bound in _lib.py but not declared in the header: {extra}'


def test_library_loads_and_exports_every_symbol():
    if not os.path.exists(_lib.LIB_PATH):
        pytest.skip('libindm_b200.so not built (run __graft_entry__.build())')
    h = ctypes.CDLL(_lib.LIB_PATH)          # no GPU needed: the driver entry point for TMA maps is resolved lazily
    missing = [n for n in _declared() if not hasattr(h, n)]
    assert not missing, missing
    h.indm_version.restype = ctypes.c_char_p
    assert b'sm_100a' in h.indm_version()


def test_igemm_descriptor_mirror_matches_the_header():
    src = open(HEADER).read()
    body = src[src.index('typedef struct indm_igemm {'):src.index('} indm_igemm_t;')]
    body = re.sub(r'/\*.*?\*/', '', body, flags=re.S)
    fields = []
    for decl in body.split(';'):
        decl = decl.strip().replace('typedef struct indm_igemm {', '')
        if not decl:
            continue
        names = re.sub(r'^(const\s+)?(void|float|int32_t|int64_t)\s*\*?', '', decl.strip())
        fields += [n.strip().lstrip('*') for n in names.split(',')]
    assert fields == [f[0] for f in _lib.IgemmDesc._fields_]
