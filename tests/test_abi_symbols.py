"""The drop-in boundary without a GPU: libsafe_mpc_b200.so loads, exports every entry point include/safe_mpc_b200.h declares (and nothing
under the smpc_ prefix that the header does not declare), and the product path fails loudly -- not into a CPU fallback -- when there
is no CUDA device.  No compute call is made here."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, 'include', 'safe_mpc_b200.h')


def _declared():
    src = open(HEADER).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)          # comments mention entry points too
    return sorted(set(re.findall(r'\b(smpc_[a-z0-9_]+)\s*\(', src)))


def _library():
    from safe_mpc_b200.build import build_cuda, LIB
    build_cuda()
    return LIB


def test_library_exports_every_declared_entry_point():
    names = _declared()
    assert len(names) >= 40 and 'smpc_rti_solve' in names and 'smpc_set_ee_trajectory' in names and 'smpc_set_stream' in names
    lib = C.CDLL(_library())
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f'declared in the header but not exported: {missing}'


def test_no_undeclared_entry_points():
    out = subprocess.run(['nm', '-D', '--defined-only', _library()], capture_output=True, text=True, check=True).stdout
    exported = sorted({ln.split()[-1] for ln in out.splitlines() if re.search(r'\sT\s+smpc_', ln)})
    extra = [n for n in exported if n not in _declared()]
    assert not extra, f'exported under the smpc_ prefix but not declared in the header: {extra}'


def test_struct_layout_matches_the_header():
    """abi.Problem (ctypes mirror) has the size the library was compiled with (smpc_problem_size, if exported) or at least the
    fields the header lists, in order"""
    from safe_mpc_b200 import abi
    src = open(HEADER).read()
    body = src[src.index('typedef struct smpc_problem'):src.index('} smpc_problem_t;')]
    body = re.sub(r'/\*.*?\*/', '', body, flags=re.S)
    fields = []
    for stmt in body[body.index('{') + 1:].split(';'):
        for decl in stmt.split(','):
            m = re.search(r'([a-zA-Z_][a-zA-Z0-9_]*)\s*(?:\[[^\]]*\]\s*)*$', decl.strip())
            if m:
                fields.append(m.group(1))
    mirror = [f[0] for f in abi.Problem._fields_]
    assert fields == mirror, (fields, mirror)


def test_create_fails_loudly_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip('a CUDA device is present')
    from safe_mpc_b200.engine import Engine
    from tests.common import make_problem
    prob, params, md = make_problem('naive', N=5)
    with pytest.raises(RuntimeError, match='no CUDA device'):
        Engine(prob, 4, 0)
