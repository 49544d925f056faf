"""CPU: the C-ABI library builds for sm_100a, loads without a GPU, and exports exactly what include/mgw.h declares."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, 'include', 'mgw.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    decls = re.findall(r'MGW_API\s+[\w\s\*]+?\b(mgw_\w+)\s*\(([^;]*?)\)\s*;', src, flags=re.S)
    return {name: [a.strip() for a in args.split(',')] if args.strip() != 'void' else [] for name, args in decls}


def test_header_and_library_agree():
    import dovs_b200
    fns = header_functions()
    assert len(fns) >= 24
    lib = ctypes.CDLL(dovs_b200._lib.SO)
    for name, args in fns.items():
        assert hasattr(lib, name), '%s declared in include/mgw.h but not exported' % name
        assert name in dovs_b200._lib.SIGNATURES, '%s has no ctypes signature' % name
        assert len(dovs_b200._lib.SIGNATURES[name][1]) == len(args), '%s: arity differs from the header' % name
    assert set(dovs_b200._lib.SIGNATURES) == set(fns)
    assert lib.mgw_version() >= 100


def test_no_torch_types_in_the_abi():
    for name, args in header_functions().items():
        for a in args:
            assert re.match(r'^(const\s+)?(float|double|int32_t|uint8_t|void|int|size_t)\s*\*?\s*\w+$', a) or a.startswith('float '), (name, a)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, 'deep-online-video-stabilization_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                txt = open(os.path.join(dirpath, f)).read()
                assert 'c_oracle' not in txt and 'mesh_warp_ref' not in txt and 'oracle/' not in txt.replace('oracle/mgw_oracle.c', ''), f


@pytest.mark.skipif(torch.cuda.is_available(), reason='checks the no-GPU behaviour')
def test_fails_loudly_without_a_gpu():
    import dovs_b200
    with pytest.raises(RuntimeError, match='no CPU path'):
        dovs_b200.transformer(torch.zeros(1, 8, 8, 3), torch.zeros(1, 5, 5, 2))
    # and the raw ABI reports a CUDA error instead of computing anything
    rc = dovs_b200._lib.lib.mgw_solve_h_fwd(1, 1, 4, 4, 1, None)
    assert rc < 0 and dovs_b200._lib.lib.mgw_last_error()


def test_reference_import_lines_work_unchanged():
    """s_net_bundle_nobm.py:16 and train_bundle_nobm.py:16 import the operator by these exact statements; with the dropin/
    directory first on sys.path they resolve to this library ("a module pair importable under those names")."""
    import subprocess
    import sys
    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "from spatial_transformer3 import transformer\n"
        "t3 = transformer\n"
        "from spatial_transformer import *\n"
        "import dovs_b200, inspect\n"
        "assert t3 is dovs_b200.spatial_transformer3.transformer\n"
        "assert transformer is dovs_b200.spatial_transformer.transformer and interpolate is dovs_b200.spatial_transformer.interpolate\n"
        "assert list(inspect.signature(t3).parameters)[:2] == ['U', 'theta']\n"
        "assert list(inspect.signature(transformer).parameters)[:3] == ['U', 'theta', 'out_size']\n"
        "assert list(inspect.signature(interpolate).parameters)[:4] == ['im', 'x', 'y', 'out_size']\n"
        "print('ok')\n") % os.path.join(ROOT, 'deep-online-video-stabilization_b200', 'dropin')
    r = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, cwd='/tmp', timeout=300)
    assert r.returncode == 0 and r.stdout.strip().endswith('ok'), r.stderr[-2000:]
