"""TEST INFRASTRUCTURE -- not product code.

Loads the UNMODIFIED reference sources from /root/reference and runs them on
the torch-CPU `tensorflow` shim (oracle/tf_shim).  Only usable inside the build
container (the GPU box has no /root/reference): used by oracle/make_golden.py
to generate tests/golden/*.npz and by the CPU tests marked `reference`.

Nothing is copied: whole modules are imported from where they lie, and the
pieces of s_net_bundle_nobm.py / train_bundle_nobm.py that cannot be imported
(they build a ResNet on tf.contrib at import time) are pulled out of the parsed
AST *by name / name_scope* and exec'd verbatim.
"""
import ast
import contextlib
import importlib
import io
import os
import sys

REF = os.environ.get('MGW_REFERENCE_DIR', '/root/reference')
_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'tf_shim')


def available():
    return os.path.isfile(os.path.join(REF, 'spatial_transformer3.py'))


@contextlib.contextmanager
def _paths():
    added = [_SHIM, REF]
    for p in added:
        sys.path.insert(0, p)
    try:
        yield
    finally:
        for p in added:
            sys.path.remove(p)


def tf():
    with _paths():
        import tensorflow          # the shim
    assert getattr(tensorflow, 'NAMED', None) is not None, 'a real tensorflow shadowed the shim'
    return tensorflow


def _import(name):
    with _paths(), contextlib.redirect_stdout(io.StringIO()):
        return importlib.import_module(name)


def st3(height, width, grid_h, grid_w):
    """reference spatial_transformer3 with its config globals set (spatial_transformer3.py:16)."""
    m = _import('spatial_transformer3')
    m.height, m.width, m.grid_h, m.grid_w = height, width, grid_h, grid_w
    return m


def st1():
    m = _import('spatial_transformer')
    # the reference also imports spatial_transformer3 under `from spatial_transformer import *`
    return m


def quiet(fn, *a, **k):
    """the operator prints at call time (spatial_transformer3.py:224-226,274-276,297-300)."""
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def _parse(fname):
    with open(os.path.join(REF, fname)) as f:
        src = f.read()
    return ast.parse(src, filename=fname)


def _scope_name(node):
    """'x' for `with tf.name_scope('x'):` nodes, else None."""
    if not isinstance(node, ast.With) or len(node.items) != 1:
        return None
    call = node.items[0].context_expr
    if isinstance(call, ast.Call) and getattr(call.func, 'attr', None) == 'name_scope' and call.args \
            and isinstance(call.args[0], ast.Constant):
        return call.args[0].value
    return None


def _exec_nodes(nodes, ns, fname):
    mod = ast.Module(body=list(nodes), type_ignores=[])
    with contextlib.redirect_stdout(io.StringIO()):
        exec(compile(mod, os.path.join(REF, fname), 'exec'), ns)   # noqa: S102
    return ns


def s_net_namespace(height, width, grid_h, grid_w, batch_size, max_matches, do_crop_rate=0.8):
    """get_4_pts / warp_pts of s_net_bundle_nobm.py:29-71,215-230 as callables."""
    tree = _parse('s_net_bundle_nobm.py')
    ns = dict(tf=tf(), height=height, width=width, grid_h=grid_h, grid_w=grid_w,
              batch_size=batch_size, max_matches=max_matches, do_crop_rate=do_crop_rate)
    wanted = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in ('get_4_pts', 'warp_pts')]
    assert len(wanted) == 2
    return _exec_nodes(wanted, ns, 's_net_bundle_nobm.py')


def s_net_regularisers(ns):
    """get_black_pos, calc_distortion_loss, get_distortion_loss, get_consistency_loss of s_net_bundle_nobm.py:139-210 as
    callables, added to a namespace made by s_net_namespace()."""
    tree = _parse('s_net_bundle_nobm.py')
    names = ('get_black_pos', 'calc_distortion_loss', 'get_distortion_loss', 'get_consistency_loss')
    wanted = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in names]
    assert len(wanted) == 4
    return _exec_nodes(wanted, dict(ns), 's_net_bundle_nobm.py')


def s_net_black_loss(ns, pts1, use_black_loss):
    """the black_loss block of inference_stable_net (s_net_bundle_nobm.py:312-317) -> black_pos_loss"""
    tree = _parse('s_net_bundle_nobm.py')
    fn = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == 'inference_stable_net'][0]
    outer = [n for n in fn.body if isinstance(n, ast.With)][0]
    blocks = [n for n in outer.body if _scope_name(n) == 'black_loss']
    assert len(blocks) == 1
    ns = dict(ns)
    ns.update(pts1=pts1)
    t = ns['tf']
    saved = t.placeholder
    t.placeholder = lambda *a, **k: use_black_loss
    try:
        _exec_nodes(blocks, ns, 's_net_bundle_nobm.py')
    finally:
        t.placeholder = saved
    return ns['black_pos_loss'], ns['black_pos']


def s_net_losses(ns, matches, mask, flow, h_trans, y, black_pix):
    """feature_loss and img_loss blocks of inference_stable_net (s_net_bundle_nobm.py:335-352)."""
    tree = _parse('s_net_bundle_nobm.py')
    fn = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == 'inference_stable_net'][0]
    outer = [n for n in fn.body if isinstance(n, ast.With)][0]          # with tf.variable_scope('stable_net')
    blocks = [n for n in outer.body if _scope_name(n) in ('feature_loss', 'img_loss')]
    assert len(blocks) == 2
    ns = dict(ns)
    ns.update(matches=matches, mask=mask, flow=flow, h_trans=h_trans, y=y, black_pix=black_pix)
    t = ns['tf']
    saved = t.placeholder
    t.placeholder = lambda *a, **k: None                                # use_feature_loss placeholder is unused
    try:
        _exec_nodes(blocks, ns, 's_net_bundle_nobm.py')
    finally:
        t.placeholder = saved
    return ns['feature_loss'], ns['img_loss'], ns['stable_warpped']


def train_temp_loss(height, width, batch_size, ret1, ret2, flow, use_temp_loss=1.0):
    """temp_loss block of train_bundle_nobm.py:115-125 (interpolate from spatial_transformer.py)."""
    tree = _parse('train_bundle_nobm.py')
    block = [n for n in tree.body if _scope_name(n) == 'temp_loss']
    assert len(block) == 1
    t = tf()
    ns = dict(tf=t, height=height, width=width, batch_size=batch_size, ret1=ret1, ret2=ret2,
              interpolate=st1().interpolate, show_image=lambda *a, **k: None,
              x_flow=flow[..., 0:1], y_flow=flow[..., 1:2])
    saved = t.placeholder
    t.placeholder = lambda *a, **k: use_temp_loss
    try:
        _exec_nodes(block, ns, 'train_bundle_nobm.py')
    finally:
        t.placeholder = saved
    return ns['temp_loss']


def deploy_warp_rev_bundle2(height, width):
    """warpRevBundle2 of deploy_bundle.py:136-146 as a callable (the script itself cannot be imported: it parses argv and
    opens a TF session at import time).  Needs cv2."""
    import cv2
    tree = _parse('deploy_bundle.py')
    fn = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == 'warpRevBundle2']
    assert len(fn) == 1
    ns = dict(cv2=cv2, width=width, height=height)
    return _exec_nodes(fn, ns, 'deploy_bundle.py')['warpRevBundle2']


def config_cvt_img2train(height, width):
    """cvt_img2train of config.py:6-21 with the config globals height / width set (needs cv2 and Pillow)."""
    m = _import('config')
    m.height, m.width = height, width
    return m.cvt_img2train


def deploy_warp_rev_bundle(height, width, grid_h, grid_w):
    """warpRevBundle and cvt_theta_mat_bundle of deploy_bundle.py:121-134,148-173 as callables (needs cv2)."""
    import math
    import cv2
    import numpy as np
    tree = _parse('deploy_bundle.py')
    fns = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in ('warpRevBundle', 'cvt_theta_mat_bundle', 'warpRev', 'cvt_theta_mat')]
    assert len(fns) == 4
    ns = dict(cv2=cv2, np=np, math=math, width=width, height=height, grid_h=grid_h, grid_w=grid_w)
    ns = _exec_nodes(fns, ns, 'deploy_bundle.py')
    ns['warpRevBundle'].warpRev = ns['warpRev']              # the single-homography variant (:100-118) rides along
    return ns['warpRevBundle'], ns['cvt_theta_mat_bundle']


def deploy_stream_blocks():
    """The per-frame state handling of deploy_bundle.py as three code objects compiled from the reference's own statements
    (picked out of the `while(True)` loop by line range; the script cannot be imported):
      'assemble' :259-283  in_x from before_masks / before_frames / after_frames (+ tmp_in_x = in_x.copy())
      'refine'   :284-295  for j in range(args.refine): sess.run -> frame = img + black*(-1) -> tmp_in_x[..., -1] = frame
      'update'   :319-328  before_frames.append(frame) / before_masks.append(black) / pop(0)
    They run in a namespace that supplies np, args, the lists, height, width, input_mask, MaxSpan, black_mask, time and a
    stand-in `sess` with a run() method."""
    tree = _parse('deploy_bundle.py')
    loops = [n for n in ast.walk(tree) if isinstance(n, ast.While)]
    assert len(loops) == 1
    body = loops[0].body

    def pick(lo, hi):
        nodes = [n for n in body if lo <= n.lineno <= hi]
        assert nodes and nodes[0].lineno == lo, (lo, [n.lineno for n in nodes])
        return compile(ast.Module(body=nodes, type_ignores=[]), os.path.join(REF, 'deploy_bundle.py'), 'exec')

    return {'assemble': pick(259, 283), 'refine': pick(284, 295), 'update': pick(319, 328)}


def deploy_crop_block():
    """The crop search of deploy_bundle.py:344-365 (summed-area table + the four nested loops that leave `max_s`, `ans`) as a
    code object compiled from the reference's own statements in the `finally:` clause.  Runs in a namespace that supplies
    np, math, height, width and all_black (int64 [height,width]); prints progress (silence it with quiet())."""
    tree = _parse('deploy_bundle.py')
    tries = [n for n in ast.walk(tree) if isinstance(n, ast.Try) and any(m.lineno == 344 for m in n.finalbody)]
    assert len(tries) == 1
    nodes = [n for n in tries[0].finalbody if 344 <= n.lineno <= 365]
    assert [type(n).__name__ for n in nodes] == ['Assign', 'For', 'Assign', 'Assign', 'For'], [type(n).__name__ for n in nodes]
    return compile(ast.Module(body=nodes, type_ignores=[]), os.path.join(REF, 'deploy_bundle.py'), 'exec')
