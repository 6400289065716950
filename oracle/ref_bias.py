"""The reference's OWN two distance -> bias epilogues, executed from its source -- TEST INFRASTRUCTURE ONLY,
usable only where /root/reference exists (the build container); it produces tests/golden/bias_golden.npz.

model/geoformer/geoformer_fs.py cannot be imported (spconv, PG_OP and faiss are absent), but the two epilogues
are plain torch code inside it.  They are cut out of the file with `ast` and compiled unmodified:
  mask_heads_forward   geoformer_fs.py:263-300   the whole method.  Called with no dynamic-conv layers
                       (weights = biases = []) and zero-channel mask features, so that what it returns is exactly
                       its `relative_coords` (Q,3,N); `Tensor.cuda` is patched to the identity for the call
                       (:270 moves geo_dist to the GPU; there is none here).
  forward_decoder      the statements of geoformer_fs.py:680-702 (from `relative_coords = torch.abs(` to the
                       assignment `geo_dist_context[cond] = ...`), wrapped into a function of the names they read.
Nothing is restated: the statements that run are the reference's.
"""
import ast
import os
import textwrap

import torch
import torch.nn.functional as F

REF_FILE = "/root/reference/model/geoformer/geoformer_fs.py"


def available():
    return os.path.exists(REF_FILE)


def _class_methods():
    src = open(REF_FILE).read()
    tree = ast.parse(src)
    out = {}
    for node in ast.walk(tree):
        if isinstance(node, ast.ClassDef):
            for f in node.body:
                if isinstance(f, ast.FunctionDef):
                    out[f.name] = f
    return src, out


def load_mask_heads_forward():
    """-> f(geo_dist (Q,N), coords (N,3), fps_sampling_coords (Q,3)) = the reference's relative_coords (Q,3,N)"""
    src, methods = _class_methods()
    node = methods["mask_heads_forward"]
    code = textwrap.dedent(ast.get_source_segment(src, node))
    ns = {"torch": torch, "F": F}
    exec(compile(code, REF_FILE + ":mask_heads_forward", "exec"), ns)
    fn = ns["mask_heads_forward"]
    lines = (node.lineno, node.end_lineno)

    def run(geo_dist, coords, fps_sampling_coords):
        Q, N = geo_dist.shape
        saved = torch.Tensor.cuda
        torch.Tensor.cuda = lambda self, *a, **k: self
        try:
            x = fn(None, geo_dist.clone(), torch.zeros(N, 0, 1), [], [], Q, coords.clone(), fps_sampling_coords.clone())
        finally:
            torch.Tensor.cuda = saved
        return x.reshape(Q, 3, N)

    return run, lines


def load_decoder_relative_pos():
    """-> f(geo_dists list[(Q,N_b)], pre_enc_inds (B,C) int, query_locs (B,Q,3), context_locs (B,C,3))
    = the reference's geo_dist_context (B,Q,C,3) after line :702"""
    src, methods = _class_methods()
    node = methods["forward_decoder"]
    first = last = None
    for st in node.body:
        seg = ast.get_source_segment(src, st) or ""
        if first is None and seg.startswith("relative_coords = torch.abs("):
            first = st
        if seg.startswith("geo_dist_context[cond] ="):
            last = st
    assert first is not None and last is not None, "forward_decoder no longer has the expected statements"
    body = [st for st in node.body if first.lineno <= st.lineno <= last.lineno]
    text = "\n".join(textwrap.dedent(ast.get_source_segment(src, st)) for st in body)
    code = ("def decoder_lines(geo_dists, pre_enc_inds, query_locs, context_locs, batch_size):\n"
            + textwrap.indent(text, "    ") + "\n    return geo_dist_context\n")
    ns = {"torch": torch, "F": F}
    exec(compile(code, REF_FILE + ":forward_decoder[%d-%d]" % (first.lineno, last.end_lineno), "exec"), ns)
    fn = ns["decoder_lines"]

    def run(geo_dists, pre_enc_inds, query_locs, context_locs):
        return fn([g.clone() for g in geo_dists], pre_enc_inds.clone(), query_locs.clone(), context_locs.clone(),
                  context_locs.shape[0])

    return run, (first.lineno, last.end_lineno)
