"""Run the reference's OWN geodesic code on CPU -- TEST INFRASTRUCTURE ONLY, build container only.

Loads /root/reference/model/geoformer/geodesic_utils.py unmodified through importlib (it imports
only torch) and feeds it a stand-in for the faiss index (`add/search/reset`, the protocol used at
geodesic_utils.py:18-21) backed by the canonical kNN of oracle.c.  CPU is the only deterministic
way to run it (scatter_ with duplicate indices, geodesic_utils.py:4-8).  /root/reference does not
exist on the GPU box, so this module is used only to (a) validate oracle.c:orc_geodesic and
(b) generate the committed fixtures in tests/golden/ (tests/golden/make_golden.py).
"""
import importlib.util
import os

import numpy as np
import torch

REF_FILE = "/root/reference/model/geoformer/geodesic_utils.py"


def available():
    return os.path.exists(REF_FILE)


def load_reference_module():
    spec = importlib.util.spec_from_file_location("_ref_geodesic_utils", REF_FILE)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class OracleFlatL2:
    """faiss.GpuIndexFlatL2 look-alike (CPU tensors): exact L2, squared distances, -1 padding."""

    def __init__(self):
        self.x = None

    def add(self, x):
        self.x = x.detach().cpu().numpy().astype(np.float32)

    def search(self, q, k, D_out, I_out):
        import oracle

        D2, I = oracle.knn_sq(self.x, k, q.detach().cpu().numpy())
        D_out.copy_(torch.from_numpy(D2))
        I_out.copy_(torch.from_numpy(I))

    def reset(self):
        self.x = None


class _TorchWithIEEESqrt:
    """`torch` as seen by the reference module, with a correctly rounded float32 sqrt.

    Why: on the reference's real platform (CUDA) torch.sqrt is IEEE correctly rounded.  On CPU,
    torch.sqrt of a large contiguous tensor goes through MKL VML whose result is off by one ulp for
    ~0.7 % of inputs (measured here: 34012 / 5e6), which would leak into every geodesic sum.  The
    bit-exact fixtures are therefore generated with this shim (numpy's sqrt is the hardware
    sqrtps = correctly rounded); the unmodified-torch run is also checked, to 1e-6 relative.
    """

    def __getattr__(self, name):
        return getattr(torch, name)

    @staticmethod
    def sqrt(t):
        return torch.from_numpy(np.sqrt(t.detach().cpu().numpy()))


def reference_cal_geodesic(pre_enc_inds, locs_float, batch_offsets, max_step, neighbor, radius, n_queries,
                           ieee_sqrt=True):
    """Returns the reference's list of (Q, N_b) float32 tensors, computed by the reference file."""
    mod = load_reference_module()
    if ieee_sqrt:
        mod.torch = _TorchWithIEEESqrt()
    torch.set_num_threads(os.cpu_count() or 1)
    return mod.cal_geodesic_vectorize(
        OracleFlatL2(), pre_enc_inds, locs_float, batch_offsets,
        max_step=max_step, neighbor=neighbor, radius=radius, n_queries=n_queries)
