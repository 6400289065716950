"""Drop-in for the reference's compiled package `pointnet2` (lib/pointnet2/setup.py:19-33):
`import geoformer_b200.pointnet2._ext as _ext` gives the nine operators of bindings.cpp:9-22.
See INTEGRATION.md for the one-line alias that makes `import pointnet2._ext` resolve here."""
from . import _ext  # noqa: F401
