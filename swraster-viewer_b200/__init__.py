"""swraster-viewer_b200 — B200-native rasterisation hot path of swraster-viewer.

Layout: csrc/ (CUDA kernels + the C ABI of include/swr.h), host/ (C++ mirror of the
reference's Renderer / RenderCamera / RenderBuffer above the C ABI), and thin Python
wrappers (renderer.py) used by tests and bench.py. Import as `swraster_viewer_b200`.
"""
from . import abi  # noqa: F401
from .renderer import Renderer, RenderCamera, RenderBuffer, load_libraries, LibraryMissing  # noqa: F401
