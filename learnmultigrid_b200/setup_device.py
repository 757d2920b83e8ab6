"""Device-side hierarchy setup (SpGEMM Galerkin products, transposes, colour permutation, SELL build).
Filled in by the setup kernels of csrc/setup_kernels.cu."""
from . import _lib


def setup_device(h, A, Q_list, colors, dense_coarse_max):
    raise _lib.MgError("device setup kernels are not built yet; use setup='host'")


def download_level_matrix(h, l):
    raise _lib.MgError("level matrices were not kept on the host")
