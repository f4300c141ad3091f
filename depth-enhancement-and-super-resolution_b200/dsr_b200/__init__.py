"""dsr_b200 - B200-native training step of the image-guided depth-enhancement network.

Host-side mirror of the reference's interface for this path (``networks.define_G``,
``translation_network.define_Gen``, ``main_model.MainModel`` with ``set_input`` /
``optimize_parameters``, ``norms.SurfaceNormals*``) over the C-ABI library ``libdsr_b200.so``
(``include/dsr_b200.h``).  Import never touches the GPU; the first op call loads the library and
fails loudly if it is not built or the tensors are not on a CUDA device.
"""
from . import _lib  # noqa: F401

__version__ = "0.1.0"
