"""B200-native LT-removert / LT-map hot path.

`binding` is the ctypes view of the C-ABI (include/ltr_b200.h); `removert` mirrors the reference's
Removerter / Session call graph on top of it.  CUDA-only: importing works anywhere, but creating a
context without libltr_b200.so or without an sm_100 GPU raises.
"""
from . import binding  # noqa: F401
from .binding import Context, LtrError, MODE_HD, MODE_ND, MODE_PD, reset_rimg_size  # noqa: F401
