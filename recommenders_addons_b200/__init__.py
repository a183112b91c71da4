"""recommenders_addons_b200 -- B200-native (sm_100a) engine for the `tfra.dynamic_embedding` table hot
path of tensorflow/recommenders-addons.

    from recommenders_addons_b200 import dynamic_embedding as de

The product path is hand-written CUDA behind the C ABI in include/detable.h
(recommenders_addons_b200/lib/libdetable.so).  There is NO CPU fallback: importing the table classes
works anywhere, but every table operation needs the CUDA library and a GPU and fails loudly otherwise.
"""
from . import dynamic_embedding  # noqa: F401

__version__ = "0.1.0"
