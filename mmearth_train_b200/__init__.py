"""mmearth_train_b200 -- B200-native MP-MAE (FCMAE) pretraining step.

Host-side mirror of the reference's ``models.fcmae`` surface over the C-ABI library
``lib/libmpmae.so`` (``include/mpmae.h``).  Importing the package loads the library and raises if it
has not been built: there is no CPU or PyTorch fallback on the product path.
"""
from . import _native                                                  # noqa: F401  (fails loudly without the .so)
from . import checkpoint, convnextv2, data, dist, engine, graph, optim, synthetic         # noqa: F401
from .fcmae import (FCMAE, UncertaintyWeightingStrategy, convnextv2_atto, convnextv2_base, convnextv2_femto,  # noqa: F401
                    convnextv2_huge, convnextv2_large, convnextv2_nano, convnextv2_pico, convnextv2_tiny)

from .graph import GraphedStep                                         # noqa: F401

__all__ = ["FCMAE", "GraphedStep", "UncertaintyWeightingStrategy", "convnextv2_atto", "convnextv2_femto", "convnextv2_pico",
           "convnextv2_nano", "convnextv2_tiny", "convnextv2_base", "convnextv2_large", "convnextv2_huge"]
