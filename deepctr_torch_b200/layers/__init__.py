from .core import DNN, PredictionLayer
from .interaction import (FM, CIN, CrossNet, CrossNetMix, SENETLayer, BilinearInteraction)

__all__ = ["DNN", "PredictionLayer", "FM", "CIN", "CrossNet", "CrossNetMix", "SENETLayer",
           "BilinearInteraction"]
