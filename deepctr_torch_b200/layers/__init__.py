from .core import DNN, PredictionLayer
from .interaction import (FM, CIN, CrossNet, CrossNetMix, SENETLayer, BilinearInteraction,
                          BiInteractionPooling, AFMLayer, InteractingLayer)

__all__ = ["DNN", "PredictionLayer", "FM", "CIN", "CrossNet", "CrossNetMix", "SENETLayer",
           "BilinearInteraction", "BiInteractionPooling", "AFMLayer", "InteractingLayer"]
