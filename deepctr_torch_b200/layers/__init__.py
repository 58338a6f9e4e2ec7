from .activation import Dice, Identity, activation_layer
from .core import DNN, PredictionLayer
from .interaction import (FM, CIN, CrossNet, CrossNetMix, SENETLayer, BilinearInteraction,
                          BiInteractionPooling, AFMLayer, InteractingLayer)

__all__ = ["Dice", "Identity", "activation_layer", "DNN", "PredictionLayer", "FM", "CIN", "CrossNet", "CrossNetMix", "SENETLayer",
           "BilinearInteraction", "BiInteractionPooling", "AFMLayer", "InteractingLayer"]
