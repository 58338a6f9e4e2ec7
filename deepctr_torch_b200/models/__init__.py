from .basemodel import BaseModel
from .deepfm import DeepFM
from .xdeepfm import xDeepFM
from .fibinet import FiBiNET
from .dcn import DCN
from .dcnmix import DCNMix

__all__ = ["BaseModel", "DeepFM", "xDeepFM", "FiBiNET", "DCN", "DCNMix"]
