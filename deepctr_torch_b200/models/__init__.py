from .basemodel import BaseModel
from .deepfm import DeepFM
from .xdeepfm import xDeepFM
from .fibinet import FiBiNET
from .dcn import DCN
from .dcnmix import DCNMix
from .wdl import WDL
from .nfm import NFM
from .afm import AFM
from .ifm import IFM
from .difm import DIFM

__all__ = ["BaseModel", "DeepFM", "xDeepFM", "FiBiNET", "DCN", "DCNMix", "WDL", "NFM", "AFM", "IFM", "DIFM"]
