"""use_b200 -- B200-native SGMSE reverse-SDE sampling path (drop-in for the reference's predict path).

Python host layer over libuse_b200.so (C ABI in include/use_b200.h).  Import name: ``use_b200`` (the directory
is ``universal-speech-enhancement_b200``; ``use_b200.py`` at the repo root aliases it).
"""
from . import _lib  # noqa: F401
from .backbones import BackboneRegistry, NCSNpp, NCSNppLarge  # noqa: F401
from .model_wrapper import ScoreModel, pad_spec  # noqa: F401
from .registry import Registry  # noqa: F401
from .sampling import CorrectorRegistry, PredictorRegistry, get_pc_sampler  # noqa: F401
from .sdes import OUVESDE, SDERegistry  # noqa: F401
from .sgmse_module import SGMSEModule  # noqa: F401
from .gan import GANModule, NCSNPP_Wrapper  # noqa: F401

__all__ = ["ScoreModel", "SGMSEModule", "NCSNpp", "NCSNppLarge", "BackboneRegistry", "SDERegistry", "PredictorRegistry",
           "CorrectorRegistry", "OUVESDE", "Registry", "get_pc_sampler", "pad_spec", "GANModule", "NCSNPP_Wrapper"]
