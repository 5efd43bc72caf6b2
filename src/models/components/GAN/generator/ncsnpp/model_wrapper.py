"""`_target_: src.models.components.GAN.generator.ncsnpp.model_wrapper.NCSNPP_Wrapper` -> B200 implementation."""
import use_b200  # noqa: F401
from use_b200.gan import NCSNPP_Wrapper  # noqa: F401
