"""`_target_: src.models.LSGAN_module.GANModule` (configs/model/LSGAN.yaml:1) -> B200 predict-side implementation."""
import use_b200  # noqa: F401
from use_b200.gan import GANModule  # noqa: F401
