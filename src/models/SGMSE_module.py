"""`_target_: src.models.SGMSE_module.SGMSEModule` (configs/model/SGMSE_Large.yaml:1) -> B200 implementation."""
import use_b200  # noqa: F401
from use_b200.sgmse_module import SGMSEModule  # noqa: F401
