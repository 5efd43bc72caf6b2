"""`_target_: src.models.components.sgmse.model_wrapper.ScoreModel` -> B200 implementation."""
import use_b200  # noqa: F401
from use_b200.model_wrapper import ScoreModel, get_window, pad_spec  # noqa: F401
