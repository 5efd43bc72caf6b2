"""`_target_: src.data.loadwav_datamodule.LoadWavDataModule` -> minimal predict-side datamodule of use_b200."""
import use_b200  # noqa: F401
from use_b200.predict import LoadWavDataModule  # noqa: F401
