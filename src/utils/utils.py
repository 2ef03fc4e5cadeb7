"""Drop-in module path of the reference (`from src.utils.utils import zero_pad_model_input`)."""
from microbeseg_b200.utils import (get_nucleus_ids, min_max_normalization,  # noqa: F401
                                   zero_pad_model_input)
