"""Drop-in module path of the reference (`from src.training.ranger2020 import Ranger`, train.py)."""
from microbeseg_b200.ranger import Ranger  # noqa: F401
