"""Drop-in module path of the reference (`from src.inference.postprocessing import ...`)."""
from microbeseg_b200.postprocessing import boundary_postprocessing, distance_postprocessing  # noqa: F401
