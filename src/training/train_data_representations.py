"""Drop-in module path of the reference (`from src.training.train_data_representations import get_label`)."""
from microbeseg_b200.labels import distance_label, get_label  # noqa: F401
