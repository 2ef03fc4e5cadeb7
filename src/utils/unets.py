"""Drop-in module path of the reference (`from src.utils.unets import build_unet, get_weights`)."""
from microbeseg_b200.unets import DUNet, UNet, build_unet, get_weights  # noqa: F401
