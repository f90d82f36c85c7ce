from .renderer import PSNetwork, Network, Normal_Network, SGBasis  # noqa: F401
from ..pipeline import split_input, merge_output  # noqa: F401
from .train import train_step  # noqa: F401
