from .network import NeuralNetwork  # noqa: F401
from .rendering import Renderer  # noqa: F401
from .common import arange_pixels  # noqa: F401
from .loss import Loss  # noqa: F401
