from .network import NeuralNetwork  # noqa: F401
from .rendering import Renderer  # noqa: F401
from .common import arange_pixels, get_tensor_values, sample_patch_points  # noqa: F401
from .loss import Loss  # noqa: F401
from .training import Trainer  # noqa: F401
