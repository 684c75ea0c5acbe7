from .renderer import AggregationRenderer, DeviceWarp
from . import utils
from . import glm_compat
