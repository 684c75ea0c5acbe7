from .renderer import AggregationRenderer, SimpleRenderer, DeviceWarp
from . import utils
from . import glm_compat
