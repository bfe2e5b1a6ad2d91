from .detection import Detect
from .prior_box import PriorBox
