from .running_stats import WelfordRunningStat
from .metrics_logger import MetricsLogger
from .kbhit import KBHit
from . import reporting, torch_functions
from .torch_functions import compute_gae
