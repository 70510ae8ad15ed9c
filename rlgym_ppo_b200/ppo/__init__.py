from .discrete_policy import DiscreteFF
from .multi_discrete_policy import MultiDiscreteFF
from .continuous_policy import ContinuousPolicy
from .value_estimator import ValueEstimator
from .experience_buffer import ExperienceBuffer
from .ppo_learner import PPOLearner
