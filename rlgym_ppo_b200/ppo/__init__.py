from .discrete_policy import DiscreteFF
from .value_estimator import ValueEstimator
from .experience_buffer import ExperienceBuffer
from .ppo_learner import PPOLearner
