from .actor_critic import actor_critic  # noqa: F401
from .robot_actor_critic import robot_actor_critic  # noqa: F401
