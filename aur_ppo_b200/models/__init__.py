from .actor_critic import actor_critic  # noqa: F401
