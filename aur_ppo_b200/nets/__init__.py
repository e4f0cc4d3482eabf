from .nets import continuous_net, critic, discrete_net, layer_init  # noqa: F401
