from . import adapter, containers, factories, feedforward, normalizer, sampling_layers, types  # noqa: F401
