from .base import SAM2AdapterConfig, cfgAMG, BaseAdapter, get_adapter  # noqa: F401
