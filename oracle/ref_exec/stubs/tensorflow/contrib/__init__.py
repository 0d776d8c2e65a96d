from . import framework, layers, slim      # noqa: F401
