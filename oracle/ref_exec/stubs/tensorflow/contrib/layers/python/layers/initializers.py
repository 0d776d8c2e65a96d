from tensorflow.contrib.layers import variance_scaling_initializer, xavier_initializer      # noqa: F401
