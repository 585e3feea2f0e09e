"""keras.layers.normalization stand-in (TEST INFRASTRUCTURE ONLY)."""
from keras.layers import BatchNormalization  # noqa: F401
