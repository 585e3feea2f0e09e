"""keras.layers.core stand-in (TEST INFRASTRUCTURE ONLY)."""
from keras.layers import Layer, Dense, Activation, Dropout, Lambda, Permute, Reshape  # noqa: F401
