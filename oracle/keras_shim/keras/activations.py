"""keras.activations stand-in (TEST INFRASTRUCTURE ONLY)."""
import torch
tanh = torch.tanh
relu = torch.relu
sigmoid = torch.sigmoid
