"""b200-cyclegan: a Blackwell-native CycleGAN training engine.

Capabilities and CLI of bryanlimy/tf2-cyclegan, re-designed for 8×B200:
PyTorch shell + hand-written sm_100a CUDA kernels + NVLink peer-memory
collectives. See README.md / DESIGN.md.
"""
__version__ = "0.1.0"
