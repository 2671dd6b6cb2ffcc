"""cianna_b200 - B200-native (sm_100a) compute core behind CIANNA's C / Python API.

    from cianna_b200 import CIANNA as cnn      # same calls as upstream's `import CIANNA as cnn`

Native code lives in-tree: libcianna_b200.so (CUDA kernels + C-ABI, include/cianna_b200.h) and
libcianna_host.so (host C library).  Build with `python -m cianna_b200.build`.
"""
__version__ = "0.1.0"
