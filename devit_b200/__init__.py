"""devit_b200 -- B200-native (sm_100a) implementation of DeViT's collaborative-inference hot path.

See DESIGN.md for the path / boundary, include/devit_b200.h for the C ABI.
"""
__version__ = "0.1.0"
