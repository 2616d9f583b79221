"""B200-native (sm_100a) implementation of the DQ-VAE stage-1 forward/backward hot path of
CrossmodalGroup/DynamicVectorQuantization.  See DESIGN.md / INTEGRATION.md."""
__version__ = "0.1.0"
