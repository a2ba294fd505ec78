"""vihds_b200: B200-native batched ODE-integration + ELBO engine behind the plugin surface of microsoft/vi-hds.

The arithmetic of the hot path lives in ``csrc/libvihds_b200.so`` (C ABI: include/vihds_b200.h); this package is the
host-side mirror of the reference's Python interface (Config, Parameters, Encoder, Decoder, BaseVAE, build_model,
models.LOOKUP, Training.cost).  Importing the package does not load the library; the first kernel launch does, and
fails loudly if it is missing -- there is no CPU / PyTorch fallback for the hot path.
"""
__version__ = "0.1.0"
