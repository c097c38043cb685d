"""sbayes_b200: B200-native (sm_100a CUDA) backend for the sBayes MCMC inner loop.

Only the hot path of SURVEY.md section 8 lives here: packed layout, the C-ABI loader, and the
host-side mirror of the reference's `MCMCChain` / `Model.likelihood` / `Model.prior` /
MC3-swap interface.  There is no CPU fallback: every compute entry point needs the CUDA
library `sbayes_b200/csrc/libsbayes_b200.so` and a B200.
"""
__version__ = "0.1.0"
