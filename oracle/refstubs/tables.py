"""Empty stub for PyTables (only needed by the HDF5 LikelihoodLogger, not used by the harness)."""
