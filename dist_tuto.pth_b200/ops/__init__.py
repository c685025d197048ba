"""Hand-written sm_100a kernels and their Python entry points (see csrc/)."""
