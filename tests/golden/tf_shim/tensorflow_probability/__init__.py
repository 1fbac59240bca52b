"""Empty stand-in: the hot path never calls tensorflow_probability (only the out-of-scope
particle-Gibbs sampler does)."""
