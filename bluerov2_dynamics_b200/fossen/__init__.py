"""Mirror of the reference's `fossen/` package: same module, class and function names, served by the CUDA engine."""
