"""Bindings of the CUDA library into the reference's own extension points."""
