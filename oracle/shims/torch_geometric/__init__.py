"""Shim of the torch_geometric 2.2.0 symbols the reference model imports."""
