"""Host-side mirror of the reference's model interface for the DiST path (registries, builders, modules)."""
