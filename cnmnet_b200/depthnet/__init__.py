"""Mirror of the reference package ``depthnet`` (same module and symbol names)."""
