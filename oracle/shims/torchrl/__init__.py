"""Stand-in for torchrl (only the spec classes the reference constructs in _make_spec)."""
