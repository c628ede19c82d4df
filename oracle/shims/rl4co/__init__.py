"""Stand-in for rl4co 0.6.0 (see ../README.md)."""
