"""Import-only stand-in (the render modules are imported at module scope by the envs)."""
cm = colormaps = None
