"""Host-side mirror of the reference's `models` package (nerf.py, rendering.py)."""
