"""Host-side mirror of the part of the reference's `datasets` package that feeds render_rays (ray_utils)."""
