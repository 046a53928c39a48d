"""Host-side helpers around the sampling hot path (mirror of the reference's `src/utils` for this path only)."""
