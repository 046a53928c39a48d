"""Mirror of the reference's `src/models` for the denoisers on the sampling hot path."""
