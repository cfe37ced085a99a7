"""Mirror of the reference's `intern` package (hot-path modules only)."""
