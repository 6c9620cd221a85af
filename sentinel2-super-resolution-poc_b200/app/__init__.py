"""Mirror of the reference's ``server/app`` modules for the hot path only."""
