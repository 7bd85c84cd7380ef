"""models package — B200 mirror of the reference src/models."""
