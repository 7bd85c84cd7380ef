"""utils package — host-side shims next to the synthesis path."""
