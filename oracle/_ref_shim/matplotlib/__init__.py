"""Stub: plotting is outside the hot path; any attribute access returns a no-op callable."""
