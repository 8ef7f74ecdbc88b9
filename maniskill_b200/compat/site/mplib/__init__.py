"""Import-level stand-in (motion planning examples; outside the hot path)."""
