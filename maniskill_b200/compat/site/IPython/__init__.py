"""Import-level stand-in."""
