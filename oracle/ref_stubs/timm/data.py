"""Empty stand-in."""
