"""Empty stand-in: saev.data.pe imports timm at module scope."""
