"""Empty stand-in: saev.data.{siglip,fake_clip} import open_clip at module scope."""
