"""saev_b200 — B200-native (sm_100a) training step for saev sparse autoencoders."""

__version__ = "0.1.0"
