"""saev_b200 — B200-native (sm_100a) training step for saev sparse autoencoders."""

__version__ = "0.1.0"


def install(**kw):
    """Rebind saev's hot-path names to this package (see saev_b200/dropin.py)."""
    from .dropin import install as _install

    return _install(**kw)


def uninstall():
    from .dropin import uninstall as _uninstall

    return _uninstall()
