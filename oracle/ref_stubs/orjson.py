"""Minimal stand-in for `orjson` so the read-only reference at /root/reference/src can be
imported in the build container (orjson is not installed and there is no network).
Test infrastructure only: used by oracle/gen_golden.py; never imported by saev_b200."""
import dataclasses
import enum
import json
import pathlib

OPT_APPEND_NEWLINE = 1
OPT_INDENT_2 = 2
OPT_SORT_KEYS = 4


def _default_chain(user_default):
    def _d(obj):
        if dataclasses.is_dataclass(obj) and not isinstance(obj, type):
            return dataclasses.asdict(obj)
        if isinstance(obj, enum.Enum):
            return obj.value
        if isinstance(obj, pathlib.PurePath):
            return str(obj)
        if user_default is not None:
            return user_default(obj)
        raise TypeError(f"not serialisable: {type(obj)}")

    return _d


def dumps(obj, default=None, option=0) -> bytes:
    option = option or 0
    s = json.dumps(
        obj,
        default=_default_chain(default),
        indent=2 if option & OPT_INDENT_2 else None,
        sort_keys=bool(option & OPT_SORT_KEYS),
        separators=None if option & OPT_INDENT_2 else (",", ":"),
    )
    if option & OPT_APPEND_NEWLINE:
        s += "\n"
    return s.encode()


def loads(b):
    return json.loads(b)
