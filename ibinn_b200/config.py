""".ini handling: `default.ini` overlaid by an experiment file (reference main.py:11-16)."""
import configparser
import os

DEFAULT_INI = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'default.ini')


def make_args(ini=None, overrides=None, default_ini=None):
    """ConfigParser = default.ini, then `ini`, then {section: {key: value}} overrides."""
    args = configparser.ConfigParser()
    args.read(default_ini or DEFAULT_INI)
    if ini:
        if not os.path.isfile(ini):
            raise FileNotFoundError(ini)
        args.read(ini)
    for sec, kv in (overrides or {}).items():
        for k, v in kv.items():
            args[sec][k] = str(v)
    return args
