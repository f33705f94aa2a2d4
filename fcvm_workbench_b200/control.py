"""Control-file (.inp) reader for fcVM analyses.

The reference keeps one 21-line text file per model under ``control files/``;
``fcVM.FCMacro`` reads it line by line (reference: source code/fcVM.FCMacro:76-98).
The field order and the parse types below are the macro's; lines that are
missing at the end of older files (gnl / maxImp / ev1 / ev2) take the values the
reference GUI writes by default.
"""
from __future__ import annotations

import dataclasses
import os

# (name, parser) in file order -- reference: fcVM.FCMacro:77-97
_FIELDS = (
    ("sig_yield", float),
    ("grav_x", float),
    ("grav_y", float),
    ("grav_z", float),
    ("nstep", int),
    ("iterat_max", int),
    ("error_max", float),
    ("relax", float),
    ("scale_re", float),
    ("scale_up", float),
    ("scale_dn", float),
    ("disp_output", str),
    ("ultimate_strain", float),
    ("Et_E", float),
    ("target_LF", float),
    ("csr_option", str),
    ("averaged_option", str),
    ("gnl", str),
    ("maxImp", str),
    ("ev1", str),
    ("ev2", str),
)

_DEFAULT_TAIL = {"gnl": "GNLN", "maxImp": "0.0", "ev1": "1.0", "ev2": "0.0"}


@dataclasses.dataclass
class Control:
    sig_yield: float = 240.0
    grav_x: float = 0.0
    grav_y: float = 0.0
    grav_z: float = 0.0
    nstep: int = 10
    iterat_max: int = 20
    error_max: float = 1.0e-3
    relax: float = 1.2
    scale_re: float = 2.0
    scale_up: float = 1.2
    scale_dn: float = 1.2
    disp_output: str = "total"
    ultimate_strain: float = 0.25
    Et_E: float = 0.0
    target_LF: float = 2.0
    csr_option: str = "PEEQ"
    averaged_option: str = "unaveraged"
    gnl: str = "GNLN"
    maxImp: str = "0.0"
    ev1: str = "1.0"
    ev2: str = "0.0"

    def write(self, path: str) -> None:
        with open(path, "w", encoding="utf8") as f:
            for name, _ in _FIELDS:
                f.write(f"{getattr(self, name)}\n")


def read_control(path: str) -> Control:
    """Parse a fcVM control file exactly as the macro does (one value per line)."""
    with open(path, encoding="utf8") as f:
        lines = [ln.strip() for ln in f.readlines()]
    # the macro never skips blank lines; trailing blanks are simply absent fields
    while lines and lines[-1] == "":
        lines.pop()
    kw = {}
    for i, (name, conv) in enumerate(_FIELDS):
        if i < len(lines) and lines[i] != "":
            kw[name] = conv(lines[i])
        elif name in _DEFAULT_TAIL:
            kw[name] = _DEFAULT_TAIL[name]
        else:
            raise ValueError(f"{os.path.basename(path)}: missing control field '{name}' (line {i + 1})")
    return Control(**kw)
