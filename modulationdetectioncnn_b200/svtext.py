"""Reader/writer for the reference's fixed-point text fixtures.

Formats, all with 18-bit Q6.12 two's-complement
binary literals:

* data line   ``18'dADDR: data <= 18'bBITS;`` (weight files) or ``... data = ...``
  (test-vector files, ``DenseWeights1.txt``) - the grammar printed by
  /root/reference/CNN.ipynb:1 (cell 24) and pasted into the ``case`` ROMs of
  /root/reference/cnn_test_latest1.sv:122-142,685-707,719-3132;
* bare bias line ``18'bBITS`` (optional ``// comment``) - the dense-bias section,
  e.g. /root/reference/12.15.latestWeights.txt:14-18, pasted into
  cnn_test_latest1.sv:260-262;
* tables are separated by blank lines / ``* header`` lines; a new table starts
  when the address does not increase.

Weight file layout (12.14.weights.txt:24-30): conv table (3F entries:
``w0,w1,bias`` per filter) -> dense bias (C bare lines) -> 2C tables of
``129*F`` entries ordered class1-I, class1-Q, class2-I, ...
Vector file layout (cnn_test_latest1.sv:88-89,102): addresses 0..127 = I,
128..255 = Q.
"""
from __future__ import annotations

import re
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from .fixedpoint import WIDTH, bits_to_int, int_to_bits

__all__ = [
    "Table", "ParsedFile", "parse_text", "parse_file", "QWeights",
    "load_qweights", "load_vectors", "format_table", "write_qweights",
    "write_vector", "parse_sv_roms",
]

_DATA_RE = re.compile(r"^\s*(\d+)'d(\d+)\s*:\s*data\s*(<=|=)\s*(\d+)'b([01]+)\s*;")
_BARE_RE = re.compile(r"^\s*(\d+)'b([01]+)\s*;?\s*(//.*)?$")


@dataclass
class Table:
    """One run of ``ADDR: data`` lines with increasing addresses."""
    addrs: List[int] = field(default_factory=list)
    values: List[int] = field(default_factory=list)
    overwide: List[int] = field(default_factory=list)  # addresses whose literal was > width
    first_line: int = 0

    def dense(self, size: Optional[int] = None) -> np.ndarray:
        n = size if size is not None else (max(self.addrs) + 1 if self.addrs else 0)
        out = np.zeros(n, dtype=np.int32)
        for a, v in zip(self.addrs, self.values):
            if a < n:
                out[a] = v
        return out


@dataclass
class ParsedFile:
    tables: List[Table]
    bare: List[int]                 # bare ``18'b...`` lines, in file order
    bare_lines: List[int]
    headers: List[Tuple[int, str]]  # (line number, text) of ``* ...`` lines
    overwidth_policy: str
    n_overwide: int


def parse_text(text: str, overwidth: str = "verilog", width: int = WIDTH) -> ParsedFile:
    tables: List[Table] = []
    bare: List[int] = []
    bare_lines: List[int] = []
    headers: List[Tuple[int, str]] = []
    cur: Optional[Table] = None
    n_over = 0
    for ln, line in enumerate(text.splitlines(), 1):
        m = _DATA_RE.match(line)
        if m:
            addr, bits = int(m.group(2)), m.group(5)
            if cur is None or (cur.addrs and addr <= cur.addrs[-1]):
                cur = Table(first_line=ln)
                tables.append(cur)
            if len(bits) > width:
                n_over += 1
                cur.overwide.append(addr)
            cur.addrs.append(addr)
            cur.values.append(bits_to_int(bits, width, overwidth))
            continue
        m = _BARE_RE.match(line)
        if m:
            bits = m.group(2)
            if len(bits) > width:
                n_over += 1
            bare.append(bits_to_int(bits, width, overwidth))
            bare_lines.append(ln)
            cur = None
            continue
        s = line.strip()
        if s.startswith("*"):
            headers.append((ln, s))
            cur = None
        elif s and not s.startswith("//"):
            # free text such as "first table": acts as a separator
            cur = None
    return ParsedFile(tables, bare, bare_lines, headers, overwidth, n_over)


def parse_file(path: str, overwidth: str = "verilog") -> ParsedFile:
    with open(path, "r") as fh:
        return parse_text(fh.read(), overwidth)


# --------------------------------------------------------------------------- weights
@dataclass
class QWeights:
    """Literal ROM contents of the integer datapath (never re-quantised).

    conv_tab   int32 [3F]        ``[w0,w1,bias]`` per filter  (rom_cov, sv:685-707)
    dense_bias int32 [C]                                       (sv:260-262)
    dense_tabs int32 [2C,129F]   rows ordered c0-I, c0-Q, c1-I, ... (sv:719-3132)
    """
    conv_tab: np.ndarray
    dense_bias: np.ndarray
    dense_tabs: np.ndarray
    source: str = ""
    n_overwide: int = 0

    @property
    def filters(self) -> int:
        return self.conv_tab.shape[0] // 3

    @property
    def classes(self) -> int:
        return self.dense_tabs.shape[0] // 2

    def validate(self) -> "QWeights":
        F, C = self.filters, self.classes
        if self.conv_tab.shape != (3 * F,) or F < 1:
            raise ValueError("conv_tab must have 3*F entries")
        if self.dense_bias.shape != (C,):
            raise ValueError(f"dense_bias must have {C} entries, got {self.dense_bias.shape}")
        if self.dense_tabs.shape != (2 * C, 129 * F):
            raise ValueError(f"dense_tabs must be [{2*C},{129*F}], got {self.dense_tabs.shape}")
        for a in (self.conv_tab, self.dense_bias, self.dense_tabs):
            if a.min(initial=0) < -(1 << 17) or a.max(initial=0) >= (1 << 17):
                raise ValueError("value outside signed 18-bit range")
        return self


def load_qweights(path: str, *, conv_from: Optional[str] = None,
                  dense_bias: Optional[Sequence[int]] = None,
                  classes: int = 3, overwidth: str = "verilog") -> QWeights:
    """Load a ``*Weights.txt`` file.

    ``conv_from``: take the conv table and dense bias from another file (weight
    set B: conv+bias in ``12.14.weights.txt``, dense tables in
    ``12.15.denseWeights.txt``).  ``dense_bias``: supply the bias when the file
    has no bias section (``am.fm.qpsk.txt``).
    """
    pf = parse_file(path, overwidth)
    n_over = pf.n_overwide
    tables = list(pf.tables)
    conv = None
    bias = list(dense_bias) if dense_bias is not None else None
    if conv_from is not None:
        cf = parse_file(conv_from, overwidth)
        n_over += cf.n_overwide
        conv = cf.tables[0].dense()
        if bias is None and cf.bare:
            bias = cf.bare[:classes]
    elif tables and len(tables[0].addrs) % 3 == 0 and len(tables[0].addrs) < 129:
        conv = tables.pop(0).dense()
    if conv is None:
        raise ValueError(f"{path}: no conv table found (pass conv_from=...)")
    if conv_from is not None and tables and len(tables[0].addrs) < 129:
        tables.pop(0)  # the file's own conv table is superseded
    if bias is None:
        if len(pf.bare) < classes:
            raise ValueError(f"{path}: no dense-bias section (pass dense_bias=...)")
        bias = pf.bare[:classes]
    F = conv.shape[0] // 3
    dense = [t for t in tables if len(t.addrs) >= 129 * F]
    if len(dense) < 2 * classes:
        raise ValueError(f"{path}: expected {2*classes} dense tables of {129*F}, found {len(dense)}")
    dt = np.stack([t.dense(129 * F) for t in dense[: 2 * classes]])
    return QWeights(conv.astype(np.int32), np.asarray(bias, dtype=np.int32), dt.astype(np.int32),
                    source=path, n_overwide=n_over).validate()


def load_vectors(path: str, overwidth: str = "verilog") -> np.ndarray:
    """All 256-entry test vectors in a file -> int32 [n,256] (0-127 I, 128-255 Q)."""
    pf = parse_file(path, overwidth)
    vecs = [t.dense(256) for t in pf.tables if t.addrs and max(t.addrs) < 256 and len(t.addrs) > 9]
    if not vecs:
        raise ValueError(f"{path}: no test vector found")
    return np.stack(vecs).astype(np.int32)


# --------------------------------------------------------------------------- writers
def format_table(values: Sequence[int], assign: str = "<=", addr_digits: Optional[int] = None,
                 width: int = WIDTH) -> str:
    """``18'dNNN: data <= 18'b...;`` lines, in the reference grammar."""
    n = len(values)
    nd = addr_digits if addr_digits is not None else max(2, len(str(max(n - 1, 0))))
    return "\n".join(
        f"{width}'d{str(i).rjust(nd, '0')}: data {assign} {width}'b{int_to_bits(v, width)};"
        for i, v in enumerate(values))


def write_qweights(qw: QWeights, path: str) -> None:
    """Self-contained weight file in the layout of ``12.15.latestWeights.txt``."""
    C = qw.classes
    parts = ["* Convolution Bias + Weights:\n", format_table(qw.conv_tab, "<=", 2), "\n\n* Dense Bias:\n"]
    parts.append("\n".join(f"{WIDTH}'b{int_to_bits(v)}" for v in qw.dense_bias))
    parts.append(f"\n\n* Dense Weights ({2*C} Tables):\n")
    for t in range(2 * C):
        parts.append(format_table(qw.dense_tabs[t], "<=", 3))
        parts.append("\n")
    with open(path, "w") as fh:
        fh.write("\n".join(parts))


def write_vector(vec: Sequence[int], path: str, header: Optional[str] = None) -> None:
    """One 256-entry vector in the grammar of CNN.ipynb cell 24 (``data = ``)."""
    v = np.asarray(vec).reshape(-1)
    if v.shape[0] != 256:
        raise ValueError("a test vector has 256 entries (128 I then 128 Q)")
    with open(path, "w") as fh:
        if header:
            fh.write(f"* {header}\n\n")
        fh.write(format_table(v, "=", 3))


# --------------------------------------------------------------------------- SV ROMs
_MODULE_RE = re.compile(r"^\s*module\s+(\w+)")
_CASE_RE = re.compile(r"^\s*(\d+)'d(\d+)\s*:\s*data\s*(<=|=)\s*(\d+)'b([01]+)\s*;")
_ASSIGN_BIAS_RE = re.compile(r"assign\s+dense_bias\s*\[(\d)\]\s*=\s*18'b([01]+)\s*;")


def parse_sv_roms(path: str, overwidth: str = "verilog") -> Dict[str, np.ndarray]:
    """Pull the ``case`` ROM contents out of ``cnn_test_latest1.sv``.

    Returns ``{module_name: int32 table}`` for every module containing
    ``ADDR: data <= ...`` lines that are not commented out, plus
    ``"dense_bias"`` from the ``assign dense_bias[i]`` constants (sv:260-262).
    """
    roms: Dict[str, Dict[int, int]] = {}
    bias: Dict[int, int] = {}
    mod = None
    with open(path, "r") as fh:
        for line in fh:
            code = line.split("//")[0]
            m = _MODULE_RE.match(code)
            if m:
                mod = m.group(1)
                continue
            m = _ASSIGN_BIAS_RE.search(code)
            if m:
                bias[int(m.group(1))] = bits_to_int(m.group(2), WIDTH, overwidth)
                continue
            m = _CASE_RE.match(code)
            if m and mod:
                roms.setdefault(mod, {})[int(m.group(2))] = bits_to_int(m.group(5), WIDTH, overwidth)
    out: Dict[str, np.ndarray] = {}
    for name, d in roms.items():
        arr = np.zeros(max(d) + 1, dtype=np.int32)
        for a, v in d.items():
            arr[a] = v
        out[name] = arr
    if bias:
        out["dense_bias"] = np.array([bias[i] for i in sorted(bias)], dtype=np.int32)
    return out
