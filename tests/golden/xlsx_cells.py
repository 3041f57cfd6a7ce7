#!/usr/bin/env python3
"""Minimal .xlsx reader (zip + XML, no third-party package): cached cell values of one sheet.
usage: xlsx_cells.py book.xlsx [sheet index] -> prints cell = value (formula)"""
import re
import sys
import zipfile
import xml.etree.ElementTree as ET

NS = {"m": "http://schemas.openxmlformats.org/spreadsheetml/2006/main"}


def read_sheet(path, sheet_index=0):
    """{cell ref: (value, formula or None)}; strings resolved through sharedStrings."""
    z = zipfile.ZipFile(path)
    shared = []
    if "xl/sharedStrings.xml" in z.namelist():
        root = ET.fromstring(z.read("xl/sharedStrings.xml"))
        for si in root.findall("m:si", NS):
            shared.append("".join(t.text or "" for t in si.iter("{%s}t" % NS["m"])))
    sheets = sorted((n for n in z.namelist() if re.match(r"xl/worksheets/sheet\d+\.xml$", n)),
                    key=lambda n: int(re.findall(r"\d+", n)[-1]))
    root = ET.fromstring(z.read(sheets[sheet_index]))
    cells = {}
    for c in root.iter("{%s}c" % NS["m"]):
        ref, typ = c.get("r"), c.get("t")
        v, f = c.find("m:v", NS), c.find("m:f", NS)
        if v is None:
            continue
        val = v.text
        if typ == "s":
            val = shared[int(val)]
        elif typ in (None, "n"):
            val = float(val)
        cells[ref] = (val, f.text if f is not None else None)
    return cells


def col_index(col):
    n = 0
    for ch in col:
        n = n * 26 + ord(ch) - 64
    return n


def block(cells, top_left, bottom_right):
    """2-D list of values of a rectangular range, None where empty."""
    c0, r0 = re.match(r"([A-Z]+)(\d+)", top_left).groups()
    c1, r1 = re.match(r"([A-Z]+)(\d+)", bottom_right).groups()
    def name(i):
        s = ""
        while i:
            i, rem = divmod(i - 1, 26)
            s = chr(65 + rem) + s
        return s
    return [[cells.get(f"{name(c)}{r}", (None, None))[0] for c in range(col_index(c0), col_index(c1) + 1)]
            for r in range(int(r0), int(r1) + 1)]


if __name__ == "__main__":
    cells = read_sheet(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0)
    for ref in sorted(cells, key=lambda r: (int(re.findall(r"\d+", r)[0]), col_index(re.findall(r"[A-Z]+", r)[0]))):
        v, f = cells[ref]
        print(ref, "=", repr(v), f"   [{f}]" if f else "")
