"""Stall samples of an ncu source-page capture per C++ source line.

    ncu -i prof.ncu-rep --page source --csv > src.csv
    cuobjdump -xelf all pywatershed_b200/libpwsb200.so; nvdisasm -g -c pws_b200.sm_100a.cubin > all.dis
    python tools/ncu_source_lines.py src.csv all.dis <mangled kernel name> [csrc dir]

The CSV holds one row per SASS instruction (address, samples, stall reasons); the
disassembly with line info (-g, the library is built with -lineinfo) maps the
same addresses to file:line.  Prints the lines by samples.
"""
import collections
import csv
import pathlib as pl
import re
import sys


def main():
    src_csv, dis, kernel = sys.argv[1:4]
    csrc = pl.Path(sys.argv[4]) if len(sys.argv) > 4 else pl.Path(__file__).resolve().parent.parent / "pywatershed_b200" / "csrc"
    lines = open(dis).read().split("\n")
    start = next(i for i, l in enumerate(lines) if l.startswith(".text." + kernel))
    cur, addr2line = None, {}
    for l in lines[start + 1:]:
        if (l.startswith(".text.") or l.startswith("//-----")) and addr2line:
            break
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*)", l)
        if m and cur:
            addr2line[int(m.group(1), 16)] = cur
    rows = list(csv.reader(open(src_csv)))
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    data = rows[2:]
    base = int(data[0][ix["Address"]], 16)
    agg, tot = collections.Counter(), 0
    for r in data:
        s = int(r[ix["# Samples"]] or 0)
        tot += s
        agg[addr2line.get(int(r[ix["Address"]], 16) - base, ("?", 0))] += s
    text = {f.name: f.read_text().split("\n") for f in csrc.glob("*.cuh")}
    print(f"{kernel}: {tot} samples, {len(data)} SASS instructions")
    for (f, ln), s in agg.most_common(40):
        t = text[f][ln - 1].strip()[:90] if f in text and 0 < ln <= len(text[f]) else ""
        print(f"{f}:{ln:5d} {s:7d} {100 * s / tot:5.2f}%  {t}")


if __name__ == "__main__":
    main()
