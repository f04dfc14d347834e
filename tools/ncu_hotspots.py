#!/usr/bin/env python
"""Per-source-line view of an ncu capture: joins `ncu --page source --csv` (SASS order, samples,
executed counts) with nvdisasm's line info of the SAME binary, instruction by instruction.

    python tools/ncu_hotspots.py report.ncu-rep library.so kernel_regex [per_unit_divisor]
"""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile


def sass_lines(so, pattern):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
    out = []
    for cubin in sorted(os.listdir(tmp)):
        dis = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, cubin)], capture_output=True,
                             text=True).stdout
        fn, cur, keep = None, None, False
        for line in dis.splitlines():
            m = re.match(r"\s*\.text\.(\S+):", line)
            if m:
                fn = m.group(1)
                keep = re.search(pattern, fn) is not None
                cur = None
                continue
            m = re.search(r'//## File "([^"]+)", line (\d+)', line)
            if m:
                cur = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
            if m and keep:
                out.append((int(m.group(1), 16), cur, m.group(2).strip()))
    return out


def main():
    rep, so, pattern = sys.argv[1], sys.argv[2], sys.argv[3]
    div = float(sys.argv[4]) if len(sys.argv) > 4 else 1.0
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr = rows[1]
    i_src, i_ex, i_s = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    prof = []
    for r in rows[2:]:
        try:
            prof.append((r[i_src], int(r[i_ex]), int(r[i_s])))
        except (ValueError, IndexError):
            continue
    sass = sass_lines(so, pattern)
    if len(sass) != len(prof):
        sys.stderr.write("warning: %d profiled instructions vs %d disassembled (binary differs from the capture?)\n"
                         % (len(prof), len(sass)))
    by_line = collections.defaultdict(lambda: [0, 0, 0])
    for (src, ex, smp), (_, loc, _) in zip(prof, sass):
        a = by_line[loc]
        a[0] += ex
        a[1] += smp
        a[2] += 1
    tot_ex = sum(a[0] for a in by_line.values())
    tot_s = sum(a[1] for a in by_line.values())
    print("# %s | %s | executed %.0f (%.1f per unit), samples %d" % (rep, pattern, tot_ex, tot_ex / div, tot_s))
    print("# file:line  static  executed/unit  samples%  cycles-per-issue")
    for loc, a in sorted(by_line.items(), key=lambda kv: -kv[1][1])[:45]:
        print("%-28s %6d %12.1f %7.2f%% %8.2f" % ("%s:%s" % loc if loc else "?", a[2], a[0] / div, 100.0 * a[1] / tot_s,
                                                  (a[1] / tot_s) / max(a[0] / tot_ex, 1e-12)))


if __name__ == "__main__":
    main()
