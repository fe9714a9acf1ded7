#!/usr/bin/env python
"""Per-source-line stall samples / instruction shares / active lanes of one kernel in an .ncu-rep (needs -lineinfo).

    python tools/ncu_lines.py gpurun_out/prof.ncu-rep <kernel regex> [top_n]
"""
import collections
import csv
import subprocess
import sys

report, kernel = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
text = subprocess.run(["ncu", "-i", report, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name",
                       "regex:" + kernel], capture_output=True, text=True).stdout
per_line = collections.defaultdict(lambda: [0, 0, 0, ""])
header = current = source_file = None


def number(cell):
    try:
        return int(cell)
    except ValueError:
        return 0


for row in csv.reader(text.splitlines()):
    if row and row[0] == "File Path":
        source_file = row[1].split("/")[-1]
    elif row and row[0] == "Function Name":
        print(row[1][:120])
    elif row and row[0] == "Line No":
        header = row
        samples, executed, threads = (header.index(c) for c in ("# Samples", "Instructions Executed", "Thread Instructions Executed"))
    elif header is not None and len(row) > threads:
        if row[0].isdigit():
            current = (source_file, int(row[0]))
            per_line[current][3] = row[1]
        if row[2] and current is not None:
            per_line[current][0] += number(row[samples])
            per_line[current][1] += number(row[executed])
            per_line[current][2] += number(row[threads])
total_samples = sum(v[0] for v in per_line.values()) or 1
total_executed = sum(v[1] for v in per_line.values()) or 1
print("samples %d, warp instructions %d" % (total_samples, total_executed))
for (name, line), v in sorted(per_line.items(), key=lambda kv: -kv[1][int(__import__('os').environ.get('NCU_SORT','0'))])[:top]:
    print("%-18s %4d %5.1f%% samples %5.1f%% instr  lanes %4.1f | %s" % (name[:18], line, 100 * v[0] / total_samples,
                                                                         100 * v[1] / total_executed, v[2] / max(1, v[1]), v[3].strip()[:90]))
