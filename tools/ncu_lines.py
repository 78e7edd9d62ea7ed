"""Per-source-line stall samples and executed instructions of one kernel launch in an ncu report (needs -lineinfo and
--import-source on).  usage: python tools/ncu_lines.py report.ncu-rep <kernel regex> [launch index] [top n]"""
import csv
import subprocess
import sys


def main():
    rep, rx = sys.argv[1], sys.argv[2]
    skip = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name", "regex:" + rx,
                          "--launch-skip", str(skip), "--launch-count", "1"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    fname, hdr, lines, cur = "", None, {}, None
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            fname = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            hdr = r
            i_s, i_i = hdr.index("# Samples"), hdr.index("Instructions Executed")
            continue
        if hdr is None or len(r) <= max(i_s, i_i):
            continue
        if r[0] != "":   # a CUDA source line; its SASS instructions follow
            cur = (fname, int(r[0]), r[1].strip()[:110])
            lines.setdefault(cur, [0, 0])
        elif cur is not None:
            try:
                lines[cur][0] += int(r[i_s]); lines[cur][1] += int(r[i_i])
            except ValueError:
                pass
    ts = sum(v[0] for v in lines.values()) or 1
    ti = sum(v[1] for v in lines.values()) or 1
    print("samples %d, warp instructions %d" % (ts, ti))
    for k, v in sorted(lines.items(), key=lambda kv: -kv[1][0])[:top]:
        print("%5.1f %% samples %5.1f %% instr  %s:%d  %s" % (100. * v[0] / ts, 100. * v[1] / ti, k[0], k[1], k[2]))


if __name__ == "__main__":
    main()
